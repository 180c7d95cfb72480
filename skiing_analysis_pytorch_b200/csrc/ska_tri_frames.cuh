// ska_tri_frames.cuh - fused triangulation + reprojection scoring with PER-FRAME extrinsics (row a2:
// process_triangulate hands every frame its own R[i], T[i] - triangulation/triangulate.py:76-82).
//
// tri_kernel_cta_frames: the CTA organisation of tri_kernel_cta (ska_triangulate_impl.cuh) plus a per-stage CAMERA
// TABLE in shared memory.  For every group of NW 64-point tiles the producer warp
//   1. (lane 0) arms the stage's mbarrier with the byte count and issues the bulk copies of the keypoints,
//   2. (all 32 lanes, a frame per lane) reads the 96 V bytes [R|t] of every frame the group touches, chooses the frame's
//      conditioning origin and forms the centred fp32 hi/lo projection rows in fp64 - the arithmetic prep_camera /
//      default_centre run on the host for a static rig - into the stage's table (16 V + 4 floats per frame),
//   3. (lane 0) arrives on the barrier, which releases the table together with the keypoints.
// A consumer lane reads the table entries of its two points' frames (one entry for most pairs: 64 points span at most
// 64 / J + 2 frames), forms the DLT rows in scalar FFMAs from those registers and then runs the same packed pair
// arithmetic as the static kernel (fast_stage_core / score_core) with the conditioning origin and the depth row per
// point.  Nothing per frame ever goes to global memory: the 632-byte-per-frame workspace of the two-kernel form
// (tri_prep_frames + tri_frames_kernel, kept below for frame-major input, V > 4, skew / thin prism and tiny skeletons)
// is not needed on this path.
#pragma once
#include "ska_tri_point.cuh"

namespace ska {

// exact n / d for 0 <= n < 2^31 (round-up method; mul == 0 encodes a power-of-two divisor)
struct FastDiv {
  uint32_t mul, shift;
};
static inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  uint32_t s = 0;
  while ((1u << s) < d) ++s;
  if ((1u << s) == d) {
    f.mul = 0;
    f.shift = s;
  } else {
    f.mul = (uint32_t)((((uint64_t)1 << 32) * (((uint64_t)1 << s) - d)) / d + 1);
    f.shift = s;
  }
  return f;
}
__device__ __forceinline__ uint32_t fastdiv(uint32_t n, FastDiv f) {
  if (f.mul == 0) return n >> f.shift;
  return (__umulhi(f.mul, n) + n) >> f.shift;
}

constexpr int kFrMaxFrames = 128;  // table rows per stage; groups of NW * 64 points touch <= NW * 64 / J + 2 frames
constexpr int kFrMinJoints = 8;    // smaller skeletons take the two-kernel form

template <int V>
struct FrameRow {  // one frame's cameras, centred on the frame's own conditioning origin
  float P[V][16];  // Ph[0..11] (row-major 3x4 hi part), Pl[3], Pl[7], Pl[11] (lo parts of the translation column), pad
  float c[4];      // conditioning origin (fp32, as the static kernel holds it), pad
};

template <int V>
struct TriFramesParams {
  CamDev ck[V];        // static intrinsics / distortion (extrinsic members unused)
  double Kn[V][5];     // K / K[2][2]: fx, skew, cx, fy, cy (fp64, for the prep)
  SkaCamera cams[V];   // static K / dist for the cold path (R, t ignored)
  const double* Rt;    // (T, V, 12)
  uint32_t weight_sqrt, pinhole;
  int32_t J;
  FastDiv divJ;
  int64_t N, n_tiles;
  int64_t k_sV, c_sV;
  const float* kpts;
  const float* conf;
  float* X;
  float* err;
  float* proj;
  uint8_t* status;
};

// The frame's conditioning origin: the point closest (least squares) to all optical axes, regularised towards the
// mean camera centre - default_centre of ska_prep.h.  fp32 is enough: ANY point near the subject conditions the solve,
// and the fp32 value chosen here enters the fp64 translation column below exactly.
template <int V>
__device__ __forceinline__ void frame_centre(const double* Rt /*V x 12, registers*/, float c[3]) {
  float A00 = 0, A01 = 0, A02 = 0, A11 = 0, A12 = 0, A22 = 0, b0 = 0, b1 = 0, b2 = 0, m0 = 0, m1 = 0, m2 = 0;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const double* Rd = Rt + 12 * v;
    float R[9], t[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = (float)Rd[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) t[k] = (float)Rd[9 + k];
    const float C0 = -(R[0] * t[0] + R[3] * t[1] + R[6] * t[2]);
    const float C1 = -(R[1] * t[0] + R[4] * t[1] + R[7] * t[2]);
    const float C2 = -(R[2] * t[0] + R[5] * t[1] + R[8] * t[2]);
    const float d0 = R[6], d1 = R[7], d2 = R[8];
    m0 += C0 * (1.0f / V);
    m1 += C1 * (1.0f / V);
    m2 += C2 * (1.0f / V);
    const float p00 = 1.0f - d0 * d0, p01 = -d0 * d1, p02 = -d0 * d2, p11 = 1.0f - d1 * d1, p12 = -d1 * d2, p22 = 1.0f - d2 * d2;
    A00 += p00; A01 += p01; A02 += p02; A11 += p11; A12 += p12; A22 += p22;
    b0 += p00 * C0 + p01 * C1 + p02 * C2;
    b1 += p01 * C0 + p11 * C1 + p12 * C2;
    b2 += p02 * C0 + p12 * C1 + p22 * C2;
  }
  const float mu = 1e-4f * V;  // (parallel axes leave A singular along them: pull those directions to the mean camera centre)
  A00 += mu; A11 += mu; A22 += mu;
  b0 += mu * m0; b1 += mu * m1; b2 += mu * m2;
  // symmetric 3x3 solve by cofactors (A is positive definite thanks to mu)
  const float c00 = A11 * A22 - A12 * A12, c01 = A02 * A12 - A01 * A22, c02 = A01 * A12 - A02 * A11;
  const float c11 = A00 * A22 - A02 * A02, c12 = A01 * A02 - A00 * A12, c22 = A00 * A11 - A01 * A01;
  const float det = A00 * c00 + A01 * c01 + A02 * c02;
  float x0 = m0, x1 = m1, x2 = m2;
  if (fabsf(det) > 0.0f && fabsf(det) <= 3.0e38f) {
    const float id = 1.0f / det;
    x0 = (c00 * b0 + c01 * b1 + c02 * b2) * id;
    x1 = (c01 * b0 + c11 * b1 + c12 * b2) * id;
    x2 = (c02 * b0 + c12 * b1 + c22 * b2) * id;
  }
  c[0] = (fabsf(x0) <= 3.0e38f) ? x0 : 0.0f;
  c[1] = (fabsf(x1) <= 3.0e38f) ? x1 : 0.0f;
  c[2] = (fabsf(x2) <= 3.0e38f) ? x2 : 0.0f;
}

// One frame's table row: P' = Kn [R | R c + t] per view.  The translation column - the ~1e4-magnitude entries whose
// fp32 rounding sets the pixel-error floor - is formed in fp64 and split into hi / lo parts; the rotation columns are
// rounded to fp32 once (fp64 products, one rounding), like the static rig's.
template <int V>
__device__ __forceinline__ void frame_row(const TriFramesParams<V>& prm, const double* Rt /*V x 12, registers*/, FrameRow<V>& o) {
  float cf[3];
  frame_centre<V>(Rt, cf);
  const double c0 = (double)cf[0], c1 = (double)cf[1], c2 = (double)cf[2];
  float4* out = reinterpret_cast<float4*>(&o);
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const double* R = Rt + 12 * v;
    const double* t = R + 9;
    const double fx = prm.Kn[v][0], sk = prm.Kn[v][1], px = prm.Kn[v][2], fy = prm.Kn[v][3], py = prm.Kn[v][4];
    const float fxf = (float)fx, skf = (float)sk, pxf = (float)px, fyf = (float)fy, pyf = (float)py;
    float h[3][4];
#pragma unroll
    for (int m = 0; m < 3; ++m) {  // rotation columns: fp32 (a 1e-7 relative change of the camera, like rounding the fp64 product)
      const float r0 = (float)R[m], r1 = (float)R[3 + m], r2 = (float)R[6 + m];
      h[0][m] = fmaf(fxf, r0, fmaf(skf, r1, pxf * r2));
      h[1][m] = fmaf(fyf, r1, pyf * r2);
      h[2][m] = r2;
    }
    const double t0 = R[0] * c0 + R[1] * c1 + R[2] * c2 + t[0];
    const double t1 = R[3] * c0 + R[4] * c1 + R[5] * c2 + t[1];
    const double t2 = R[6] * c0 + R[7] * c1 + R[8] * c2 + t[2];
    const double p0 = fx * t0 + sk * t1 + px * t2, p1 = fy * t1 + py * t2, p2 = t2;
    h[0][3] = (float)p0;
    h[1][3] = (float)p1;
    h[2][3] = (float)p2;
    out[4 * v + 0] = make_float4(h[0][0], h[0][1], h[0][2], h[0][3]);
    out[4 * v + 1] = make_float4(h[1][0], h[1][1], h[1][2], h[1][3]);
    out[4 * v + 2] = make_float4(h[2][0], h[2][1], h[2][2], h[2][3]);
    out[4 * v + 3] = make_float4((float)(p0 - (double)h[0][3]), (float)(p1 - (double)h[1][3]), (float)(p2 - (double)h[2][3]), 0.f);
  }
  out[4 * V] = make_float4(cf[0], cf[1], cf[2], 0.f);
}

// the 12 V doubles [R|t] of one frame, global -> registers (128-bit loads)
template <int V>
__device__ __forceinline__ void load_frame_rt(const double* __restrict__ src, double* dst) {
  const double2* s2 = reinterpret_cast<const double2*>(src);
#pragma unroll
  for (int k = 0; k < 6 * V; ++k) {
    const double2 q = __ldg(s2 + k);
    dst[2 * k] = q.x;
    dst[2 * k + 1] = q.y;
  }
}

// DLT rows of one view of one point from a table row's registers (LO = 1: translation column restored from the lo parts)
__device__ __forceinline__ void rows_from_table(const float* P /*16*/, float u, float v, float a[4], float b[4]) {
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    a[m] = fmaf(u, P[8 + m], -P[m]);
    b[m] = fmaf(v, P[8 + m], -P[4 + m]);
  }
  a[3] += fmaf(u, P[14], -P[12]);
  b[3] += fmaf(v, P[14], -P[13]);
}

#ifndef SKA_FR_WARPS
#define SKA_FR_WARPS 10  // consumer warps; with the producers 12 warps x 168 registers (the per-point camera rows cost ~30 registers
                         // over the static kernel)
#endif
#ifndef SKA_FR_PRODUCERS
#define SKA_FR_PRODUCERS 2  // producer warps (measured: 10 + 2 at 4 stages slightly ahead of 9 + 3): building a group's camera table is ~1000 mostly dependent instructions (fp64 products,
                            // fp64 -> fp32 conversions), ~3 us for one warp against ~1 us of consumer time per group
#endif
#ifndef SKA_FR_MAXREG
#define SKA_FR_MAXREG 168
#endif
#ifndef SKA_FR_STAGES
#define SKA_FR_STAGES 4
#endif

template <int V, int NW, bool CONF, int STAGES>
struct FrSmem {
  static constexpr int kViewK = NW * 64 * 2;
  static constexpr int kViewC = CONF ? NW * 64 : 0;
  static constexpr int kKptFloats = V * kViewK;
  static constexpr int kStageFloats = kKptFloats + V * kViewC;
  static constexpr size_t oTab = (size_t)STAGES * kStageFloats * sizeof(float);
  static constexpr size_t oX = oTab + (size_t)STAGES * kFrMaxFrames * sizeof(FrameRow<V>);
  static constexpr size_t oFull = oX + (size_t)NW * 64 * 3 * sizeof(float);
  static constexpr size_t oEmpty = oFull + (size_t)STAGES * sizeof(uint64_t);
  static constexpr size_t oCold = oEmpty + (size_t)STAGES * sizeof(uint64_t);
  static constexpr size_t bytes = oCold + (size_t)NW * 32 * sizeof(uint32_t);
};

// The rare point (and the < 64-point tail): build the point's frame cameras with the host's own prep code and run the
// general per-point path; per-view outputs are written here, X is returned.
template <int V, bool CONF>
__device__ __noinline__ void frames_point_cold(const TriFramesParams<V>& prm, uint32_t i, float* Xout) {
  const uint32_t f = fastdiv(i, prm.divJ);
  SkaCamera cams[V];
#pragma unroll 1
  for (int v = 0; v < V; ++v) {
    cams[v] = prm.cams[v];
    const double* s = prm.Rt + ((int64_t)f * V + v) * 12;
    for (int k = 0; k < 9; ++k) cams[v].R[k] = s[k];
    for (int k = 0; k < 3; ++k) cams[v].t[k] = s[9 + k];
  }
  double c[3];
  default_centre(cams, V, c);
  CamDev cd[V];
  double P64[V][12];
#pragma unroll 1
  for (int v = 0; v < V; ++v) {
    int d = 0;
    const char* why = "";
    prep_camera(cams[v], c, prm.pinhole != 0, cd[v], P64[v], d, &why);
  }
  float u[1][V], vv[1][V], w2[1][V], du[1][V], dv[1][V], X[1][3];
  uint8_t st[1];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const float2 q = *reinterpret_cast<const float2*>(prm.kpts + (int64_t)k * prm.k_sV + 2 * (int64_t)i);
    u[0][k] = q.x;
    vv[0][k] = q.y;
    const float cf = (CONF && prm.conf != nullptr) ? prm.conf[(int64_t)k * prm.c_sV + i] : 1.0f;
    w2[0][k] = prm.weight_sqrt ? cf : cf * cf;
  }
  PointSource src;
  src.kpts = prm.kpts + 2 * (int64_t)i;
  src.conf = (CONF && prm.conf != nullptr) ? prm.conf + i : nullptr;
  src.k_sV = prm.k_sV;
  src.c_sV = prm.c_sV;
  src.weight_sqrt = prm.weight_sqrt;
  tri_points<V, 1, true, 2, kSolverSecular>(cd, P64, (float)c[0], (float)c[1], (float)c[2], u, vv, w2, src, X, du, dv, st);
#pragma unroll
  for (int k = 0; k < V; ++k) {
    if (prm.err != nullptr) prm.err[(int64_t)k * prm.c_sV + i] = sqrt_fast(fmaf(du[0][k], du[0][k], dv[0][k] * dv[0][k]));
    if (prm.proj != nullptr)
      *reinterpret_cast<float2*>(prm.proj + (int64_t)k * prm.k_sV + 2 * (int64_t)i) = make_float2(u[0][k] + du[0][k], vv[0][k] + dv[0][k]);
  }
  if (prm.status != nullptr) prm.status[i] = st[0];
  Xout[0] = X[0][0];
  Xout[1] = X[0][1];
  Xout[2] = X[0][2];
}

// the < 64-point tail of the clip: a thread per point (every lane of the last warp stays alive for the votes inside)
template <int V, bool CONF>
__global__ void __launch_bounds__(64) tri_frames_tail_kernel(const __grid_constant__ TriFramesParams<V> prm, uint32_t first) {
  const uint32_t N = (uint32_t)prm.N;
  const uint32_t i_raw = first + threadIdx.x;
  const bool live = i_raw < N;
  float X[3];
  frames_point_cold<V, CONF>(prm, live ? i_raw : N - 1, X);
  if (live) {
    prm.X[3 * (int64_t)i_raw] = X[0];
    prm.X[3 * (int64_t)i_raw + 1] = X[1];
    prm.X[3 * (int64_t)i_raw + 2] = X[2];
  }
}

template <int V, bool CONF, int DIST, int NW, int NP, int STAGES, bool LEAN>
__global__ void __maxnreg__(SKA_FR_MAXREG) tri_kernel_cta_frames(const __grid_constant__ TriFramesParams<V> prm) {
  static_assert(DIST <= 1, "skew / thin prism take the two-kernel form");
  using L = FrSmem<V, NW, CONF, STAGES>;
  constexpr int kPts = 64;
  extern __shared__ __align__(128) unsigned char fr_smem[];
  float* sK = reinterpret_cast<float*>(fr_smem);
  FrameRow<V>* sTab = reinterpret_cast<FrameRow<V>*>(fr_smem + L::oTab);  // [STAGES][kFrMaxFrames]
  float(*sX)[kPts * 3] = reinterpret_cast<float(*)[kPts * 3]>(fr_smem + L::oX);
  uint64_t* sFull = reinterpret_cast<uint64_t*>(fr_smem + L::oFull);
  uint64_t* sEmpty = reinterpret_cast<uint64_t*>(fr_smem + L::oEmpty);
  uint32_t(*sCold)[32] = reinterpret_cast<uint32_t(*)[32]>(fr_smem + L::oCold);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t n_wt = (uint32_t)prm.n_tiles;
  const uint32_t n_groups = (n_wt + NW - 1) / NW;
  if (threadIdx.x < STAGES) {
    mbar_init(sFull + threadIdx.x, 1);
    mbar_init(sEmpty + threadIdx.x, NW);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();

  if (warp >= NW) {
    // ---------------------------------------------------------------- producer warps: bulk copies + the stage's camera table
    // Producer p takes the CTA's groups p, p + NP, ... (group q of the CTA lives in stage q % STAGES).
    // A lane builds the table rows of frames f0 + lane and f0 + 32 + lane of the group (a group of NW tiles touches at most
    // NW * 64 / J + 2 <= 128 frames: up to four rounds, two in the common case).  The [R|t] of the NEXT group's first two
    // rounds is loaded into registers right after this group's rows are built, so its HBM latency hides behind the wait
    // for the stage.
    static_assert(STAGES % NP == 0, "a producer warp must always meet the same stages");
    const uint32_t pw = (uint32_t)(warp - NW);
    const uint32_t T = (uint32_t)(prm.N / prm.J);
    double rt0[12 * V], rt1[12 * V];
    auto prefetch = [&](uint32_t gq) {  // [R|t] of group gq's first two rounds -> registers (clamped: the loads are unconditional)
      if (gq < n_groups) {
        const uint32_t f = fastdiv(gq * NW * kPts, prm.divJ) + lane;
        load_frame_rt<V>(prm.Rt + (int64_t)(f < T ? f : T - 1) * (12 * V), rt0);
        load_frame_rt<V>(prm.Rt + (int64_t)(f + 32 < T ? f + 32 : T - 1) * (12 * V), rt1);
      }
    };
    prefetch(blockIdx.x + pw * gridDim.x);
    uint32_t q = pw;  // the CTA's q-th group
    for (uint32_t g = blockIdx.x + pw * gridDim.x; g < n_groups; g += NP * gridDim.x, q += NP) {
      const int st = (int)(q % STAGES);
      const uint32_t round = q / STAGES;
      const uint32_t t0 = g * NW;
      const uint32_t nt = (n_wt - t0 < (uint32_t)NW) ? (n_wt - t0) : (uint32_t)NW;
      const uint32_t f0 = fastdiv(t0 * kPts, prm.divJ), f1 = fastdiv((t0 + nt) * kPts - 1, prm.divJ);
      if (round > 0) mbar_wait(sEmpty + st, (round - 1) & 1u);  // (every lane waits: all of them write the table)
      float* stage = sK + (size_t)st * L::kStageFloats;
      if (lane == 0) {
        asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(sFull + st)),
                     "r"((uint32_t)V * nt * (uint32_t)(kPts * (CONF ? 12 : 8)))
                     : "memory");
#pragma unroll
        for (int k = 0; k < V; ++k) {
          bulk_g2s(stage + k * L::kViewK, prm.kpts + (int64_t)k * prm.k_sV + (int64_t)t0 * (kPts * 2), nt * (kPts * 8), sFull + st);
          if (CONF)
            bulk_g2s(stage + L::kKptFloats + k * L::kViewC, prm.conf + (int64_t)k * prm.c_sV + (int64_t)t0 * kPts, nt * (kPts * 4), sFull + st);
        }
      }
      FrameRow<V>* tab = sTab + (size_t)st * kFrMaxFrames;
      if (f0 + lane <= f1) frame_row<V>(prm, rt0, tab[lane]);
      if (f0 + 32 + lane <= f1) frame_row<V>(prm, rt1, tab[32 + lane]);
      for (uint32_t f = f0 + 64 + lane; f <= f1; f += 32) {  // small skeletons only
        load_frame_rt<V>(prm.Rt + (int64_t)f * (12 * V), rt0);
        frame_row<V>(prm, rt0, tab[f - f0]);
      }
      prefetch(g + NP * gridDim.x);
      __syncwarp();
      if (lane == 0) mbar_arrive(sFull + st);  // release: the table is complete; the phase ends when the copies have landed too
    }
    return;
  }

  // ------------------------------------------------------------------ consumer warps
  int st = 0;
  uint32_t par = 0;
  float* sx = sX[warp];
  uint32_t* cold = sCold[warp];
  uint32_t g = blockIdx.x;
  for (;;) {
    uint32_t n_cold = 0;
    for (; g < n_groups && n_cold < 32u; g += gridDim.x) {
      const uint32_t wt = g * NW + warp;
      if (wt >= n_wt) {
        g = n_groups;
        break;
      }
      mbar_wait(sFull + st, par);
      const float* stage = sK + (size_t)st * L::kStageFloats;
      const FrameRow<V>* tab = sTab + (size_t)st * kFrMaxFrames;
      const uint32_t i0 = wt * kPts + 2u * lane;
      const uint32_t f0 = fastdiv(g * NW * kPts, prm.divJ);
      const uint32_t ea = fastdiv(i0, prm.divJ) - f0, eb = fastdiv(i0 + 1, prm.divJ) - f0;
      F2 ut[V], vt[V], w2[V], a[V][4], b[V][4], zr[V][4], cc[3];
      {
        float a0[V][4], b0[V][4], a1[V][4], b1[V][4];
        float4 q[V];
#pragma unroll
        for (int k = 0; k < V; ++k) {
          q[k] = *reinterpret_cast<const float4*>(stage + k * L::kViewK + warp * (kPts * 2) + 4 * lane);
          ut[k] = mk2(q[k].x, q[k].z);
          vt[k] = mk2(q[k].y, q[k].w);
          if (CONF) {
            const float2 cf = *reinterpret_cast<const float2*>(stage + L::kKptFloats + k * L::kViewC + warp * kPts + 2 * lane);
            w2[k] = prm.weight_sqrt ? mk2(cf.x, cf.y) : mk2(cf.x * cf.x, cf.y * cf.y);
          } else {
            w2[k] = mk2(1.f, 1.f);
          }
        }
        // point 0: its frame's row; point 1: the same registers unless the pair straddles a frame boundary
        float E[V][16], c0[3], c1[3], z0[V][4], z1[V][4];
        const float4* ra = reinterpret_cast<const float4*>(tab + ea);
#pragma unroll
        for (int k = 0; k < V; ++k)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 e = ra[4 * k + j];
            E[k][4 * j] = e.x; E[k][4 * j + 1] = e.y; E[k][4 * j + 2] = e.z; E[k][4 * j + 3] = e.w;
          }
        {
          const float4 e = ra[4 * V];
          c0[0] = e.x; c0[1] = e.y; c0[2] = e.z;
        }
#pragma unroll
        for (int k = 0; k < V; ++k) {
          rows_from_table(E[k], q[k].x, q[k].y, a0[k], b0[k]);
#pragma unroll
          for (int m = 0; m < 4; ++m) z0[k][m] = E[k][8 + m];
        }
        c1[0] = c0[0]; c1[1] = c0[1]; c1[2] = c0[2];
        if (eb != ea) {
          const float4* rb = reinterpret_cast<const float4*>(tab + eb);
#pragma unroll
          for (int k = 0; k < V; ++k)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 e = rb[4 * k + j];
              E[k][4 * j] = e.x; E[k][4 * j + 1] = e.y; E[k][4 * j + 2] = e.z; E[k][4 * j + 3] = e.w;
            }
          const float4 e = rb[4 * V];
          c1[0] = e.x; c1[1] = e.y; c1[2] = e.z;
        }
#pragma unroll
        for (int k = 0; k < V; ++k) {
          rows_from_table(E[k], q[k].z, q[k].w, a1[k], b1[k]);
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            z1[k][m] = E[k][8 + m];
            a[k][m] = mk2(a0[k][m], a1[k][m]);
            b[k][m] = mk2(b0[k][m], b1[k][m]);
            zr[k][m] = mk2(z0[k][m], z1[k][m]);
          }
        }
#pragma unroll
        for (int m = 0; m < 3; ++m) cc[m] = mk2(c0[m], c1[m]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(sEmpty + st);
      if (++st == STAGES) {
        st = 0;
        par ^= 1u;
      }
      Sym4T<F2> M;
      FastStage<F2> fs;
      fast_stage_core<V, CONF, F2, F2>(cc[0], cc[1], cc[2], w2, a, b, M, fs);
      const bool good = mall(fs.conv) && mall(fs.well);
      if (!__all_sync(0xffffffffu, good)) {
        if (lane == 0) cold[n_cold] = wt;
        ++n_cold;
        continue;
      }
      F2 ra2[V], rb2[V];
      row_residuals<V, F2>(a, b, fs.y0, fs.y1, fs.y2, ra2, rb2);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const F2 z = vfma(zr[k][0], fs.y0, vfma(zr[k][1], fs.y1, vfma(zr[k][2], fs.y2, zr[k][3])));
        F2 du, dv;
        score_core<DIST, F2>(prm.ck[k], prm.ck[k], z, fs.y0, fs.y1, fs.y2, ra2[k], rb2[k], ut[k], vt[k], du, dv);
        if (LEAN || prm.err != nullptr) {
          const F2 e = sqrt_fast(vfma(du, du, vmul(dv, dv)));
          __stcs(reinterpret_cast<float2*>(prm.err + (int64_t)k * prm.c_sV + i0), make_float2(e.x, e.y));
        }
        if (!LEAN && prm.proj != nullptr) {
          const F2 pu = vadd(ut[k], du), pv = vadd(vt[k], dv);
          __stcs(reinterpret_cast<float4*>(prm.proj + (int64_t)k * prm.k_sV + 2 * (int64_t)i0), make_float4(pu.x, pv.x, pu.y, pv.y));
        }
      }
      if (!LEAN && prm.status != nullptr) *reinterpret_cast<uchar2*>(prm.status + i0) = make_uchar2(0, 0);
      const F2 X0 = vadd(fs.y0, cc[0]), X1 = vadd(fs.y1, cc[1]), X2 = vadd(fs.y2, cc[2]);
      *reinterpret_cast<float2*>(sx + lane * 6) = make_float2(X0.x, X1.x);
      *reinterpret_cast<float2*>(sx + lane * 6 + 2) = make_float2(X2.x, X0.y);
      *reinterpret_cast<float2*>(sx + lane * 6 + 4) = make_float2(X1.y, X2.y);
      __syncwarp();
      float* gx = prm.X + (int64_t)wt * (kPts * 3);
      __stcs(reinterpret_cast<float4*>(gx) + lane, reinterpret_cast<const float4*>(sx)[lane]);
      if (lane < 16) __stcs(reinterpret_cast<float4*>(gx) + 32 + lane, reinterpret_cast<const float4*>(sx)[32 + lane]);
      __syncwarp();
    }
    __syncwarp();
#pragma unroll 1
    for (uint32_t ci = 0; ci < n_cold; ++ci) {
      const uint32_t wt = cold[ci];
      frames_point_cold<V, CONF>(prm, wt * kPts + 2u * lane, sx + lane * 6);
      frames_point_cold<V, CONF>(prm, wt * kPts + 2u * lane + 1u, sx + lane * 6 + 3);
      __syncwarp();
      float* gx = prm.X + (int64_t)wt * (kPts * 3);
      __stcs(reinterpret_cast<float4*>(gx) + lane, reinterpret_cast<const float4*>(sx)[lane]);
      if (lane < 16) __stcs(reinterpret_cast<float4*>(gx) + 32 + lane, reinterpret_cast<const float4*>(sx)[32 + lane]);
      __syncwarp();
    }
    if (g >= n_groups) break;
  }
}

template <int V, bool CONF, int DIST, bool LEAN>
static cudaError_t launch_frames_impl(TriFramesParams<V>& prm, cudaStream_t stream) {
  constexpr int NW = SKA_FR_WARPS, NP = SKA_FR_PRODUCERS, STAGES = (V <= 2) ? SKA_FR_STAGES : NP, BLOCK = 32 * (NW + NP);  // (V = 3, 4: the tables are larger)
  auto kern = tri_kernel_cta_frames<V, CONF, DIST, NW, NP, STAGES, LEAN>;
  constexpr size_t smem = FrSmem<V, NW, CONF, STAGES>::bytes;
  static_assert(smem <= 227 * 1024, "tri_kernel_cta_frames staging does not fit the SM's shared memory");
  int dev = 0, sms = 0;
  cudaError_t ce = cudaGetDevice(&dev);
  if (ce != cudaSuccess) return ce;
  ce = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (ce != cudaSuccess) return ce;
  ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (ce != cudaSuccess) return ce;
  const int64_t n_groups = (prm.n_tiles + NW - 1) / NW;
  int64_t grid = sms;
  if (grid > n_groups) grid = n_groups;
  kern<<<(unsigned)grid, BLOCK, smem, stream>>>(prm);
  return cudaGetLastError();
}

// frames a group of NW tiles can touch: NW * 64 consecutive points
template <int NW>
static bool frames_table_fits(int J) {
  return J >= kFrMinJoints && (NW * 64 + J - 1) / J + 1 <= kFrMaxFrames;
}

}  // namespace ska
