// ska_ba_wide.cu - BA linearisation for C >= 3 cameras (reduced system too large for per-thread
// registers): the Schur accumulation is a skinny SYRK, done warp-cooperatively from shared memory.
//
// Each warp owns 32 points per tile.
//   phase 1  thread = point: pass A over the cameras builds the damped 3x3 point block, its Cholesky
//            and dp0 = -Hd^-1 gp (the point's own step with the cameras frozen); pass B re-projects
//            each free camera c and
//              - writes Y_c = L^-1 W_c (3 x 6) transposed into the warp's staging buffer
//                Yt[col][3*lane + k] (the only shared-memory staging: 16.8 KB per warp at C = 8),
//              - forms the camera's 33 per-point sums-to-be (Hcc upper triangle 21, gc 6, bw 6) in
//                registers and reduces them over the 32 points with a TRANSPOSING butterfly: 5
//                exchange steps of 16/8/4/2/1 values leave lane i holding the warp total of value i
//                (31 shuffles for 32 values instead of 160).
//   phase 2  lane = one 6x6 block (a, b), a <= b, of Sw = sum Y^T Y: 36 fp32 accumulators, operands
//            read as 128-bit shared loads of 4 consecutive rows of one column (padded column stride:
//            conflict free), 144 FMA per 12 loads.
// Every kFlush tiles the per-lane fp32 accumulators are folded into the CTA's fp64 packed reduced
// system in shared memory in a fixed warp order (no atomics, deterministic); the CTA writes one
// fp64 partial row, ba_reduce_columns (ska_ba.cu) sums the rows in a fixed order.
//
// The same kernel also runs C == 2 (SKA_BA_FORCE_WIDE) so the two paths can be tested against
// each other.
#include <cuda_runtime.h>

#include "ska_ba.cuh"
#include "ska_internal.h"

namespace ska {

struct BaKernelArgsW {
  ObsLayout lay;
  int64_t N;
  int64_t n_tiles;
  const float* x2d;
  const float* conf;
  const double* cams;
  const double* ctrl;
  const float* Xpp;
  double* partials;
};

#ifndef SKA_BA_WIDE_WARPS
#define SKA_BA_WIDE_WARPS 12  // 384 threads x <= 168 registers = the whole register file, one CTA per SM
#endif

template <int C>
struct Wide {
  static constexpr int NC = C - 1, n = 6 * NC, nS = n * (n + 1) / 2;
  static constexpr int oBw = nS, oGc = nS + n, oHcc = nS + 2 * n, oCost = oHcc + 21 * NC, oClamp = oCost + 1, size = oClamp + 1;
  static constexpr int NP = NC * (NC + 1) / 2;              // 6x6 block pairs a <= b
  static constexpr int SL_A = largest_div(24, 32 / NP);     // row slices in phase 2 (24 chunks of 4 rows)
  static constexpr int CH_A = 24 / SL_A;
  static constexpr int W = SKA_BA_WIDE_WARPS;               // warps per CTA
  static constexpr int warp_floats = n * kYStride;
  static constexpr size_t smem = (size_t)W * warp_floats * sizeof(float) + (size_t)size * sizeof(double) + C * sizeof(CamF);
};

template <int C>
__global__ void __launch_bounds__(32 * Wide<C>::W, 1) ba_linearize_wide_kernel(const BaKernelArgsW a) {
  using L = Wide<C>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* s_warp = reinterpret_cast<float*>(smem_raw);  // W * warp_floats, a multiple of 16 bytes
  double* s_red = reinterpret_cast<double*>(s_warp + (size_t)L::W * L::warp_floats);
  CamF* s_cam = reinterpret_cast<CamF*>(s_red + L::size);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* Yt = s_warp + (size_t)warp * L::warp_floats;

  if (threadIdx.x < C) load_cam(a.cams + threadIdx.x * kCamStride, s_cam[threadIdx.x]);
  for (int k = threadIdx.x; k < L::size; k += blockDim.x) s_red[k] = 0.0;
  __syncthreads();
  const float lam = (float)a.ctrl[kCtrlLambda];
  const int cur = (int)a.ctrl[kCtrlCur];
  const float* X = a.Xpp + (int64_t)cur * 3 * a.N;

  // phase-2 role: block pair and row slice
  int pa = 0, pb_ = 0;
  const int pair = lane % L::NP, slice_a = lane / L::NP;
  {
    int q = pair;
    for (int r = 0; r < L::NC; ++r) {
      const int len = L::NC - r;
      if (q < len) {
        pa = r;
        pb_ = r + q;
        break;
      }
      q -= len;
    }
  }
  const bool act_a = slice_a < L::SL_A;

  float accS[6][6], accH[L::NC], accX[L::NC];  // accH: lane i = entry i of the camera's 33 sums; accX: entry 32
#pragma unroll
  for (int k = 0; k < 6; ++k)
#pragma unroll
    for (int l = 0; l < 6; ++l) accS[k][l] = 0.f;
#pragma unroll
  for (int c = 0; c < L::NC; ++c) accH[c] = accX[c] = 0.f;
  float cost = 0.f, nclamp = 0.f;
  int since = 0;

  for (int64_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int64_t i = (tile * L::W + warp) * 32 + lane;
    const bool valid = i < a.N;
    // ---------------------------------------------------------------- phase 1
    float u[C], v[C], cw[C], Xp[3] = {0.f, 0.f, 1.f};
    if (valid) {
      int64_t koff, coff;
      obs_offsets(a.lay, i, koff, coff);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float2 q = __ldg(reinterpret_cast<const float2*>(a.x2d + koff + c * a.lay.k_sV));
        u[c] = q.x;
        v[c] = q.y;
        cw[c] = __ldg(a.conf + coff + c * a.lay.c_sV);
      }
      Xp[0] = __ldg(X + 3 * i);
      Xp[1] = __ldg(X + 3 * i + 1);
      Xp[2] = __ldg(X + 3 * i + 2);
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) u[c] = v[c] = cw[c] = 0.f;
    }
    PointBlock pb;
    pb_zero(pb);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      ObsLin ol;
      project_lin(s_cam[c], Xp, u[c], v[c], ol);
      float au[3], av[3];
      point_rows(s_cam[c], ol, au, av);
      pb_add(pb, cw[c], au, av, ol.eu, ol.ev);
      cost = fmaf(cw[c], fmaf(ol.eu, ol.eu, ol.ev * ol.ev), cost);
      nclamp += (valid && ol.clamped) ? 1.f : 0.f;
    }
    const Chol3 f = chol3_damped(pb, lam);
    float d0, d1, d2;  // dp0 = -Hd^-1 gp (zero for a point without a factor: chol3 zeroes it)
    {
      float y0, y1, y2;
      chol3_fwd(f, -pb.g0, -pb.g1, -pb.g2, y0, y1, y2);
      chol3_bwd(f, y0, y1, y2, d0, d1, d2);
      d0 = f.ok ? d0 : 0.f;
      d1 = f.ok ? d1 : 0.f;
      d2 = f.ok ? d2 : 0.f;
    }
    __syncwarp();  // previous tile's phase 2 is done with the staging buffer
#pragma unroll
    for (int c = 1; c < C; ++c) {
      ObsLin ol;
      project_lin(s_cam[c], Xp, u[c], v[c], ol);
      float au[3], av[3], bu[6], bv[6];
      point_rows(s_cam[c], ol, au, av);
      camera_rows(ol, bu, bv);
      const float w = cw[c];
      float s33[32], s32;
      const float adu = -fmaf(au[0], d0, fmaf(au[1], d1, au[2] * d2));
      const float adv = -fmaf(av[0], d0, fmaf(av[1], d1, av[2] * d2));
      int q = 0;
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        const float su = w * bu[r], sv = w * bv[r];
        float y0, y1, y2;  // a point without a factor has a zeroed one: Y = 0 without a branch
        chol3_fwd(f, fmaf(su, au[0], sv * av[0]), fmaf(su, au[1], sv * av[1]), fmaf(su, au[2], sv * av[2]), y0, y1, y2);
        float* yt = Yt + (6 * (c - 1) + r) * kYStride + 3 * lane;
        yt[0] = y0;
        yt[1] = y1;
        yt[2] = y2;
#pragma unroll
        for (int s2 = r; s2 < 6; ++s2, ++q) s33[q] = fmaf(su, bu[s2], sv * bv[s2]);  // Hcc upper triangle, 21 entries
        s33[21 + r] = fmaf(su, ol.eu, sv * ol.ev);                                   // gc
        if (r < 5) s33[27 + r] = fmaf(su, adu, sv * adv);                            // bw (entries 27..31)
        else s32 = fmaf(su, adu, sv * adv);                                          // bw[5] = entry 32
      }
      accH[c - 1] += transpose_reduce32(s33, lane);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s32 += __shfl_xor_sync(0xffffffffu, s32, o);
      accX[c - 1] += s32;
    }
    __syncwarp();
    // ---------------------------------------------------------------- phase 2: Sw block (pa, pb_)
    if (act_a) {
      // scalar FFMA on purpose: the packed form (rows 0-1 / 2-3 of a chunk in the halves of an F2, 6x3
      // half blocks) was measured 10 % slower - half as many independent accumulator chains
      const float* ya = Yt + (6 * pa) * kYStride + 4 * slice_a * L::CH_A;
      const float* yb = Yt + (6 * pb_) * kYStride + 4 * slice_a * L::CH_A;
#pragma unroll 2
      for (int ch = 0; ch < L::CH_A; ++ch) {
        float4 A4[6], B4[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          A4[k] = *reinterpret_cast<const float4*>(ya + k * kYStride + 4 * ch);
          B4[k] = *reinterpret_cast<const float4*>(yb + k * kYStride + 4 * ch);
        }
#pragma unroll
        for (int k = 0; k < 6; ++k)
#pragma unroll
          for (int l = 0; l < 6; ++l)
            accS[k][l] = fmaf(A4[k].x, B4[l].x, fmaf(A4[k].y, B4[l].y, fmaf(A4[k].z, B4[l].z, fmaf(A4[k].w, B4[l].w, accS[k][l]))));
      }
    }
    // ---------------------------------------------------------------- periodic fp64 fold
    const bool last = tile + gridDim.x >= a.n_tiles;
    if (++since == kFlush || last) {
      since = 0;
      double dc = (double)cost, dn = (double)nclamp;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        dc += __shfl_down_sync(0xffffffffu, dc, o);
        dn += __shfl_down_sync(0xffffffffu, dn, o);
      }
      cost = nclamp = 0.f;
      for (int w = 0; w < L::W; ++w) {
        __syncthreads();
        if (warp == w) {
          for (int sl = 0; sl < L::SL_A; ++sl) {
            if (act_a && slice_a == sl) {
#pragma unroll
              for (int k = 0; k < 6; ++k)
#pragma unroll
                for (int l = 0; l < 6; ++l) {
                  const int ra = 6 * pa + k, rb = 6 * pb_ + l;
                  if (rb >= ra) s_red[ra * L::n - (ra * (ra - 1)) / 2 + (rb - ra)] += (double)accS[k][l];
                }
            }
            __syncwarp();
          }
          // lane i holds entry i of every camera's [Hcc (21) | gc (6) | bw (6)]; entry 32 is on every lane
          const int o33 = lane < 21 ? L::oHcc + lane : (lane < 27 ? L::oGc + (lane - 21) : L::oBw + (lane - 27));
          const int s33 = lane < 21 ? 21 : 6;
#pragma unroll
          for (int c = 0; c < L::NC; ++c) {
            s_red[o33 + s33 * c] += (double)accH[c];
            if (lane == 0) s_red[L::oBw + 6 * c + 5] += (double)accX[c];
          }
          if (lane == 0) {
            s_red[L::oCost] += dc;
            s_red[L::oClamp] += dn;
          }
        }
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 6; ++k)
#pragma unroll
        for (int l = 0; l < 6; ++l) accS[k][l] = 0.f;
#pragma unroll
      for (int c = 0; c < L::NC; ++c) accH[c] = accX[c] = 0.f;
    }
  }
  __syncthreads();
  double* out = a.partials + (int64_t)blockIdx.x * L::size;
  for (int k = threadIdx.x; k < L::size; k += blockDim.x) out[k] = s_red[k];
}

template <int C>
static int launch_wide(const SkaBaProblem& in, BaKernelArgsW& a, cudaStream_t s) {
  using L = Wide<C>;
  auto kern = ba_linearize_wide_kernel<C>;
  static bool attr_set[64] = {};
  int dev = 0, sms = 0;
  cudaError_t ce = cudaGetDevice(&dev);
  if (ce == cudaSuccess) ce = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (ce != cudaSuccess) return set_error((int)ce, cudaGetErrorString(ce));
  if (dev < 64 && !attr_set[dev]) {  // idempotent; a benign race sets it twice
    ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::smem);
    if (ce != cudaSuccess) return set_error((int)ce, cudaGetErrorString(ce));
    attr_set[dev] = true;
  }
  int per_sm = 1;
  ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * L::W, L::smem);
  if (ce != cudaSuccess) return set_error((int)ce, cudaGetErrorString(ce));
  if (per_sm < 1) per_sm = 1;
  const int64_t per_tile = 32 * L::W;
  a.n_tiles = (a.N + per_tile - 1) / per_tile;
  int64_t grid = (int64_t)sms * per_sm;
  if (grid > a.n_tiles) grid = a.n_tiles;
  if (grid < 1) grid = 1;
  if (in.ws_bytes < (size_t)grid * L::size * sizeof(double)) return set_error(SKA_EWORKSPACE, "workspace too small (see ska_ba_workspace_bytes)");
  kern<<<(unsigned)grid, 32 * L::W, L::smem, s>>>(a);
  ce = cudaGetLastError();
  if (ce != cudaSuccess) return set_error((int)ce, cudaGetErrorString(ce));
  return launch_reduce(a.partials, (int)grid, L::size, in.d_red, s);
}

int ba_linearize_wide(const SkaBaProblem& in, cudaStream_t s) {
  BaKernelArgsW a;
  a.N = in.T * (int64_t)in.J;
  a.lay.J = in.J;
  if (in.layout == SKA_LAYOUT_FRAME_MAJOR) {
    a.lay.k_sV = 2 * (int64_t)in.J;
    a.lay.k_sT = 2 * (int64_t)in.J * in.C;
    a.lay.c_sV = in.J;
    a.lay.c_sT = (int64_t)in.J * in.C;
  } else {
    a.lay.k_sV = 2 * a.N;
    a.lay.k_sT = 2 * (int64_t)in.J;
    a.lay.c_sV = a.N;
    a.lay.c_sT = in.J;
  }
  a.x2d = in.d_x2d;
  a.conf = in.d_conf;
  a.cams = in.d_cams;
  a.ctrl = in.d_ctrl;
  a.Xpp = in.d_Xpp;
  a.partials = (double*)in.d_workspace;
  switch (in.C) {
    case 2: return launch_wide<2>(in, a, s);
    case 3: return launch_wide<3>(in, a, s);
    case 4: return launch_wide<4>(in, a, s);
    case 5: return launch_wide<5>(in, a, s);
    case 6: return launch_wide<6>(in, a, s);
    case 7: return launch_wide<7>(in, a, s);
    case 8: return launch_wide<8>(in, a, s);
    default: return set_error(SKA_EINVAL, "C must be in 2..8");
  }
}

}  // namespace ska
