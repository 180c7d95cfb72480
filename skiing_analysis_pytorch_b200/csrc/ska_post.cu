// ska_post.cu - post-triangulation triage and temporal smoothing: the step the
// reference runs right after triangulation, as streaming kernels over the whole clip.
//
// Reference anchors (file:line relative to the reference checkout):
//   post_triage_single / post_triage_sequence   triangulation/postprocess.py:71-170
//     undistort pixels (cv2.undistortPoints(x, K, d, P=K), :93-98), pinhole reprojection error with the
//     1e-12 guard (:32-44), positive depth in both cameras (:47-52), confidence gate (:108-110),
//     error gate (:112), rejected joints -> NaN (:115-116), per-frame report (:118-124)
//   smooth_skeleton                             triangulation/postprocess.py:54-68
//     Savitzky-Golay (scipy.signal.savgol_filter, mode="interp") over the FINITE samples of every
//     (joint, coordinate) series, compacted in time and scattered back
//
// Arithmetic is fp64 per point like the reference's numpy (the inputs are fp32 and the undistorted
// pixels are rounded to fp32 exactly where cv2 returns fp32), so the accept / reject decisions at the
// thresholds agree with the reference's; traffic is 12 + 16 (+8) bytes in, 12 + 4 + 1 bytes out per joint.
#include <cuda_runtime.h>
#include <math.h>

#include "ska_internal.h"

namespace ska {

// ------------------------------------------------------------------------------------------------
struct TriageCam {
  double R[9], t[3];
  double fx, fy, cx, cy;   // undistortion normalises with these (cv2 ignores skew there)
  double P[12];            // K [R|t] with the full K (postprocess.py:28-29 build_P)
  double d[12];
  int undistort;
};

struct TriageArgs {
  TriageCam cam[2];
  int64_t N;
  const float* X;
  const float* kpts;  // (2, N, 2) view-major
  const float* conf;  // (2, N) or nullptr
  double conf_thr, err_thr;
  float* Xc;
  float* em;
  uint8_t* flags;
};

// cv2.undistortPoints(x, K, d, P=K): five fixed-point iterations (default criterion MAX_ITER = 5), fp32 out
__device__ __forceinline__ void undistort_px(const TriageCam& c, float u, float v, double& uo, double& vo) {
  const double x0 = ((double)u - c.cx) / c.fx, y0 = ((double)v - c.cy) / c.fy;
  double x = x0, y = y0;
  const double* k = c.d;  // k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4
#pragma unroll 1
  for (int it = 0; it < 5; ++it) {
    const double r2 = x * x + y * y;
    const double icdist = (1.0 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1.0 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
    if (icdist < 0.0) {  // cv2 gives up on the point and returns the distorted normalised coordinates
      x = x0;
      y = y0;
      break;
    }
    const double dx = 2.0 * k[2] * x * y + k[3] * (r2 + 2.0 * x * x) + k[8] * r2 + k[9] * r2 * r2;
    const double dy = k[2] * (r2 + 2.0 * y * y) + 2.0 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
    x = (x0 - dx) * icdist;
    y = (y0 - dy) * icdist;
  }
  uo = (double)(float)(c.fx * x + c.cx);  // cv2 returns float32 pixels for float32 input
  vo = (double)(float)(c.fy * y + c.cy);
}

__global__ void __launch_bounds__(256) post_triage_kernel(const __grid_constant__ TriageArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.N) return;
  const float Xf[3] = {__ldg(a.X + 3 * i), __ldg(a.X + 3 * i + 1), __ldg(a.X + 3 * i + 2)};
  const double X = Xf[0], Y = Xf[1], Z = Xf[2];
  double e[2];
  bool pos = true;
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    const TriageCam& c = a.cam[v];
    const float2 k = __ldg(reinterpret_cast<const float2*>(a.kpts + ((int64_t)v * a.N + i) * 2));
    double u = k.x, w = k.y;
    if (c.undistort) undistort_px(c, k.x, k.y, u, w);
    const double yx = c.P[0] * X + c.P[1] * Y + c.P[2] * Z + c.P[3];
    const double yy = c.P[4] * X + c.P[5] * Y + c.P[6] * Z + c.P[7];
    const double yz = c.P[8] * X + c.P[9] * Y + c.P[10] * Z + c.P[11] + 1e-12;
    const double du = yx / yz - u, dv = yy / yz - w;
    e[v] = sqrt(du * du + dv * dv);
    const double zc = c.R[6] * X + c.R[7] * Y + c.R[8] * Z + c.t[2];
    pos = pos && (zc > 0.0);
  }
  const double em = 0.5 * (e[0] + e[1]);
  bool cf = true;
  if (a.conf != nullptr) cf = ((double)__ldg(a.conf + i) >= a.conf_thr) && ((double)__ldg(a.conf + a.N + i) >= a.conf_thr);
  const bool eok = isfinite(em) && (em <= a.err_thr);
  const bool keep = pos && eok && cf;
  const float nanv = __int_as_float(0x7fc00000);
  a.Xc[3 * i] = keep ? Xf[0] : nanv;
  a.Xc[3 * i + 1] = keep ? Xf[1] : nanv;
  a.Xc[3 * i + 2] = keep ? Xf[2] : nanv;
  if (a.em != nullptr) a.em[i] = (float)em;
  if (a.flags != nullptr) a.flags[i] = (uint8_t)((pos ? 1 : 0) | (eok ? 2 : 0) | (cf ? 4 : 0) | (keep ? 8 : 0));
}

int post_triage(const SkaCamera* cams, const float* X, const float* kpts, const float* conf, int64_t T, int J, uint32_t flags,
                double conf_thr, double err_thr, float* Xc, float* em, uint8_t* fl, cudaStream_t s) {
  TriageArgs a;
  for (int v = 0; v < 2; ++v) {
    const SkaCamera& c = cams[v];
    TriageCam& o = a.cam[v];
    if (c.dist[12] != 0.0 || c.dist[13] != 0.0) return set_error(SKA_EUNSUPPORTED, "tilted sensor model (taux, tauy) is not implemented");
    for (int k = 0; k < 9; ++k) o.R[k] = c.R[k];
    for (int k = 0; k < 3; ++k) o.t[k] = c.t[k];
    o.fx = c.K[0];
    o.fy = c.K[4];
    o.cx = c.K[2];
    o.cy = c.K[5];
    for (int r = 0; r < 3; ++r)
      for (int m = 0; m < 4; ++m) {
        double acc = 0.0;
        for (int k = 0; k < 3; ++k) acc += c.K[3 * r + k] * (m < 3 ? c.R[3 * k + m] : c.t[k]);
        o.P[4 * r + m] = acc;
      }
    for (int k = 0; k < 12; ++k) o.d[k] = c.dist[k];
    o.undistort = (flags >> v) & 1u;
  }
  a.N = T * (int64_t)J;
  a.X = X;
  a.kpts = kpts;
  a.conf = conf;
  a.conf_thr = conf_thr;
  a.err_thr = err_thr;
  a.Xc = Xc;
  a.em = em;
  a.flags = fl;
  post_triage_kernel<<<(unsigned)((a.N + 255) / 256), 256, 0, s>>>(a);
  const cudaError_t ce = cudaGetLastError();
  return ce == cudaSuccess ? SKA_OK : set_error((int)ce, cudaGetErrorString(ce));
}

// per-frame counts of the four flag bits: THREAD per frame (a frame's J flag bytes are contiguous; a warp-per-frame
// shuffle reduction of 17 bytes cost 113 warp instructions per frame, this costs ~3)
__global__ void __launch_bounds__(128) flag_counts_kernel(const uint8_t* __restrict__ flags, int64_t T, int J, int32_t* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const uint8_t* f = flags + t * J;
  int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  for (int j = 0; j < J; ++j) {
    const unsigned b = f[j];
    c0 += b & 1u;
    c1 += (b >> 1) & 1u;
    c2 += (b >> 2) & 1u;
    c3 += (b >> 3) & 1u;
  }
  *reinterpret_cast<int4*>(out + 4 * t) = make_int4(c0, c1, c2, c3);
}

int flag_counts(const uint8_t* flags, int64_t T, int J, int32_t* out, cudaStream_t s) {
  if (T == 0) return SKA_OK;
  flag_counts_kernel<<<(unsigned)((T + 127) / 128), 128, 0, s>>>(flags, T, J, out);
  const cudaError_t ce = cudaGetLastError();
  return ce == cudaSuccess ? SKA_OK : set_error((int)ce, cudaGetErrorString(ce));
}

// ------------------------------------------------------------------------------------------------
// Savitzky-Golay over the finite samples of every series.  X is (T, S) row-major (S = J*3 series).
// Tiles of kSgRows frames are staged in shared memory (coalesced), one warp per series inside a tile
// ranks its finite samples with ballots.  Passes: count -> per-series scan of the tile counts ->
// compact (series-major) -> filter + scatter.
constexpr int kSgRows = 128;   // frames per tile (4 ballots per series)
constexpr int kSgMaxWin = 25;
constexpr int kSgThreads = 256;

struct SgWeights {
  int win, h;
  float w[kSgMaxWin][kSgMaxWin];  // row p: weights producing output position p of a window (rows < h: first-edge
};                                // interp; row h: interior; rows > h: last-edge interp)

// weights for all positions of one window: w_p = A (A^T A)^-1 a(p), A = Vandermonde in the centred sample index
static int sg_weights(int win, int poly, SgWeights& W) {
  if (win < 1 || win > kSgMaxWin || (win % 2) == 0) return set_error(SKA_EUNSUPPORTED, "Savitzky-Golay window must be odd and <= 25");
  if (poly < 0 || poly >= win || poly > 6) return set_error(SKA_EINVAL, "polyorder must be < window and <= 6");
  const int m = poly + 1, h = win / 2;
  double G[7][14];  // [A^T A | I] -> Gauss-Jordan
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) {
      double acc = 0.0;
      for (int z = 0; z < win; ++z) acc += pow((double)(z - h), i) * pow((double)(z - h), j);
      G[i][j] = acc;
      G[i][m + j] = (i == j) ? 1.0 : 0.0;
    }
  for (int c = 0; c < m; ++c) {
    int piv = c;
    for (int r = c + 1; r < m; ++r)
      if (fabs(G[r][c]) > fabs(G[piv][c])) piv = r;
    if (!(fabs(G[piv][c]) > 0.0)) return set_error(SKA_EINVAL, "singular Savitzky-Golay design matrix");
    for (int j = 0; j < 2 * m; ++j) {
      const double tmp = G[c][j];
      G[c][j] = G[piv][j];
      G[piv][j] = tmp;
    }
    const double inv = 1.0 / G[c][c];
    for (int j = 0; j < 2 * m; ++j) G[c][j] *= inv;
    for (int r = 0; r < m; ++r)
      if (r != c) {
        const double f = G[r][c];
        for (int j = 0; j < 2 * m; ++j) G[r][j] -= f * G[c][j];
      }
  }
  W.win = win;
  W.h = h;
  for (int p = 0; p < win; ++p)
    for (int z = 0; z < win; ++z) {
      double acc = 0.0;
      for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) acc += pow((double)(z - h), i) * G[i][m + j] * pow((double)(p - h), j);
      W.w[p][z] = (float)acc;
    }
  return SKA_OK;
}

// stage rows [f0, f0 + kSgRows) of X into shared memory, coalesced
__device__ __forceinline__ void sg_stage(const float* __restrict__ X, int64_t T, int S, int64_t f0, float* tile) {
  const int64_t n = ((T - f0) < kSgRows ? (T - f0) : kSgRows) * S;
  const float* src = X + f0 * S;
  for (int64_t q = threadIdx.x; q < (int64_t)kSgRows * S; q += blockDim.x) tile[q] = q < n ? src[q] : __int_as_float(0x7fc00000);
}

// pass A: counts[tile][s] = finite samples of series s in the tile
__global__ void __launch_bounds__(kSgThreads) sg_count_kernel(const float* __restrict__ X, int64_t T, int S, int32_t* __restrict__ counts) {
  extern __shared__ float tile[];
  const int64_t f0 = (int64_t)blockIdx.x * kSgRows;
  sg_stage(X, T, S, f0, tile);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int s = warp; s < S; s += kSgThreads / 32) {
    int c = 0;
#pragma unroll
    for (int r = 0; r < kSgRows / 32; ++r) {
      const float x = tile[(r * 32 + lane) * S + s];
      c += __popc(__ballot_sync(0xffffffffu, fabsf(x) <= 3.4e38f));
    }
    if (lane == 0) counts[(int64_t)blockIdx.x * S + s] = c;
  }
}

// pass B: per series, exclusive scan of the tile counts (in place) and the series total
__global__ void __launch_bounds__(1024) sg_scan_kernel(int32_t* __restrict__ counts, int64_t n_tiles, int S, int32_t* __restrict__ total) {
  __shared__ int32_t part[1024];
  const int s = blockIdx.x;
  const int64_t per = (n_tiles + blockDim.x - 1) / blockDim.x;
  const int64_t a = (int64_t)threadIdx.x * per, b = (a + per < n_tiles) ? a + per : n_tiles;
  int32_t sum = 0;
  for (int64_t i = a; i < b; ++i) sum += counts[i * S + s];
  part[threadIdx.x] = sum;
  __syncthreads();
  for (int o = 1; o < (int)blockDim.x; o <<= 1) {  // Hillis-Steele inclusive scan of the per-thread sums
    const int32_t v = (threadIdx.x >= (unsigned)o) ? part[threadIdx.x - o] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  int32_t run = part[threadIdx.x] - sum;  // exclusive prefix of this thread's segment
  for (int64_t i = a; i < b; ++i) {
    const int32_t c = counts[i * S + s];
    counts[i * S + s] = run;
    run += c;
  }
  if (threadIdx.x == blockDim.x - 1) total[s] = part[threadIdx.x];
}

// pass C (FILTER = false): comp[s][rank] = value;  pass D (FILTER = true): out = filtered / passthrough
template <bool FILTER>
__global__ void __launch_bounds__(kSgThreads) sg_apply_kernel(const float* __restrict__ X, int64_t T, int S,
                                                             const int32_t* __restrict__ offsets, const int32_t* __restrict__ total,
                                                             float* __restrict__ comp, float* __restrict__ out,
                                                             const __grid_constant__ SgWeights W) {
  extern __shared__ float tile[];
  const int64_t f0 = (int64_t)blockIdx.x * kSgRows;
  sg_stage(X, T, S, f0, tile);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int s = warp; s < S; s += kSgThreads / 32) {
    int base = offsets[(int64_t)blockIdx.x * S + s];
    const int n = total[s];
    const float* cs = comp + (int64_t)s * T;
#pragma unroll 1
    for (int r = 0; r < kSgRows / 32; ++r) {
      const int row = r * 32 + lane;
      const float x = tile[row * S + s];
      const bool fin = fabsf(x) <= 3.4e38f;
      const unsigned m = __ballot_sync(0xffffffffu, fin);
      const int rank = base + __popc(m & ((1u << lane) - 1u));
      base += __popc(m);
      if (!FILTER) {
        if (fin) comp[(int64_t)s * T + rank] = x;
      } else {
        float y = x;
        if (fin && n >= W.win) {
          // window of compacted samples that produces this output, and the output's position inside it
          int start = rank - W.h, p = W.h;
          if (rank < W.h) {
            start = 0;
            p = rank;
          } else if (rank >= n - W.h) {
            start = n - W.win;
            p = rank - start;
          }
          float acc = 0.f;
          for (int z = 0; z < W.win; ++z) acc = fmaf(W.w[p][z], cs[start + z], acc);
          y = acc;
        }
        tile[row * S + s] = y;
      }
    }
  }
  if (FILTER) {
    __syncthreads();
    const int64_t nq = ((T - f0) < kSgRows ? (T - f0) : kSgRows) * S;
    float* dst = out + f0 * S;
    for (int64_t q = threadIdx.x; q < nq; q += blockDim.x) dst[q] = tile[q];
  }
}

size_t savgol_workspace_bytes(int64_t T, int S) {
  const int64_t n_tiles = (T + kSgRows - 1) / kSgRows;
  size_t b = (size_t)n_tiles * S * sizeof(int32_t);          // tile counts -> offsets
  b = (b + 255) / 256 * 256 + (size_t)S * sizeof(int32_t);   // totals
  b = (b + 255) / 256 * 256 + (size_t)S * T * sizeof(float); // compacted series, series-major
  return b + 256;
}

int savgol(const float* X, int64_t T, int S, int win, int poly, float* out, void* ws, size_t ws_bytes, cudaStream_t s) {
  SgWeights W;
  const int rc = sg_weights(win, poly, W);
  if (rc != SKA_OK) return rc;
  if (S < 1 || S > 4096) return set_error(SKA_EINVAL, "1 <= series count <= 4096");
  if (ws_bytes < savgol_workspace_bytes(T, S)) return set_error(SKA_EWORKSPACE, "workspace too small (see ska_savgol_workspace_bytes)");
  if (T == 0) return SKA_OK;
  const int64_t n_tiles = (T + kSgRows - 1) / kSgRows;
  unsigned char* p = (unsigned char*)ws;
  int32_t* counts = (int32_t*)p;
  size_t off = ((size_t)n_tiles * S * sizeof(int32_t) + 255) / 256 * 256;
  int32_t* total = (int32_t*)(p + off);
  off = (off + (size_t)S * sizeof(int32_t) + 255) / 256 * 256;
  float* comp = (float*)(p + off);
  const size_t smem = (size_t)kSgRows * S * sizeof(float);
  if (smem > 200 * 1024) return set_error(SKA_EUNSUPPORTED, "too many series for the shared-memory tile (J*3 <= 400)");
  cudaError_t ce;
  if (smem > 48 * 1024) {
    if ((ce = cudaFuncSetAttribute(sg_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess ||
        (ce = cudaFuncSetAttribute(sg_apply_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess ||
        (ce = cudaFuncSetAttribute(sg_apply_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
      return set_error((int)ce, cudaGetErrorString(ce));
  }
  sg_count_kernel<<<(unsigned)n_tiles, kSgThreads, smem, s>>>(X, T, S, counts);
  sg_scan_kernel<<<S, 1024, 0, s>>>(counts, n_tiles, S, total);
  sg_apply_kernel<false><<<(unsigned)n_tiles, kSgThreads, smem, s>>>(X, T, S, counts, total, comp, nullptr, W);
  sg_apply_kernel<true><<<(unsigned)n_tiles, kSgThreads, smem, s>>>(X, T, S, counts, total, comp, out, W);
  ce = cudaGetLastError();
  return ce == cudaSuccess ? SKA_OK : set_error((int)ce, cudaGetErrorString(ce));
}

}  // namespace ska
