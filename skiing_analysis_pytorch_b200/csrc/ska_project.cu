// ska_project.cu - standalone projection entry points (everything that is NOT the fused
// triangulate+reproject kernel or the LM): cv2-style reprojection of given 3D points with per-view
// pixel error, nan-aware per-frame error statistics, loss.py's differentiable projection and the
// fused reprojection loss + analytic gradient.  sm_100a; all of it is streaming, HBM-bound work:
// coalesced thread-per-point kernels, cameras through kernel parameters or L1, fp64 fixed-order
// reductions (no atomics).
//
// Reference anchors (file:line relative to the reference checkout):
//   reproject_points            triangulation/reproject.py:49-83; bundle_adjustment/reproject.py:74-153
//   error statistics            triangulation/reproject.py:243-261 (nan-aware, quirk Q6)
//   project_points              bundle_adjustment/loss.py:17-84
//   reprojection_loss           bundle_adjustment/loss.py:90-94 (backward = what torch.autograd derives)
#include <cuda_runtime.h>
#include <math.h>

#include "ska_internal.h"
#include "ska_project.cuh"

namespace ska {

constexpr int kPB = 256;

// ------------------------------------------------------------------------------------------------
// cv2-style reprojection of N = T*J given points into V cameras (+ optional pixel error)
struct CvArgs {
  CamCv64 cam[SKA_MAX_VIEWS];
  int32_t V, J;
  uint32_t frame_major;
  int64_t N;
  int64_t k_sV, k_sT, c_sV, c_sT;
  const float* X;
  const float* kpts;
  float* proj;
  float* err;
};

__global__ void __launch_bounds__(kPB) project_cv_kernel(const __grid_constant__ CvArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.N) return;
  int64_t koff = 2 * i, coff = i;
  if (a.frame_major) {
    const uint32_t t = (uint32_t)i / (uint32_t)a.J, j = (uint32_t)i - t * (uint32_t)a.J;
    koff = (int64_t)t * a.k_sT + 2 * (int64_t)j;
    coff = (int64_t)t * a.c_sT + (int64_t)j;
  }
  const double X = (double)__ldg(a.X + 3 * i), Y = (double)__ldg(a.X + 3 * i + 1), Z = (double)__ldg(a.X + 3 * i + 2);
  for (int v = 0; v < a.V; ++v) {
    double u, w;
    project_cv64(a.cam[v], X, Y, Z, u, w);
    const float uf = (float)u, wf = (float)w;  // cv2.projectPoints returns float32 pixels
    if (a.proj != nullptr) *reinterpret_cast<float2*>(a.proj + koff + (int64_t)v * a.k_sV) = make_float2(uf, wf);
    if (a.err != nullptr) {
      // |proj_f32 - float64(kpt)| like reproject.py:243-244 (quirk Q6)
      const float2 k = __ldg(reinterpret_cast<const float2*>(a.kpts + koff + (int64_t)v * a.k_sV));
      const double du = (double)uf - (double)k.x, dv = (double)wf - (double)k.y;
      a.err[coff + (int64_t)v * a.c_sV] = (float)sqrt(du * du + dv * dv);
    }
  }
}

int project_cv(const SkaCamera* cams, int V, const float* X, const float* kpts, int64_t T, int J, int layout,
               float* proj, float* err, cudaStream_t s) {
  CvArgs a;
  for (int v = 0; v < V; ++v) {
    const SkaCamera& c = cams[v];
    const double k22 = c.K[8];
    if (!(fabs(k22) > 0.0) || !isfinite(k22)) return set_error(SKA_EINVAL, "K[2][2] must be finite and non-zero");
    if (c.dist[12] != 0.0 || c.dist[13] != 0.0) return set_error(SKA_EUNSUPPORTED, "tilted sensor model (taux, tauy) is not implemented");
    for (int k = 0; k < 9; ++k) a.cam[v].R[k] = c.R[k];
    for (int k = 0; k < 3; ++k) a.cam[v].t[k] = c.t[k];
    a.cam[v].fx = c.K[0] / k22;  // cv2.projectPoints reads fx, fy, cx, cy only (skew ignored)
    a.cam[v].fy = c.K[4] / k22;
    a.cam[v].cx = c.K[2] / k22;
    a.cam[v].cy = c.K[5] / k22;
    for (int k = 0; k < 12; ++k) a.cam[v].d[k] = c.dist[k];
  }
  a.V = V;
  a.J = J;
  a.N = T * (int64_t)J;
  const bool fm = layout == SKA_LAYOUT_FRAME_MAJOR;
  a.frame_major = fm ? 1u : 0u;
  if (fm) {
    a.k_sV = 2 * (int64_t)J;
    a.k_sT = 2 * (int64_t)J * V;
    a.c_sV = J;
    a.c_sT = (int64_t)J * V;
  } else {
    a.k_sV = 2 * a.N;
    a.k_sT = 2 * (int64_t)J;
    a.c_sV = a.N;
    a.c_sT = J;
  }
  a.X = X;
  a.kpts = kpts;
  a.proj = proj;
  a.err = err;
  project_cv_kernel<<<(unsigned)((a.N + kPB - 1) / kPB), kPB, 0, s>>>(a);
  const cudaError_t ce = cudaGetLastError();
  return ce == cudaSuccess ? SKA_OK : set_error((int)ce, cudaGetErrorString(ce));
}

// ------------------------------------------------------------------------------------------------
// nan-aware per-(frame, view) error statistics: rmse, mean, median, max  (np.nanmean / nanmedian /
// nanmax of reproject.py:254-261).  One warp per (frame, view); J values staged in shared memory,
// the median by rank counting (J <= 1024, O(J^2/32) compares per warp).
constexpr int kStatsMaxJ = 1024;
constexpr int kStatsWarps = 4;

__global__ void __launch_bounds__(32 * kStatsWarps) frame_stats_kernel(const float* __restrict__ err, int64_t T, int V, int J,
                                                                     int64_t e_sT, int64_t e_sV, float* __restrict__ out) {
  extern __shared__ float s_e[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * kStatsWarps + warp;  // = t * V + v
  if (row >= T * V) return;
  const int64_t t = row / V;
  const int v = (int)(row - t * V);
  const float* e = err + t * e_sT + (int64_t)v * e_sV;
  float* se = s_e + warp * J;
  double s1 = 0.0, s2 = 0.0;
  float mx = -INFINITY;
  int n = 0;
  for (int j = lane; j < J; j += 32) {
    const float x = e[j];
    se[j] = x;
    if (x == x) {
      s1 += (double)x;
      s2 += (double)x * (double)x;
      mx = fmaxf(mx, x);
      ++n;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    n += __shfl_xor_sync(0xffffffffu, n, o);
  }
  __syncwarp();
  // the two middle order statistics among the non-NaN values
  const int r_lo = (n - 1) / 2, r_hi = n / 2;
  float m_lo = 0.f, m_hi = 0.f;
  bool has_lo = false, has_hi = false;
  for (int j = lane; j < J; j += 32) {
    const float x = se[j];
    if (x == x) {
      int rank = 0;
      for (int k = 0; k < J; ++k) {
        const float y = se[k];
        rank += (y < x || (y == x && k < j)) ? 1 : 0;
      }
      if (rank == r_lo) {
        m_lo = x;
        has_lo = true;
      }
      if (rank == r_hi) {
        m_hi = x;
        has_hi = true;
      }
    }
  }
  // ranks are unique, so exactly one lane holds each middle value
  float a_lo = 0.f, a_hi = 0.f;
  const unsigned b_lo = __ballot_sync(0xffffffffu, has_lo), b_hi = __ballot_sync(0xffffffffu, has_hi);
  if (b_lo) a_lo = __shfl_sync(0xffffffffu, m_lo, __ffs(b_lo) - 1);
  if (b_hi) a_hi = __shfl_sync(0xffffffffu, m_hi, __ffs(b_hi) - 1);
  if (lane == 0) {
    float* o = out + row * 4;
    if (n == 0) {
      o[0] = o[1] = o[2] = o[3] = NAN;
    } else {
      o[0] = (float)sqrt(s2 / n);
      o[1] = (float)(s1 / n);
      o[2] = (a_lo == a_hi) ? a_lo : 0.5f * a_lo + 0.5f * a_hi;
      o[3] = mx;
    }
  }
}

// Small skeletons (J <= JM <= 32: COCO-17 is the common case): THREAD per (frame, view) row, the row in registers, ranks
// by J^2 register compares.  The warp-per-row kernel above spends 435 warp instructions on a 17-joint row (15 idle lanes,
// shared-memory rank loop); this one ~27 per row, and a warp's 32 rows are one contiguous 32 * J * 4-byte read.
template <int JM>
__global__ void __launch_bounds__(128) frame_stats_small_kernel(const float* __restrict__ err, int64_t T, int V, int J, int64_t e_sT,
                                                                int64_t e_sV, float* __restrict__ out) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // = t * V + v
  if (row >= T * V) return;
  const int64_t t = row / V;
  const int v = (int)(row - t * V);
  const float* e = err + t * e_sT + (int64_t)v * e_sV;
  float x[JM];
  double s1 = 0.0, s2 = 0.0;
  float mx = -INFINITY;
  int n = 0;
#pragma unroll
  for (int j = 0; j < JM; ++j) {
    x[j] = (j < J) ? e[j] : NAN;
    if (x[j] == x[j]) {
      s1 += (double)x[j];
      s2 += (double)x[j] * (double)x[j];
      mx = fmaxf(mx, x[j]);
      ++n;
    }
  }
  const int r_lo = (n - 1) / 2, r_hi = n / 2;
  float a_lo = 0.f, a_hi = 0.f;
#pragma unroll
  for (int j = 0; j < JM; ++j) {
    int rank = 0;
#pragma unroll
    for (int k = 0; k < JM; ++k) rank += (x[k] < x[j] || (x[k] == x[j] && k < j)) ? 1 : 0;  // NaN compares false: never counted
    const bool ok = x[j] == x[j];
    a_lo = (ok && rank == r_lo) ? x[j] : a_lo;
    a_hi = (ok && rank == r_hi) ? x[j] : a_hi;
  }
  float* o = out + row * 4;
  if (n == 0) {
    o[0] = o[1] = o[2] = o[3] = NAN;
  } else {
    o[0] = (float)sqrt(s2 / n);
    o[1] = (float)(s1 / n);
    o[2] = (a_lo == a_hi) ? a_lo : 0.5f * a_lo + 0.5f * a_hi;
    o[3] = mx;
  }
}

int frame_stats(const float* err, int64_t T, int J, int V, int layout, float* out, cudaStream_t s) {
  if (J > kStatsMaxJ) return set_error(SKA_EUNSUPPORTED, "frame statistics support J <= 1024");
  const bool fm = layout == SKA_LAYOUT_FRAME_MAJOR;
  const int64_t e_sT = fm ? (int64_t)J * V : J, e_sV = fm ? J : T * (int64_t)J;
  const int64_t rows = T * V;
  if (rows == 0) return SKA_OK;
  if (J <= 32) {
    const unsigned grid = (unsigned)((rows + 127) / 128);
    if (J <= 17) frame_stats_small_kernel<17><<<grid, 128, 0, s>>>(err, T, V, J, e_sT, e_sV, out);
    else frame_stats_small_kernel<32><<<grid, 128, 0, s>>>(err, T, V, J, e_sT, e_sV, out);
    const cudaError_t ce = cudaGetLastError();
    return ce == cudaSuccess ? SKA_OK : set_error((int)ce, cudaGetErrorString(ce));
  }
  const size_t smem = (size_t)kStatsWarps * J * sizeof(float);
  frame_stats_kernel<<<(unsigned)((rows + kStatsWarps - 1) / kStatsWarps), 32 * kStatsWarps, smem, s>>>(err, T, V, J, e_sT, e_sV, out);
  const cudaError_t ce = cudaGetLastError();
  return ce == cudaSuccess ? SKA_OK : set_error((int)ce, cudaGetErrorString(ce));
}

// ------------------------------------------------------------------------------------------------
// loss.py project_points: X (T,J,3) -> (T,C,J,2) in dtype S.  R/t/K may be shared over the clip
// (frame stride 0) or per frame; camera stride is 9 / 3 / 9 elements.
template <typename S>
struct LossCams {
  const S* R;
  const S* t;
  const S* K;
  int64_t R_sT, t_sT, K_sT;
};

template <typename S>
__global__ void __launch_bounds__(kPB) project_points_kernel(const S* __restrict__ X, int64_t T, int J, int C, LossCams<S> cm,
                                                            S* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * (int64_t)J) return;
  const int64_t t = i / J;
  const int j = (int)(i - t * J);
  const S Xp[3] = {X[3 * i], X[3 * i + 1], X[3 * i + 2]};
  for (int c = 0; c < C; ++c) {
    LossObs<S> o;
    project_loss<S>(cm.R + t * cm.R_sT + 9 * c, cm.t + t * cm.t_sT + 3 * c, cm.K + t * cm.K_sT + 9 * c, Xp, o);
    S* p = out + ((t * C + c) * (int64_t)J + j) * 2;
    p[0] = o.u;
    p[1] = o.v;
  }
}

template <typename S>
int project_points(const S* X, int64_t T, int J, int C, const S* R, int64_t R_sT, const S* t, int64_t t_sT, const S* K,
                   int64_t K_sT, S* out, cudaStream_t s) {
  const int64_t N = T * (int64_t)J;
  if (N == 0) return SKA_OK;
  LossCams<S> cm{R, t, K, R_sT, t_sT, K_sT};
  project_points_kernel<S><<<(unsigned)((N + kPB - 1) / kPB), kPB, 0, s>>>(X, T, J, C, cm, out);
  const cudaError_t ce = cudaGetLastError();
  return ce == cudaSuccess ? SKA_OK : set_error((int)ce, cudaGetErrorString(ce));
}
template int project_points<float>(const float*, int64_t, int, int, const float*, int64_t, const float*, int64_t, const float*, int64_t, float*, cudaStream_t);
template int project_points<double>(const double*, int64_t, int, int, const double*, int64_t, const double*, int64_t, const double*, int64_t, double*, cudaStream_t);

// ------------------------------------------------------------------------------------------------
// reprojection_loss forward + analytic gradient.
//   sums[0] = sum conf |proj - x2d|^2, sums[1] = sum conf, sums[2] = #clamped observations
//   loss = w * sums[0] / (sums[1] + 1e-6)
// Gradients are returned UNSCALED (d sums[0] / d param / 2 = sum conf J^T e); the caller multiplies
// by 2 w / (sums[1] + 1e-6) (and autograd's grad_output).
template <int NV, typename S>
__device__ __forceinline__ void block_reduce_rows(const S (&v)[NV], double* scratch, double* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double x = (double)v[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (lane == 0) scratch[warp * NV + i] = x;
  }
  __syncthreads();
  const int nw = (int)(blockDim.x >> 5);
  for (int i = threadIdx.x; i < NV; i += blockDim.x) {
    double s = 0.0;
    for (int w = 0; w < nw; ++w) s += scratch[w * NV + i];
    out[i] = s;
  }
  __syncthreads();
}

template <typename S>
struct LossArgs {
  const S* X;
  const S* x2d;   // (T,C,J,2)
  const S* conf;  // (T,C,J)
  int64_t T;
  int32_t J, C;
  LossCams<S> cm;
  S* gX;          // nullable (T,J,3)
  double* partials;
};

// pass 1: thread per point, loop over cameras: loss sums and gX
template <typename S>
__global__ void __launch_bounds__(kPB) loss_points_kernel(const LossArgs<S> a) {
  __shared__ double scratch[(kPB / 32) * 3];
  const int64_t N = a.T * (int64_t)a.J;
  S acc[3] = {S(0), S(0), S(0)};
  double dacc[3] = {0.0, 0.0, 0.0};
  int run = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
    const int64_t t = i / a.J;
    const int j = (int)(i - t * a.J);
    const S Xp[3] = {a.X[3 * i], a.X[3 * i + 1], a.X[3 * i + 2]};
    S g[3] = {S(0), S(0), S(0)};
    for (int c = 0; c < a.C; ++c) {
      const S* R = a.cm.R + t * a.cm.R_sT + 9 * c;
      const S* K = a.cm.K + t * a.cm.K_sT + 9 * c;
      LossObs<S> o;
      project_loss<S>(R, a.cm.t + t * a.cm.t_sT + 3 * c, K, Xp, o);
      const int64_t q = (t * a.C + c) * (int64_t)a.J + j;
      // conf == nullptr: vector-Jacobian mode, x2d holds the cotangent of proj (project_points' backward)
      const bool vjp = a.conf == nullptr;
      const S eu = vjp ? a.x2d[2 * q] : o.u - a.x2d[2 * q], ev = vjp ? a.x2d[2 * q + 1] : o.v - a.x2d[2 * q + 1];
      const S cw = vjp ? S(1) : a.conf[q];
      acc[0] += cw * (eu * eu + ev * ev);
      acc[1] += cw;
      acc[2] += o.clamped ? S(1) : S(0);
      if (a.gX != nullptr) {
        S gXc[3], gK[6];
        project_loss_adjoint<S>(K, o, cw * eu, cw * ev, gXc, gK);
        g[0] += R[0] * gXc[0] + R[3] * gXc[1] + R[6] * gXc[2];
        g[1] += R[1] * gXc[0] + R[4] * gXc[1] + R[7] * gXc[2];
        g[2] += R[2] * gXc[0] + R[5] * gXc[1] + R[8] * gXc[2];
      }
    }
    if (a.gX != nullptr) {
      a.gX[3 * i] = g[0];
      a.gX[3 * i + 1] = g[1];
      a.gX[3 * i + 2] = g[2];
    }
    if (++run == 16) {  // bound the run length of the low-precision accumulation
      run = 0;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        dacc[k] += (double)acc[k];
        acc[k] = S(0);
      }
    }
  }
  double v[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) v[k] = dacc[k] + (double)acc[k];
  block_reduce_rows<3, double>(v, scratch, a.partials + (int64_t)blockIdx.x * 3);
}

// pass 2a: cameras shared over the clip: blockIdx.y = camera, 18 sums over every point
//   [0..8] gR row-major, [9..11] gt, [12..17] gK rows 0 and 1
template <typename S>
__global__ void __launch_bounds__(kPB) loss_cam_shared_kernel(const LossArgs<S> a) {
  __shared__ double scratch[(kPB / 32) * 18];
  const int c = blockIdx.y;
  const int64_t N = a.T * (int64_t)a.J;
  S acc[18];
  double dacc[18];
#pragma unroll
  for (int k = 0; k < 18; ++k) {
    acc[k] = S(0);
    dacc[k] = 0.0;
  }
  int run = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
    const int64_t t = i / a.J;
    const int j = (int)(i - t * a.J);
    const S Xp[3] = {a.X[3 * i], a.X[3 * i + 1], a.X[3 * i + 2]};
    const S* K = a.cm.K + t * a.cm.K_sT + 9 * c;
    LossObs<S> o;
    project_loss<S>(a.cm.R + t * a.cm.R_sT + 9 * c, a.cm.t + t * a.cm.t_sT + 3 * c, K, Xp, o);
    const int64_t q = (t * a.C + c) * (int64_t)a.J + j;
    const bool vjp = a.conf == nullptr;
    const S cw = vjp ? S(1) : a.conf[q];
    const S eu = vjp ? a.x2d[2 * q] : o.u - a.x2d[2 * q], ev = vjp ? a.x2d[2 * q + 1] : o.v - a.x2d[2 * q + 1];
    S gXc[3], gK[6];
    project_loss_adjoint<S>(K, o, cw * eu, cw * ev, gXc, gK);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int k = 0; k < 3; ++k) acc[3 * r + k] += gXc[r] * Xp[k];
      acc[9 + r] += gXc[r];
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) acc[12 + k] += gK[k];
    if (++run == 16) {
      run = 0;
#pragma unroll
      for (int k = 0; k < 18; ++k) {
        dacc[k] += (double)acc[k];
        acc[k] = S(0);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 18; ++k) dacc[k] += (double)acc[k];
  block_reduce_rows<18, double>(dacc, scratch, a.partials + ((int64_t)blockIdx.x * a.C + c) * 18);
}

// fixed-order sum of the per-CTA rows, scattered into gR (C,3,3) / gt (C,3) / gK (C,3,3; row 2 = 0)
template <typename S>
__global__ void loss_cam_finish_kernel(const double* __restrict__ partials, int rows, int C, S* gR, S* gt, S* gK) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * 18) return;
  double s = 0.0;
  for (int r = 0; r < rows; ++r) s += partials[(int64_t)r * C * 18 + idx];
  const int c = idx / 18, k = idx % 18;
  if (k < 9) {
    if (gR != nullptr) gR[9 * c + k] = (S)s;
  } else if (k < 12) {
    if (gt != nullptr) gt[3 * c + (k - 9)] = (S)s;
  } else if (gK != nullptr) {
    gK[9 * c + (k - 12)] = (S)s;
    if (k == 12) gK[9 * c + 6] = gK[9 * c + 7] = gK[9 * c + 8] = S(0);
  }
}

// pass 2b: per-frame cameras: one thread per (frame, camera) sums its J joints and writes
// gR (T,C,3,3) / gt (T,C,3) / gK (T,C,3,3) directly
template <typename S>
__global__ void __launch_bounds__(128) loss_cam_frames_kernel(const LossArgs<S> a, S* __restrict__ gR, S* __restrict__ gt,
                                                             S* __restrict__ gK) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // t * C + c
  if (row >= a.T * a.C) return;
  const int64_t t = row / a.C;
  const int c = (int)(row - t * a.C);
  const S* R = a.cm.R + t * a.cm.R_sT + 9 * c;
  const S* tv = a.cm.t + t * a.cm.t_sT + 3 * c;
  const S* K = a.cm.K + t * a.cm.K_sT + 9 * c;
  S acc[18];
#pragma unroll
  for (int k = 0; k < 18; ++k) acc[k] = S(0);
  for (int j = 0; j < a.J; ++j) {
    const int64_t i = t * a.J + j;
    const S Xp[3] = {a.X[3 * i], a.X[3 * i + 1], a.X[3 * i + 2]};
    LossObs<S> o;
    project_loss<S>(R, tv, K, Xp, o);
    const int64_t q = row * (int64_t)a.J + j;
    const bool vjp = a.conf == nullptr;
    const S cw = vjp ? S(1) : a.conf[q];
    const S eu = vjp ? a.x2d[2 * q] : o.u - a.x2d[2 * q], ev = vjp ? a.x2d[2 * q + 1] : o.v - a.x2d[2 * q + 1];
    S gXc[3], gKk[6];
    project_loss_adjoint<S>(K, o, cw * eu, cw * ev, gXc, gKk);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int k = 0; k < 3; ++k) acc[3 * r + k] += gXc[r] * Xp[k];
      acc[9 + r] += gXc[r];
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) acc[12 + k] += gKk[k];
  }
  if (gR != nullptr)
    for (int k = 0; k < 9; ++k) gR[row * 9 + k] = acc[k];
  if (gt != nullptr)
    for (int k = 0; k < 3; ++k) gt[row * 3 + k] = acc[9 + k];
  if (gK != nullptr) {
    for (int k = 0; k < 6; ++k) gK[row * 9 + k] = acc[12 + k];
    for (int k = 6; k < 9; ++k) gK[row * 9 + k] = S(0);
  }
}

static int loss_grid(int64_t N) {
  int g = ba_max_grid() / 2;
  const int64_t need = (N + kPB - 1) / kPB;
  if (g > need) g = (int)need;
  return g < 1 ? 1 : g;
}

size_t loss_workspace_bytes(int C) { return (size_t)(ba_max_grid() / 2) * (size_t)(18 * C > 3 ? 18 * C : 3) * sizeof(double); }

template <typename S>
int reprojection_loss(const S* X, int64_t T, int J, int C, const S* R, int64_t R_sT, const S* t, int64_t t_sT, const S* K,
                      int64_t K_sT, const S* x2d, const S* conf, double* sums, S* gX, S* gR, S* gt, S* gK, void* ws,
                      size_t ws_bytes, cudaStream_t s) {
  if (ws_bytes < loss_workspace_bytes(C)) return set_error(SKA_EWORKSPACE, "workspace too small (see ska_loss_workspace_bytes)");
  const int64_t N = T * (int64_t)J;
  LossArgs<S> a;
  a.X = X;
  a.x2d = x2d;
  a.conf = conf;
  a.T = T;
  a.J = J;
  a.C = C;
  a.cm = LossCams<S>{R, t, K, R_sT, t_sT, K_sT};
  a.gX = gX;
  a.partials = (double*)ws;
  const int grid = loss_grid(N);
  cudaError_t ce;
  loss_points_kernel<S><<<grid, kPB, 0, s>>>(a);
  if ((ce = cudaGetLastError()) != cudaSuccess) return set_error((int)ce, cudaGetErrorString(ce));
  int rc = launch_reduce(a.partials, grid, 3, sums, s);
  if (rc != SKA_OK) return rc;
  // camera gradients: a tensor shared over the clip (frame stride 0) is reduced over every point,
  // a per-frame tensor over the frame's joints only
  S* sR = (gR != nullptr && R_sT == 0) ? gR : nullptr;
  S* st = (gt != nullptr && t_sT == 0) ? gt : nullptr;
  S* sK = (gK != nullptr && K_sT == 0) ? gK : nullptr;
  S* fR = (gR != nullptr && R_sT != 0) ? gR : nullptr;
  S* ft = (gt != nullptr && t_sT != 0) ? gt : nullptr;
  S* fK = (gK != nullptr && K_sT != 0) ? gK : nullptr;
  if (sR != nullptr || st != nullptr || sK != nullptr) {
    dim3 g2(grid, C);
    loss_cam_shared_kernel<S><<<g2, kPB, 0, s>>>(a);
    if ((ce = cudaGetLastError()) != cudaSuccess) return set_error((int)ce, cudaGetErrorString(ce));
    loss_cam_finish_kernel<S><<<(C * 18 + 127) / 128, 128, 0, s>>>(a.partials, grid, C, sR, st, sK);
    if ((ce = cudaGetLastError()) != cudaSuccess) return set_error((int)ce, cudaGetErrorString(ce));
  }
  if (fR != nullptr || ft != nullptr || fK != nullptr) {
    const int64_t rows = T * C;
    if (rows > 0) {  // an empty clip has no per-frame camera gradients: a 0-block grid is a launch error
      loss_cam_frames_kernel<S><<<(unsigned)((rows + 127) / 128), 128, 0, s>>>(a, fR, ft, fK);
      if ((ce = cudaGetLastError()) != cudaSuccess) return set_error((int)ce, cudaGetErrorString(ce));
    }
  }
  return SKA_OK;
}
template int reprojection_loss<float>(const float*, int64_t, int, int, const float*, int64_t, const float*, int64_t, const float*, int64_t,
                                      const float*, const float*, double*, float*, float*, float*, float*, void*, size_t, cudaStream_t);
template int reprojection_loss<double>(const double*, int64_t, int, int, const double*, int64_t, const double*, int64_t, const double*,
                                       int64_t, const double*, const double*, double*, double*, double*, double*, double*, void*, size_t,
                                       cudaStream_t);

}  // namespace ska
