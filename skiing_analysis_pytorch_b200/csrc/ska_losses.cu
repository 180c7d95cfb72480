// ska_losses.cu - the regularisers of bundle_adjustment/loss.py as fused value + analytic-gradient
// kernels (forward-only callers pass NULL gradient pointers).  All four are O(T*J) or O(T*C)
// streaming sums: thread-per-element kernels, fp64 fixed-order reductions, no atomics.
// Gradients are returned UNSCALED (derivative of the returned raw sum); the Python side applies
// w / count and autograd's grad_output.
//
// Reference anchors (file:line relative to the reference checkout):
//   camera_center_from_Rt  bundle_adjustment/loss.py:97-100     C = -R^T t
//   camera_smooth_loss     :103-106   w * mean((C[1:] - C[:-1])^2)
//   baseline_reg_loss      :109-114   w * mean((|C0 - C1| - mean.detach())^2)
//   bone_length_loss       :134-150   w * mean((|X_i - X_j| - ref)^2), ref = mean over T (detached) or given
//   pose_temporal_loss     :153-155   w * mean((X[1:] - X[:-1])^2)
#include <cuda_runtime.h>
#include <math.h>

#include "ska_internal.h"

namespace ska {

constexpr int kLB = 256;

template <int NV>
__device__ __forceinline__ void block_sum_store(const double (&v)[NV], double* scratch, double* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double x = v[i];
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

static int small_grid(int64_t n) {
  int g = ba_max_grid() / 4;
  const int64_t need = (n + kLB - 1) / kLB;
  if (g > need) g = (int)need;
  return g < 1 ? 1 : g;
}

size_t reg_workspace_bytes() { return (size_t)ba_max_grid() * SKA_MAX_BONES * sizeof(double); }  // one partial row per block

#define SKA_LAUNCH_CHECK()                                                     \
  do {                                                                         \
    const cudaError_t ce_ = cudaGetLastError();                                \
    if (ce_ != cudaSuccess) return set_error((int)ce_, cudaGetErrorString(ce_)); \
  } while (0)

// ------------------------------------------------------------------------------------------------
// pose_temporal: sum over t < T-1 of (X[t+1] - X[t])^2, M = J*3 values per frame.
// A thread owns one of the M columns over a chunk of kPtChunk frames and walks down the chunk with the previous / current
// / next value in registers: every element is loaded once, no per-element index division (the grid-stride form with
// i / M and three loads per element ran at 0.5 TB/s).  Adjacent threads own adjacent columns: rows are read coalesced.
constexpr int kPtChunk = 128;

template <typename S>
__global__ void __launch_bounds__(kLB) pose_temporal_kernel(const S* __restrict__ X, int64_t T, int64_t M, int64_t chunk, S* __restrict__ gX,
                                                           double* __restrict__ partials) {
  __shared__ double scratch[kLB / 32];
  double acc[1] = {0.0};
  const int64_t n_chunks = (T + chunk - 1) / chunk;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid < n_chunks * M) {
    const int64_t c = tid / M, m = tid - c * M;
    const int64_t t0 = c * chunk, t1 = (t0 + chunk < T) ? t0 + chunk : T;
    const S* p = X + t0 * M + m;
    S* g = (gX != nullptr) ? gX + t0 * M + m : nullptr;
    S prev = (t0 > 0) ? p[-M] : S(0);
    S cur = p[0];
    for (int64_t t = t0; t < t1; ++t) {
      const bool has_next = t + 1 < T;
      const S nxt = has_next ? p[M] : S(0);
      const S dn = has_next ? nxt - cur : S(0);
      const S dp = (t > 0) ? cur - prev : S(0);
      acc[0] += (double)dn * (double)dn;
      if (g != nullptr) {
        *g = S(2) * (dp - dn);
        g += M;
      }
      prev = cur;
      cur = nxt;
      p += M;
    }
  }
  block_sum_store<1>(acc, scratch, partials + blockIdx.x);
}

template <typename S>
int pose_temporal(const S* X, int64_t T, int J, double* sum, S* gX, void* ws, size_t ws_bytes, cudaStream_t s) {
  if (ws_bytes < reg_workspace_bytes()) return set_error(SKA_EWORKSPACE, "workspace too small (see ska_reg_workspace_bytes)");
  const int64_t M = (int64_t)J * 3;
  // one partial per block in the workspace: keep the block count within it by lengthening the chunks of long clips
  const int64_t max_blocks = (int64_t)(reg_workspace_bytes() / sizeof(double));
  int64_t chunk = kPtChunk;
  while (chunk > 8 && (T + chunk - 1) / chunk * M < 300000) chunk /= 2;  // short clips: enough threads to fill the GPU
  while (((T + chunk - 1) / chunk * M + kLB - 1) / kLB > max_blocks) chunk *= 2;
  const int64_t threads = (T + chunk - 1) / chunk * M;
  const int grid = (int)((threads + kLB - 1) / kLB < 1 ? 1 : (threads + kLB - 1) / kLB);
  pose_temporal_kernel<S><<<grid, kLB, 0, s>>>(X, T, M, chunk, gX, (double*)ws);
  SKA_LAUNCH_CHECK();
  return launch_reduce((const double*)ws, grid, 1, sum, s);
}
template int pose_temporal<float>(const float*, int64_t, int, double*, float*, void*, size_t, cudaStream_t);
template int pose_temporal<double>(const double*, int64_t, int, double*, double*, void*, size_t, cudaStream_t);

// ------------------------------------------------------------------------------------------------
// bone_length.  pass 0 (ref == nullptr): sums[b] = sum_t len[t][b].
// pass 1: sums[0] = sum_{t,b} (len - ref[b])^2 and gX (thread-private frame rows, zero-initialised here).
struct Bones {
  int32_t n;
  int32_t i[SKA_MAX_BONES], j[SKA_MAX_BONES];
};

template <typename S>
__global__ void __launch_bounds__(kLB) bone_length_kernel(const S* __restrict__ X, int64_t T, int J, const Bones bones,
                                                         const double* __restrict__ ref, S* __restrict__ gX,
                                                         double* __restrict__ partials) {
  __shared__ double scratch[(kLB / 32) * SKA_MAX_BONES];
  double acc[SKA_MAX_BONES];
#pragma unroll
  for (int b = 0; b < SKA_MAX_BONES; ++b) acc[b] = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < T; t += stride) {
    const S* x = X + t * (int64_t)J * 3;
    S* g = (gX != nullptr) ? gX + t * (int64_t)J * 3 : nullptr;
    if (g != nullptr)
      for (int k = 0; k < 3 * J; ++k) g[k] = S(0);
#pragma unroll
    for (int b = 0; b < SKA_MAX_BONES; ++b) {
      if (b < bones.n) {
        const int i = bones.i[b], j = bones.j[b];
        const S d0 = x[3 * i] - x[3 * j], d1 = x[3 * i + 1] - x[3 * j + 1], d2 = x[3 * i + 2] - x[3 * j + 2];
        const double len = sqrt((double)d0 * d0 + (double)d1 * d1 + (double)d2 * d2);
        if (ref == nullptr) {
          acc[b] += len;
        } else {
          const double r = len - ref[b];
          acc[0] += r * r;
          if (g != nullptr && len > 0.0) {  // torch.norm's subgradient at 0 is 0
            const double f = 2.0 * r / len;
            g[3 * i] += (S)(f * d0); g[3 * i + 1] += (S)(f * d1); g[3 * i + 2] += (S)(f * d2);
            g[3 * j] -= (S)(f * d0); g[3 * j + 1] -= (S)(f * d1); g[3 * j + 2] -= (S)(f * d2);
          }
        }
      }
    }
  }
  block_sum_store<SKA_MAX_BONES>(acc, scratch, partials + (int64_t)blockIdx.x * SKA_MAX_BONES);
}

// Staged form: a block of FB threads owns FB consecutive frames; their rows (one contiguous run in global memory) are
// copied into shared memory coalesced, every thread works on its frame's row there (gradient accumulation included - the
// direct kernel above read-modify-writes its gradient row in GLOBAL memory, 12 bones x 6 entries per frame), and the
// gradient rows leave coalesced.  Same arithmetic in the same order per frame: bit-identical results.  Row stride is odd
// (3J | 1): the FB threads of a warp hit distinct banks.
template <typename S>
__global__ void __launch_bounds__(128) bone_length_staged_kernel(const S* __restrict__ X, int64_t T, int J, const Bones bones,
                                                                const double* __restrict__ ref, S* __restrict__ gX,
                                                                double* __restrict__ partials) {
  extern __shared__ __align__(16) unsigned char bl_smem[];
  __shared__ double scratch[4 * SKA_MAX_BONES];
  const int FB = blockDim.x, M = 3 * J, RS = M | 1;
  S* xs = reinterpret_cast<S*>(bl_smem);
  S* gs = xs + (size_t)FB * RS;
  double acc[SKA_MAX_BONES];
#pragma unroll
  for (int b = 0; b < SKA_MAX_BONES; ++b) acc[b] = 0.0;
  const bool want_g = gX != nullptr && ref != nullptr;
  for (int64_t f0 = (int64_t)blockIdx.x * FB; f0 < T; f0 += (int64_t)gridDim.x * FB) {
    const int nv = (int)((T - f0) < FB ? (T - f0) : FB);
    const int64_t base = f0 * M;
    {
      int f = 0, k = threadIdx.x;
      while (k >= M) {
        k -= M;
        ++f;
      }
      for (int e = threadIdx.x; e < nv * M; e += FB) {
        xs[f * RS + k] = X[base + e];
        if (want_g) gs[f * RS + k] = S(0);
        k += FB;
        while (k >= M) {
          k -= M;
          ++f;
        }
      }
    }
    __syncthreads();
    if ((int)threadIdx.x < nv) {
      const S* x = xs + threadIdx.x * RS;
      S* g = gs + threadIdx.x * RS;
#pragma unroll
      for (int b = 0; b < SKA_MAX_BONES; ++b) {
        if (b < bones.n) {
          const int i = bones.i[b], j = bones.j[b];
          const S d0 = x[3 * i] - x[3 * j], d1 = x[3 * i + 1] - x[3 * j + 1], d2 = x[3 * i + 2] - x[3 * j + 2];
          const double len = sqrt((double)d0 * d0 + (double)d1 * d1 + (double)d2 * d2);
          if (ref == nullptr) {
            acc[b] += len;
          } else {
            const double r = len - ref[b];
            acc[0] += r * r;
            if (want_g && len > 0.0) {  // torch.norm's subgradient at 0 is 0
              const double fq = 2.0 * r / len;
              g[3 * i] += (S)(fq * d0); g[3 * i + 1] += (S)(fq * d1); g[3 * i + 2] += (S)(fq * d2);
              g[3 * j] -= (S)(fq * d0); g[3 * j + 1] -= (S)(fq * d1); g[3 * j + 2] -= (S)(fq * d2);
            }
          }
        }
      }
    }
    __syncthreads();
    if (gX != nullptr) {
      int f = 0, k = threadIdx.x;
      while (k >= M) {
        k -= M;
        ++f;
      }
      for (int e = threadIdx.x; e < nv * M; e += FB) {
        gX[base + e] = want_g ? gs[f * RS + k] : S(0);
        k += FB;
        while (k >= M) {
          k -= M;
          ++f;
        }
      }
    }
    __syncthreads();
  }
  block_sum_store<SKA_MAX_BONES>(acc, scratch, partials + (int64_t)blockIdx.x * SKA_MAX_BONES);
}

template <typename S>
int bone_length(const S* X, int64_t T, int J, const int32_t* bi, const int32_t* bj, int nb, const double* ref, double* sums,
                S* gX, void* ws, size_t ws_bytes, cudaStream_t s) {
  if (ws_bytes < reg_workspace_bytes()) return set_error(SKA_EWORKSPACE, "workspace too small (see ska_reg_workspace_bytes)");
  if (nb < 0 || nb > SKA_MAX_BONES) return set_error(SKA_EINVAL, "at most SKA_MAX_BONES bones");
  Bones b;
  b.n = nb;
  for (int k = 0; k < SKA_MAX_BONES; ++k) {
    b.i[k] = k < nb ? bi[k] : 0;
    b.j[k] = k < nb ? bj[k] : 0;
    if (k < nb && (bi[k] < 0 || bi[k] >= J || bj[k] < 0 || bj[k] >= J)) return set_error(SKA_EINVAL, "bone index out of range");
  }
  // gradient pass: staged kernel when FB rows of x and g fit in shared memory (J <= ~140 in fp32); the direct kernel otherwise
  const int RS = (3 * J) | 1;
  int FB = 128;
  while (FB > 32 && (size_t)2 * FB * RS * sizeof(S) > 96 * 1024) FB /= 2;
  const size_t smem = (size_t)2 * FB * RS * sizeof(S);
  if (gX != nullptr && ref != nullptr && smem <= 200 * 1024) {  // forward-only passes: the direct kernel's row reads are already L1-coalesced (measured faster)
    static size_t attr_bytes[64] = {};
    int dev = 0;
    if (smem > 48 * 1024 && cudaGetDevice(&dev) == cudaSuccess && dev < 64 && attr_bytes[dev] < smem) {  // idempotent; benign race
      const cudaError_t ca = cudaFuncSetAttribute(bone_length_staged_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (ca != cudaSuccess) return set_error((int)ca, cudaGetErrorString(ca));
      attr_bytes[dev] = smem;
    }
    const int64_t need = (T + FB - 1) / FB;
    int grid = ba_max_grid();  // one partial row per block in the workspace (reg_workspace_bytes)
    if (grid > need) grid = (int)need;
    if (grid < 1) grid = 1;
    bone_length_staged_kernel<S><<<grid, FB, smem, s>>>(X, T, J, b, ref, gX, (double*)ws);
    SKA_LAUNCH_CHECK();
    return launch_reduce((const double*)ws, grid, SKA_MAX_BONES, sums, s);
  }
  const int grid = small_grid(T);
  bone_length_kernel<S><<<grid, kLB, 0, s>>>(X, T, J, b, ref, gX, (double*)ws);
  SKA_LAUNCH_CHECK();
  return launch_reduce((const double*)ws, grid, SKA_MAX_BONES, sums, s);
}
template int bone_length<float>(const float*, int64_t, int, const int32_t*, const int32_t*, int, const double*, double*, float*, void*, size_t, cudaStream_t);
template int bone_length<double>(const double*, int64_t, int, const int32_t*, const int32_t*, int, const double*, double*, double*, void*, size_t, cudaStream_t);

// ------------------------------------------------------------------------------------------------
// camera centres C = -R^T t and the adjoint of that map
template <typename S>
__device__ __forceinline__ void cam_centre(const S* __restrict__ R, const S* __restrict__ t, double C[3]) {
#pragma unroll
  for (int k = 0; k < 3; ++k) C[k] = -((double)R[k] * t[0] + (double)R[3 + k] * t[1] + (double)R[6 + k] * t[2]);
}
// gR[i][k] = -t_i gC_k,  gt_i = -(R gC)_i
template <typename S>
__device__ __forceinline__ void cam_centre_adjoint(const S* __restrict__ R, const S* __restrict__ t, const double gC[3], S* gR, S* gt) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (gR != nullptr) {
#pragma unroll
      for (int k = 0; k < 3; ++k) gR[3 * i + k] = (S)(-(double)t[i] * gC[k]);
    }
    if (gt != nullptr) gt[i] = (S)(-((double)R[3 * i] * gC[0] + (double)R[3 * i + 1] * gC[1] + (double)R[3 * i + 2] * gC[2]));
  }
}

template <typename S>
__global__ void __launch_bounds__(kLB) camera_centre_kernel(const S* __restrict__ R, const S* __restrict__ t, int64_t n, S* __restrict__ C) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double c[3];
  cam_centre<S>(R + 9 * i, t + 3 * i, c);
  C[3 * i] = (S)c[0];
  C[3 * i + 1] = (S)c[1];
  C[3 * i + 2] = (S)c[2];
}

template <typename S>
int camera_centre(const S* R, const S* t, int64_t n, S* C, cudaStream_t s) {
  if (n == 0) return SKA_OK;
  camera_centre_kernel<S><<<(unsigned)((n + kLB - 1) / kLB), kLB, 0, s>>>(R, t, n, C);
  SKA_LAUNCH_CHECK();
  return SKA_OK;
}
template int camera_centre<float>(const float*, const float*, int64_t, float*, cudaStream_t);
template int camera_centre<double>(const double*, const double*, int64_t, double*, cudaStream_t);

// camera_smooth: cameras (D0, M): sum over d < D0-1 of |C[d+1][m] - C[d][m]|^2
template <typename S>
__global__ void __launch_bounds__(kLB) camera_smooth_kernel(const S* __restrict__ R, const S* __restrict__ t, int64_t D0, int64_t M,
                                                           S* __restrict__ gR, S* __restrict__ gt, double* __restrict__ partials) {
  __shared__ double scratch[kLB / 32];
  double acc[1] = {0.0};
  const int64_t n = D0 * M, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t d = i / M;
    double c[3], cn[3], cp[3], g[3];
    cam_centre<S>(R + 9 * i, t + 3 * i, c);
    if (d + 1 < D0) cam_centre<S>(R + 9 * (i + M), t + 3 * (i + M), cn);
    if (d > 0) cam_centre<S>(R + 9 * (i - M), t + 3 * (i - M), cp);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double dn = (d + 1 < D0) ? cn[k] - c[k] : 0.0;
      const double dp = (d > 0) ? c[k] - cp[k] : 0.0;
      acc[0] += dn * dn;
      g[k] = 2.0 * (dp - dn);
    }
    cam_centre_adjoint<S>(R + 9 * i, t + 3 * i, g, gR ? gR + 9 * i : nullptr, gt ? gt + 3 * i : nullptr);
  }
  block_sum_store<1>(acc, scratch, partials + blockIdx.x);
}

template <typename S>
int camera_smooth(const S* R, const S* t, int64_t D0, int64_t M, double* sum, S* gR, S* gt, void* ws, size_t ws_bytes, cudaStream_t s) {
  if (ws_bytes < reg_workspace_bytes()) return set_error(SKA_EWORKSPACE, "workspace too small (see ska_reg_workspace_bytes)");
  const int grid = small_grid(D0 * M);
  camera_smooth_kernel<S><<<grid, kLB, 0, s>>>(R, t, D0, M, gR, gt, (double*)ws);
  SKA_LAUNCH_CHECK();
  return launch_reduce((const double*)ws, grid, 1, sum, s);
}
template int camera_smooth<float>(const float*, const float*, int64_t, int64_t, double*, float*, float*, void*, size_t, cudaStream_t);
template int camera_smooth<double>(const double*, const double*, int64_t, int64_t, double*, double*, double*, void*, size_t, cudaStream_t);

// baseline_reg: cameras (T, C>=2).  pass 0 (mean == nullptr): sums[0] = sum_t |C0 - C1|.
// pass 1: sums[0] = sum_t (|C0 - C1| - *mean)^2 and gradients for cameras 0 and 1 (others zero).
template <typename S>
__global__ void __launch_bounds__(kLB) baseline_kernel(const S* __restrict__ R, const S* __restrict__ t, int64_t T, int Cn,
                                                      const double* __restrict__ mean, S* __restrict__ gR, S* __restrict__ gt,
                                                      double* __restrict__ partials) {
  __shared__ double scratch[kLB / 32];
  double acc[1] = {0.0};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < T; f += stride) {
    const int64_t i0 = f * Cn, i1 = f * Cn + 1;
    double c0[3], c1[3];
    cam_centre<S>(R + 9 * i0, t + 3 * i0, c0);
    cam_centre<S>(R + 9 * i1, t + 3 * i1, c1);
    const double d[3] = {c0[0] - c1[0], c0[1] - c1[1], c0[2] - c1[2]};
    const double b = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    if (mean == nullptr) {
      acc[0] += b;
    } else {
      const double r = b - mean[0];
      acc[0] += r * r;
      if (gR != nullptr || gt != nullptr) {
        const double fct = b > 0.0 ? 2.0 * r / b : 0.0;
        const double g0[3] = {fct * d[0], fct * d[1], fct * d[2]}, g1[3] = {-g0[0], -g0[1], -g0[2]};
        cam_centre_adjoint<S>(R + 9 * i0, t + 3 * i0, g0, gR ? gR + 9 * i0 : nullptr, gt ? gt + 3 * i0 : nullptr);
        cam_centre_adjoint<S>(R + 9 * i1, t + 3 * i1, g1, gR ? gR + 9 * i1 : nullptr, gt ? gt + 3 * i1 : nullptr);
        for (int c = 2; c < Cn; ++c) {
          if (gR != nullptr)
            for (int k = 0; k < 9; ++k) gR[9 * (f * Cn + c) + k] = S(0);
          if (gt != nullptr)
            for (int k = 0; k < 3; ++k) gt[3 * (f * Cn + c) + k] = S(0);
        }
      }
    }
  }
  block_sum_store<1>(acc, scratch, partials + blockIdx.x);
}

template <typename S>
int baseline_reg(const S* R, const S* t, int64_t T, int Cn, const double* mean, double* sum, S* gR, S* gt, void* ws, size_t ws_bytes,
                 cudaStream_t s) {
  if (ws_bytes < reg_workspace_bytes()) return set_error(SKA_EWORKSPACE, "workspace too small (see ska_reg_workspace_bytes)");
  if (Cn < 2) return set_error(SKA_EINVAL, "baseline_reg needs at least two cameras");
  const int grid = small_grid(T);
  baseline_kernel<S><<<grid, kLB, 0, s>>>(R, t, T, Cn, mean, gR, gt, (double*)ws);
  SKA_LAUNCH_CHECK();
  return launch_reduce((const double*)ws, grid, 1, sum, s);
}
template int baseline_reg<float>(const float*, const float*, int64_t, int, const double*, double*, float*, float*, void*, size_t, cudaStream_t);
template int baseline_reg<double>(const double*, const double*, int64_t, int, const double*, double*, double*, double*, void*, size_t, cudaStream_t);

}  // namespace ska
