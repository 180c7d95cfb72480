// ska_optim.cu - the update kernels of the first-order (Adam) form of the regularised bundle adjustment
// (specification: oracle/first_order.py).  The objective and its analytic gradient come from the
// loss kernels (ska_project.cu / ska_losses.cu = bundle_adjustment/loss.py:17-155); what the reference configures for
// the optimiser it never defines is `lr`, `num_iters` and one weight per loss term (configs/vggt.yaml:43-52;
// call site vggt/multi_view_process.py:553-564).  Three element-wise kernels, templated on the caller's dtype:
//   adam_kernel         torch.optim.Adam's update: m <- m + (1-b1)(g-m); v <- b2 v + (1-b2) g^2;
//                       step = step_size * m / (sqrt(v) * inv_sqrt_bc2 + eps); p <- p - step (and / or step_out <- step);
//                       step_size / inv_sqrt_bc2 optionally come from device memory so that one captured CUDA graph of an
//                       iteration can be replayed for every k
//   so3_tangent_grad    g_omega = (B21 - B12, B02 - B20, B10 - B01), B = (dL/dR) R^T   (left perturbation R = exp([w]x) R)
//   so3_retract         R <- exp([-step]x) R   (Rodrigues in fp64, series below |w|^2 < 1e-16)
#include <cuda_runtime.h>
#include <math.h>

#include "ska_internal.h"

namespace ska {

// up to three gradient terms g = s0 g0 + s1 g1 + s2 g2 (the loss kernels return UNSCALED gradients of their raw sums: the
// weights / counts are applied here instead of in separate element-wise passes)
template <typename S>
struct GradTerms {
  const S* g0;
  const S* g1;
  const S* g2;
  double s0, s1, s2;
};

template <typename S>
__global__ void __launch_bounds__(256) adam_kernel(S* __restrict__ p, const GradTerms<S> gt, S* __restrict__ m, S* __restrict__ v,
                                                   int64_t n, double step_size, double b1, double b2, double eps, double inv_sqrt_bc2,
                                                   S* __restrict__ step_out, const double* __restrict__ scalars) {
  if (scalars != nullptr) {  // per-iteration scalars from device memory: lets a captured CUDA graph be replayed for every k
    step_size = scalars[0];
    inv_sqrt_bc2 = scalars[1];
  }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double gi = gt.s0 * (double)gt.g0[i];
    if (gt.g1 != nullptr) gi += gt.s1 * (double)gt.g1[i];
    if (gt.g2 != nullptr) gi += gt.s2 * (double)gt.g2[i];
    const double mi = (double)m[i] + (1.0 - b1) * (gi - (double)m[i]);
    const double vi = b2 * (double)v[i] + (1.0 - b2) * gi * gi;
    m[i] = (S)mi;
    v[i] = (S)vi;
    const double st = step_size * mi / (sqrt(vi) * inv_sqrt_bc2 + eps);
    if (p != nullptr) p[i] = (S)((double)p[i] - st);
    if (step_out != nullptr) step_out[i] = (S)st;
  }
}

template <typename S>
__global__ void __launch_bounds__(256) so3_tangent_grad_kernel(const S* __restrict__ R, const S* __restrict__ gR, int64_t n, S* __restrict__ gw) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double r[9], q[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    r[k] = (double)R[9 * i + k];
    q[k] = (double)gR[9 * i + k];
  }
  // B = gR R^T: B_ab = sum_k gR_ak R_bk
  auto B = [&](int a, int b) { return q[3 * a] * r[3 * b] + q[3 * a + 1] * r[3 * b + 1] + q[3 * a + 2] * r[3 * b + 2]; };
  gw[3 * i] = (S)(B(2, 1) - B(1, 2));
  gw[3 * i + 1] = (S)(B(0, 2) - B(2, 0));
  gw[3 * i + 2] = (S)(B(1, 0) - B(0, 1));
}

template <typename S>
__global__ void __launch_bounds__(256) so3_retract_kernel(S* __restrict__ R, const S* __restrict__ step, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double wx = -(double)step[3 * i], wy = -(double)step[3 * i + 1], wz = -(double)step[3 * i + 2];
  const double th2 = wx * wx + wy * wy + wz * wz;
  double A, Bc;
  if (th2 < 1e-16) {
    A = 1.0 - th2 / 6.0;
    Bc = 0.5 - th2 / 24.0;
  } else {
    const double th = sqrt(th2);
    A = sin(th) / th;
    Bc = (1.0 - cos(th)) / th2;
  }
  double E[9];
  E[0] = 1.0 - Bc * (wy * wy + wz * wz);
  E[1] = -A * wz + Bc * wx * wy;
  E[2] = A * wy + Bc * wx * wz;
  E[3] = A * wz + Bc * wx * wy;
  E[4] = 1.0 - Bc * (wx * wx + wz * wz);
  E[5] = -A * wx + Bc * wy * wz;
  E[6] = -A * wy + Bc * wx * wz;
  E[7] = A * wx + Bc * wy * wz;
  E[8] = 1.0 - Bc * (wx * wx + wy * wy);
  double r[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) r[k] = (double)R[9 * i + k];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int c = 0; c < 3; ++c) R[9 * i + 3 * a + c] = (S)(E[3 * a] * r[c] + E[3 * a + 1] * r[3 + c] + E[3 * a + 2] * r[6 + c]);
}

// Scalar bookkeeping of one first-order iteration in ONE thread (instead of ~20 one-element torch kernels):
// record: hist[k] = [total, c0 s0 / (den + 1e-6) (den == nullptr: c0 s0), c1 s1, c2 s2, c3 s3, c4 s4] for the five loss terms
// (a NULL sum pointer = term switched off), then k <- k + 1 and the Adam scalars of step k:
// scal = {lr / (1 - b1^k), 1 / sqrt(1 - b2^k)}.
struct RecordArgs {
  const double* s[5];
  const double* den;
  double c[5];
};
__global__ void first_order_record_kernel(double* __restrict__ k, double* __restrict__ scal, double* __restrict__ hist, int64_t max_rows,
                                          const RecordArgs a, double lr, double b1, double b2) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  const int64_t it = (int64_t)k[0];
  double tot = 0.0, term[5];
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    term[q] = 0.0;
    if (a.s[q] != nullptr) term[q] = (q == 0 && a.den != nullptr) ? a.c[q] * a.s[q][0] / (a.den[0] + 1e-6) : a.c[q] * a.s[q][0];
    tot += term[q];
  }
  if (hist != nullptr && it < max_rows) {
    double* h = hist + it * 6;
    h[0] = tot;
#pragma unroll
    for (int q = 0; q < 5; ++q) h[1 + q] = term[q];
  }
  const double kk = (double)(it + 1);
  k[0] = kk;
  scal[0] = lr / (1.0 - pow(b1, kk));
  scal[1] = rsqrt(1.0 - pow(b2, kk));
}

int first_order_record(double* k, double* scal, double* hist, int64_t max_rows, const double* const* sums, const double* den,
                       const double* coef, double lr, double b1, double b2, cudaStream_t s) {
  RecordArgs a;
  for (int q = 0; q < 5; ++q) {
    a.s[q] = sums[q];
    a.c[q] = coef[q];
  }
  a.den = den;
  first_order_record_kernel<<<1, 32, 0, s>>>(k, scal, hist, max_rows, a, lr, b1, b2);
  const cudaError_t ce = cudaGetLastError();
  return ce == cudaSuccess ? SKA_OK : set_error((int)ce, cudaGetErrorString(ce));
}

static int grid_of(int64_t n) {
  int64_t g = (n + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  return (int)(g < 1 ? 1 : g);
}

template <typename S>
int adam_step(S* p, const S* g, S* m, S* v, int64_t n, double step_size, double b1, double b2, double eps, double inv_sqrt_bc2, S* step_out,
              const double* scalars, const S* g1, const S* g2, double s0, double s1, double s2, cudaStream_t s) {
  if (n == 0) return SKA_OK;
  const GradTerms<S> gt{g, g1, g2, s0, s1, s2};
  adam_kernel<S><<<grid_of(n), 256, 0, s>>>(p, gt, m, v, n, step_size, b1, b2, eps, inv_sqrt_bc2, step_out, scalars);
  const cudaError_t ce = cudaGetLastError();
  return ce == cudaSuccess ? SKA_OK : set_error((int)ce, cudaGetErrorString(ce));
}
template <typename S>
int so3_tangent_grad(const S* R, const S* gR, int64_t n, S* gw, cudaStream_t s) {
  if (n == 0) return SKA_OK;
  so3_tangent_grad_kernel<S><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(R, gR, n, gw);
  const cudaError_t ce = cudaGetLastError();
  return ce == cudaSuccess ? SKA_OK : set_error((int)ce, cudaGetErrorString(ce));
}
template <typename S>
int so3_retract(S* R, const S* step, int64_t n, cudaStream_t s) {
  if (n == 0) return SKA_OK;
  so3_retract_kernel<S><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(R, step, n);
  const cudaError_t ce = cudaGetLastError();
  return ce == cudaSuccess ? SKA_OK : set_error((int)ce, cudaGetErrorString(ce));
}

template int adam_step<float>(float*, const float*, float*, float*, int64_t, double, double, double, double, double, float*, const double*, const float*,
                              const float*, double, double, double, cudaStream_t);
template int adam_step<double>(double*, const double*, double*, double*, int64_t, double, double, double, double, double, double*, const double*,
                               const double*, const double*, double, double, double, cudaStream_t);
template int so3_tangent_grad<float>(const float*, const float*, int64_t, float*, cudaStream_t);
template int so3_tangent_grad<double>(const double*, const double*, int64_t, double*, cudaStream_t);
template int so3_retract<float>(float*, const float*, int64_t, cudaStream_t);
template int so3_retract<double>(double*, const double*, int64_t, cudaStream_t);

}  // namespace ska
