// ska_fuse.cu - two-view 3D-3D fusion of monocular pose estimates + adaptive EMA smoothing, sm_100a.
//
// Replaces, for a whole clip at once, the per-frame Python / numpy of the reference's `fuse` pipeline:
//   _kabsch_rigid_align, _align_right_to_left        fuse/main_raw.py:48-95
//   fit_weakpersp_3d_to_2d, weakpersp_reproj_confidence   fuse/confidence.py:9-108
//   canonicalize_pose_3d, crossview_consistency_confidence fuse/confidence.py:118-224
//   softmax2, fuse_frame_3d                          fuse/fuse.py:87-94, 289-326
//   temporal_smooth_ema                              fuse/fuse.py:329-412
// Arithmetic is fp64 like the reference's numpy.  The two SVDs (3x3 cross-covariance of the rigid alignment, 3x2 of the
// weak-perspective fit) are one-sided Jacobi (Hestenes) iterations on the columns - no squaring of the matrix, so the
// rotation / the orthonormal map agree with LAPACK's to rounding for every non-degenerate frame.
//
// fuse_frames_kernel: ONE WARP PER FRAME, lane l owns joints l, l + 32, l + 64 (J <= 96: the SAM-3D-Body skeleton has 70);
//   all frame-level sums are fixed-order xor-butterfly reductions (deterministic), the small factorizations run
//   redundantly on every lane.  Reads 80 B, writes 24 B per joint (+ optional q / aligned outputs).
// ema_kernel: thread = (chunk of frames, joint).  The recurrence is a contraction (every step multiplies the state
//   error by at most rho = max(1 - alpha_min, |1 + alpha_min - 2 alpha_max|) < 1), so a chunk started `halo` valid samples
//   early from the "first frame" rule reproduces the sequential scan to below fp64 rounding; the host picks halo so that
//   rho^halo < 1e-18, or runs one chunk (halo < 0: exact sequential) when rho is too close to 1.  The update uses
//   explicit round-to-nearest multiplies and adds (no FMA contraction) so the sequential mode is bit-identical to numpy.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "ska_internal.h"

namespace ska {

constexpr int kFuseMaxJ = 96;
constexpr int kFuseJPL = 3;  // joints per lane
constexpr double kFuseEps = 1e-8;  // fuse/fuse.py:19

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ bool fin3(const double* p) { return isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]); }
__device__ __forceinline__ bool fin2(const double* p) { return isfinite(p[0]) && isfinite(p[1]); }

// one Hestenes rotation of columns p, q of the 3-row matrices A (and V with NV rows)
template <int NV>
__device__ __forceinline__ bool hestenes_rotate(double* ap, double* aq, double* vp, double* vq) {
  const double alpha = ap[0] * ap[0] + ap[1] * ap[1] + ap[2] * ap[2];
  const double beta = aq[0] * aq[0] + aq[1] * aq[1] + aq[2] * aq[2];
  const double gamma = ap[0] * aq[0] + ap[1] * aq[1] + ap[2] * aq[2];
  if (!(fabs(gamma) > 1e-300) || !(gamma * gamma > 1e-31 * alpha * beta)) return false;
  const double zeta = (beta - alpha) / (2.0 * gamma);
  const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
  const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double x = ap[k], y = aq[k];
    ap[k] = c * x - s * y;
    aq[k] = s * x + c * y;
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const double x = vp[k], y = vq[k];
    vp[k] = c * x - s * y;
    vq[k] = s * x + c * y;
  }
  return true;
}

__device__ __forceinline__ void cross3(const double* a, const double* b, double* o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

// Rotation of the Kabsch problem for H = src_c^T dst_c (row-major 3x3): R = V diag(1, 1, det) U^T  (main_raw.py:58-65)
__device__ void kabsch_rotation(const double H[9], double R[9], double* sum_sv = nullptr) {
  double a[3][3], v[3][3];  // columns
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      a[c][r] = H[3 * r + c];
      v[c][r] = (r == c) ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 12; ++sweep) {
    bool any = false;
    any |= hestenes_rotate<3>(a[0], a[1], v[0], v[1]);
    any |= hestenes_rotate<3>(a[0], a[2], v[0], v[2]);
    any |= hestenes_rotate<3>(a[1], a[2], v[1], v[2]);
    if (!any) break;
  }
  double s[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) s[c] = sqrt(a[c][0] * a[c][0] + a[c][1] * a[c][1] + a[c][2] * a[c][2]);
  if (sum_sv != nullptr) *sum_sv = s[0] + s[1] + s[2];
  // the two largest singular values (order between them is irrelevant for the sum of outer products)
  int i0 = 0, i1 = 1, i2 = 2;
  if (s[i2] > s[i0]) { const int q = i0; i0 = i2; i2 = q; }
  if (s[i2] > s[i1]) { const int q = i1; i1 = i2; i2 = q; }
  double u0[3], u1[3], u2[3], v2[3];
  const double r0 = s[i0] > 0.0 ? 1.0 / s[i0] : 0.0, r1 = s[i1] > 0.0 ? 1.0 / s[i1] : 0.0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    u0[k] = a[i0][k] * r0;
    u1[k] = a[i1][k] * r1;
  }
  // third pair by cross products: this IS the det-fixed solution (flip of the smallest singular direction)
  cross3(u0, u1, u2);
  cross3(v[i0], v[i1], v2);
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) R[3 * r + c] = v[i0][r] * u0[c] + v[i1][r] * u1[c] + v2[r] * u2[c];
}

// Fast path of the same rotation: R = V U^T is the transpose of the orthogonal polar factor of H whenever det(H) > 0
// (no reflection to repair), and Newton's iteration X <- (X + X^-T) / 2 reaches it quadratically - six iterations of ~80
// fp64 operations for a body-shaped point set instead of ~5 Jacobi sweeps with two roots and two divisions per rotation.
// Returns false (caller runs kabsch_rotation) for a reflected or nearly planar / collinear cross-covariance.
__device__ __forceinline__ bool polar_rotation(const double H[9], double R[9], double* sum_sv = nullptr) {
  double f = 0.0;
#pragma unroll
  for (int q = 0; q < 9; ++q) f += H[q] * H[q];
  if (!(f > 0.0) || !isfinite(f)) return false;
  const double sc = sqrt(3.0 / f);  // singular values of X have unit root mean square
  double X[9], Cf[9];
#pragma unroll
  for (int q = 0; q < 9; ++q) X[q] = H[q] * sc;
  for (int it = 0; it < 16; ++it) {
    Cf[0] = X[4] * X[8] - X[5] * X[7];
    Cf[1] = X[5] * X[6] - X[3] * X[8];
    Cf[2] = X[3] * X[7] - X[4] * X[6];
    Cf[3] = X[2] * X[7] - X[1] * X[8];
    Cf[4] = X[0] * X[8] - X[2] * X[6];
    Cf[5] = X[1] * X[6] - X[0] * X[7];
    Cf[6] = X[1] * X[5] - X[2] * X[4];
    Cf[7] = X[2] * X[3] - X[0] * X[5];
    Cf[8] = X[0] * X[4] - X[1] * X[3];
    const double det = X[0] * Cf[0] + X[1] * Cf[1] + X[2] * Cf[2];
    if (it == 0 && !(det > 0.02)) return false;
    const double hid = 0.5 / det;
    double diff = 0.0;
#pragma unroll
    for (int q = 0; q < 9; ++q) {
      const double xn = 0.5 * X[q] + Cf[q] * hid;  // X^-T = cofactor / det
      diff += (xn - X[q]) * (xn - X[q]);
      X[q] = xn;
    }
    if (diff < 1e-26) break;  // quadratic convergence: the step before this one was already below 1e-13
  }
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) R[3 * r + c] = X[3 * c + r];
  if (sum_sv != nullptr) {  // H = Q P with P = Q^T H symmetric positive definite: trace(P) = sum of singular values
    double tr = 0.0;
#pragma unroll
    for (int q = 0; q < 9; ++q) tr += X[q] * H[q];
    *sum_sv = tr;
  }
  return true;
}

struct FuseArgs {
  const double* Xl;
  const double* Xr;
  const double* Ul;
  const double* Ur;
  int64_t T;
  int32_t J;
  SkaFuseParams prm;
  double* fused;
  double* ql;
  double* qr;
  double* aligned;
  uint8_t* status;
};

// weak-perspective confidence of one view for the lane's joints (confidence.py:9-108); returns false if the fit fails.
// The confidence is returned as its exponent: conf = exp(-arg), arg = err^2 / (2 sigma^2) >= 0; arg < 0 marks conf = 0
// (undefined residual) - the caller folds the exponents of the two confidences into ONE exp per quality value.
__device__ __forceinline__ bool weakpersp_conf(const double (&X)[kFuseJPL][3], const double (&U)[kFuseJPL][2], const bool (&in)[kFuseJPL],
                                               double sigma_px, int min_points, double (&conf)[kFuseJPL]) {
  bool val[kFuseJPL];
  double cnt = 0.0, sx[3] = {0, 0, 0}, su[2] = {0, 0};
#pragma unroll
  for (int k = 0; k < kFuseJPL; ++k) {
    val[k] = in[k] && fin3(X[k]) && fin2(U[k]);
    if (val[k]) {
      cnt += 1.0;
#pragma unroll
      for (int d = 0; d < 3; ++d) sx[d] += X[k][d];
      su[0] += U[k][0];
      su[1] += U[k][1];
    }
  }
  cnt = wsum(cnt);
#pragma unroll
  for (int k = 0; k < kFuseJPL; ++k) conf[k] = -1.0;
  if (cnt < (double)min_points) return false;
  double muX[3], muU[2];
#pragma unroll
  for (int d = 0; d < 3; ++d) muX[d] = wsum(sx[d]) / cnt;
  muU[0] = wsum(su[0]) / cnt;
  muU[1] = wsum(su[1]) / cnt;
  double c1[3] = {0, 0, 0}, c2[3] = {0, 0, 0}, den = 0.0;  // columns of C = Xc^T Uc (3x2)
#pragma unroll
  for (int k = 0; k < kFuseJPL; ++k)
    if (val[k]) {
      const double u0 = U[k][0] - muU[0], u1 = U[k][1] - muU[1];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const double xc = X[k][d] - muX[d];
        c1[d] += xc * u0;
        c2[d] += xc * u1;
        den += xc * xc;
      }
    }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    c1[d] = wsum(c1[d]);
    c2[d] = wsum(c2[d]);
  }
  den = wsum(den);
  if (den < 1e-12) return false;
  // 3x2 SVD by one Hestenes rotation: C V = [a1 a2] with orthogonal columns; M = u1 v1^T + u2 v2^T
  double v1[2] = {1.0, 0.0}, v2[2] = {0.0, 1.0};
  hestenes_rotate<2>(c1, c2, v1, v2);
  const double s1 = sqrt(c1[0] * c1[0] + c1[1] * c1[1] + c1[2] * c1[2]);
  const double s2 = sqrt(c2[0] * c2[0] + c2[1] * c2[1] + c2[2] * c2[2]);
  const double r1 = s1 > 0.0 ? 1.0 / s1 : 0.0, r2 = s2 > 0.0 ? 1.0 / s2 : 0.0;
  double M[3][2];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    M[d][0] = c1[d] * r1 * v1[0] + c2[d] * r2 * v2[0];
    M[d][1] = c1[d] * r1 * v1[1] + c2[d] * r2 * v2[1];
  }
  const double s = (s1 + s2) / den;
  const double t0 = muU[0] - s * (muX[0] * M[0][0] + muX[1] * M[1][0] + muX[2] * M[2][0]);
  const double t1 = muU[1] - s * (muX[0] * M[0][1] + muX[1] * M[1][1] + muX[2] * M[2][1]);
  const double sig = sigma_px > 1e-12 ? sigma_px : 1e-12;
  const double inv2s = 1.0 / (2.0 * sig * sig);
#pragma unroll
  for (int k = 0; k < kFuseJPL; ++k) {
    if (!in[k]) continue;
    const double h0 = s * (X[k][0] * M[0][0] + X[k][1] * M[1][0] + X[k][2] * M[2][0]) + t0;
    const double h1 = s * (X[k][0] * M[0][1] + X[k][1] * M[1][1] + X[k][2] * M[2][1]) + t1;
    const double d0 = h0 - U[k][0], d1 = h1 - U[k][1];
    const double e2 = d0 * d0 + d1 * d1;  // err^2; NaN / inf exactly when the reference's err is undefined
    if (isfinite(e2)) conf[k] = e2 * inv2s;
  }
  return true;
}

__device__ __forceinline__ void normalize3(double* v, double eps) {
  const double n = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  if (n < eps) {
    v[0] = v[1] = v[2] = 0.0;  // v * 0.0 of the reference (a NaN norm takes the division branch there too)
  } else {
    const double r = 1.0 / n;
    v[0] *= r;
    v[1] *= r;
    v[2] *= r;
  }
}

// canonical frame of one view (confidence.py:118-181): rows of Rc, origin, scale; ok = false -> every joint undefined
struct Canon {
  double R[3][3], root[3], s;
  bool ok;
};
__device__ __forceinline__ Canon canonical_frame(const double* __restrict__ Xf, const SkaFuseParams& p) {
  Canon c;
  const double* r = Xf + 3 * p.root;
  const double* lh = Xf + 3 * p.lhip;
  const double* rh = Xf + 3 * p.rhip;
  const double* ls = Xf + 3 * p.lsho;
  const double* rs = Xf + 3 * p.rsho;
  c.ok = fin3(r) && fin3(lh) && fin3(rh) && fin3(ls) && fin3(rs);
  double x[3], y[3], z[3], hip[3], torso[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    c.root[d] = r[d];
    const double Lh = lh[d] - r[d], Rh = rh[d] - r[d], Ls = ls[d] - r[d], Rs = rs[d] - r[d];
    hip[d] = Rh - Lh;
    torso[d] = 0.5 * (Ls + Rs) - 0.5 * (Lh + Rh);
    x[d] = hip[d];
    y[d] = torso[d];
  }
  const double eps = 1e-9;
  normalize3(x, eps);
  normalize3(y, eps);
  cross3(x, y, z);
  normalize3(z, eps);
  cross3(z, x, y);
  normalize3(y, eps);
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    c.R[0][d] = x[d];
    c.R[1][d] = y[d];
    c.R[2][d] = z[d];
  }
  const double* sv = p.scale_mode == 0 ? hip : torso;
  c.s = sqrt(sv[0] * sv[0] + sv[1] * sv[1] + sv[2] * sv[2]);
  if (!isfinite(c.s) || c.s < eps) c.ok = false;
  return c;
}

#ifndef SKA_FUSE_MINB
#define SKA_FUSE_MINB 3  // 162 registers, no spills (4 CTAs per SM needs 128 registers and spills 280 B: measured 6 % slower)
#endif
__global__ void __launch_bounds__(128, SKA_FUSE_MINB) fuse_frames_kernel(const FuseArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t frame = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (frame >= a.T) return;
  const int J = a.J;
  const double* Xlf = a.Xl + frame * J * 3;
  const double* Xrf = a.Xr + frame * J * 3;
  const double* Ulf = a.Ul + frame * J * 2;
  const double* Urf = a.Ur + frame * J * 2;

  double Xl[kFuseJPL][3], Xr[kFuseJPL][3], Ul[kFuseJPL][2], Ur[kFuseJPL][2];
  bool in[kFuseJPL];
#pragma unroll
  for (int k = 0; k < kFuseJPL; ++k) {
    const int j = lane + 32 * k;
    in[k] = j < J;
    const int jj = in[k] ? j : 0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      Xl[k][d] = Xlf[3 * jj + d];
      Xr[k][d] = Xrf[3 * jj + d];
    }
    Ul[k][0] = Ulf[2 * jj];
    Ul[k][1] = Ulf[2 * jj + 1];
    Ur[k][0] = Urf[2 * jj];
    Ur[k][1] = Urf[2 * jj + 1];
  }
  uint8_t st = 0;

  // ---- rigid alignment right -> left over the joints finite in both views (main_raw.py:71-95)
  bool both[kFuseJPL];
  double cnt = 0.0, sr[3] = {0, 0, 0}, sl[3] = {0, 0, 0};
#pragma unroll
  for (int k = 0; k < kFuseJPL; ++k) {
    both[k] = in[k] && fin3(Xl[k]) && fin3(Xr[k]);
    if (both[k]) {
      cnt += 1.0;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        sr[d] += Xr[k][d];
        sl[d] += Xl[k][d];
      }
    }
  }
  cnt = wsum(cnt);
  double Xa[kFuseJPL][3];
#pragma unroll
  for (int k = 0; k < kFuseJPL; ++k)
#pragma unroll
    for (int d = 0; d < 3; ++d) Xa[k][d] = Xr[k][d];
  if ((a.prm.pad_ & 2) != 0) {
    // flags bit 1: the views already share a coordinate system (fuse/main_unity.py:96-132): no rigid alignment
  } else if (cnt >= 3.0) {
    double mr[3], ml[3], H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      mr[d] = wsum(sr[d]) / cnt;
      ml[d] = wsum(sl[d]) / cnt;
    }
#pragma unroll
    for (int k = 0; k < kFuseJPL; ++k)
      if (both[k]) {
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) H[3 * r + c] += (Xr[k][r] - mr[r]) * (Xl[k][c] - ml[c]);
      }
#pragma unroll
    for (int q = 0; q < 9; ++q) H[q] = wsum(H[q]);
    double R[9];
    if ((a.prm.pad_ & 1) || !polar_rotation(H, R)) kabsch_rotation(H, R);  // pad_ bit 0: test hook forcing the Jacobi SVD path
    double tr[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) tr[r] = ml[r] - (R[3 * r] * mr[0] + R[3 * r + 1] * mr[1] + R[3 * r + 2] * mr[2]);
#pragma unroll
    for (int k = 0; k < kFuseJPL; ++k)
      if (both[k]) {
#pragma unroll
        for (int r = 0; r < 3; ++r) Xa[k][r] = (R[3 * r] * Xr[k][0] + R[3 * r + 1] * Xr[k][1] + R[3 * r + 2] * Xr[k][2]) + tr[r];
      }
  } else {
    st |= SKA_FUSE_NO_ALIGN;
  }

  // ---- per-view weak-perspective reprojection confidence (confidence.py:62-108)
  double c1l[kFuseJPL], c1r[kFuseJPL];
  if (!weakpersp_conf(Xl, Ul, in, a.prm.sigma_px, a.prm.min_points, c1l)) st |= SKA_FUSE_FIT_LEFT_FAILED;
  if (!weakpersp_conf(Xr, Ur, in, a.prm.sigma_px, a.prm.min_points, c1r)) st |= SKA_FUSE_FIT_RIGHT_FAILED;

  // ---- cross-view consistency in the canonical body frame (confidence.py:183-224), RAW right view
  const Canon ca = canonical_frame(Xlf, a.prm), cb = canonical_frame(Xrf, a.prm);
  const double sig3 = a.prm.sigma_3d > 1e-12 ? a.prm.sigma_3d : 1e-12;
  const double inv2s3 = 1.0 / (2.0 * sig3 * sig3);
  const double isa = ca.ok ? 1.0 / ca.s : 0.0, isb = cb.ok ? 1.0 / cb.s : 0.0;
  const bool fit_failed = (st & (SKA_FUSE_FIT_LEFT_FAILED | SKA_FUSE_FIT_RIGHT_FAILED)) != 0;
#pragma unroll
  for (int k = 0; k < kFuseJPL; ++k) {
    if (!in[k]) continue;
    const int j = lane + 32 * k;
    double b2 = -1.0;  // exponent of the cross-view confidence, < 0 = undefined (confidence 0)
    if (ca.ok && cb.ok) {
      double d2 = 0.0;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const double pa = (ca.R[r][0] * (Xl[k][0] - ca.root[0]) + ca.R[r][1] * (Xl[k][1] - ca.root[1]) + ca.R[r][2] * (Xl[k][2] - ca.root[2])) * isa;
        const double pb = (cb.R[r][0] * (Xr[k][0] - cb.root[0]) + cb.R[r][1] * (Xr[k][1] - cb.root[1]) + cb.R[r][2] * (Xr[k][2] - cb.root[2])) * isb;
        d2 += (pa - pb) * (pa - pb);
      }
      if (isfinite(d2)) b2 = d2 * inv2s3;
    }
    // ---- q = sqrt(conf1 * conf2) = exp(-(arg1 + arg2) / 2): one exp per view instead of three exps and two roots
    const double ql = (c1l[k] >= 0.0 && b2 >= 0.0) ? exp(-0.5 * (c1l[k] + b2)) : 0.0;
    const double qr = (c1r[k] >= 0.0 && b2 >= 0.0) ? exp(-0.5 * (c1r[k] + b2)) : 0.0;
    // ---- softmax fusion (fuse.py:87-94, 289-326): exp(q - max) is 1 for the larger quality
    const double eo = exp(-fabs(ql - qr));
    const double ea = ql >= qr ? 1.0 : eo, eb = ql >= qr ? eo : 1.0;
    const double iss = 1.0 / (ea + eb + kFuseEps);
    const double wl = ea * iss, wr = eb * iss;
    const double iw = 1.0 / (wl + wr + kFuseEps);
    const bool okl = fin3(Xl[k]), okr = fin3(Xa[k]);
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    double f[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      if (okl && okr) f[d] = (wl * Xl[k][d] + wr * Xa[k][d]) * iw;
      else if (okl) f[d] = Xl[k][d];
      else if (okr) f[d] = Xa[k][d];
      else f[d] = nan;
      if (fit_failed) f[d] = nan;  // the reference raises ValueError for such a frame (confidence.py:31-32, 52-53)
    }
    const int64_t o = frame * J + j;
#pragma unroll
    for (int d = 0; d < 3; ++d) a.fused[3 * o + d] = f[d];
    if (a.ql != nullptr) a.ql[o] = fit_failed ? nan : ql;
    if (a.qr != nullptr) a.qr[o] = fit_failed ? nan : qr;
    if (a.aligned != nullptr) {
#pragma unroll
      for (int d = 0; d < 3; ++d) a.aligned[3 * o + d] = okr ? Xa[k][d] : nan;  // _array_to_dict drops non-finite rows
    }
  }
  if (a.status != nullptr && lane == 0) a.status[frame] = st;
}

// ------------------------------------------------------------------------------------------------
// Product path: the same arithmetic as fuse_frames_kernel with the FRAME-LEVEL work de-duplicated and every stage at its
// own register budget.  In the warp-per-frame kernel above every lane repeats the frame's 3x3 polar / SVD iteration, the
// two weak-perspective fits and the two canonical frames (~2 600 of its 4 840 warp instructions per frame), at 12 warps
// per SM.  Three launches over a caller-provided workspace of kFuseRow doubles per frame:
//   fuse_moments_kernel  warp per frame: the 42 raw moments (counts, sums, second moments) of the frame's joints,
//                        fixed-order butterfly sums;
//   fuse_params_kernel   thread per frame: moments -> rigid alignment, both weak-perspective fits, both canonical frames
//                        (32 frames' factorisations side by side in a warp instead of one repeated 32 times); the 53
//                        parameters overwrite the frame's row;
//   fuse_joints_kernel   thread per (frame, joint): confidences, softmax fusion, stores - fully coalesced, ~70 registers.
// Centred second moments come from raw ones (sum x y^T - n mx my^T): coordinates are O(10) with O(0.3) spread, so the
// cancellation costs < 3 of fp64's 16 digits - far inside the 1e-9 parity tolerance.
constexpr int kFuseRow = 56;   // doubles per frame in the workspace (42 moments, then 53 parameters)
constexpr int kMomK = 0, kMomL = 16, kMomR = 29, kMomN = 42;

struct FitWP {  // weak-perspective map of one view: uhat = s X M + t
  double M[3][2], s, t[2];
  bool ok;
};

// moments m[0..12] = cnt, sum X (3), sum U (2), sum X_a U_b (6, a-major), sum |X|^2
__device__ __forceinline__ FitWP fit_from_moments(const double* m, int min_points) {
  FitWP f;
  f.ok = false;
  f.s = 0.0;
  f.t[0] = f.t[1] = 0.0;
#pragma unroll
  for (int d = 0; d < 3; ++d) f.M[d][0] = f.M[d][1] = 0.0;
  const double n = m[0];
  if (n < (double)min_points) return f;
  const double in = 1.0 / n;
  const double muX[3] = {m[1] * in, m[2] * in, m[3] * in}, muU[2] = {m[4] * in, m[5] * in};
  double c1[3], c2[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    c1[d] = m[6 + 2 * d] - n * muX[d] * muU[0];
    c2[d] = m[7 + 2 * d] - n * muX[d] * muU[1];
  }
  const double den = m[12] - n * (muX[0] * muX[0] + muX[1] * muX[1] + muX[2] * muX[2]);
  if (!(den >= 1e-12)) return f;
  double v1[2] = {1.0, 0.0}, v2[2] = {0.0, 1.0};
  hestenes_rotate<2>(c1, c2, v1, v2);
  const double s1 = sqrt(c1[0] * c1[0] + c1[1] * c1[1] + c1[2] * c1[2]);
  const double s2 = sqrt(c2[0] * c2[0] + c2[1] * c2[1] + c2[2] * c2[2]);
  const double r1 = s1 > 0.0 ? 1.0 / s1 : 0.0, r2 = s2 > 0.0 ? 1.0 / s2 : 0.0;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    f.M[d][0] = c1[d] * r1 * v1[0] + c2[d] * r2 * v2[0];
    f.M[d][1] = c1[d] * r1 * v1[1] + c2[d] * r2 * v2[1];
  }
  f.s = (s1 + s2) / den;
  f.t[0] = muU[0] - f.s * (muX[0] * f.M[0][0] + muX[1] * f.M[1][0] + muX[2] * f.M[2][0]);
  f.t[1] = muU[1] - f.s * (muX[0] * f.M[0][1] + muX[1] * f.M[1][1] + muX[2] * f.M[2][1]);
  f.ok = true;
  return f;
}

struct FuseArgs3 {
  FuseArgs a;
  double* ws;  // [T][kFuseRow]
};

// kLPF lanes share a frame (32 / kLPF frames per warp): the butterfly sums need log2(kLPF) steps instead of 5 and run for
// all of the warp's frames at once - with 8 lanes per frame the 42 reductions cost 95 warp instructions per frame
// instead of 630 (measured: 2.49 -> see profiles/README.md).
constexpr int kLPF = 8;

__global__ void __launch_bounds__(256) fuse_moments_kernel(const FuseArgs3 g) {
  const FuseArgs& a = g.a;
  const int lane = threadIdx.x & 31, sub = lane & (kLPF - 1);
  const int64_t warp_global = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t frame = warp_global * (32 / kLPF) + (lane / kLPF);
  const bool live = frame < a.T;   // no early return: the shuffles below are warp-wide
  const int64_t fr = live ? frame : a.T - 1;
  const int J = a.J;
  const double* Xlf = a.Xl + fr * J * 3;
  const double* Xrf = a.Xr + fr * J * 3;
  const double* Ulf = a.Ul + fr * J * 2;
  const double* Urf = a.Ur + fr * J * 2;
  double m[kMomN];
#pragma unroll
  for (int q = 0; q < kMomN; ++q) m[q] = 0.0;
#pragma unroll 1
  for (int j = sub; j < J; j += kLPF) {
    double xl[3], xr[3], ul[2], ur[2];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      xl[d] = Xlf[3 * j + d];
      xr[d] = Xrf[3 * j + d];
    }
    ul[0] = Ulf[2 * j];
    ul[1] = Ulf[2 * j + 1];
    ur[0] = Urf[2 * j];
    ur[1] = Urf[2 * j + 1];
    const bool fl = fin3(xl), fr_ = fin3(xr);
    if (fl && fr_) {
      m[kMomK] += 1.0;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        m[kMomK + 1 + d] += xr[d];
        m[kMomK + 4 + d] += xl[d];
#pragma unroll
        for (int e = 0; e < 3; ++e) m[kMomK + 7 + 3 * d + e] += xr[d] * xl[e];
      }
    }
    if (fl && fin2(ul)) {
      m[kMomL] += 1.0;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        m[kMomL + 1 + d] += xl[d];
        m[kMomL + 6 + 2 * d] += xl[d] * ul[0];
        m[kMomL + 7 + 2 * d] += xl[d] * ul[1];
        m[kMomL + 12] += xl[d] * xl[d];
      }
      m[kMomL + 4] += ul[0];
      m[kMomL + 5] += ul[1];
    }
    if (fr_ && fin2(ur)) {
      m[kMomR] += 1.0;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        m[kMomR + 1 + d] += xr[d];
        m[kMomR + 6 + 2 * d] += xr[d] * ur[0];
        m[kMomR + 7 + 2 * d] += xr[d] * ur[1];
        m[kMomR + 12] += xr[d] * xr[d];
      }
      m[kMomR + 4] += ur[0];
      m[kMomR + 5] += ur[1];
    }
  }
  double* row = g.ws + fr * kFuseRow;
#pragma unroll
  for (int q = 0; q < kMomN; ++q) {
    double v = m[q];
#pragma unroll
    for (int o = kLPF / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);  // fixed order within the frame's lanes
    if (live && sub == 0) row[q] = v;
  }
}

// The block's 64 workspace rows (one contiguous run) pass through shared memory: a thread reading / writing its own
// 448-byte row directly scatters every access over 32 sectors (the kernel ran at 12 % issue utilisation on those loads).
// Row stride in shared memory is 57 doubles (odd in 8-byte words: the 32 threads of a warp hit distinct bank pairs).
constexpr int kParamThreads = 64, kParamRowS = kFuseRow + 1;

__global__ void __launch_bounds__(kParamThreads) fuse_params_kernel(const FuseArgs3 g) {
  __shared__ double s_rows[kParamThreads * kParamRowS];
  const FuseArgs& a = g.a;
  const int64_t f0 = (int64_t)blockIdx.x * kParamThreads;
  const int64_t frame = f0 + threadIdx.x;
  const int nv = (int)((a.T - f0) < kParamThreads ? (a.T - f0) : kParamThreads);
  {
    const double* src = g.ws + f0 * kFuseRow;
    for (int e = threadIdx.x; e < nv * kFuseRow; e += kParamThreads) s_rows[(e / kFuseRow) * kParamRowS + (e % kFuseRow)] = src[e];
  }
  __syncthreads();
  if (frame < a.T) {
  const int J = a.J;
  double* out = s_rows + threadIdx.x * kParamRowS;
  double m[kMomN];
#pragma unroll
  for (int q = 0; q < kMomN; ++q) m[q] = out[q];
  uint8_t st = 0;
  // rigid alignment right -> left (main_raw.py:48-95); flags bit 1: views already share a frame (main_unity.py:96-132)
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, tr[3] = {0, 0, 0};
  bool aligned = false;
  if ((a.prm.pad_ & 2) == 0) {
    const double n = m[kMomK];
    if (n >= 3.0) {
      const double in = 1.0 / n;
      double mr[3], ml[3], H[9];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        mr[d] = m[kMomK + 1 + d] * in;
        ml[d] = m[kMomK + 4 + d] * in;
      }
#pragma unroll
      for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int e = 0; e < 3; ++e) H[3 * d + e] = m[kMomK + 7 + 3 * d + e] - n * mr[d] * ml[e];
      if ((a.prm.pad_ & 1) || !polar_rotation(H, R)) kabsch_rotation(H, R);
#pragma unroll
      for (int r = 0; r < 3; ++r) tr[r] = ml[r] - (R[3 * r] * mr[0] + R[3 * r + 1] * mr[1] + R[3 * r + 2] * mr[2]);
      aligned = true;
    } else {
      st |= SKA_FUSE_NO_ALIGN;
    }
  }
  const FitWP fl = fit_from_moments(m + kMomL, a.prm.min_points), fr = fit_from_moments(m + kMomR, a.prm.min_points);
  if (!fl.ok) st |= SKA_FUSE_FIT_LEFT_FAILED;
  if (!fr.ok) st |= SKA_FUSE_FIT_RIGHT_FAILED;
  const Canon ca = canonical_frame(a.Xl + frame * J * 3, a.prm), cb = canonical_frame(a.Xr + frame * J * 3, a.prm);
  int o = 0;
#pragma unroll
  for (int q = 0; q < 9; ++q) out[o++] = R[q];
#pragma unroll
  for (int q = 0; q < 3; ++q) out[o++] = tr[q];
  const FitWP* fits[2] = {&fl, &fr};
#pragma unroll
  for (int v = 0; v < 2; ++v) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      out[o++] = fits[v]->s * fits[v]->M[d][0];  // s folded into the map
      out[o++] = fits[v]->s * fits[v]->M[d][1];
    }
    out[o++] = fits[v]->t[0];
    out[o++] = fits[v]->t[1];
  }
  const Canon* cans[2] = {&ca, &cb};
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    const double is = cans[v]->ok ? 1.0 / cans[v]->s : 0.0;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) out[o++] = cans[v]->R[r][c] * is;  // scale folded into the rotation rows
#pragma unroll
    for (int c = 0; c < 3; ++c) out[o++] = cans[v]->root[c];
  }
  out[o++] = (double)(st | (aligned ? 8 : 0) | (ca.ok && cb.ok ? 16 : 0));  // 28 + 24 + 1 = 53 doubles
  if (a.status != nullptr) a.status[frame] = st;
  }
  __syncthreads();
  {
    double* dst = g.ws + f0 * kFuseRow;
    for (int e = threadIdx.x; e < nv * kFuseRow; e += kParamThreads) dst[e] = s_rows[(e / kFuseRow) * kParamRowS + (e % kFuseRow)];
  }
}

__global__ void __launch_bounds__(256) fuse_joints_kernel(const FuseArgs3 g) {
  const FuseArgs& a = g.a;
  const int J = a.J;
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= a.T * J) return;
  const int64_t frame = o / J;
  const int j = (int)(o - frame * J);
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  const double sig = a.prm.sigma_px > 1e-12 ? a.prm.sigma_px : 1e-12;
  const double inv2s = 1.0 / (2.0 * sig * sig);
  const double sig3 = a.prm.sigma_3d > 1e-12 ? a.prm.sigma_3d : 1e-12;
  const double inv2s3 = 1.0 / (2.0 * sig3 * sig3);
  const double* P = g.ws + frame * kFuseRow;
  const int flags = (int)P[52];
  const bool fit_failed = (flags & (SKA_FUSE_FIT_LEFT_FAILED | SKA_FUSE_FIT_RIGHT_FAILED)) != 0;
  const bool aligned = (flags & 8) != 0, canon_ok = (flags & 16) != 0;
  const bool fit_l = !(flags & SKA_FUSE_FIT_LEFT_FAILED), fit_r = !(flags & SKA_FUSE_FIT_RIGHT_FAILED);
  const double* Xlf = a.Xl + frame * J * 3;
  const double* Xrf = a.Xr + frame * J * 3;
  const double* Ulf = a.Ul + frame * J * 2;
  const double* Urf = a.Ur + frame * J * 2;
  double xl[3], xr[3], ul[2], ur[2];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    xl[d] = Xlf[3 * j + d];
    xr[d] = Xrf[3 * j + d];
  }
  ul[0] = Ulf[2 * j];
  ul[1] = Ulf[2 * j + 1];
  ur[0] = Urf[2 * j];
  ur[1] = Urf[2 * j + 1];
  const bool okl = fin3(xl), okr_raw = fin3(xr);
  // right view in the left frame
  double xa[3] = {xr[0], xr[1], xr[2]};
  if (aligned && okl && okr_raw) {
#pragma unroll
    for (int r = 0; r < 3; ++r) xa[r] = (P[3 * r] * xr[0] + P[3 * r + 1] * xr[1] + P[3 * r + 2] * xr[2]) + P[9 + r];
  }
  // weak-perspective residual exponents (confidence.py:62-108); < 0 = confidence 0
  double a1l = -1.0, a1r = -1.0;
  if (fit_l) {
    const double* W = P + 12;
    const double d0 = (xl[0] * W[0] + xl[1] * W[2] + xl[2] * W[4]) + W[6] - ul[0];
    const double d1 = (xl[0] * W[1] + xl[1] * W[3] + xl[2] * W[5]) + W[7] - ul[1];
    const double e2 = d0 * d0 + d1 * d1;
    if (isfinite(e2)) a1l = e2 * inv2s;
  }
  if (fit_r) {
    const double* W = P + 20;
    const double d0 = (xr[0] * W[0] + xr[1] * W[2] + xr[2] * W[4]) + W[6] - ur[0];
    const double d1 = (xr[0] * W[1] + xr[1] * W[3] + xr[2] * W[5]) + W[7] - ur[1];
    const double e2 = d0 * d0 + d1 * d1;
    if (isfinite(e2)) a1r = e2 * inv2s;
  }
  // cross-view distance in the canonical frames (confidence.py:183-224), RAW right view
  double b2 = -1.0;
  if (canon_ok) {
    const double* A = P + 28;
    const double* B = P + 40;
    double d2 = 0.0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const double pa = A[3 * r] * (xl[0] - A[9]) + A[3 * r + 1] * (xl[1] - A[10]) + A[3 * r + 2] * (xl[2] - A[11]);
      const double pb = B[3 * r] * (xr[0] - B[9]) + B[3 * r + 1] * (xr[1] - B[10]) + B[3 * r + 2] * (xr[2] - B[11]);
      d2 += (pa - pb) * (pa - pb);
    }
    if (isfinite(d2)) b2 = d2 * inv2s3;
  }
  const double ql = (a1l >= 0.0 && b2 >= 0.0) ? exp(-0.5 * (a1l + b2)) : 0.0;
  const double qr = (a1r >= 0.0 && b2 >= 0.0) ? exp(-0.5 * (a1r + b2)) : 0.0;
  const double eo = exp(-fabs(ql - qr));
  const double ea = ql >= qr ? 1.0 : eo, eb = ql >= qr ? eo : 1.0;
  const double iss = 1.0 / (ea + eb + kFuseEps);
  const double wl = ea * iss, wr = eb * iss;
  const double iw = 1.0 / (wl + wr + kFuseEps);
  const bool okr = fin3(xa);
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    double fv;
    if (okl && okr) fv = (wl * xl[d] + wr * xa[d]) * iw;
    else if (okl) fv = xl[d];
    else if (okr) fv = xa[d];
    else fv = nan;
    a.fused[3 * o + d] = fit_failed ? nan : fv;  // the reference raises ValueError for such a frame
  }
  if (a.ql != nullptr) a.ql[o] = fit_failed ? nan : ql;
  if (a.qr != nullptr) a.qr[o] = fit_failed ? nan : qr;
  if (a.aligned != nullptr) {
#pragma unroll
    for (int d = 0; d < 3; ++d) a.aligned[3 * o + d] = okr ? xa[d] : nan;
  }
}

// ------------------------------------------------------------------------------------------------
// rigid_transform_3D of bundle_adjustment/fuse/fuse.py:96-232 (the fusion the bundle_adjustment / fuse.side /
// front_side.side pipelines call) for a whole clip: Umeyama alignment of the right view onto the left from the five torso
// joints (fuse_check.py:26-78: cross-covariance / N, SVD with det fix, optional scale sum(S) / var), threshold fusion per
// joint (fuse.py:55-93) and the per-frame diagnostics (plain means, NaN when a joint is missing - like the reference).
//   rigid_params_kernel  thread per frame -> [R (9) | t (3) | s] and status
//   rigid_fuse_kernel    kLPF lanes per frame: aligned right view, fused joints, the three mean distances
struct RigidArgs {
  const double* L;
  const double* R;
  int64_t T;
  int32_t J;
  int32_t torso[5];
  int32_t allow_scale;
  double tau;
  const double* tau_j;  // nullable (J,)
  const double* wL;     // nullable, (T,J) or (J,) with w_sT = 0
  const double* wR;
  int64_t w_sT;
  double* fused;
  double* Rts;          // (T,13)
  double* diag;         // nullable (T,4)
  uint8_t* status;      // nullable
};

__global__ void __launch_bounds__(64) rigid_params_kernel(const RigidArgs a) {
  const int64_t frame = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (frame >= a.T) return;
  const double* Lf = a.L + frame * a.J * 3;
  const double* Rf = a.R + frame * a.J * 3;
  double tl[5][3], sr[5][3];
  bool ok[5];
  double n = 0.0, tm[3] = {0, 0, 0}, sm[3] = {0, 0, 0};
#pragma unroll
  for (int k = 0; k < 5; ++k) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      tl[k][d] = Lf[3 * a.torso[k] + d];
      sr[k][d] = Rf[3 * a.torso[k] + d];
    }
    ok[k] = fin3(tl[k]) && fin3(sr[k]);
    if (ok[k]) {
      n += 1.0;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        tm[d] += tl[k][d];
        sm[d] += sr[k][d];
      }
    }
  }
  double* out = a.Rts + frame * 13;
  if (n < 3.0) {  // the reference raises ValueError (fuse_check.py:43-44)
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    for (int q = 0; q < 13; ++q) out[q] = nan;
    if (a.status != nullptr) a.status[frame] = 1;
    return;
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    tm[d] /= n;
    sm[d] /= n;
  }
  double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, var = 0.0;
#pragma unroll
  for (int k = 0; k < 5; ++k)
    if (ok[k]) {
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const double sc = sr[k][r] - sm[r];
        var += sc * sc;
#pragma unroll
        for (int c = 0; c < 3; ++c) H[3 * r + c] += sc * (tl[k][c] - tm[c]);
      }
    }
#pragma unroll
  for (int q = 0; q < 9; ++q) H[q] /= n;
  double Rm[9], sumS = 0.0;
  if (!polar_rotation(H, Rm, &sumS)) kabsch_rotation(H, Rm, &sumS);
  const double s = a.allow_scale ? sumS / (var / n + 1e-12) : 1.0;
#pragma unroll
  for (int q = 0; q < 9; ++q) out[q] = Rm[q];
#pragma unroll
  for (int r = 0; r < 3; ++r) out[9 + r] = tm[r] - s * (Rm[3 * r] * sm[0] + Rm[3 * r + 1] * sm[1] + Rm[3 * r + 2] * sm[2]);
  out[12] = s;
  if (a.status != nullptr) a.status[frame] = 0;
}

__global__ void __launch_bounds__(256) rigid_fuse_kernel(const RigidArgs a) {
  const int lane = threadIdx.x & 31, sub = lane & (kLPF - 1);
  const int64_t warp_global = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t frame = warp_global * (32 / kLPF) + (lane / kLPF);
  const bool live = frame < a.T;
  const int64_t fr = live ? frame : a.T - 1;
  const int J = a.J;
  const double* Lf = a.L + fr * J * 3;
  const double* Rf = a.R + fr * J * 3;
  const double* P = a.Rts + fr * 13;
  double Rm[9], tv[3];
#pragma unroll
  for (int q = 0; q < 9; ++q) Rm[q] = P[q];
#pragma unroll
  for (int q = 0; q < 3; ++q) tv[q] = P[9 + q];
  const double s = P[12];
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  double d_lr = 0.0, d_fl = 0.0, d_fr = 0.0;
#pragma unroll 1
  for (int j = sub; j < J; j += kLPF) {
    double l[3], r[3], ra[3], f[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      l[d] = Lf[3 * j + d];
      r[d] = Rf[3 * j + d];
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) ra[q] = s * (Rm[3 * q] * r[0] + Rm[3 * q + 1] * r[1] + Rm[3 * q + 2] * r[2]) + tv[q];
    const bool okl = fin3(l), okr = fin3(ra);
    const double wl = a.wL != nullptr ? a.wL[fr * a.w_sT + j] : 1.0, wr = a.wR != nullptr ? a.wR[fr * a.w_sT + j] : 1.0;
    const double tau = a.tau_j != nullptr ? a.tau_j[j] : a.tau;
    if (okl && okr) {
      const double e0 = l[0] - ra[0], e1 = l[1] - ra[1], e2 = l[2] - ra[2];
      const double dist = sqrt(e0 * e0 + e1 * e1 + e2 * e2);
      if (dist > tau) {
        const bool pick_l = wl >= wr;
#pragma unroll
        for (int d = 0; d < 3; ++d) f[d] = pick_l ? l[d] : ra[d];
      } else {
        const double iw = wl + wr + 1e-9;
#pragma unroll
        for (int d = 0; d < 3; ++d) f[d] = (wl * l[d] + wr * ra[d]) / iw;
      }
    } else {
#pragma unroll
      for (int d = 0; d < 3; ++d) f[d] = okl ? l[d] : (okr ? ra[d] : nan);
    }
    if (live) {
#pragma unroll
      for (int d = 0; d < 3; ++d) a.fused[(fr * J + j) * 3 + d] = f[d];
    }
    // diagnostics against the RAW right view (fuse.py:205-208): plain means, NaN propagates
    {
      const double a0 = l[0] - r[0], a1 = l[1] - r[1], a2 = l[2] - r[2];
      const double b0 = f[0] - l[0], b1 = f[1] - l[1], b2 = f[2] - l[2];
      const double c0 = f[0] - r[0], c1 = f[1] - r[1], c2 = f[2] - r[2];
      d_lr += sqrt(a0 * a0 + a1 * a1 + a2 * a2);
      d_fl += sqrt(b0 * b0 + b1 * b1 + b2 * b2);
      d_fr += sqrt(c0 * c0 + c1 * c1 + c2 * c2);
    }
  }
#pragma unroll
  for (int o = kLPF / 2; o > 0; o >>= 1) {
    d_lr += __shfl_xor_sync(0xffffffffu, d_lr, o);
    d_fl += __shfl_xor_sync(0xffffffffu, d_fl, o);
    d_fr += __shfl_xor_sync(0xffffffffu, d_fr, o);
  }
  if (live && sub == 0 && a.diag != nullptr) {
    const double lr = d_lr / J, fl = d_fl / J, frr = d_fr / J;
    double* o = a.diag + fr * 4;
    o[0] = lr;
    o[1] = fl;
    o[2] = frr;
    o[3] = lr - 0.5 * (fl + frr);
  }
}

int rigid_fuse(const double* L, const double* R, int64_t T, int J, const int32_t* torso5, double tau, const double* tau_j, int allow_scale,
               const double* wL, const double* wR, int64_t w_sT, double* fused, double* Rts, double* diag, uint8_t* status,
               cudaStream_t s) {
  if (T == 0) return SKA_OK;
  RigidArgs a;
  a.L = L;
  a.R = R;
  a.T = T;
  a.J = J;
  for (int k = 0; k < 5; ++k) a.torso[k] = torso5[k];
  a.allow_scale = allow_scale;
  a.tau = tau;
  a.tau_j = tau_j;
  a.wL = wL;
  a.wR = wR;
  a.w_sT = w_sT;
  a.fused = fused;
  a.Rts = Rts;
  a.diag = diag;
  a.status = status;
  rigid_params_kernel<<<(unsigned)((T + 63) / 64), 64, 0, s>>>(a);
  const int64_t fpb = 8 * (32 / kLPF);
  rigid_fuse_kernel<<<(unsigned)((T + fpb - 1) / fpb), 256, 0, s>>>(a);
  const cudaError_t ce = cudaGetLastError();
  return ce == cudaSuccess ? SKA_OK : set_error((int)ce, cudaGetErrorString(ce));
}

// ------------------------------------------------------------------------------------------------
struct EmaArgs {
  const double* X;
  double* Y;
  const double* alpha_joint;
  int64_t T;
  int32_t J;
  int32_t adaptive;
  double alpha, alpha_min, alpha_max, speed_gain;
  int64_t chunk;
  int32_t halo;  // valid samples to replay before a chunk; < 0: exact sequential scan (one chunk)
  int64_t n_chunks;
};

__device__ __forceinline__ void ema_step(double (&y)[3], bool& oky, const double x[3], double aj, const EmaArgs& a) {
  const bool okx = isfinite(x[0]) && isfinite(x[1]) && isfinite(x[2]);
  if (okx && oky) {
    double al = a.alpha;
    if (a.adaptive) {
      const double d0 = __dsub_rn(x[0], y[0]), d1 = __dsub_rn(x[1], y[1]), d2 = __dsub_rn(x[2], y[2]);
      const double speed = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2)));
      al = __dadd_rn(aj, __dmul_rn(a.speed_gain, speed));
      al = fmin(fmax(al, a.alpha_min), a.alpha_max);
    }
    const double be = __dsub_rn(1.0, al);
#pragma unroll
    for (int d = 0; d < 3; ++d) y[d] = __dadd_rn(__dmul_rn(al, x[d]), __dmul_rn(be, y[d]));
  } else if (okx) {
    y[0] = x[0];
    y[1] = x[1];
    y[2] = x[2];
    oky = true;
  }  // else hold: y unchanged (stays NaN until the first finite sample)
}

__global__ void __launch_bounds__(128) ema_kernel(const EmaArgs a) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int J = a.J;
  const int64_t c = tid / J;
  const int j = (int)(tid - c * J);
  if (c >= a.n_chunks) return;
  const int64_t t0 = c * a.chunk;
  const int64_t t1 = (t0 + a.chunk < a.T) ? t0 + a.chunk : a.T;
  const int64_t row = (int64_t)J * 3;
  const double* Xj = a.X + 3 * j;
  double* Yj = a.Y + 3 * j;
  const double aj = a.alpha_joint[j];
  // start: `halo` finite samples before the chunk (or frame 0)
  int64_t ts = t0;
  if (a.halo >= 0) {
    int seen = 0;
    while (ts > 0 && seen < a.halo) {
      --ts;
      const double* p = Xj + ts * row;
      if (isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2])) ++seen;
    }
  } else {
    ts = 0;
  }
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  double y[3];
  bool oky;
  {
    const double* p = Xj + ts * row;
    oky = isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]);  // array_to_dict keeps finite rows only (fuse.py:76-82)
    y[0] = oky ? p[0] : nan;
    y[1] = oky ? p[1] : nan;
    y[2] = oky ? p[2] : nan;
    if (ts >= t0) {
      double* q = Yj + ts * row;
      q[0] = y[0];
      q[1] = y[1];
      q[2] = y[2];
    }
  }
  // pointer-increment addressing (the row stride is loop invariant), eight rows in flight: the loads do not depend on
  // the recurrence, so the only exposed latency is the first batch
  constexpr int kU = 8;
  int64_t t = ts + 1;
  const double* px = Xj + t * row;
  double* py = Yj + t * row;
  for (; t + kU <= t1; t += kU) {
    double x[kU][3];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      x[u][0] = px[0];
      x[u][1] = px[1];
      x[u][2] = px[2];
      px += row;
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      ema_step(y, oky, x[u], aj, a);
      if (t + u >= t0) {
        py[0] = y[0];
        py[1] = y[1];
        py[2] = y[2];
      }
      py += row;
    }
  }
  for (; t < t1; ++t) {
    const double x[3] = {px[0], px[1], px[2]};
    ema_step(y, oky, x, aj, a);
    if (t >= t0) {
      py[0] = y[0];
      py[1] = y[1];
      py[2] = y[2];
    }
    px += row;
    py += row;
  }
}

// ------------------------------------------------------------------------------------------------
size_t fuse_workspace_bytes(int64_t T) { return (size_t)(T > 0 ? T : 0) * kFuseRow * sizeof(double); }

int fuse_frames(const double* Xl, const double* Xr, const double* Ul, const double* Ur, int64_t T, int J, const SkaFuseParams& prm,
                double* fused, double* ql, double* qr, double* aligned, uint8_t* status, void* ws, size_t ws_bytes, cudaStream_t s) {
  if (T == 0) return SKA_OK;
  FuseArgs a{Xl, Xr, Ul, Ur, T, J, prm, fused, ql, qr, aligned, status};
  const int wpb = 4;
  if ((prm.pad_ & 4) == 0) {  // product path: moments -> per-frame parameters -> per-joint fusion
    if (ws == nullptr || ws_bytes < (size_t)T * kFuseRow * sizeof(double))
      return set_error(SKA_EWORKSPACE, "workspace too small (see ska_fuse_workspace_bytes)");
    if ((T * (int64_t)J + 255) / 256 > 0x7fffffffLL) return set_error(SKA_EINVAL, "too many joints for one launch; shard the clip");
    const FuseArgs3 g{a, (double*)ws};
    {
      const int64_t fpb = 8 * (32 / kLPF);  // frames per 256-thread block
      fuse_moments_kernel<<<(unsigned)((T + fpb - 1) / fpb), 256, 0, s>>>(g);
    }
    fuse_params_kernel<<<(unsigned)((T + 63) / 64), 64, 0, s>>>(g);
    fuse_joints_kernel<<<(unsigned)((T * (int64_t)J + 255) / 256), 256, 0, s>>>(g);
    const cudaError_t ce = cudaGetLastError();
    return ce == cudaSuccess ? SKA_OK : set_error((int)ce, cudaGetErrorString(ce));
  }
  const int64_t grid = (T + wpb - 1) / wpb;  // flags bit 2: the warp-per-frame kernel (A/B testing)
  if (grid > 0x7fffffffLL) return set_error(SKA_EINVAL, "too many frames for one launch; shard the clip");
  fuse_frames_kernel<<<(unsigned)grid, 32 * wpb, 0, s>>>(a);
  const cudaError_t ce = cudaGetLastError();
  return ce == cudaSuccess ? SKA_OK : set_error((int)ce, cudaGetErrorString(ce));
}

int ema_smooth(const double* X, int64_t T, int J, const double* alpha_joint, int adaptive, double alpha, double alpha_min,
               double alpha_max, double speed_gain, int64_t chunk, int halo, double* Y, cudaStream_t s) {
  if (T == 0) return SKA_OK;
  EmaArgs a;
  a.X = X;
  a.Y = Y;
  a.alpha_joint = alpha_joint;
  a.T = T;
  a.J = J;
  a.adaptive = adaptive;
  a.alpha = alpha;
  a.alpha_min = alpha_min;
  a.alpha_max = alpha_max;
  a.speed_gain = speed_gain;
  if (halo < 0 || chunk <= 0 || chunk >= T) {
    a.chunk = T;
    a.halo = -1;
    a.n_chunks = 1;
  } else {
    a.chunk = chunk;
    a.halo = halo;
    a.n_chunks = (T + chunk - 1) / chunk;
  }
  const int64_t threads = a.n_chunks * J;
  const int64_t grid = (threads + 127) / 128;
  if (grid > 0x7fffffffLL) return set_error(SKA_EINVAL, "too many chunks for one launch");
  ema_kernel<<<(unsigned)grid, 128, 0, s>>>(a);
  const cudaError_t ce = cudaGetLastError();
  return ce == cudaSuccess ? SKA_OK : set_error((int)ce, cudaGetErrorString(ce));
}

}  // namespace ska
