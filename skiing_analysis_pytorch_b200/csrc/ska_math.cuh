// ska_math.cuh - per-point arithmetic shared by the sm_100a kernels and the host-emulation test
// harness (tests/hostemu).  Everything here is __host__ __device__ so the exact instruction
// sequence the GPU runs (fmaf chains, hi/lo splits, secular iteration) can be exercised on the
// CPU-only authoring box; the product only ever calls it from device code.
//
// Math restated from SURVEY.md appendix A:
//   A.1 weighted DLT rows w(u P2 - P0), w(v P2 - P1)   (vggt/triangulate.py:23-31, cv2.triangulatePoints)
//   A.2 rational/tangential/thin-prism distortion      (cv2.projectPoints; triangulation/reproject.py:77-78)
#pragma once
#include <math.h>
#include <stdint.h>

#include "ska_vec.cuh"

namespace ska {

// ---------------------------------------------------------------------------------------------
// Kernel-side camera, fp32, prepared in fp64 (prep_camera in ska_prep.h).
// All quantities are expressed in CENTRED coordinates Y = X - c: P' = K [R | R c + t].
// Centring is numerical conditioning only; the eigenproblem solved is the reference's
// (un-centred, ||X~|| = 1) one - see secular_solve below.
struct CamDev {
  float Ph[12];   // P' hi part, row-major 3x4, scaled so that P'[2] = [r3 | t'_z] (K[2][2] == 1)
  float Pl[12];   // P' lo part (P' - (double)Ph)
  float Rxy[6];   // rows 0 and 1 of R           (normalised coordinates for the distortion model)
  float txy[2];   // (R c + t).x, .y
  float fx, fy, skew;
  float ifx, ify;  // 1/fx, 1/fy
  float ncx, ncy;  // -cx/fx, -cy/fy: normalised coordinate of pixel u = fma(u, ifx, ncx)   (zero skew)
  float dk[3];    // k1-k4, k2-k5, k3-k6  (numerator minus denominator of the rational term)
  float kd[3];    // k4, k5, k6
  float p1, p2;
  float tp1, tp2; // 2*p1, 2*p2
  float s[4];     // s1..s4
};

// Everything below is templated on T = float (one point) or F2 (two points in lockstep, packed
// FFMA2 on sm_100a - ska_vec.cuh); camera coefficients stay scalar operands.

// One DLT row pair for one view: a = u*P2 - P0, b = v*P2 - P1 (4 entries each).  The products are
// fused (single rounding of the exact u*P2h - P0h, so the cancellation between u*r3 and cx*r3
// costs nothing).  LO selects how much of the fp32-rounding of P' is restored from the lo parts:
// 0 = none, 1 = translation column only (the ~1e4-magnitude entries that set the px-error floor),
// 2 = all columns.
template <int LO, typename T>
SKA_HD void dlt_rows(const CamDev& c, T u, T v, T a[4], T b[4]) {
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    T ah = vfma(u, c.Ph[8 + m], -c.Ph[m]);
    T bh = vfma(v, c.Ph[8 + m], -c.Ph[4 + m]);
    if (LO == 2 || (LO == 1 && m == 3)) {
      ah = vadd(ah, vfma(u, c.Pl[8 + m], -c.Pl[m]));
      bh = vadd(bh, vfma(v, c.Pl[8 + m], -c.Pl[4 + m]));
    }
    a[m] = ah;
    b[m] = bh;
  }
}

// Symmetric 4x4 accumulator, upper triangle: 00 01 02 03 11 12 13 22 23 33.
template <typename T>
struct Sym4T {
  T m00, m01, m02, m03, m11, m12, m13, m22, m23, m33;
};
using Sym4 = Sym4T<float>;
template <typename T>
SKA_HD void sym4_zero(Sym4T<T>& M) {
  M.m00 = M.m01 = M.m02 = M.m03 = M.m11 = M.m12 = M.m13 = M.m22 = M.m23 = M.m33 = Vec<T>::splat(0.f);
}
template <typename T>
SKA_HD void sym4_rank1(Sym4T<T>& M, const T r[4], T w2) {
  const T s0 = vmul(r[0], w2), s1 = vmul(r[1], w2), s2 = vmul(r[2], w2), s3 = vmul(r[3], w2);
  M.m00 = vfma(s0, r[0], M.m00);
  M.m01 = vfma(s0, r[1], M.m01);
  M.m02 = vfma(s0, r[2], M.m02);
  M.m03 = vfma(s0, r[3], M.m03);
  M.m11 = vfma(s1, r[1], M.m11);
  M.m12 = vfma(s1, r[2], M.m12);
  M.m13 = vfma(s1, r[3], M.m13);
  M.m22 = vfma(s2, r[2], M.m22);
  M.m23 = vfma(s2, r[3], M.m23);
  M.m33 = vfma(s3, r[3], M.m33);
}
template <typename T>
SKA_HD void sym4_rank1_unit(Sym4T<T>& M, const T r[4]) {
  M.m00 = vfma(r[0], r[0], M.m00);
  M.m01 = vfma(r[0], r[1], M.m01);
  M.m02 = vfma(r[0], r[2], M.m02);
  M.m03 = vfma(r[0], r[3], M.m03);
  M.m11 = vfma(r[1], r[1], M.m11);
  M.m12 = vfma(r[1], r[2], M.m12);
  M.m13 = vfma(r[1], r[3], M.m13);
  M.m22 = vfma(r[2], r[2], M.m22);
  M.m23 = vfma(r[2], r[3], M.m23);
  M.m33 = vfma(r[3], r[3], M.m33);
}
template <int I, typename T>
SKA_HD Sym4 sym4_lane(const Sym4T<T>& M) {
  Sym4 r;
  r.m00 = lane<I>(M.m00); r.m01 = lane<I>(M.m01); r.m02 = lane<I>(M.m02); r.m03 = lane<I>(M.m03);
  r.m11 = lane<I>(M.m11); r.m12 = lane<I>(M.m12); r.m13 = lane<I>(M.m13);
  r.m22 = lane<I>(M.m22); r.m23 = lane<I>(M.m23); r.m33 = lane<I>(M.m33);
  return r;
}

// LDL^T of a symmetric 3x3 (a00 a01 a02 a11 a12 a22), reciprocal pivots kept.
template <typename T>
struct Ldl3T {
  T l10, l20, l21, r0, r1, r2;
  typename Vec<T>::Mask pos;  // all pivots > 0  <=>  matrix positive definite
};
using Ldl3 = Ldl3T<float>;
template <typename T>
SKA_HD Ldl3T<T> ldl3(T a00, T a01, T a02, T a11, T a12, T a22) {
  Ldl3T<T> f;
  f.r0 = rcp_fast(a00);
  f.l10 = vmul(a01, f.r0);
  f.l20 = vmul(a02, f.r0);
  const T d1 = vfma(vneg(f.l10), a01, a11);
  f.r1 = rcp_fast(d1);
  const T t21 = vfma(vneg(f.l20), a01, a12);
  f.l21 = vmul(t21, f.r1);
  const T d2 = vfma(vneg(f.l21), t21, vfma(vneg(f.l20), a02, a22));
  f.r2 = rcp_fast(d2);
  f.pos = mand(mand(vgt(a00, 0.f), vgt(d1, 0.f)), vgt(d2, 0.f));
  return f;
}
// trace(M33^-1) from the LDL^T factors: an upper bound of 1/lambda_min(M33).
template <typename T>
SKA_HD T ldl3_inv_trace(const Ldl3T<T>& f) {
  // M^-1 = L^-T D^-1 L^-1 with L^-1 = [[1,0,0],[-l10,1,0],[l10*l21-l20,-l21,1]]
  const T q = vfma(f.l10, f.l21, vneg(f.l20));
  return vfma(f.r2, vfma(q, q, vfma(f.l21, f.l21, 1.0f)), vfma(f.r1, vfma(f.l10, f.l10, 1.0f), f.r0));
}
template <typename T>
SKA_HD void ldl3_solve(const Ldl3T<T>& f, T b0, T b1, T b2, T& x0, T& x1, T& x2) {
  const T y1 = vfma(vneg(f.l10), b0, b1);
  const T y2 = vfma(vneg(f.l21), y1, vfma(vneg(f.l20), b0, b2));
  x2 = vmul(y2, f.r2);
  x1 = vfma(vneg(f.l21), x2, vmul(y1, f.r1));
  x0 = vfma(vneg(f.l20), x2, vfma(vneg(f.l10), x1, vmul(b0, f.r0)));
}

// ---------------------------------------------------------------------------------------------
// Two views' cameras interleaved: every member is the pair (view 2i, view 2i+1).  Used by the view-pair
// form of the per-point arithmetic (tri_point_vp): one packed instruction does the same step for two views of
// ONE point, and each pair of coefficients is one 64-bit constant load.  No skew / thin prism (DIST <= 1).
struct CamPairDev {
  F2 Ph[12];
  F2 Pl3[3];      // lo parts of the translation column: Pl[3], Pl[7], Pl[11]
  F2 fx, fy, ifx, ify, ncx, ncy;
  F2 dk[3], kd[3];
  F2 p1, p2, tp1, tp2;
  F2 s[4];        // unused (PRISM = false); keeps distort_delta generic
};
SKA_HD void make_cam_pair(const CamDev& a, const CamDev& b, CamPairDev& o) {
  for (int i = 0; i < 12; ++i) o.Ph[i] = mk2(a.Ph[i], b.Ph[i]);
  o.Pl3[0] = mk2(a.Pl[3], b.Pl[3]);
  o.Pl3[1] = mk2(a.Pl[7], b.Pl[7]);
  o.Pl3[2] = mk2(a.Pl[11], b.Pl[11]);
  o.fx = mk2(a.fx, b.fx); o.fy = mk2(a.fy, b.fy);
  o.ifx = mk2(a.ifx, b.ifx); o.ify = mk2(a.ify, b.ify);
  o.ncx = mk2(a.ncx, b.ncx); o.ncy = mk2(a.ncy, b.ncy);
  for (int i = 0; i < 3; ++i) { o.dk[i] = mk2(a.dk[i], b.dk[i]); o.kd[i] = mk2(a.kd[i], b.kd[i]); }
  o.p1 = mk2(a.p1, b.p1); o.p2 = mk2(a.p2, b.p2);
  o.tp1 = mk2(a.tp1, b.tp1); o.tp2 = mk2(a.tp2, b.tp2);
  for (int i = 0; i < 4; ++i) o.s[i] = mk2(a.s[i], b.s[i]);
}
// DLT rows of two views of one point (LO = 1: translation column restored from the lo parts)
SKA_HD void dlt_rows_vp(const CamPairDev& c, F2 u, F2 v, F2 a[4], F2 b[4]) {
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    a[m] = vfma(u, c.Ph[8 + m], vneg(c.Ph[m]));
    b[m] = vfma(v, c.Ph[8 + m], vneg(c.Ph[4 + m]));
  }
  a[3] = vadd(a[3], vfma(u, c.Pl3[2], vneg(c.Pl3[0])));
  b[3] = vadd(b[3], vfma(v, c.Pl3[2], vneg(c.Pl3[1])));
}

// ---------------------------------------------------------------------------------------------
// Smallest eigenpair of the un-centred DLT normal matrix M = A^T A, computed in centred
// coordinates.  With X~ = [c + Y; 1] the eigen-equations (M - lam I) X~ = 0 read
//     (M'33 - lam I) Y = -m'34 + lam c            (rows 1..3, M' = T^T M T, T = [I c; 0 1])
//     lam = |A' [Y;1]|^2 / (|c + Y|^2 + 1)        (Rayleigh quotient, evaluated from the rows so
//                                                  the 1e8 -> 1 cancellation never happens in M)
// Fixed-point iteration on lam starting from lam = 0 (the inhomogeneous least-squares point);
// quadratically convergent because the Rayleigh quotient is stationary at eigenvectors.
// Certificate: if LDL^T of (M'33 - lam I) has all pivots > 0 then lam < lambda_min(M33) and by
// Cauchy interlacing M has exactly ONE eigenvalue below lambda_min(M33) - the converged pair is
// the smallest one, i.e. the vector np.linalg.svd / cv2.triangulatePoints return.
// Callers fall back to the fp64 Jacobi solver when `ok` comes back false.
struct SecularState {
  float y0, y1, y2;  // centred solution
  float lam;
  float step2;       // squared length of the last step (contraction monitor)
  bool ok;
};

constexpr int kSecularMaxIter = 4;
constexpr float kSecularTol2 = 1e-8f;  // (relative step)^2 below which the NEXT iterate is exact to fp32

SKA_HD bool secular_step(const Sym4& M, const float cx, const float cy, const float cz, float lam, SecularState& s) {
  // one update for a given Rayleigh quotient; returns the converged flag
  const Ldl3 f = ldl3(M.m00 - lam, M.m01, M.m02, M.m11 - lam, M.m12, M.m22 - lam);
  float n0, n1, n2;
  ldl3_solve(f, fmaf(lam, cx, -M.m03), fmaf(lam, cy, -M.m13), fmaf(lam, cz, -M.m23), n0, n1, n2);
  const float d0 = n0 - s.y0, d1 = n1 - s.y1, d2 = n2 - s.y2;
  const float step2 = fmaf(d0, d0, fmaf(d1, d1, d2 * d2));
  const float X0 = n0 + cx, X1 = n1 + cy, X2 = n2 + cz;
  const float nrm2 = fmaf(X0, X0, fmaf(X1, X1, fmaf(X2, X2, 1.0f)));
  s.y0 = n0;
  s.y1 = n1;
  s.y2 = n2;
  // accept only in the fast-contraction regime (this step < 1/10 of the previous one): a small
  // step alone proves nothing when convergence is slow (near-degenerate geometry)
  const bool contracting = step2 <= 0.01f * s.step2;
  s.lam = lam;
  s.step2 = step2;
  s.ok = f.pos;
  return f.pos && contracting && (step2 <= kSecularTol2 * nrm2);
}

// ---------------------------------------------------------------------------------------------
// Cyclic Jacobi on a symmetric 4x4, register resident; returns the eigenvector of the smallest
// eigenvalue.  T = float (north-star design point) or double (exact mode / fallback).
template <typename T>
SKA_HD void jacobi4_smallest(T a[4][4], T vec[4], int sweeps) {
  T v[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) v[i][j] = (i == j) ? T(1) : T(0);
  for (int sw = 0; sw < sweeps; ++sw) {
#pragma unroll
    for (int p = 0; p < 3; ++p) {
#pragma unroll
      for (int q = p + 1; q < 4; ++q) {
        const T apq = a[p][q];
        if (apq != T(0)) {
          const T theta = (a[q][q] - a[p][p]) / (T(2) * apq);
          const T tt = (theta >= T(0) ? T(1) : T(-1)) / (fabs(theta) + sqrt(theta * theta + T(1)));
          const T c = T(1) / sqrt(tt * tt + T(1));
          const T s = tt * c;
          if (isfinite(c) && isfinite(s)) {
            a[p][p] -= tt * apq;
            a[q][q] += tt * apq;
            a[p][q] = a[q][p] = T(0);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              if (r != p && r != q) {
                const T arp = a[r][p], arq = a[r][q];
                a[r][p] = a[p][r] = c * arp - s * arq;
                a[r][q] = a[q][r] = s * arp + c * arq;
              }
              const T vrp = v[r][p], vrq = v[r][q];
              v[r][p] = c * vrp - s * vrq;
              v[r][q] = s * vrp + c * vrq;
            }
          }
        }
      }
    }
  }
  int best = 0;
  T lo = a[0][0];
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (a[i][i] < lo) {
      lo = a[i][i];
      best = i;
    }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    T x = v[i][0];
    if (best == 1) x = v[i][1];
    if (best == 2) x = v[i][2];
    if (best == 3) x = v[i][3];
    vec[i] = x;
  }
}

// ---------------------------------------------------------------------------------------------
// cv2.projectPoints distortion increment on normalised coordinates: returns (x'' - x, y'' - y).
// rad - 1 is formed as ((k1-k4) r2 + (k2-k5) r4 + (k3-k6) r6) / den so no bits are lost to 1 + ...
// PRISM adds the thin-prism terms s1..s4.
template <bool PRISM, typename T, typename CamT>
SKA_HD void distort_delta(const CamT& c, T x, T y, T& dx, T& dy) {
  // dx = x (rad-1) + 2 p1 x y + p2 (r2 + 2 x^2) = x s + p2 r2,  dy = y s + p1 r2  with  s = (rad-1) + 2 p1 y + 2 p2 x
  const T r2 = vfma(y, y, vmul(x, x));
  const T num = vmul(r2, vfma(r2, vfma(r2, c.dk[2], c.dk[1]), c.dk[0]));
  const T den = vfma(r2, vfma(r2, vfma(r2, c.kd[2], c.kd[1]), c.kd[0]), 1.0f);
  const T s = vfma(x, c.tp2, vfma(y, c.tp1, vmul(num, rcp_fast(den))));
  T tx = vmul(r2, c.p2);
  T ty = vmul(r2, c.p1);
  if (PRISM) {
    const T r4 = vmul(r2, r2);
    tx = vfma(r2, c.s[0], vfma(r4, c.s[1], tx));
    ty = vfma(r2, c.s[2], vfma(r4, c.s[3], ty));
  }
  dx = vfma(x, s, tx);
  dy = vfma(y, s, ty);
}

SKA_HD void distort64(const double* d /*12*/, double x, double y, double& xd, double& yd) {
  const double r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
  const double rad = (1.0 + d[0] * r2 + d[1] * r4 + d[4] * r6) / (1.0 + d[5] * r2 + d[6] * r4 + d[7] * r6);
  xd = x * rad + 2.0 * d[2] * x * y + d[3] * (r2 + 2.0 * x * x) + d[8] * r2 + d[9] * r4;
  yd = y * rad + d[2] * (r2 + 2.0 * y * y) + 2.0 * d[3] * x * y + d[10] * r2 + d[11] * r4;
}

}  // namespace ska
