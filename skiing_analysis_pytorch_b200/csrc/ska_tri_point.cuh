// ska_tri_point.cuh - PTS (frame, joint) points per thread: weighted V-view DLT + fused
// reprojection scoring.  __host__ __device__ so tests/hostemu can run the very same code path on
// the CPU-only box.
//
// Reference arithmetic replaced (file:line relative to the reference checkout):
//   rows / SVD / dehomogenise  vggt/triangulate.py:23-34, triangulation/triangulate.py:65-68
//   reprojection + pixel error triangulation/reproject.py:63-83, :243-244
#pragma once
#include "ska_math.cuh"

namespace ska {

#if defined(__CUDA_ARCH__)
#define SKA_WARP_ALL(p) __all_sync(0xffffffffu, (p))
#define SKA_WARP_ANY(p) __any_sync(0xffffffffu, (p))
#else
#define SKA_WARP_ALL(p) (p)
#define SKA_WARP_ANY(p) (p)
#endif

enum : uint32_t { kSolverSecular = 0, kSolverJacobi64 = 1, kSolverJacobi32 = 2 };

// Where one point's observations live in global memory; the rare fp64 path re-reads them from
// here instead of receiving register arrays (which would force local-memory spills on the hot path).
struct PointSource {
  const float* kpts;   // address of view 0's (u,v) of the thread's first point (next point: +2)
  const float* conf;   // address of view 0's confidence of the first point (next: +1), or nullptr
  int64_t k_sV, c_sV;  // view strides in floats
  uint32_t weight_sqrt;
};

struct Vec3d {
  double x, y, z;
};

// fp64 rows from the un-centred fp64 P, fp64 A^T A, fp64 cyclic Jacobi: exact-mode solver and the
// fallback of the fp32 fast path.  Mirrors the reference (fp64 SVD of the same A) to ~1e-12.
// Scalar arguments and a by-value result keep the (noinline, rare) call free of local memory.
template <int V>
SKA_HD_NOINLINE Vec3d solve_jacobi64(const double (*P64)[12], const float* kp, const float* cp, int64_t k_sV,
                                     int64_t c_sV, uint32_t weight_sqrt) {
  double a[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) a[i][j] = 0.0;
#pragma unroll 1
  for (int k = 0; k < V; ++k) {
    const double* P = P64[k];
    const double u = (double)kp[k * k_sV], v = (double)kp[k * k_sV + 1];
    double ww = 1.0;
    if (cp != nullptr) {
      const double c = (double)cp[k * c_sV];
      ww = weight_sqrt ? c : c * c;
    }
    double ra[4], rb[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      ra[m] = u * P[8 + m] - P[m];
      rb[m] = v * P[8 + m] - P[4 + m];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = i; j < 4; ++j) a[i][j] += ww * (ra[i] * ra[j] + rb[i] * rb[j]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < i; ++j) a[i][j] = a[j][i];
  double h[4];
  jacobi4_smallest<double>(a, h, 10);
  const double ih = 1.0 / h[3];
  Vec3d X;
  X.x = h[0] * ih;
  X.y = h[1] * ih;
  X.z = h[2] * ih;
  return X;
}

// fp32 Jacobi on the un-centred fp32 normal matrix (north-star design point, measurement only).
static SKA_HD_NOINLINE void solve_jacobi32(const Sym4& M, float cx, float cy, float cz, float Y[3]) {
  // un-centre: M = T^-T M' T^-1, T = [I c; 0 1]
  const float q0 = fmaf(M.m00, cx, fmaf(M.m01, cy, M.m02 * cz));
  const float q1 = fmaf(M.m01, cx, fmaf(M.m11, cy, M.m12 * cz));
  const float q2 = fmaf(M.m02, cx, fmaf(M.m12, cy, M.m22 * cz));
  float a[4][4];
  a[0][0] = M.m00; a[0][1] = a[1][0] = M.m01; a[0][2] = a[2][0] = M.m02;
  a[1][1] = M.m11; a[1][2] = a[2][1] = M.m12; a[2][2] = M.m22;
  a[0][3] = a[3][0] = M.m03 - q0;
  a[1][3] = a[3][1] = M.m13 - q1;
  a[2][3] = a[3][2] = M.m23 - q2;
  a[3][3] = M.m33 - 2.0f * (M.m03 * cx + M.m13 * cy + M.m23 * cz) + (q0 * cx + q1 * cy + q2 * cz);
  float h[4];
  jacobi4_smallest<float>(a, h, 6);
  const float ih = 1.0f / h[3];
  Y[0] = h[0] * ih - cx;
  Y[1] = h[1] * ih - cy;
  Y[2] = h[2] * ih - cz;
}

// Rayleigh quotient of the centred iterate, evaluated from the rows; also returns the row
// residuals (ra, rb) so the caller can reuse them.  T = float or F2 (two points in lockstep).
template <int V, bool CONF, typename T>
SKA_HD T rayleigh(const T (*a)[4], const T (*b)[4], const T* w2, T y0, T y1, T y2, float cx, float cy, float cz, T* ra,
                  T* rb, T& den) {
  T num = Vec<T>::splat(0.f);
#pragma unroll
  for (int k = 0; k < V; ++k) {
    ra[k] = vfma(a[k][0], y0, vfma(a[k][1], y1, vfma(a[k][2], y2, a[k][3])));
    rb[k] = vfma(b[k][0], y0, vfma(b[k][1], y1, vfma(b[k][2], y2, b[k][3])));
    const T rr = vfma(ra[k], ra[k], vmul(rb[k], rb[k]));
    num = CONF ? vfma(w2[k], rr, num) : vadd(num, rr);
  }
  const T X0 = vadd(y0, cx), X1 = vadd(y1, cy), X2 = vadd(y2, cz);
  den = vfma(X0, X0, vfma(X1, X1, vfma(X2, X2, 1.0f)));
  return vmul(num, rcp_fast(den));
}

constexpr float kFastTol2 = 1e-9f;   // one-factorisation path accepted if the correction is < 3.2e-5 relative
constexpr float kFastLamTr = 1e-3f;  // and lam * trace(M33^-1) < 1e-3  (=> lam < 1e-3 * lambda_min(M33))
constexpr float kCondMax = 3e4f;     // fp32 path only while trace(M33) trace(M33^-1) <= 3e4

SKA_HD float pick(float a, int) { return a; }
SKA_HD float pick(F2 a, int i) { return i == 0 ? a.x : a.y; }
SKA_HD bool pick(bool a, int) { return a; }
SKA_HD bool pick(B2 a, int i) { return i == 0 ? a.x : a.y; }

// DLT rows of a point (float) or of a packed pair (F2).  The pair's rows are formed with SCALAR FFMAs per point and
// packed afterwards: the two points' pixels arrive interleaved from memory (u0 v0 u1 v1), so a packed FFMA2 would
// first need the pair (u0, u1) assembled with two moves - which the register allocator re-does in front of every
// use (measured: ~40 moves per tile) - and its two camera coefficients loaded (LDC), whereas the scalar FFMA takes
// both coefficients straight from the constant bank and its result lands in the pair's half for free.
template <int LO>
SKA_HD void dlt_rows_pts(const CamDev& c, float u, float v, float a[4], float b[4]) {
  dlt_rows<LO>(c, u, v, a, b);
}
template <int LO>
SKA_HD void dlt_rows_pts(const CamDev& c, F2 u, F2 v, F2 a[4], F2 b[4]) {
  float a0[4], b0[4], a1[4], b1[4];
  dlt_rows<LO>(c, u.x, v.x, a0, b0);
  dlt_rows<LO>(c, u.y, v.y, a1, b1);
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    a[m] = mk2(a0[m], a1[m]);
    b[m] = mk2(b0[m], b1[m]);
  }
}

// Result of the one-factorisation fast stage for one point (T = float) or a pair (T = F2)
template <typename T>
struct FastStage {
  T y0, y1, y2;  // centred solution after the first-order secular step
  T lam, step2;
  typename Vec<T>::Mask conv, well, ok;
};

// fp32 error allowance of the Rayleigh numerator formed from the normal matrix (below): its terms are bounded by
// m33 and carry the rounding of ~2V accumulations + 3 FMAs each
template <int V>
struct LamEps {
  static constexpr float value = (2.0f * V + 8.0f) * 6e-8f;
};

// rows -> normal matrix -> LDL^T at lam = 0 -> inhomogeneous least-squares point Y0 -> first-order secular step
// with the SAME factorisation -> certificate.  The rows a, b and M are kept for the caller (residuals at the
// final point; the rare general path).
//
// The Rayleigh quotient at Y0 comes from the normal matrix: [Y0;1]^T M [Y0;1] = m33 + m3.Y0 (since M33 Y0 = -m3).
// In CENTRED coordinates m33 = sum (row . [0;1])^2 is only (|Y|/sigma)^2 ~ 1e4 times the quotient, so the fp32
// cancellation costs ~1e-3 of lam - irrelevant for a correction that is itself < 3.2e-5 of |X| - and the
// certificate uses the upper bound lam_c = (num + eps m33) / den, which keeps it rigorous.  (Evaluating the
// quotient from the row residuals, as the general path does, costs 9 more FMAs per view.)
// fast_stage_core: from the rows on.  CT = float (one centre for every point) or T (a centre per point: per-frame cameras).
template <int V, bool CONF, typename T, typename CT>
SKA_HD void fast_stage_core(const CT cx, const CT cy, const CT cz, const T* w2, const T (*a)[4], const T (*b)[4], Sym4T<T>& M,
                            FastStage<T>& o) {
  sym4_zero(M);
#pragma unroll
  for (int k = 0; k < V; ++k) {
    if (CONF) {
      sym4_rank1(M, a[k], w2[k]);
      sym4_rank1(M, b[k], w2[k]);
    } else {
      sym4_rank1_unit(M, a[k]);
      sym4_rank1_unit(M, b[k]);
    }
  }
  // lam = 0: the inhomogeneous least-squares point
  const Ldl3T<T> f0 = ldl3(M.m00, M.m01, M.m02, M.m11, M.m12, M.m22);
  T y0, y1, y2;
  ldl3_solve(f0, vneg(M.m03), vneg(M.m13), vneg(M.m23), y0, y1, y2);
  const T num = vfma(M.m03, y0, vfma(M.m13, y1, vfma(M.m23, y2, M.m33)));
  const T X0 = vadd(y0, cx), X1 = vadd(y1, cy), X2 = vadd(y2, cz);
  const T den = vfma(X0, X0, vfma(X1, X1, vfma(X2, X2, 1.0f)));
  const T rden = rcp_fast(den);
  const T lam = vmul(num, rden);
  const T lam_c = vmul(vfma(M.m33, LamEps<V>::value, num), rden);
  // first-order step of the secular equation with the SAME factorisation:
  //   Y(lam) - Y(0) = lam (M33 - lam I)^-1 (c + Y0) = lam z + O(lam^2),  z = M33^-1 (c + Y0)
  T z0, z1, z2;
  ldl3_solve(f0, X0, X1, X2, z0, z1, z2);
  const T d0 = vmul(lam, z0), d1 = vmul(lam, z1), d2 = vmul(lam, z2);
  const T step2 = vfma(d0, d0, vfma(d1, d1, vmul(d2, d2)));
  // certificate: lam * trace(M33^-1) < 1e-3 => lam < 1e-3 lambda_min(M33): M33 - lam I is positive
  // definite (Cauchy interlacing: this is the smallest eigenpair) and the dropped second-order
  // term is < 1e-3 of a step that is itself < 3.2e-5 relative.
  const T itr = ldl3_inv_trace(f0);
  o.conv = mand(mand(f0.pos, vlt(vmul(lam_c, itr), kFastLamTr)), vle(step2, vmul(den, kFastTol2)));
  // conditioning gate: trace(M33) trace(M33^-1) bounds cond(M33); beyond kCondMax (rays nearly
  // parallel, point near infinity) fp32 cannot hold the north-star tolerance -> fp64 path
  o.well = vle(vmul(vadd(vadd(M.m00, M.m11), M.m22), itr), kCondMax);
  o.y0 = vadd(y0, d0);
  o.y1 = vadd(y1, d1);
  o.y2 = vadd(y2, d2);
  o.lam = lam;
  o.step2 = step2;
  o.ok = f0.pos;
}

template <int V, bool CONF, int LO, typename T>
SKA_HD void fast_stage(const CamDev* __restrict__ cam, const float cx, const float cy, const float cz, const T* u, const T* v,
                       const T* w2, T (*a)[4], T (*b)[4], Sym4T<T>& M, FastStage<T>& o) {
#pragma unroll
  for (int k = 0; k < V; ++k) dlt_rows_pts<LO>(cam[k], u[k], v[k], a[k], b[k]);
  fast_stage_core<V, CONF, T, float>(cx, cy, cz, w2, a, b, M, o);
}

// row residuals (a . [Y;1], b . [Y;1]) of every view at the point Y
template <int V, typename T>
SKA_HD void row_residuals(const T (*a)[4], const T (*b)[4], T Y0, T Y1, T Y2, T* ra, T* rb) {
#pragma unroll
  for (int k = 0; k < V; ++k) {
    ra[k] = vfma(a[k][0], Y0, vfma(a[k][1], Y1, vfma(a[k][2], Y2, a[k][3])));
    rb[k] = vfma(b[k][0], Y0, vfma(b[k][1], Y1, vfma(b[k][2], Y2, b[k][3])));
  }
}

// Fused reprojection scoring, differential form:
//   proj - obs = -(row . [Y;1]) / z  (+ fx*dx_distortion - skew*y)
// so the ~1e3 px magnitudes of proj and obs never meet in fp32.
SKA_HD float pix_norm(float u, float s, float o) { return fmaf(u, s, o); }
SKA_HD F2 pix_norm(F2 u, float s, float o) { return mk2(fmaf(u.x, s, o), fmaf(u.y, s, o)); }  // scalar: see dlt_rows_pts

// u, v: the observed pixel (DIST == 1 derives the normalised coordinates from it).
// c: the view's projection rows; ck: the camera whose intrinsics / distortion apply (the view's own, or view 0's
// when the caller knows that all views share them - the constant loads then fold across views).
template <int DIST, typename T>
SKA_HD void score_core(const CamDev& c, const CamDev& ck, T z, T Y0, T Y1, T Y2, T ra, T rb, T u, T v, T& eu, T& ev) {
  const T iz = rcp_fast(z);
  eu = vmul(vneg(ra), iz);
  ev = vmul(vneg(rb), iz);
  if (DIST) {
    T x, y;
    if (DIST == 1) {
      // zero skew: the pinhole reprojection is u + eu, so x = (u + eu - cx) / fx - two FMAs per coordinate
      // instead of a second 3x4 row product (the cancellation u - cx costs 6e-8 in x: nothing after the
      // distortion polynomial's ~0.1 sensitivity)
      x = vfma(eu, ck.ifx, pix_norm(u, ck.ifx, ck.ncx));
      y = vfma(ev, ck.ify, pix_norm(v, ck.ify, ck.ncy));
    } else {
      x = vmul(vfma(c.Rxy[0], Y0, vfma(c.Rxy[1], Y1, vfma(c.Rxy[2], Y2, c.txy[0]))), iz);
      y = vmul(vfma(c.Rxy[3], Y0, vfma(c.Rxy[4], Y1, vfma(c.Rxy[5], Y2, c.txy[1]))), iz);
    }
    T dx, dy;
    distort_delta<(DIST >= 2)>(ck, x, y, dx, dy);
    eu = vfma(dx, ck.fx, eu);
    ev = vfma(dy, ck.fy, ev);
    if (DIST >= 2) eu = vfma(y, -ck.skew, eu);
  }
}
template <int DIST, typename T>
SKA_HD void score_view(const CamDev& c, const CamDev& ck, T Y0, T Y1, T Y2, T ra, T rb, T u, T v, T& eu, T& ev) {
  // depth: the hi part alone is good to 6e-8 relative, which is all a divisor of a ~1 px quantity needs
  const T z = vfma(c.Ph[8], Y0, vfma(c.Ph[9], Y1, vfma(c.Ph[10], Y2, c.Ph[11])));
  score_core<DIST, T>(c, ck, z, Y0, Y1, Y2, ra, rb, u, v, eu, ev);
}

template <int V, int DIST, bool SAMEK, typename T>
SKA_HD void score_views(const CamDev* __restrict__ cam, T Y0, T Y1, T Y2, const T* ra, const T* rb, const T* u, const T* v,
                        T* du, T* dv) {
#pragma unroll
  for (int k = 0; k < V; ++k) score_view<DIST, T>(cam[k], cam[SAMEK ? 0 : k], Y0, Y1, Y2, ra[k], rb[k], u[k], v[k], du[k], dv[k]);
}

template <bool PACK>
struct PointVec {
  using type = float;
};
template <>
struct PointVec<true> {
  using type = F2;
};

// PTS points per thread.  PTS == 2 runs the two points as ONE packed (F2) computation - every
// FFMA of the hot path becomes an FFMA2; any other PTS runs them one by one in scalar fp32.
// u,v,w2: [PTS][V] pixel coordinates and squared row weights (w2 unused if !CONF).
// LO: see dlt_rows.  DIST: 0 pinhole scoring, 1 rational+tangential, 2 + thin prism + skew.
// SAMEK: every view shares view 0's intrinsics / distortion (checked by the caller).
// Outputs: X (un-centred), du/dv = reprojected minus observed pixel per view, status.
template <int V, int PTS, bool CONF, int DIST, uint32_t SOLVER, int LO = 1, bool SAMEK = false>
SKA_HD void tri_points(const CamDev* __restrict__ cam, const double (*P64)[12], const float cx, const float cy,
                       const float cz, const float (*u)[V], const float (*v)[V], const float (*w2)[V],
                       const PointSource& src, float (*X)[3], float (*du)[V], float (*dv)[V], uint8_t* status) {
  constexpr bool PACK = (PTS == 2);
  using T = typename PointVec<PACK>::type;
  constexpr int NG = PACK ? 1 : PTS;   // lockstep groups
  // ---- gather the inputs of each group
  T ut[NG][V], vt[NG][V], wt[NG][V];
#pragma unroll
  for (int g = 0; g < NG; ++g)
#pragma unroll
    for (int k = 0; k < V; ++k) {
      if constexpr (PACK) {
        ut[g][k] = mk2(u[0][k], u[1][k]);
        vt[g][k] = mk2(v[0][k], v[1][k]);
        wt[g][k] = CONF ? mk2(w2[0][k], w2[1][k]) : mk2(1.f, 1.f);
      } else {
        ut[g][k] = u[g][k];
        vt[g][k] = v[g][k];
        wt[g][k] = CONF ? w2[g][k] : 1.f;
      }
    }
  T a[NG][V][4], b[NG][V][4];
  Sym4T<T> M[NG];
  float Y[PTS][3];
  bool need64[PTS];
#pragma unroll
  for (int p = 0; p < PTS; ++p) {
    need64[p] = (SOLVER == kSolverJacobi64);
    status[p] = 0;
  }
  if (SOLVER == kSolverSecular) {
    FastStage<T> fs[NG];
    bool all_fast = true;
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      fast_stage<V, CONF, LO, T>(cam, cx, cy, cz, ut[g], vt[g], wt[g], a[g], b[g], M[g], fs[g]);
      all_fast = all_fast && mall(fs[g].conv);
    }
    SecularState s[PTS];
    bool conv[PTS], well[PTS];
#pragma unroll
    for (int p = 0; p < PTS; ++p) {
      const int g = PACK ? 0 : p, i = PACK ? p : 0;
      s[p].y0 = pick(fs[g].y0, i);
      s[p].y1 = pick(fs[g].y1, i);
      s[p].y2 = pick(fs[g].y2, i);
      s[p].lam = pick(fs[g].lam, i);
      s[p].step2 = pick(fs[g].step2, i);
      s[p].ok = pick(fs[g].ok, i);
      conv[p] = pick(fs[g].conv, i);
      well[p] = pick(fs[g].well, i);
    }
    if (!SKA_WARP_ALL(all_fast)) {
      // general path (rare): full secular iteration per point in scalar fp32 (refactorise at every
      // lam, Rayleigh quotient from the row residuals), quadratically convergent
#pragma unroll 1
      for (int it = 0; it < kSecularMaxIter; ++it) {
        bool done = true;
#pragma unroll
        for (int p = 0; p < PTS; ++p) {
          const int g = PACK ? 0 : p, i = PACK ? p : 0;
          if (!conv[p]) {
            float as[V][4], bs[V][4], ws[V], r1[V], r2[V], den;
#pragma unroll
            for (int k = 0; k < V; ++k) {
#pragma unroll
              for (int m = 0; m < 4; ++m) {
                as[k][m] = pick(a[g][k][m], i);
                bs[k][m] = pick(b[g][k][m], i);
              }
              ws[k] = pick(wt[g][k], i);
            }
            const float lam = rayleigh<V, CONF, float>(as, bs, ws, s[p].y0, s[p].y1, s[p].y2, cx, cy, cz, r1, r2, den);
            const Sym4 Ms = PACK ? (i == 0 ? sym4_lane<0>(M[g]) : sym4_lane<1>(M[g])) : sym4_lane<0>(M[g]);
            conv[p] = secular_step(Ms, cx, cy, cz, lam, s[p]);
          }
          // a lane that lost positive-definiteness can never certify: do not wait for it
          done = done && (conv[p] || !s[p].ok);
        }
        if (SKA_WARP_ALL(done)) break;
      }
    }
#pragma unroll
    for (int p = 0; p < PTS; ++p) {
      const int g = PACK ? 0 : p, i = PACK ? p : 0;
      Y[p][0] = s[p].y0;
      Y[p][1] = s[p].y1;
      Y[p][2] = s[p].y2;
      const bool finite_in = fabsf(pick(M[g].m33, i)) <= 3.0e38f;  // false for NaN / inf inputs
      need64[p] = !(conv[p] && s[p].ok && well[p]) && finite_in;
      if (!finite_in) status[p] = 2;
    }
  } else {
    // measurement / exact solvers: rows and the normal matrix only
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      sym4_zero(M[g]);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        dlt_rows_pts<LO>(cam[k], ut[g][k], vt[g][k], a[g][k], b[g][k]);
        sym4_rank1(M[g], a[g][k], wt[g][k]);
        sym4_rank1(M[g], b[g][k], wt[g][k]);
      }
    }
    if (SOLVER == kSolverJacobi32) {
#pragma unroll
      for (int p = 0; p < PTS; ++p) {
        const int g = PACK ? 0 : p, i = PACK ? p : 0;
        const Sym4 Ms = PACK ? (i == 0 ? sym4_lane<0>(M[g]) : sym4_lane<1>(M[g])) : sym4_lane<0>(M[g]);
        solve_jacobi32(Ms, cx, cy, cz, Y[p]);
      }
    }
  }

  if (SOLVER != kSolverJacobi32) {
    bool any64 = false;
#pragma unroll
    for (int p = 0; p < PTS; ++p) any64 = any64 || need64[p];
    if (SKA_WARP_ANY(any64)) {
#pragma unroll  // static indexing only: a rolled loop would push Y/need64/status into local memory
      for (int p = 0; p < PTS; ++p) {
        if (need64[p]) {
          const Vec3d Xd = solve_jacobi64<V>(P64, src.kpts + 2 * p, src.conf ? src.conf + p : nullptr, src.k_sV,
                                             src.c_sV, src.weight_sqrt);
          Y[p][0] = (float)(Xd.x - (double)cx);
          Y[p][1] = (float)(Xd.y - (double)cy);
          Y[p][2] = (float)(Xd.z - (double)cz);
          if (SOLVER == kSolverSecular) status[p] = 1;
        }
      }
    }
  }

  // ---- un-centre, flag non-finite results, residuals at the final point, score every view
#pragma unroll
  for (int p = 0; p < PTS; ++p) {
    X[p][0] = Y[p][0] + cx;
    X[p][1] = Y[p][1] + cy;
    X[p][2] = Y[p][2] + cz;
    if (!(fabsf(X[p][0]) <= 3.0e38f && fabsf(X[p][1]) <= 3.0e38f && fabsf(X[p][2]) <= 3.0e38f)) status[p] = 2;
  }
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    T Yt[3];
    if constexpr (PACK) {
#pragma unroll
      for (int m = 0; m < 3; ++m) Yt[m] = mk2(Y[0][m], Y[1][m]);
    } else {
#pragma unroll
      for (int m = 0; m < 3; ++m) Yt[m] = Y[g][m];
    }
    T ra[V], rb[V], dut[V], dvt[V];
    row_residuals<V, T>(a[g], b[g], Yt[0], Yt[1], Yt[2], ra, rb);
    score_views<V, DIST, SAMEK, T>(cam, Yt[0], Yt[1], Yt[2], ra, rb, ut[g], vt[g], dut, dvt);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      if constexpr (PACK) {
        du[0][k] = dut[k].x;
        du[1][k] = dut[k].y;
        dv[0][k] = dvt[k].x;
        dv[1][k] = dvt[k].y;
      } else {
        du[g][k] = dut[k];
        dv[g][k] = dvt[k];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// View-pair form for many views (even V >= 4): ONE point per call, the per-view work - DLT rows, normal-matrix
// accumulation, final residuals, scoring - packed over PAIRS OF VIEWS (F2 = views 2i, 2i+1)
// with interleaved camera constants (CamPairDev); the 3x3 solve, the secular step and the certificate are scalar.
// Same arithmetic per view as tri_points<V, 1, ...> (each packed half is the scalar IEEE operation); sums over the
// views are formed as (even views) + (odd views), so results agree with the scalar form to rounding, not bit for bit.
// The rare general path and the fp64 fallback reuse the scalar per-view code on `cam`.
// ROWS: where the DLT rows live between the normal-matrix pass and the final-residual pass
//   kRowsRegs   in registers (16 per view pair: 206 registers at 8 views, 8 warps per SM)
//   kRowsRecomp formed a second time (14 more packed operations + 15 constant loads per view pair)
//   kRowsSmem   parked in the caller's shared-memory slab: one 8-byte chunk per (view pair, row entry) and thread,
//               rowbuf[(8 i + m) * row_stride + thread] - conflict-free 64-bit accesses, no register shuffling
// SAMEK: every view shares view 0's intrinsics / distortion: those constants are loaded once, not per view pair.
enum : int { kRowsRegs = 0, kRowsRecomp = 1, kRowsSmem = 2 };
#if defined(__CUDA_ARCH__)
#define SKA_OPAQUE(x) asm volatile("" : "+f"(x))
#else
#define SKA_OPAQUE(x) (void)(x)
#endif
// read a chunk back from the slab: volatile on the device, otherwise the compiler forwards the stored registers to
// the loads and the rows stay live in registers after all
SKA_HD F2 slab_load(const F2* p) {
#if defined(__CUDA_ARCH__)
  F2 r;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "r"((uint32_t)__cvta_generic_to_shared(p)));
  return r;
#else
  return *p;
#endif
}
// DLT rows of one view pair: SCALAR per view (both coefficients of every FFMA come straight from the constant bank
// and the result lands in its half of the pair for free; a packed row needs the pair (u_a, u_b) assembled from two
// separate loads - the register allocator re-copies it in front of every use - plus 15 constant loads)
template <int LO>
SKA_HD void dlt_rows_pair(const CamDev& ca, const CamDev& cb, float ua, float va, float ub, float vb, F2 a[4], F2 b[4]) {
  float a0[4], b0[4], a1[4], b1[4];
  dlt_rows<LO>(ca, ua, va, a0, b0);
  dlt_rows<LO>(cb, ub, vb, a1, b1);
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    a[m] = mk2(a0[m], a1[m]);
    b[m] = mk2(b0[m], b1[m]);
  }
}

// rows of the view pairs kept in registers between the two halves (all pairs, or one transient pair)
template <int V, int ROWS>
struct VpRows {
  static constexpr int H = V / 2;
  static constexpr int HK = (ROWS == kRowsRegs) ? H : 1;
  F2 a[HK][4], b[HK][4];
};

// First half: rows -> normal matrix -> LDL^T -> least-squares point -> first-order secular step -> certificate.
// u2, v2, w22: the observations / squared weights packed by view pairs.
template <int V, bool CONF, int ROWS>
SKA_HD void vp_fast(const CamDev* __restrict__ cam, const float cx, const float cy, const float cz, const float* u, const float* v,
                    const F2* w22, VpRows<V, ROWS>& rows, F2* rowbuf, int row_stride, Sym4& M, SecularState& s, bool& conv,
                    bool& well) {
  constexpr int H = V / 2;
  Sym4T<F2> M2;
  sym4_zero(M2);
#pragma unroll
  for (int i = 0; i < H; ++i) {
    F2(&ai)[4] = rows.a[ROWS == kRowsRegs ? i : 0];
    F2(&bi)[4] = rows.b[ROWS == kRowsRegs ? i : 0];
    dlt_rows_pair<1>(cam[2 * i], cam[2 * i + 1], u[2 * i], v[2 * i], u[2 * i + 1], v[2 * i + 1], ai, bi);
    if (ROWS == kRowsSmem) {
      F2* r = rowbuf + (8 * i) * row_stride;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        r[m * row_stride] = ai[m];
        r[(4 + m) * row_stride] = bi[m];
      }
    }
    if (CONF) {
      sym4_rank1(M2, ai, w22[i]);
      sym4_rank1(M2, bi, w22[i]);
    } else {
      sym4_rank1_unit(M2, ai);
      sym4_rank1_unit(M2, bi);
    }
  }
  M.m00 = M2.m00.x + M2.m00.y; M.m01 = M2.m01.x + M2.m01.y; M.m02 = M2.m02.x + M2.m02.y; M.m03 = M2.m03.x + M2.m03.y;
  M.m11 = M2.m11.x + M2.m11.y; M.m12 = M2.m12.x + M2.m12.y; M.m13 = M2.m13.x + M2.m13.y;
  M.m22 = M2.m22.x + M2.m22.y; M.m23 = M2.m23.x + M2.m23.y; M.m33 = M2.m33.x + M2.m33.y;
  // lam = 0: the inhomogeneous least-squares point; Rayleigh numerator from the normal matrix (see fast_stage)
  const Ldl3 f0 = ldl3(M.m00, M.m01, M.m02, M.m11, M.m12, M.m22);
  float y0, y1, y2;
  ldl3_solve(f0, -M.m03, -M.m13, -M.m23, y0, y1, y2);
  const float num = fmaf(M.m03, y0, fmaf(M.m13, y1, fmaf(M.m23, y2, M.m33)));
  const float X0 = y0 + cx, X1 = y1 + cy, X2 = y2 + cz;
  const float den = fmaf(X0, X0, fmaf(X1, X1, fmaf(X2, X2, 1.0f)));
  const float rden = rcp_fast(den);
  const float lam = num * rden;
  const float lam_c = fmaf(M.m33, LamEps<V>::value, num) * rden;
  float z0, z1, z2;
  ldl3_solve(f0, X0, X1, X2, z0, z1, z2);
  const float d0 = lam * z0, d1 = lam * z1, d2 = lam * z2;
  const float step2 = fmaf(d0, d0, fmaf(d1, d1, d2 * d2));
  const float itr = ldl3_inv_trace(f0);
  conv = f0.pos && (lam_c * itr < kFastLamTr) && (step2 <= kFastTol2 * den);
  well = (M.m00 + M.m11 + M.m22) * itr <= kCondMax;
  s.y0 = y0 + d0;
  s.y1 = y1 + d1;
  s.y2 = y2 + d2;
  s.lam = lam;
  s.step2 = step2;
  s.ok = f0.pos;
}

// Second half: residuals of every view pair at the final point Y (rows from registers / recomputed / from the slab)
// and the fused reprojection scoring; eu, ev: reprojected minus observed pixel, packed by view pairs.
template <int V, int DIST, int ROWS, bool SAMEK>
SKA_HD void vp_score(const CamPairDev* __restrict__ camp, const CamDev* __restrict__ cam, const float* Y, float* u, float* v,
                     VpRows<V, ROWS>& rows, const F2* rowbuf, int row_stride, F2* eu2, F2* ev2) {
  constexpr int H = V / 2;
#pragma unroll
  for (int i = 0; i < H; ++i) {
    const CamPairDev& c = camp[i];
    const CamPairDev& ck = camp[SAMEK ? 0 : i];  // intrinsics / distortion
    const CamDev& cka = cam[SAMEK ? 0 : 2 * i];
    const CamDev& ckb = cam[SAMEK ? 0 : 2 * i + 1];
    F2(&ai)[4] = rows.a[ROWS == kRowsRegs ? i : 0];
    F2(&bi)[4] = rows.b[ROWS == kRowsRegs ? i : 0];
    if (ROWS == kRowsRecomp) {
      SKA_OPAQUE(u[2 * i]);  // a new value as far as the compiler knows: the rows are recomputed, not kept
      SKA_OPAQUE(v[2 * i]);
      SKA_OPAQUE(u[2 * i + 1]);
      SKA_OPAQUE(v[2 * i + 1]);
      dlt_rows_pair<1>(cam[2 * i], cam[2 * i + 1], u[2 * i], v[2 * i], u[2 * i + 1], v[2 * i + 1], ai, bi);
    } else if (ROWS == kRowsSmem) {
      const F2* r = rowbuf + (8 * i) * row_stride;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        ai[m] = slab_load(r + m * row_stride);
        bi[m] = slab_load(r + (4 + m) * row_stride);
      }
    }
    const F2 ra = vfma(ai[0], Y[0], vfma(ai[1], Y[1], vfma(ai[2], Y[2], ai[3])));
    const F2 rb = vfma(bi[0], Y[0], vfma(bi[1], Y[1], vfma(bi[2], Y[2], bi[3])));
    const F2 z = vfma(c.Ph[8], Y[0], vfma(c.Ph[9], Y[1], vfma(c.Ph[10], Y[2], c.Ph[11])));
    const F2 iz = rcp_fast(z);
    F2 eu = vmul(vneg(ra), iz), ev = vmul(vneg(rb), iz);
    if (DIST) {
      // normalised coordinate of the observed pixel: scalar per view (see dlt_rows_pair)
      const F2 xn = mk2(fmaf(u[2 * i], cka.ifx, cka.ncx), fmaf(u[2 * i + 1], ckb.ifx, ckb.ncx));
      const F2 yn = mk2(fmaf(v[2 * i], cka.ify, cka.ncy), fmaf(v[2 * i + 1], ckb.ify, ckb.ncy));
      const F2 x = vfma(eu, ck.ifx, xn);
      const F2 y = vfma(ev, ck.ify, yn);
      F2 dx, dy;
      distort_delta<false>(ck, x, y, dx, dy);
      eu = vfma(dx, ck.fx, eu);
      ev = vfma(dy, ck.fy, ev);
    }
    eu2[i] = eu;
    ev2[i] = ev;
  }
}

template <int V, bool CONF, int DIST, int ROWS = kRowsRegs, bool SAMEK = false, int LO = 1>
SKA_HD void tri_point_vp(const CamPairDev* __restrict__ camp, const CamDev* __restrict__ cam, const double (*P64)[12],
                         const float cx, const float cy, const float cz, const float* u, const float* v, const float* w2,
                         const PointSource& src, float* X, float* du, float* dv, uint8_t& status, F2* rowbuf = nullptr,
                         int row_stride = 0) {
  static_assert(V % 2 == 0 && DIST <= 1 && LO == 1, "view-pair form: even V, no skew / thin prism");
  constexpr int H = V / 2;
  F2 w22[H];
  float uu[V], vv[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    uu[k] = u[k];
    vv[k] = v[k];
  }
#pragma unroll
  for (int i = 0; i < H; ++i) w22[i] = CONF ? mk2(w2[2 * i], w2[2 * i + 1]) : mk2(1.f, 1.f);
  VpRows<V, ROWS> rows;
  Sym4 M;
  SecularState s;
  bool conv, well;
  vp_fast<V, CONF, ROWS>(cam, cx, cy, cz, uu, vv, w22, rows, rowbuf, row_stride, M, s, conv, well);
  status = 0;
  if (!SKA_WARP_ALL(conv)) {
    // general path (rare): full secular iteration in scalar fp32 on the scalar cameras
#pragma unroll 1
    for (int it = 0; it < kSecularMaxIter; ++it) {
      if (!conv) {
        float nm = 0.f;
#pragma unroll 1
        for (int k = 0; k < V; ++k) {
          float as[4], bs[4];
          dlt_rows<LO>(cam[k], u[k], v[k], as, bs);
          const float r1 = fmaf(as[0], s.y0, fmaf(as[1], s.y1, fmaf(as[2], s.y2, as[3])));
          const float r2 = fmaf(bs[0], s.y0, fmaf(bs[1], s.y1, fmaf(bs[2], s.y2, bs[3])));
          const float rr = fmaf(r1, r1, r2 * r2);
          nm = CONF ? fmaf(w2[k], rr, nm) : nm + rr;
        }
        const float Xa = s.y0 + cx, Xb = s.y1 + cy, Xc = s.y2 + cz;
        const float lm = nm * rcp_fast(fmaf(Xa, Xa, fmaf(Xb, Xb, fmaf(Xc, Xc, 1.0f))));
        conv = secular_step(M, cx, cy, cz, lm, s);
      }
      if (SKA_WARP_ALL(conv || !s.ok)) break;
    }
  }
  float Y[3] = {s.y0, s.y1, s.y2};
  const bool finite_in = fabsf(M.m33) <= 3.0e38f;  // false for NaN / inf inputs
  const bool need64 = !(conv && s.ok && well) && finite_in;
  if (!finite_in) status = 2;
  if (SKA_WARP_ANY(need64)) {
    if (need64) {
      const Vec3d Xd = solve_jacobi64<V>(P64, src.kpts, src.conf, src.k_sV, src.c_sV, src.weight_sqrt);
      Y[0] = (float)(Xd.x - (double)cx);
      Y[1] = (float)(Xd.y - (double)cy);
      Y[2] = (float)(Xd.z - (double)cz);
      status = 1;
    }
  }
  X[0] = Y[0] + cx;
  X[1] = Y[1] + cy;
  X[2] = Y[2] + cz;
  if (!(fabsf(X[0]) <= 3.0e38f && fabsf(X[1]) <= 3.0e38f && fabsf(X[2]) <= 3.0e38f)) status = 2;
  F2 eu2[H], ev2[H];
  vp_score<V, DIST, ROWS, SAMEK>(camp, cam, Y, uu, vv, rows, rowbuf, row_stride, eu2, ev2);
#pragma unroll
  for (int i = 0; i < H; ++i) {
    du[2 * i] = eu2[i].x;
    du[2 * i + 1] = eu2[i].y;
    dv[2 * i] = ev2[i].x;
    dv[2 * i + 1] = ev2[i].y;
  }
}

}  // namespace ska
