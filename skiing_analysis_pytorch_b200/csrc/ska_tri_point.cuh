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
// residuals (ra, rb) so the caller can reuse them.
template <int V, bool CONF>
SKA_HD float rayleigh(const float (*a)[4], const float (*b)[4], const float* w2, float y0, float y1, float y2, float cx,
                      float cy, float cz, float* ra, float* rb, float& den) {
  float num = 0.f;
#pragma unroll
  for (int k = 0; k < V; ++k) {
    ra[k] = fmaf(a[k][0], y0, fmaf(a[k][1], y1, fmaf(a[k][2], y2, a[k][3])));
    rb[k] = fmaf(b[k][0], y0, fmaf(b[k][1], y1, fmaf(b[k][2], y2, b[k][3])));
    const float rr = fmaf(ra[k], ra[k], rb[k] * rb[k]);
    num = CONF ? fmaf(w2[k], rr, num) : (num + rr);
  }
  const float X0 = y0 + cx, X1 = y1 + cy, X2 = y2 + cz;
  den = fmaf(X0, X0, fmaf(X1, X1, fmaf(X2, X2, 1.0f)));
  return num * rcp_fast(den);
}

constexpr float kFastTol2 = 1e-9f;   // one-factorisation path accepted if the correction is < 3.2e-5 relative
constexpr float kFastLamTr = 1e-3f;  // and lam * trace(M33^-1) < 1e-3  (=> lam < 1e-3 * lambda_min(M33))
constexpr float kCondMax = 3e4f;     // fp32 path only while trace(M33) trace(M33^-1) <= 3e4

// PTS points in lockstep (independent dependency chains interleave -> ILP).
// u,v,w2: [PTS][V] pixel coordinates and squared row weights (w2 unused if !CONF).
// LO: see dlt_rows.  DIST: 0 pinhole scoring, 1 rational+tangential, 2 + thin prism + skew.
// Outputs: X (un-centred), du/dv = reprojected minus observed pixel per view, status.
template <int V, int PTS, bool CONF, int DIST, uint32_t SOLVER, int LO = 1>
SKA_HD void tri_points(const CamDev* __restrict__ cam, const double (*P64)[12], const float cx, const float cy,
                       const float cz, const float (*u)[V], const float (*v)[V], const float (*w2)[V],
                       const PointSource& src, float (*X)[3], float (*du)[V], float (*dv)[V], uint8_t* status) {
  // ---- rows (kept in registers) and the normal matrix in centred coordinates
  float a[PTS][V][4], b[PTS][V][4];
  Sym4 M[PTS];
#pragma unroll
  for (int p = 0; p < PTS; ++p) sym4_zero(M[p]);
#pragma unroll
  for (int k = 0; k < V; ++k) {
#pragma unroll
    for (int p = 0; p < PTS; ++p) {
      dlt_rows<LO>(cam[k], u[p][k], v[p][k], a[p][k], b[p][k]);
      if (CONF) {
        sym4_rank1(M[p], a[p][k], w2[p][k]);
        sym4_rank1(M[p], b[p][k], w2[p][k]);
      } else {
        sym4_rank1_unit(M[p], a[p][k]);
        sym4_rank1_unit(M[p], b[p][k]);
      }
    }
  }

  float Y[PTS][3], ra[PTS][V], rb[PTS][V];
  bool need64[PTS], have_res[PTS];
#pragma unroll
  for (int p = 0; p < PTS; ++p) {
    need64[p] = (SOLVER == kSolverJacobi64);
    have_res[p] = false;
    status[p] = 0;
  }
  if (SOLVER == kSolverSecular) {
    SecularState s[PTS];
    bool conv[PTS], well[PTS];
    bool all_fast = true;
#pragma unroll
    for (int p = 0; p < PTS; ++p) {
      // lam = 0: the inhomogeneous least-squares point
      const Ldl3 f0 = ldl3(M[p].m00, M[p].m01, M[p].m02, M[p].m11, M[p].m12, M[p].m22);
      float y0, y1, y2, den;
      ldl3_solve(f0, -M[p].m03, -M[p].m13, -M[p].m23, y0, y1, y2);
      const float lam = rayleigh<V, CONF>(a[p], b[p], w2[p], y0, y1, y2, cx, cy, cz, ra[p], rb[p], den);
      // first-order step of the secular equation with the SAME factorisation:
      //   Y(lam) - Y(0) = lam (M33 - lam I)^-1 (c + Y0) = lam z + O(lam^2),  z = M33^-1 (c + Y0)
      float z0, z1, z2;
      ldl3_solve(f0, y0 + cx, y1 + cy, y2 + cz, z0, z1, z2);
      const float d0 = lam * z0, d1 = lam * z1, d2 = lam * z2;
      const float step2 = fmaf(d0, d0, fmaf(d1, d1, d2 * d2));
      // certificate: lam * trace(M33^-1) < 1e-3 => lam < 1e-3 lambda_min(M33): M33 - lam I is positive
      // definite (Cauchy interlacing: this is the smallest eigenpair) and the dropped second-order
      // term is < 1e-3 of a step that is itself < 3.2e-5 relative.
      const float itr = ldl3_inv_trace(f0);
      const bool fast = f0.pos && (lam * itr < kFastLamTr) && (step2 <= kFastTol2 * den);
      // conditioning gate: trace(M33) trace(M33^-1) bounds cond(M33); beyond kCondMax (rays nearly
      // parallel, point near infinity) fp32 cannot hold the north-star tolerance -> fp64 path
      well[p] = (M[p].m00 + M[p].m11 + M[p].m22) * itr <= kCondMax;
      s[p].y0 = y0 + d0;
      s[p].y1 = y1 + d1;
      s[p].y2 = y2 + d2;
      s[p].lam = lam;
      s[p].step2 = step2;
      s[p].ok = f0.pos;
      conv[p] = fast;
      all_fast = all_fast && fast;
      // row residuals at the corrected point: r(Y0 + d) = r(Y0) + a[0:3] . d
#pragma unroll
      for (int k = 0; k < V; ++k) {
        ra[p][k] = fmaf(a[p][k][0], d0, fmaf(a[p][k][1], d1, fmaf(a[p][k][2], d2, ra[p][k])));
        rb[p][k] = fmaf(b[p][k][0], d0, fmaf(b[p][k][1], d1, fmaf(b[p][k][2], d2, rb[p][k])));
      }
      have_res[p] = true;
    }
    if (!SKA_WARP_ALL(all_fast)) {
      // general path: full secular iteration (refactorise at every lam), quadratically convergent
#pragma unroll 1
      for (int it = 0; it < kSecularMaxIter; ++it) {
        bool done = true;
#pragma unroll
        for (int p = 0; p < PTS; ++p) {
          float den, r1[V], r2[V];
          const float lam = rayleigh<V, CONF>(a[p], b[p], w2[p], s[p].y0, s[p].y1, s[p].y2, cx, cy, cz, r1, r2, den);
          if (!conv[p]) {
            const bool c1 = secular_step(M[p], cx, cy, cz, lam, s[p]);
            conv[p] = c1;
            have_res[p] = false;
          }
          // a lane that lost positive-definiteness can never certify: do not wait for it
          done = done && (conv[p] || !s[p].ok);
        }
        if (SKA_WARP_ALL(done)) break;
      }
    }
#pragma unroll
    for (int p = 0; p < PTS; ++p) {
      Y[p][0] = s[p].y0;
      Y[p][1] = s[p].y1;
      Y[p][2] = s[p].y2;
      const bool finite_in = fabsf(M[p].m33) <= 3.0e38f;  // false for NaN / inf inputs
      need64[p] = !(conv[p] && s[p].ok && well[p]) && finite_in;
      if (!finite_in) status[p] = 2;
    }
  } else if (SOLVER == kSolverJacobi32) {
#pragma unroll
    for (int p = 0; p < PTS; ++p) solve_jacobi32(M[p], cx, cy, cz, Y[p]);
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
          have_res[p] = false;
          if (SOLVER == kSolverSecular) status[p] = 1;
        }
      }
    }
  }

  // ---- fused reprojection scoring, differential form:
  //   proj - obs = -(row . [Y;1]) / z  (+ fx*dx_distortion - skew*y)
  // so the ~1e3 px magnitudes of proj and obs never meet in fp32.
#pragma unroll
  for (int p = 0; p < PTS; ++p) {
    X[p][0] = Y[p][0] + cx;
    X[p][1] = Y[p][1] + cy;
    X[p][2] = Y[p][2] + cz;
    if (!(fabsf(X[p][0]) <= 3.0e38f && fabsf(X[p][1]) <= 3.0e38f && fabsf(X[p][2]) <= 3.0e38f)) status[p] = 2;
    if (!have_res[p]) {
#pragma unroll
      for (int k = 0; k < V; ++k) {
        ra[p][k] = fmaf(a[p][k][0], Y[p][0], fmaf(a[p][k][1], Y[p][1], fmaf(a[p][k][2], Y[p][2], a[p][k][3])));
        rb[p][k] = fmaf(b[p][k][0], Y[p][0], fmaf(b[p][k][1], Y[p][1], fmaf(b[p][k][2], Y[p][2], b[p][k][3])));
      }
    }
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const CamDev& c = cam[k];
      const float z = fmaf(c.Ph[8], Y[p][0], fmaf(c.Ph[9], Y[p][1], fmaf(c.Ph[10], Y[p][2], c.Ph[11]))) + c.Pl[11];
      const float iz = rcp_fast(z);
      float eu = -ra[p][k] * iz, ev = -rb[p][k] * iz;
      if (DIST) {
        const float x = fmaf(c.Rxy[0], Y[p][0], fmaf(c.Rxy[1], Y[p][1], fmaf(c.Rxy[2], Y[p][2], c.txy[0]))) * iz;
        const float y = fmaf(c.Rxy[3], Y[p][0], fmaf(c.Rxy[4], Y[p][1], fmaf(c.Rxy[5], Y[p][2], c.txy[1]))) * iz;
        float dx, dy;
        distort_delta<(DIST >= 2)>(c, x, y, dx, dy);
        eu = fmaf(c.fx, dx, eu);
        ev = fmaf(c.fy, dy, ev);
        if (DIST >= 2) eu = fmaf(-c.skew, y, eu);
      }
      du[p][k] = eu;
      dv[p][k] = ev;
    }
  }
}

}  // namespace ska
