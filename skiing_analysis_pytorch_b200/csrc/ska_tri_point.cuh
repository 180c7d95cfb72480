// ska_tri_point.cuh - one (frame, joint): weighted V-view DLT + fused reprojection scoring.
// __host__ __device__ so tests/hostemu can run the very same code path on the CPU-only box.
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

// fp64 rows from the un-centred fp64 P, fp64 A^T A, fp64 cyclic Jacobi; exact-mode solver and the
// fallback of the fp32 fast path.  Mirrors the reference (fp64 SVD of the same A) to ~1e-12.
template <int V>
SKA_HD_NOINLINE void solve_jacobi64(const double (*P64)[12], const float* u, const float* v, const float* w2, double X[3]) {
  double a[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) a[i][j] = 0.0;
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const double* P = P64[k];
    double ra[4], rb[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      ra[m] = (double)u[k] * P[8 + m] - P[m];
      rb[m] = (double)v[k] * P[8 + m] - P[4 + m];
    }
    const double ww = (double)w2[k];
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
  X[0] = h[0] * ih;
  X[1] = h[1] * ih;
  X[2] = h[2] * ih;
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

// PTS points in lockstep (independent dependency chains interleave -> ILP).
// u,v,w2: [PTS][V] pixel coordinates and squared row weights (w2 unused if !CONF).
// SOLVER: kSolverSecular (fast path + certified fallback), kSolverJacobi64, kSolverJacobi32.
// DIST: 0 = pinhole scoring, 1 = cv2 distortion model (also carries the skew correction).
// Outputs: X (un-centred), du/dv = reprojected minus observed pixel per view, status.
template <int V, int PTS, bool CONF, int DIST, uint32_t SOLVER>
SKA_HD void tri_points(const CamDev* __restrict__ cam, const double (*P64)[12], const float cx, const float cy,
                       const float cz, const float (*u)[V], const float (*v)[V], const float (*w2)[V],
                       float (*X)[3], float (*du)[V], float (*dv)[V], uint8_t* status) {
  // ---- normal matrices in centred coordinates
  Sym4 M[PTS];
#pragma unroll
  for (int p = 0; p < PTS; ++p) sym4_zero(M[p]);
#pragma unroll
  for (int k = 0; k < V; ++k) {
#pragma unroll
    for (int p = 0; p < PTS; ++p) {
      float a[4], b[4];
      dlt_rows<true>(cam[k], u[p][k], v[p][k], a, b);
      if (CONF) {
        sym4_rank1(M[p], a, w2[p][k]);
        sym4_rank1(M[p], b, w2[p][k]);
      } else {
        sym4_rank1_unit(M[p], a);
        sym4_rank1_unit(M[p], b);
      }
    }
  }

  float Y[PTS][3];
  bool need64[PTS];
#pragma unroll
  for (int p = 0; p < PTS; ++p) {
    need64[p] = (SOLVER == kSolverJacobi64);
    status[p] = 0;
  }
  if (SOLVER == kSolverSecular) {
    SecularState s[PTS];
    bool conv[PTS];
#pragma unroll
    for (int p = 0; p < PTS; ++p) {
      // lam = 0: the inhomogeneous least-squares point
      const Ldl3 f0 = ldl3(M[p].m00, M[p].m01, M[p].m02, M[p].m11, M[p].m12, M[p].m22);
      ldl3_solve(f0, -M[p].m03, -M[p].m13, -M[p].m23, s[p].y0, s[p].y1, s[p].y2);
      s[p].lam = 0.f;
      s[p].ok = f0.pos;
      conv[p] = false;
    }
#pragma unroll 1
    for (int it = 0; it < kSecularMaxIter; ++it) {
      bool done = true;
#pragma unroll
      for (int p = 0; p < PTS; ++p) {
        // Rayleigh quotient from the rows (never from M: that would cancel 1e8 -> 1 in fp32)
        float num = 0.f;
#pragma unroll
        for (int k = 0; k < V; ++k) {
          float a[4], b[4];
          dlt_rows<true>(cam[k], u[p][k], v[p][k], a, b);
          const float ra = fmaf(a[0], s[p].y0, fmaf(a[1], s[p].y1, fmaf(a[2], s[p].y2, a[3])));
          const float rb = fmaf(b[0], s[p].y0, fmaf(b[1], s[p].y1, fmaf(b[2], s[p].y2, b[3])));
          const float rr = fmaf(ra, ra, rb * rb);
          num = CONF ? fmaf(w2[p][k], rr, num) : (num + rr);
        }
        const float X0 = s[p].y0 + cx, X1 = s[p].y1 + cy, X2 = s[p].y2 + cz;
        const float den = fmaf(X0, X0, fmaf(X1, X1, fmaf(X2, X2, 1.0f)));
        const float lam = num * rcp_fast(den);
        const bool c1 = secular_step(M[p], cx, cy, cz, lam, s[p]);
        conv[p] = conv[p] || c1;
        // a lane that lost positive-definiteness can never certify: do not wait for it
        done = done && (conv[p] || !s[p].ok);
      }
      // converged lanes keep iterating harmlessly until the whole warp agrees
      if (SKA_WARP_ALL(done)) break;
    }
#pragma unroll
    for (int p = 0; p < PTS; ++p) {
      Y[p][0] = s[p].y0;
      Y[p][1] = s[p].y1;
      Y[p][2] = s[p].y2;
      const bool finite_in = fabsf(M[p].m33) <= 3.0e38f;  // false for NaN / inf inputs
      need64[p] = !(conv[p] && s[p].ok) && finite_in;
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
#pragma unroll 1
      for (int p = 0; p < PTS; ++p) {
        if (need64[p]) {
          double Xd[3];
          solve_jacobi64<V>(P64, u[p], v[p], w2[p], Xd);
          Y[p][0] = (float)(Xd[0] - (double)cx);
          Y[p][1] = (float)(Xd[1] - (double)cy);
          Y[p][2] = (float)(Xd[2] - (double)cz);
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
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const CamDev& c = cam[k];
      float a[4], b[4];
      dlt_rows<true>(c, u[p][k], v[p][k], a, b);
      const float ra = fmaf(a[0], Y[p][0], fmaf(a[1], Y[p][1], fmaf(a[2], Y[p][2], a[3])));
      const float rb = fmaf(b[0], Y[p][0], fmaf(b[1], Y[p][1], fmaf(b[2], Y[p][2], b[3])));
      const float z = fmaf(c.Ph[8], Y[p][0], fmaf(c.Ph[9], Y[p][1], fmaf(c.Ph[10], Y[p][2], c.Ph[11]))) + c.Pl[11];
      const float iz = rcp_fast(z);
      float eu = -ra * iz, ev = -rb * iz;
      if (DIST) {
        const float x = fmaf(c.Rxy[0], Y[p][0], fmaf(c.Rxy[1], Y[p][1], fmaf(c.Rxy[2], Y[p][2], c.txy[0]))) * iz;
        const float y = fmaf(c.Rxy[3], Y[p][0], fmaf(c.Rxy[4], Y[p][1], fmaf(c.Rxy[5], Y[p][2], c.txy[1]))) * iz;
        float dx, dy;
        distort_delta(c, x, y, dx, dy);
        eu = fmaf(c.fx, dx, fmaf(-c.skew, y, eu));
        ev = fmaf(c.fy, dy, ev);
      }
      du[p][k] = eu;
      dv[p][k] = ev;
    }
  }
}

}  // namespace ska
