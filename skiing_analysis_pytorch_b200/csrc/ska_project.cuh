// ska_project.cuh - per-observation projection arithmetic of the standalone (non-fused) entry
// points: cv2-style reprojection and the loss.py pinhole with its analytic adjoint.
// __host__ __device__ so tests/hostemu can run the same code on the CPU-only authoring box.
//
// Reference arithmetic replaced (file:line relative to the reference checkout):
//   cv2.projectPoints call sites   triangulation/reproject.py:63-83, bundle_adjustment/reproject.py:135-153
//   project_points                 bundle_adjustment/loss.py:17-84   (Z clamp :67, full K :74-82)
//   reprojection_loss (+ autograd) bundle_adjustment/loss.py:90-94
#pragma once
#include "ska_math.cuh"

namespace ska {

// cv2.projectPoints camera, fp64: cv2 itself computes in double after the caller's float32
// down-cast (quirk Q2), so the standalone reprojection does the same.
struct CamCv64 {
  double R[9], t[3];
  double fx, fy, cx, cy;
  double d[12];  // k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4 (all zero = pinhole)
};

SKA_HD void project_cv64(const CamCv64& c, double X, double Y, double Z, double& u, double& v) {
  const double xc = c.R[0] * X + c.R[1] * Y + c.R[2] * Z + c.t[0];
  const double yc = c.R[3] * X + c.R[4] * Y + c.R[5] * Z + c.t[1];
  const double zc = c.R[6] * X + c.R[7] * Y + c.R[8] * Z + c.t[2];
  const double iz = (zc != 0.0) ? 1.0 / zc : 1.0;  // cv2: z = z ? 1./z : 1
  double xd, yd;
  distort64(c.d, xc * iz, yc * iz, xd, yd);
  u = c.fx * xd + c.cx;
  v = c.fy * yd + c.cy;
}

// ---------------------------------------------------------------------------------------------
// loss.py projection in the caller's dtype S (float or double): X_c = R X + t, Z = max(z, 1e-6),
// (x, y) = X_c.xy / Z, (u, v) = rows 0 and 1 of K [x, y, 1].
template <typename S>
struct LossObs {
  S x, y, iz;     // normalised coordinates and 1 / max(z, 1e-6)
  S u, v;         // projected pixel
  bool clamped;
};

template <typename S>
SKA_HD void project_loss(const S* __restrict__ R, const S* __restrict__ t, const S* __restrict__ K, const S X[3],
                         LossObs<S>& o) {
  const S xc = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
  const S yc = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
  const S zc = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
  o.clamped = zc < S(1e-6);
  o.iz = S(1) / (o.clamped ? S(1e-6) : zc);
  o.x = xc * o.iz;
  o.y = yc * o.iz;
  o.u = K[0] * o.x + K[1] * o.y + K[2];
  o.v = K[3] * o.x + K[4] * o.y + K[5];
}

// Adjoint of project_loss for one observation: given g = d(loss)/d(u, v), returns
//   gXc = d/d(X_c)  (so that gX += R^T gXc, gt += gXc, gR += gXc (x) X)
//   gK[6] = d/d(K00 K01 K02 K10 K11 K12)
// d/dz vanishes while the clamp is active (autograd of clamp(min=1e-6), loss.py:67).
template <typename S>
SKA_HD void project_loss_adjoint(const S* __restrict__ K, const LossObs<S>& o, S gu, S gv, S gXc[3], S gK[6]) {
  const S gx = K[0] * gu + K[3] * gv;
  const S gy = K[1] * gu + K[4] * gv;
  gXc[0] = gx * o.iz;
  gXc[1] = gy * o.iz;
  gXc[2] = o.clamped ? S(0) : -(gx * o.x + gy * o.y) * o.iz;
  gK[0] = gu * o.x;
  gK[1] = gu * o.y;
  gK[2] = gu;
  gK[3] = gv * o.x;
  gK[4] = gv * o.y;
  gK[5] = gv;
}

}  // namespace ska
