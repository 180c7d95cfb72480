// ska_prep.h - fp64 preparation of kernel-side cameras.  Runs on the host for a static rig (once per
// call) and on the device for per-frame extrinsics (tri_prep_frames kernel): same code, SKA_HD.
#pragma once
#include <math.h>
#include <string.h>

#include "../../include/ska.h"
#include "ska_math.cuh"

namespace ska {

// returns SKA_OK / SKA_EINVAL / SKA_EUNSUPPORTED; fills the fp32 centred camera and the fp64 P.
// needs_dist_path is set when scoring must take the distortion/skew branch.
// dist_level: 0 = pinhole scoring, 1 = rational + tangential, 2 = thin prism and/or skew as well.
SKA_HD int prep_camera(const SkaCamera& in, const double c[3], bool pinhole_reproj, CamDev& out, double P64[12],
                       int& dist_level, const char** why) {
  const double k22 = in.K[8];
  if (!(fabs(k22) > 0.0) || !isfinite(k22)) {
    *why = "K[2][2] must be finite and non-zero";
    return SKA_EINVAL;
  }
  if (in.K[6] != 0.0 || in.K[7] != 0.0 || in.K[3] != 0.0) {
    *why = "K must be upper triangular (K[1][0], K[2][0], K[2][1] == 0)";
    return SKA_EUNSUPPORTED;
  }
  if (in.dist[12] != 0.0 || in.dist[13] != 0.0) {
    *why = "tilted sensor model (taux, tauy) is not implemented";
    return SKA_EUNSUPPORTED;
  }
  double K[9];
  for (int i = 0; i < 9; ++i) K[i] = in.K[i] / k22;
  const double* R = in.R;
  // un-centred fp64 P = K [R|t]
  for (int r = 0; r < 3; ++r)
    for (int m = 0; m < 4; ++m) {
      double acc = 0.0;
      for (int k = 0; k < 3; ++k) acc += K[3 * r + k] * (m < 3 ? R[3 * k + m] : in.t[k]);
      P64[4 * r + m] = acc;
    }
  // centred translation t' = R c + t and P' = K [R|t']
  double tc[3];
  for (int k = 0; k < 3; ++k) tc[k] = R[3 * k] * c[0] + R[3 * k + 1] * c[1] + R[3 * k + 2] * c[2] + in.t[k];
  for (int r = 0; r < 3; ++r)
    for (int m = 0; m < 4; ++m) {
      double acc = 0.0;
      for (int k = 0; k < 3; ++k) acc += K[3 * r + k] * (m < 3 ? R[3 * k + m] : tc[k]);
      const float hi = (float)acc;
      out.Ph[4 * r + m] = hi;
      out.Pl[4 * r + m] = (float)(acc - (double)hi);
    }
  for (int i = 0; i < 6; ++i) out.Rxy[i] = (float)R[i];
  out.txy[0] = (float)tc[0];
  out.txy[1] = (float)tc[1];
  out.fx = (float)K[0];
  out.fy = (float)K[4];
  out.skew = (float)K[1];
  out.ifx = (float)(1.0 / K[0]);
  out.ify = (float)(1.0 / K[4]);
  out.ncx = (float)(-K[2] / K[0]);
  out.ncy = (float)(-K[5] / K[4]);
  const double* d = in.dist;
  bool any = false;
  for (int i = 0; i < 12; ++i) any = any || (d[i] != 0.0);
  if (pinhole_reproj) any = false;
  const double z = any ? 1.0 : 0.0;
  out.dk[0] = (float)(z * (d[0] - d[5]));
  out.dk[1] = (float)(z * (d[1] - d[6]));
  out.dk[2] = (float)(z * (d[4] - d[7]));
  out.kd[0] = (float)(z * d[5]);
  out.kd[1] = (float)(z * d[6]);
  out.kd[2] = (float)(z * d[7]);
  out.p1 = (float)(z * d[2]);
  out.p2 = (float)(z * d[3]);
  out.tp1 = (float)(2.0 * z * d[2]);
  out.tp2 = (float)(2.0 * z * d[3]);
  for (int i = 0; i < 4; ++i) out.s[i] = (float)(z * d[8 + i]);
  const bool prism = any && (d[8] != 0.0 || d[9] != 0.0 || d[10] != 0.0 || d[11] != 0.0);
  dist_level = (prism || K[1] != 0.0) ? 2 : (any ? 1 : 0);
  return SKA_OK;
}

// Default conditioning origin: the point closest (least squares) to all optical axes, regularised
// towards the mean camera centre along directions the axes do not determine (parallel axes).
SKA_HD void default_centre(const SkaCamera* cams, int V, double c[3]) {
  double A[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, b[3] = {0, 0, 0}, mean[3] = {0, 0, 0};
  for (int v = 0; v < V; ++v) {
    const double* R = cams[v].R;
    const double* t = cams[v].t;
    double C[3], d[3];
    for (int k = 0; k < 3; ++k) {
      C[k] = -(R[k] * t[0] + R[3 + k] * t[1] + R[6 + k] * t[2]);  // -R^T t
      d[k] = R[6 + k];                                             // optical axis = third row of R
      mean[k] += C[k] / V;
    }
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) {
        const double pij = (i == j ? 1.0 : 0.0) - d[i] * d[j];
        A[i][j] += pij;
        b[i] += pij * C[j];
      }
    }
  }
  const double mu = 1e-6 * V;
  for (int i = 0; i < 3; ++i) {
    A[i][i] += mu;
    b[i] += mu * mean[i];
  }
  // 3x3 solve by Cramer (A is SPD thanks to mu)
  const double det = A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
                     A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
  if (!(fabs(det) > 0.0) || !isfinite(det)) {
    for (int k = 0; k < 3; ++k) c[k] = isfinite(mean[k]) ? mean[k] : 0.0;
    return;
  }
  double M[3][3];
  for (int col = 0; col < 3; ++col) {
    for (int r = 0; r < 3; ++r)
      for (int q = 0; q < 3; ++q) M[r][q] = A[r][q];
    for (int r = 0; r < 3; ++r) M[r][col] = b[r];
    const double dc = M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
                      M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
    c[col] = dc / det;
  }
  for (int k = 0; k < 3; ++k)
    if (!isfinite(c[k])) c[k] = 0.0;
  // the kernel holds c in fp32: round here so host prep and kernel agree exactly
  for (int k = 0; k < 3; ++k) c[k] = (double)(float)c[k];
}

}  // namespace ska
