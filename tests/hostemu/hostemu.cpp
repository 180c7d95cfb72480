// Host emulation of the per-point device arithmetic (TEST INFRASTRUCTURE, CPU-only box).
// Compiles ska_tri_point.cuh with g++ so the fp32 secular solver / differential reprojection can
// be checked against the fp64 oracle without a GPU.  Never linked into libska.so, never imported
// by the package: the product path is CUDA only.
#include <stdint.h>
#include <vector>

#include "../../skiing_analysis_pytorch_b200/csrc/ska_prep.h"
#include "../../skiing_analysis_pytorch_b200/csrc/ska_project.cuh"
#include "../../skiing_analysis_pytorch_b200/csrc/ska_tri_point.cuh"

using namespace ska;

template <int V>
static int run(const SkaCamera* cams, const double* centre, const float* kpts, const float* conf, int64_t N,
               uint32_t flags, float* X, float* err, uint8_t* status) {
  double c[3];
  if (centre) {
    for (int k = 0; k < 3; ++k) c[k] = (double)(float)centre[k];
  } else {
    default_centre(cams, V, c);
  }
  CamDev cam[V];
  double P64[V][12];
  int dist = 0;
  for (int v = 0; v < V; ++v) {
    int d;
    const char* why = "";
    int rc = prep_camera(cams[v], c, (flags & SKA_PINHOLE_REPROJ) != 0, cam[v], P64[v], d, &why);
    if (rc) return rc;
    dist = d > dist ? d : dist;
  }
  const int lo = (flags >> 8) & 3;  // test hook: lo-part level of the DLT rows (default 1)
  const uint32_t solver = flags & SKA_SOLVER_MASK;
  if ((flags >> 12) & 1) {
    // test hook: the view-pair form (tri_point_vp) the kernel uses for even V >= 4
    if constexpr (V >= 4 && V % 2 == 0) {
      CamPairDev camp[V / 2];
      for (int i = 0; i < V / 2; ++i) make_cam_pair(cam[2 * i], cam[2 * i + 1], camp[i]);
      if (dist > 1) return SKA_EUNSUPPORTED;
      for (int64_t i = 0; i < N; ++i) {
        float u[V], vv[V], w2[V], du[V], dv[V], Xp[3];
        for (int v = 0; v < V; ++v) {
          u[v] = kpts[(v * N + i) * 2];
          vv[v] = kpts[(v * N + i) * 2 + 1];
          const float cf = conf ? conf[v * N + i] : 1.0f;
          w2[v] = (flags & SKA_WEIGHT_SQRT) ? cf : cf * cf;
        }
        uint8_t st;
        PointSource src;
        src.kpts = kpts + 2 * i;
        src.conf = conf ? conf + i : nullptr;
        src.k_sV = 2 * N;
        src.c_sV = N;
        src.weight_sqrt = (flags & SKA_WEIGHT_SQRT) ? 1u : 0u;
        const float cx = (float)c[0], cy = (float)c[1], cz = (float)c[2];
        const int rows = (flags >> 13) & 3;  // where the rows live between the passes (kRowsRegs / Recomp / Smem)
        const bool samek = (flags >> 15) & 1; // read every view's intrinsics from view 0 (caller knows they are equal)
        F2 slab[4 * V];
#define VP2(CONF, DIST, ROWS, SK) tri_point_vp<V, CONF, DIST, ROWS, SK>(camp, cam, P64, cx, cy, cz, u, vv, w2, src, Xp, du, dv, st, slab, 1)
#define VP(CONF, DIST)                                          \
  do {                                                          \
    if (samek) {                                                \
      if (rows == 1) VP2(CONF, DIST, kRowsRecomp, true);        \
      else if (rows == 2) VP2(CONF, DIST, kRowsSmem, true);     \
      else VP2(CONF, DIST, kRowsRegs, true);                    \
    } else {                                                    \
      if (rows == 1) VP2(CONF, DIST, kRowsRecomp, false);       \
      else if (rows == 2) VP2(CONF, DIST, kRowsSmem, false);    \
      else VP2(CONF, DIST, kRowsRegs, false);                   \
    }                                                           \
  } while (0)
        if (conf) {
          if (dist) VP(true, 1); else VP(true, 0);
        } else {
          if (dist) VP(false, 1); else VP(false, 0);
        }
#undef VP
#undef VP2
        for (int k = 0; k < 3; ++k) X[3 * i + k] = Xp[k];
        if (err)
          for (int v = 0; v < V; ++v) err[v * N + i] = sqrtf(du[v] * du[v] + dv[v] * dv[v]);
        if (status) status[i] = st;
      }
      return 0;
    } else {
      return SKA_EUNSUPPORTED;
    }
  }
  if ((flags >> 10) & 1) {
    // test hook: the kernel's PTS = 2 path (two points as one packed F2 computation), pairs (i, i+1)
    for (int64_t i = 0; i + 1 < N + (N & 1); i += 2) {
      const int64_t i0 = (i + 1 < N) ? i : N - 2;  // odd tail: redo the last pair
      float u[2][V], vv[2][V], w2[2][V], du[2][V], dv[2][V], Xp[2][3];
      for (int p = 0; p < 2; ++p)
        for (int v = 0; v < V; ++v) {
          u[p][v] = kpts[(v * N + i0 + p) * 2];
          vv[p][v] = kpts[(v * N + i0 + p) * 2 + 1];
          const float cf = conf ? conf[v * N + i0 + p] : 1.0f;
          w2[p][v] = (flags & SKA_WEIGHT_SQRT) ? cf : cf * cf;
        }
      uint8_t st[2];
      PointSource src;
      src.kpts = kpts + 2 * i0;
      src.conf = conf ? conf + i0 : nullptr;
      src.k_sV = 2 * N;
      src.c_sV = N;
      src.weight_sqrt = (flags & SKA_WEIGHT_SQRT) ? 1u : 0u;
      const float cx = (float)c[0], cy = (float)c[1], cz = (float)c[2];
      if (conf) {
        if (dist) tri_points<V, 2, true, 1, kSolverSecular>(cam, P64, cx, cy, cz, u, vv, w2, src, Xp, du, dv, st);
        else tri_points<V, 2, true, 0, kSolverSecular>(cam, P64, cx, cy, cz, u, vv, w2, src, Xp, du, dv, st);
      } else {
        if (dist) tri_points<V, 2, false, 1, kSolverSecular>(cam, P64, cx, cy, cz, u, vv, w2, src, Xp, du, dv, st);
        else tri_points<V, 2, false, 0, kSolverSecular>(cam, P64, cx, cy, cz, u, vv, w2, src, Xp, du, dv, st);
      }
      for (int p = 0; p < 2; ++p) {
        for (int k = 0; k < 3; ++k) X[3 * (i0 + p) + k] = Xp[p][k];
        if (err)
          for (int v = 0; v < V; ++v) err[v * N + i0 + p] = sqrtf(du[p][v] * du[p][v] + dv[p][v] * dv[p][v]);
        if (status) status[i0 + p] = st[p];
      }
    }
    return 0;
  }
  for (int64_t i = 0; i < N; ++i) {
    float u[V], vv[V], w2[V], du[V], dv[V], Xp[3];
    for (int v = 0; v < V; ++v) {
      u[v] = kpts[(v * N + i) * 2];
      vv[v] = kpts[(v * N + i) * 2 + 1];
      const float cf = conf ? conf[v * N + i] : 1.0f;
      w2[v] = (flags & SKA_WEIGHT_SQRT) ? cf : cf * cf;
    }
    uint8_t st;
    const float cx = (float)c[0], cy = (float)c[1], cz = (float)c[2];
    const float(*U)[V] = (const float(*)[V])u;
    const float(*VV)[V] = (const float(*)[V])vv;
    const float(*W)[V] = (const float(*)[V])w2;
    float(*XO)[3] = (float(*)[3])Xp;
    float(*DU)[V] = (float(*)[V])du;
    float(*DV)[V] = (float(*)[V])dv;
    PointSource src;
    src.kpts = kpts + 2 * i;
    src.conf = conf ? conf + i : nullptr;
    src.k_sV = 2 * N;
    src.c_sV = N;
    src.weight_sqrt = (flags & SKA_WEIGHT_SQRT) ? 1u : 0u;
#define GO3(CONF, DIST, LO)                                                                                    \
  do {                                                                                                         \
    if (solver == kSolverSecular) tri_points<V, 1, CONF, DIST, kSolverSecular, LO>(cam, P64, cx, cy, cz, U, VV, W, src, XO, DU, DV, &st); \
    else if (solver == kSolverJacobi64) tri_points<V, 1, CONF, DIST, kSolverJacobi64, LO>(cam, P64, cx, cy, cz, U, VV, W, src, XO, DU, DV, &st); \
    else tri_points<V, 1, CONF, DIST, kSolverJacobi32, LO>(cam, P64, cx, cy, cz, U, VV, W, src, XO, DU, DV, &st); \
  } while (0)
#define GO(CONF, DIST)                                  \
  do {                                                  \
    if (lo == 1) GO3(CONF, DIST, 0);                    \
    else if (lo == 2) GO3(CONF, DIST, 2);               \
    else GO3(CONF, DIST, 1);                            \
  } while (0)
    if (conf) {
      if (dist == 2) GO(true, 2); else if (dist == 1) GO(true, 1); else GO(true, 0);
    } else {
      if (dist == 2) GO(false, 2); else if (dist == 1) GO(false, 1); else GO(false, 0);
    }
#undef GO
#undef GO3
    for (int k = 0; k < 3; ++k) X[3 * i + k] = Xp[k];
    if (err)
      for (int v = 0; v < V; ++v) err[v * N + i] = sqrtf(du[v] * du[v] + dv[v] * dv[v]);
    if (status) status[i] = st;
  }
  return 0;
}

extern "C" int hostemu_triangulate(const SkaCamera* cams, int32_t V, const double* centre, const float* kpts,
                                   const float* conf, int64_t N, uint32_t flags, float* X, float* err,
                                   uint8_t* status) {
  switch (V) {
    case 2: return run<2>(cams, centre, kpts, conf, N, flags, X, err, status);
    case 3: return run<3>(cams, centre, kpts, conf, N, flags, X, err, status);
    case 4: return run<4>(cams, centre, kpts, conf, N, flags, X, err, status);
    case 8: return run<8>(cams, centre, kpts, conf, N, flags, X, err, status);
    default: return SKA_EUNSUPPORTED;
  }
}

// cv2-style projection (ska_project.cuh project_cv64): cam = R(9) t(3) fx fy cx cy d(12)
extern "C" void hostemu_project_cv(const double* cam, const double* X, int64_t N, double* uv) {
  CamCv64 c;
  for (int k = 0; k < 9; ++k) c.R[k] = cam[k];
  for (int k = 0; k < 3; ++k) c.t[k] = cam[9 + k];
  c.fx = cam[12]; c.fy = cam[13]; c.cx = cam[14]; c.cy = cam[15];
  for (int k = 0; k < 12; ++k) c.d[k] = cam[16 + k];
  for (int64_t i = 0; i < N; ++i) project_cv64(c, X[3 * i], X[3 * i + 1], X[3 * i + 2], uv[2 * i], uv[2 * i + 1]);
}

// loss.py projection + adjoint for one camera (ska_project.cuh): per point outputs
//   uv (N,2), gXc (N,3), gK (N,6) for the cotangent g (N,2)
extern "C" void hostemu_project_loss(const double* R, const double* t, const double* K, const double* X, const double* g,
                                     int64_t N, double* uv, double* gXc, double* gK, uint8_t* clamped) {
  for (int64_t i = 0; i < N; ++i) {
    LossObs<double> o;
    project_loss<double>(R, t, K, X + 3 * i, o);
    uv[2 * i] = o.u;
    uv[2 * i + 1] = o.v;
    clamped[i] = o.clamped ? 1 : 0;
    project_loss_adjoint<double>(K, o, g[2 * i], g[2 * i + 1], gXc + 3 * i, gK + 6 * i);
  }
}

// ---- calibrating BA: one observation's linearisation rows (ska_ba_calib.cuh) -------------------------------
#include "../../skiing_analysis_pytorch_b200/csrc/ska_ba_calib.cuh"

extern "C" int hostemu_calib_obs(const double* cam24, const float* X, const float* uv, int64_t N, float* au, float* av, float* bu,
                                 float* bv, uint8_t* clamped, float* err2) {
  CamC c;
  for (int i = 0; i < 9; ++i) c.R[i] = (float)cam24[i];
  for (int i = 0; i < 3; ++i) c.t[i] = (float)cam24[9 + i];
  c.fx = (float)cam24[12]; c.fy = (float)cam24[13]; c.cx = (float)cam24[14]; c.cy = (float)cam24[15];
  c.k1 = (float)cam24[16]; c.k2 = (float)cam24[17]; c.p1 = (float)cam24[18]; c.p2 = (float)cam24[19]; c.k3 = (float)cam24[20];
  for (int64_t i = 0; i < N; ++i) {
    ObsCalib o;
    calib_obs(c, X + 3 * i, uv[2 * i], uv[2 * i + 1], o);
    for (int k = 0; k < 3; ++k) {
      au[3 * i + k] = o.au[k];
      av[3 * i + k] = o.av[k];
    }
    for (int k = 0; k < kCalibRow; ++k) {
      bu[kCalibRow * i + k] = o.bu[k];
      bv[kCalibRow * i + k] = o.bv[k];
    }
    clamped[i] = o.clamped ? 1 : 0;
    bool cl;
    err2[i] = calib_err2(c, X + 3 * i, uv[2 * i], uv[2 * i + 1], cl);
  }
  return 0;
}

extern "C" int hostemu_calib_tri(int r, int s) { return calib_tri(r, s); }
extern "C" int hostemu_calib_tri_row(int q) { return calib_tri_row(q); }
extern "C" int hostemu_calib_tri_col(int q) { return calib_tri_col(q); }
