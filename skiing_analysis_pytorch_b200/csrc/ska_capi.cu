// ska_capi.cu - the extern "C" boundary of libska.so: argument validation and dispatch only.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "ska_internal.h"

namespace ska {

static thread_local char g_err[512] = "";

int set_error(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg ? msg : "");
  return code;
}

}  // namespace ska

using namespace ska;

extern "C" {
#pragma GCC visibility push(default)

int ska_abi_version(void) { return SKA_ABI_VERSION; }
const char* ska_last_error(void) { return g_err; }
const char* ska_build_arch(void) { return "sm_100a"; }

int ska_triangulate_reproject_f32(const SkaCamera* cams, int32_t V, const double* centre, const double* d_Rt_frames,
                                  const float* d_kpts, const float* d_conf, int64_t T, int32_t J, int32_t layout,
                                  uint32_t flags, float* d_X, float* d_err, float* d_proj, uint8_t* d_status,
                                  void* stream) {
  if (cams == nullptr || d_kpts == nullptr || d_X == nullptr) return set_error(SKA_EINVAL, "cams, d_kpts and d_X must not be NULL");
  if (V < 2 || V > SKA_MAX_VIEWS) return set_error(SKA_EINVAL, "V must be in 2..SKA_MAX_VIEWS");
  if (T < 0 || J < 1) return set_error(SKA_EINVAL, "T must be >= 0 and J >= 1");
  if (layout != SKA_LAYOUT_VIEW_MAJOR && layout != SKA_LAYOUT_FRAME_MAJOR) return set_error(SKA_EINVAL, "bad layout");
  if ((flags & SKA_SOLVER_MASK) == 3u) return set_error(SKA_EINVAL, "bad solver");
  if (T * (int64_t)J >= (int64_t)1 << 31) return set_error(SKA_EINVAL, "T*J must be < 2^31 per call; shard the clip");
  if (d_Rt_frames != nullptr) return set_error(SKA_EUNSUPPORTED, "per-frame extrinsics: use ska_triangulate_reproject_perframe_f32");
  auto al = [](const void* p, uintptr_t n) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) % n) == 0; };
  if (!al(d_kpts, 8) || !al(d_proj, 8)) return set_error(SKA_EALIGN, "d_kpts / d_proj must be 8-byte aligned");
  if (!al(d_conf, 4) || !al(d_err, 4) || !al(d_X, 4)) return set_error(SKA_EALIGN, "float pointers must be 4-byte aligned");
  if (T == 0) return SKA_OK;
  TriArgs a;
  a.cams = cams;
  a.V = V;
  if (centre != nullptr) {
    for (int k = 0; k < 3; ++k) {
      if (!isfinite(centre[k])) return set_error(SKA_EINVAL, "centre must be finite");
      a.centre[k] = (double)(float)centre[k];
    }
  } else {
    default_centre(cams, V, a.centre);
  }
  a.Rt_frames = d_Rt_frames;
  a.kpts = d_kpts;
  a.conf = d_conf;
  a.T = T;
  a.J = J;
  a.layout = layout;
  a.flags = flags;
  a.X = d_X;
  a.err = d_err;
  a.proj = d_proj;
  a.status = d_status;
  a.stream = stream;
  return triangulate_dispatch(a);
}

static int check_problem(const SkaBaProblem* p, bool need_points) {
  if (p == nullptr) return set_error(SKA_EINVAL, "problem must not be NULL");
  if (p->C < 2 || p->C > SKA_MAX_VIEWS) return set_error(SKA_EINVAL, "C must be in 2..SKA_MAX_VIEWS");
  if (p->T < 0 || p->J < 1) return set_error(SKA_EINVAL, "T must be >= 0 and J >= 1");
  if (p->T * (int64_t)p->J >= (int64_t)1 << 31) return set_error(SKA_EINVAL, "T*J must be < 2^31 per rank; shard the clip");
  if (p->layout != SKA_LAYOUT_VIEW_MAJOR && p->layout != SKA_LAYOUT_FRAME_MAJOR) return set_error(SKA_EINVAL, "bad layout");
  if (p->d_cams == nullptr || p->d_ctrl == nullptr || p->d_red == nullptr || p->d_red2 == nullptr || p->d_delta == nullptr)
    return set_error(SKA_EINVAL, "d_cams, d_ctrl, d_red, d_red2 and d_delta must not be NULL");
  if (need_points) {
    if (p->T > 0 && (p->d_x2d == nullptr || p->d_conf == nullptr || p->d_Xpp == nullptr))
      return set_error(SKA_EINVAL, "d_x2d, d_conf and d_Xpp must not be NULL");
    if (p->d_workspace == nullptr) return set_error(SKA_EINVAL, "d_workspace must not be NULL");
    if (reinterpret_cast<uintptr_t>(p->d_x2d) % 8 != 0) return set_error(SKA_EALIGN, "d_x2d must be 8-byte aligned");
    if (reinterpret_cast<uintptr_t>(p->d_workspace) % 16 != 0) return set_error(SKA_EALIGN, "d_workspace must be 16-byte aligned");
  }
  return SKA_OK;
}

int32_t ska_ba_red_doubles(int32_t C) { return (C < 2 || C > SKA_MAX_VIEWS) ? 0 : ba_red_size(C); }

size_t ska_ba_workspace_bytes(int32_t C) {
  if (C < 2 || C > SKA_MAX_VIEWS) return 0;
  return (size_t)ba_max_grid() * (size_t)ba_red_size(C) * sizeof(double);
}

int ska_ba_sum_f32(const float* d_x, int64_t count, double* d_out, void* d_workspace, size_t ws_bytes, void* stream) {
  if (d_out == nullptr || d_workspace == nullptr || count < 0 || (count > 0 && d_x == nullptr))
    return set_error(SKA_EINVAL, "d_x, d_out and d_workspace must not be NULL; count >= 0");
  return ba_sum(d_x, count, d_out, d_workspace, ws_bytes, stream);
}

int ska_ba_linearize_f32(const SkaBaProblem* p, void* stream) {
  const int rc = check_problem(p, true);
  if (rc != SKA_OK) return rc;
  return ba_linearize(*p, (cudaStream_t)stream);
}

int ska_ba_solve_f64(const SkaBaProblem* p, uint64_t free_mask, void* stream) {
  const int rc = check_problem(p, false);
  if (rc != SKA_OK) return rc;
  return ba_solve(p->C, free_mask, p->d_red, p->d_cams, p->d_ctrl, p->d_delta, stream);
}

int ska_ba_backsub_f32(const SkaBaProblem* p, void* stream) {
  const int rc = check_problem(p, true);
  if (rc != SKA_OK) return rc;
  return ba_backsub(*p, (cudaStream_t)stream);
}

int ska_ba_control_f64(const SkaBaProblem* p, void* stream) {
  const int rc = check_problem(p, false);
  if (rc != SKA_OK) return rc;
  return ba_control(p->C, p->d_red, p->d_red2, p->d_cams, p->d_ctrl, p->d_hist, stream);
}

#pragma GCC visibility pop
}  // extern "C"
