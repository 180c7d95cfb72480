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

#pragma GCC visibility pop
}  // extern "C"
