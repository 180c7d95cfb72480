/* ska.h - C ABI of libska.so: B200 (sm_100a) kernels for batched multi-view 3D keypoint
 * reconstruction (weighted DLT triangulation + fused reprojection scoring + Schur LM bundle
 * adjustment).
 *
 * The reference (ChenKaiXuSan/Skiing_Analysis_PyTorch) is pure Python and has no FFI layer; its
 * boundary for this path is a set of Python call signatures.  Each entry
 * point below names the reference interface whose arithmetic it replaces (file:line relative to
 * the reference checkout); INTEGRATION.md shows the ctypes stub a maintainer adds on that side.
 *
 * Conventions
 *  - plain C, no torch / C++ types; every `d_*` pointer is a DEVICE pointer owned by the caller
 *    (e.g. the PyTorch allocator); the library never allocates, frees or retains caller memory.
 *  - `SkaCamera` arrays and `centre` are HOST pointers, read during the call only.
 *  - all work is enqueued on `stream` (a cudaStream_t passed as void*); no hidden synchronisation.
 *  - return value: 0 = ok, negative = SKA_E* argument error, positive = cudaError_t.
 *    `ska_last_error()` returns a thread-local message for the last non-zero return.
 *  - re-entrant and thread safe: no mutable globals, per-call constants travel as kernel
 *    parameters (never through a shared __constant__ symbol).
 */
#ifndef SKA_H_
#define SKA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SKA_ABI_VERSION 1

#define SKA_OK 0
#define SKA_EINVAL -1       /* null pointer / bad size / bad enum */
#define SKA_EUNSUPPORTED -2 /* valid request outside what the kernels implement (e.g. tilted sensor) */
#define SKA_EALIGN -3       /* pointer not aligned as documented */
#define SKA_EWORKSPACE -4   /* workspace too small */

#define SKA_MAX_VIEWS 8

/* memory layout of per-observation tensors */
#define SKA_LAYOUT_VIEW_MAJOR 0  /* kpts (V,T,J,2), conf/err (V,T,J): the reference's separate
                                    left_kpts/right_kpts arrays, triangulation/main.py:99-109 */
#define SKA_LAYOUT_FRAME_MAJOR 1 /* kpts (T,V,J,2), conf/err (T,V,J): bundle_adjustment/loss.py x2d/conf2d */

/* flags for ska_triangulate_reproject_f32 */
#define SKA_SOLVER_SECULAR 0u    /* default: fp32 secular/Rayleigh iteration with interlacing certificate,
                                    fp64 Jacobi fallback for points that fail it */
#define SKA_SOLVER_JACOBI64 1u   /* fp64 rows + fp64 register-resident cyclic Jacobi for every point */
#define SKA_SOLVER_JACOBI32 2u   /* fp32 register-resident cyclic Jacobi (north-star design point; measurement) */
#define SKA_SOLVER_MASK 3u
#define SKA_WEIGHT_SQRT 4u       /* DLT row weight = sqrt(conf) (matches loss.py's conf*err^2); default = conf */
#define SKA_PINHOLE_REPROJ 8u    /* ignore dist[] when scoring (quirk Q1 switch: dist=None) */

/* One calibrated camera, world->camera x_c = R x_w + t, fp64, row-major. HOST struct.
 * K is the full 3x3 intrinsic matrix: triangulation uses P = K [R|t] (triangulate.py:65-66,
 * vggt/triangulate.py:13-16); cv2-style reprojection uses fx,fy,cx,cy only (skew ignored, like
 * cv2.projectPoints); loss-style projection uses all of K (loss.py:74-82).
 * dist: OpenCV order k1,k2,p1,p2,k3,k4,k5,k6,s1,s2,s3,s4,taux,tauy; all zero = pinhole.
 * taux/tauy != 0 -> SKA_EUNSUPPORTED. */
typedef struct SkaCamera {
  double K[9];
  double R[9];
  double t[3];
  double dist[14];
} SkaCamera;

int ska_abi_version(void);
const char* ska_last_error(void);
/* compiled-for architecture string, e.g. "sm_100a" */
const char* ska_build_arch(void);

/* Fused weighted V-view DLT triangulation + reprojection scoring.
 * Replaces, for a whole clip in ONE launch:
 *   triangulate_joints         triangulation/triangulate.py:60-68   (cv2.triangulatePoints)
 *   triangulate_point / loop   vggt/triangulate.py:19-34, :64-71    (np.linalg.svd per joint)
 *   reproject_points + errors  triangulation/reproject.py:49-83, :243-244 (cv2.projectPoints, |proj-kpt|)
 *   the per-frame loop         triangulation/triangulate.py:76-116
 * cams[V]   host; static rig.  If d_Rt_frames != NULL it holds per-frame world->camera extrinsics
 *           (T,V,12) fp64 device = [R row-major (9), t (3)] and cams[v].R/t are ignored
 *           (process_triangulate passes per-frame R[i],T[i]: triangulate.py:76-82).
 * centre    host[3] or NULL: numerical conditioning origin only (results do not depend on it
 *           beyond rounding); NULL = least-squares intersection of the optical axes.
 * d_kpts    pixels, layout per `layout`; d_conf nullable (NULL = unit weights = reference behaviour).
 * d_X       (T,J,3) f32 out.  d_err: per-view pixel error |proj-kpt|, same layout as conf, nullable.
 * d_proj    reprojected pixels, same layout as kpts, nullable.
 * d_status  nullable (T,J) uint8 out: 0 = fast path certified, 1 = Jacobi fallback used,
 *           2 = non-finite result.
 * Alignment: d_kpts 8 B, everything else 4 B (16 B on all of them enables the 128-bit path). */
int ska_triangulate_reproject_f32(const SkaCamera* cams, int32_t V, const double* centre,
                                  const double* d_Rt_frames, const float* d_kpts, const float* d_conf,
                                  int64_t T, int32_t J, int32_t layout, uint32_t flags, float* d_X,
                                  float* d_err, float* d_proj, uint8_t* d_status, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SKA_H_ */
