// ska_triangulate_impl.cuh - fused weighted V-view DLT triangulation + reprojection scoring, sm_100a.
//
// One thread owns PTS consecutive (frame, joint) points; a warp therefore owns a contiguous group
// of 32*PTS points whose keypoints are one 128-bit (PTS=2) or 64-bit (PTS=1) load per lane and
// view, fully coalesced.  Cameras are compile-time-indexed kernel parameters (constant bank), so
// every P'/K/distortion coefficient is an immediate FFMA operand - no loads, no shared memory.
// The 4x4 normal matrix, the secular iteration and the per-view scoring stay in registers; X is
// staged through shared memory per warp so the 12-byte/point output leaves as 128-bit stores.
// HBM-bound by design: 8V (+4V conf) bytes in, 12 + 4V bytes out per point, nothing re-read.
//
// Replaces: triangulation/triangulate.py:60-68,76-116; vggt/triangulate.py:19-34,64-71;
//           triangulation/reproject.py:49-83,243-244 (file:line in the reference checkout).
#pragma once
#include <cuda_runtime.h>

#include "ska_internal.h"
#include "ska_tri_point.cuh"

namespace ska {

template <int V>
struct TriParams {
  CamDev cam[V];
  double P64[V][12];
  float c[3];
  uint32_t solver;
  uint32_t weight_sqrt;
  uint32_t frame_major;  // 1: offsets need (t, j); 0: flat point index
  uint32_t x_vec;        // X base 16-byte aligned -> staged 128-bit stores
  int32_t J;
  int64_t N;             // T*J points
  int64_t k_sV, k_sT;    // kpts / proj strides in floats (view, frame)
  int64_t c_sV, c_sT;    // conf / err strides in floats
  const float* kpts;
  const float* conf;
  float* X;
  float* err;
  float* proj;
  uint8_t* status;
};

constexpr int kBlock = 256;

template <int V, int PTS, bool CONF, int DIST, uint32_t SOLVER>
__global__ void __launch_bounds__(kBlock) tri_kernel(const __grid_constant__ TriParams<V> prm) {
  __shared__ __align__(16) float sX[kBlock / 32][32 * PTS * 3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t N = prm.N;
  const int64_t warp_first = ((int64_t)blockIdx.x * kBlock + warp * 32) * PTS;
  const int64_t i_raw = warp_first + (int64_t)lane * PTS;
  // out-of-range lanes recompute the last point(s): every lane stays alive for the warp votes
  const int64_t i0 = (i_raw + PTS <= N) ? i_raw : (N - PTS);
  const bool live = (i_raw + PTS <= N);

  int64_t koff, coff;  // offsets of point i0 within view 0
  if (prm.frame_major) {
    const uint32_t t = (uint32_t)i0 / (uint32_t)prm.J;
    const uint32_t j = (uint32_t)i0 - t * (uint32_t)prm.J;
    koff = (int64_t)t * prm.k_sT + 2 * (int64_t)j;
    coff = (int64_t)t * prm.c_sT + (int64_t)j;
  } else {
    koff = 2 * i0;
    coff = i0;
  }

  float u[PTS][V], v[PTS][V], w2[PTS][V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const float* kp = prm.kpts + koff + (int64_t)k * prm.k_sV;
    if (PTS == 2) {
      const float4 q = __ldcs(reinterpret_cast<const float4*>(kp));
      u[0][k] = q.x; v[0][k] = q.y; u[PTS - 1][k] = q.z; v[PTS - 1][k] = q.w;
    } else {
      const float2 q = __ldcs(reinterpret_cast<const float2*>(kp));
      u[0][k] = q.x; v[0][k] = q.y;
    }
    if (CONF && prm.conf != nullptr) {
      const float* cp = prm.conf + coff + (int64_t)k * prm.c_sV;
      float c0, c1 = 0.f;
      if (PTS == 2) {
        const float2 q = __ldcs(reinterpret_cast<const float2*>(cp));
        c0 = q.x; c1 = q.y;
      } else {
        c0 = __ldcs(cp);
      }
      w2[0][k] = prm.weight_sqrt ? c0 : c0 * c0;
      if (PTS == 2) w2[PTS - 1][k] = prm.weight_sqrt ? c1 : c1 * c1;
    } else {
#pragma unroll
      for (int p = 0; p < PTS; ++p) w2[p][k] = 1.0f;
    }
  }

  float X[PTS][3], du[PTS][V], dv[PTS][V];
  uint8_t st[PTS];
  tri_points<V, PTS, CONF, DIST, SOLVER>(prm.cam, prm.P64, prm.c[0], prm.c[1], prm.c[2], u, v, w2, X, du, dv, st);

  // ---- per-view error / reprojection, coalesced
#pragma unroll
  for (int k = 0; k < V; ++k) {
    if (prm.err != nullptr && live) {
      float* ep = prm.err + coff + (int64_t)k * prm.c_sV;
      const float e0 = sqrt_fast(fmaf(du[0][k], du[0][k], dv[0][k] * dv[0][k]));
      if (PTS == 2) {
        const float e1 = sqrt_fast(fmaf(du[PTS - 1][k], du[PTS - 1][k], dv[PTS - 1][k] * dv[PTS - 1][k]));
        __stcs(reinterpret_cast<float2*>(ep), make_float2(e0, e1));
      } else {
        __stcs(ep, e0);
      }
    }
    if (prm.proj != nullptr && live) {
      float* pp = prm.proj + koff + (int64_t)k * prm.k_sV;
      if (PTS == 2) {
        __stcs(reinterpret_cast<float4*>(pp), make_float4(u[0][k] + du[0][k], v[0][k] + dv[0][k],
                                                          u[PTS - 1][k] + du[PTS - 1][k], v[PTS - 1][k] + dv[PTS - 1][k]));
      } else {
        __stcs(reinterpret_cast<float2*>(pp), make_float2(u[0][k] + du[0][k], v[0][k] + dv[0][k]));
      }
    }
  }
  if (prm.status != nullptr && live) {
#pragma unroll
    for (int p = 0; p < PTS; ++p) prm.status[i0 + p] = st[p];
  }

  // ---- X: (N,3) f32.  Stage the warp's 32*PTS*3 floats, then 128-bit stores.
  float* sx = sX[warp];
#pragma unroll
  for (int p = 0; p < PTS; ++p) {
#pragma unroll
    for (int k = 0; k < 3; ++k) sx[(lane * PTS + p) * 3 + k] = X[p][k];
  }
  __syncwarp();
  const int64_t remain = N - warp_first;  // points of this warp that exist (may be <= 0)
  if (remain > 0) {
    const int npts = remain < 32 * PTS ? (int)remain : 32 * PTS;
    const int nfl = npts * 3;
    float* gx = prm.X + warp_first * 3;
    if (prm.x_vec) {  // warp_first*3 floats is a multiple of 96 floats -> 16-byte aligned
      const int nvec = nfl >> 2;
      for (int q = lane; q < nvec; q += 32)
        __stcs(reinterpret_cast<float4*>(gx) + q, reinterpret_cast<const float4*>(sx)[q]);
      for (int q = (nvec << 2) + lane; q < nfl; q += 32) gx[q] = sx[q];
    } else {
      for (int q = lane; q < nfl; q += 32) gx[q] = sx[q];
    }
  }
}

template <int V, int PTS, bool CONF, int DIST, uint32_t SOLVER = kSolverSecular>
static cudaError_t launch(const TriParams<V>& prm, cudaStream_t stream) {
  const int64_t per_block = (int64_t)kBlock * PTS;
  const int64_t blocks = (prm.N + per_block - 1) / per_block;
  tri_kernel<V, PTS, CONF, DIST, SOLVER><<<(unsigned)blocks, kBlock, 0, stream>>>(prm);
  return cudaGetLastError();
}

template <int V>
static int dispatch(const TriArgs& a) {
  TriParams<V> prm;
  bool dist = false;
  for (int v = 0; v < V; ++v) {
    bool d = false;
    const char* why = "";
    const int rc = prep_camera(a.cams[v], a.centre, (a.flags & SKA_PINHOLE_REPROJ) != 0, prm.cam[v], prm.P64[v], d, &why);
    if (rc != SKA_OK) return set_error(rc, why);
    dist = dist || d;
  }
  for (int k = 0; k < 3; ++k) prm.c[k] = (float)a.centre[k];
  prm.solver = a.flags & SKA_SOLVER_MASK;
  prm.weight_sqrt = (a.flags & SKA_WEIGHT_SQRT) ? 1u : 0u;
  prm.J = a.J;
  prm.N = a.T * (int64_t)a.J;
  const bool fm = (a.layout == SKA_LAYOUT_FRAME_MAJOR);
  prm.frame_major = fm ? 1u : 0u;
  if (fm) {
    prm.k_sV = 2 * (int64_t)a.J;
    prm.k_sT = 2 * (int64_t)a.J * V;
    prm.c_sV = a.J;
    prm.c_sT = (int64_t)a.J * V;
  } else {
    prm.k_sV = 2 * prm.N;
    prm.k_sT = 2 * (int64_t)a.J;
    prm.c_sV = prm.N;
    prm.c_sT = a.J;
  }
  prm.kpts = a.kpts;
  prm.conf = a.conf;
  prm.X = a.X;
  prm.err = a.err;
  prm.proj = a.proj;
  prm.status = a.status;
  auto al = [](const void* p, uintptr_t n) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) % n) == 0; };
  prm.x_vec = al(a.X, 16) ? 1u : 0u;
  const bool conf = (a.conf != nullptr);
  // 128-bit path: flat point index, even point count, vector-aligned streams, modest register need
  const bool pts2 = !fm && (V <= 4) && (prm.N % 2 == 0) && al(a.kpts, 16) && al(a.conf, 8) && al(a.err, 8) && al(a.proj, 16);
  cudaError_t ce;
  cudaStream_t s = (cudaStream_t)a.stream;
#define SKA_GO(PTS)                                                                      \
  (conf ? (dist ? launch<V, PTS, true, 1>(prm, s) : launch<V, PTS, true, 0>(prm, s))     \
        : (dist ? launch<V, PTS, false, 1>(prm, s) : launch<V, PTS, false, 0>(prm, s)))
  if (prm.solver == kSolverJacobi64) {
    // exact / measurement solvers: one generic instantiation (weights and distortion always on)
    ce = launch<V, 1, true, 1, kSolverJacobi64>(prm, s);
  } else if (prm.solver == kSolverJacobi32) {
    ce = launch<V, 1, true, 1, kSolverJacobi32>(prm, s);
  } else if constexpr (V <= 4) {
    ce = pts2 ? SKA_GO(2) : SKA_GO(1);
  } else {
    (void)pts2;
    ce = SKA_GO(1);
  }
#undef SKA_GO
  if (ce != cudaSuccess) return set_error((int)ce, cudaGetErrorString(ce));
  return SKA_OK;
}

}  // namespace ska
