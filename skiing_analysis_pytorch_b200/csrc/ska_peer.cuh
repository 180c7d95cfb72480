// ska_peer.cuh - the peer-memory exchange as a device function, so that a single-CTA consumer kernel (the reduced-system
// solve, the LM controller) performs its own all-reduce in its prologue: push / release flag / poll / rank-ordered sum
// and the solve are ONE kernel (protocol and buffers: ska_peer.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ska_internal.h"

namespace ska {

struct PeerDev {
  int world, rank, slot;
  uint32_t max_polls;
  double* recv[SKA_MAX_PEERS];     // recv[r]: rank r's receive area [2][world][slot]
  uint64_t* flags[SKA_MAX_PEERS];  // flags[r]: rank r's arrival flags [world]
  uint64_t* state;                 // local: [0] exchange counter, [1] first exchange that timed out
  const double* skip;              // nullable device flag, identical on every rank: non-zero = skip
};

int peer_fill(const SkaPeerComm& c, PeerDev& a);  // validates; world == 0 in `a` means "no exchange"

__device__ __forceinline__ void peer_st_release_sys(uint64_t* p, uint64_t v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t peer_ld_acquire_sys(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Called by EVERY thread of a single CTA (any block size >= world).  in / out may alias for the all-reduce.  Ends with a
// CTA barrier: afterwards every thread sees `out`.
__device__ inline void peer_exchange_block(const PeerDev& a, const double* in, int n, double* out, int gather) {
  __shared__ int s_fail;
  if (a.skip != nullptr && *a.skip != 0.0) return;  // every rank holds the same flag: all skip, the counters stay in step
  const int tid = threadIdx.x, nt = blockDim.x;
  const uint64_t epoch = a.state[0] + 1;  // this exchange's number (1, 2, ...)
  const int par = (int)(epoch & 1);
  if (tid == 0) s_fail = 0;
  // push my payload into slot `rank` of every rank's receive area (my own included: one code path, one summation order)
  for (int r = 0; r < a.world; ++r) {
    double* dst = a.recv[r] + ((size_t)par * a.world + a.rank) * a.slot;
    for (int i = tid; i < n; i += nt) dst[i] = in[i];
  }
  __threadfence_system();
  __syncthreads();
  if (tid < a.world) peer_st_release_sys(a.flags[tid] + a.rank, epoch);
  if (tid < a.world) {  // wait for the world's payloads: local polling
    const uint64_t* f = a.flags[a.rank] + tid;
    uint32_t polls = 0;
    while (peer_ld_acquire_sys(f) < epoch) {
      if (++polls > a.max_polls) {
        s_fail = 1;
        break;
      }
    }
  }
  __syncthreads();
  const double* mine = a.recv[a.rank] + (size_t)par * a.world * a.slot;
  if (gather) {
    for (int i = tid; i < a.world * n; i += nt) out[i] = mine[(size_t)(i / n) * a.slot + (i % n)];
  } else {
    for (int i = tid; i < n; i += nt) {
      double s = 0.0;
      for (int r = 0; r < a.world; ++r) s += mine[(size_t)r * a.slot + i];
      out[i] = s;
    }
  }
  __syncthreads();
  if (tid == 0) {
    a.state[0] = epoch;
    if (s_fail && a.state[1] == 0) a.state[1] = epoch;
  }
  __syncthreads();
}

}  // namespace ska
