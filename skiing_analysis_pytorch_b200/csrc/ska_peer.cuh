// ska_peer.cuh - the peer-memory exchange as a device function, so that a single-CTA consumer kernel (the reduced-system
// solve, the LM controller) performs its own all-reduce in its prologue: push / poll / rank-ordered sum and the solve are
// ONE kernel (buffers: ska_peer.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ska_internal.h"

namespace ska {

struct PeerDev {
  int world, rank, slot;
  uint32_t max_polls;
  uint64_t* recv[SKA_MAX_PEERS];   // recv[r]: rank r's receive area [2][world][slot][2] 64-bit words (see below)
  uint64_t* state;                 // local: [0] exchange counter, [1] first exchange that timed out
  const double* skip;              // nullable device flag, identical on every rank: non-zero = skip
};

int peer_fill(const SkaPeerComm& c, PeerDev& a);  // validates; world == 0 in `a` means "no exchange"

// Every double travels as TWO 64-bit words {low half | number of the exchange} and {high half | number of the exchange}: an
// aligned 8-byte store is delivered whole, so the word IS its own arrival flag - the receiver spins on the words it is
// about to sum until they carry this exchange's number.  No fence, no separate flag round trip (one NVLink flight
// instead of payload -> __threadfence_system -> flag), at twice the (tiny) payload.
__device__ __forceinline__ void peer_st_word(uint64_t* p, uint64_t v) { asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ uint64_t peer_ld_word(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double peer_join(uint64_t lo, uint64_t hi) { return __longlong_as_double((long long)((hi << 32) | (lo & 0xffffffffull))); }
// one double: both words are loaded together (independent loads), re-read until both carry this exchange's number
__device__ __forceinline__ double peer_wait_double(const uint64_t* w, uint32_t tag, uint32_t max_polls, int& fail) {
  uint64_t lo, hi;
  uint32_t polls = 0;
  for (;;) {
    lo = peer_ld_word(w);
    hi = peer_ld_word(w + 1);
    if ((uint32_t)(lo >> 32) == tag && (uint32_t)(hi >> 32) == tag) break;
    if (++polls > max_polls) {
      fail = 1;
      break;
    }
  }
  return peer_join(lo, hi);
}
// element i of every rank: the 2 x world words are loaded in ONE batch of independent loads per attempt (polling them one
// after the other costs a local-L2 round trip each: 16 dependent ones at 8 ranks), then summed in rank order
__device__ __forceinline__ double peer_wait_sum(const uint64_t* mine, int world, int slot, int i, uint32_t tag, uint32_t max_polls, int& fail) {
  uint64_t w[2 * SKA_MAX_PEERS];
  uint32_t polls = 0;
  for (;;) {
    bool ok = true;
#pragma unroll
    for (int r = 0; r < SKA_MAX_PEERS; ++r)
      if (r < world) {
        const uint64_t* q = mine + ((size_t)r * slot + i) * 2;
        w[2 * r] = peer_ld_word(q);
        w[2 * r + 1] = peer_ld_word(q + 1);
      }
#pragma unroll
    for (int r = 0; r < SKA_MAX_PEERS; ++r)
      if (r < world) ok = ok && (uint32_t)(w[2 * r] >> 32) == tag && (uint32_t)(w[2 * r + 1] >> 32) == tag;
    if (ok) break;
    if (++polls > max_polls) {
      fail = 1;
      break;
    }
  }
  double s = 0.0;
#pragma unroll
  for (int r = 0; r < SKA_MAX_PEERS; ++r)
    if (r < world) s += peer_join(w[2 * r], w[2 * r + 1]);
  return s;
}

// Called by EVERY thread of a single CTA.  in / out may alias for the all-reduce.  Ends with a CTA barrier: afterwards every
// thread sees `out`.  Areas alternate with the parity of the exchange number: overwriting parity k+1 of a peer is safe once
// its words of exchange k have been seen - it finished reading exchange k-1 (the same parity) before it pushed k.
__device__ inline void peer_exchange_block(const PeerDev& a, const double* in, int n, double* out, int gather) {
  __shared__ int s_fail;
  if (a.skip != nullptr && *a.skip != 0.0) return;  // every rank holds the same flag: all skip, the counters stay in step
  const int tid = threadIdx.x, nt = blockDim.x;
  const uint64_t epoch = a.state[0] + 1;  // this exchange's number (1, 2, ...)
  const uint32_t tag = (uint32_t)epoch;
  const int par = (int)(epoch & 1);
  if (tid == 0) s_fail = 0;
  __syncthreads();
  // push my payload into slot `rank` of every rank's receive area (my own included: one code path, one summation order)
  for (int i = tid; i < n; i += nt) {
    const uint64_t bits = (uint64_t)__double_as_longlong(in[i]);
    const uint64_t lo = (bits & 0xffffffffull) | ((uint64_t)tag << 32), hi = (bits >> 32) | ((uint64_t)tag << 32);
    for (int r = 0; r < a.world; ++r) {
      uint64_t* dst = a.recv[r] + (((size_t)par * a.world + a.rank) * a.slot + i) * 2;
      peer_st_word(dst, lo);
      peer_st_word(dst + 1, hi);
    }
  }
  // collect: every word is waited for where it is used
  int fail = 0;
  const uint64_t* mine = a.recv[a.rank] + (size_t)par * a.world * a.slot * 2;
  if (gather) {
    for (int i = tid; i < a.world * n; i += nt) out[i] = peer_wait_double(mine + ((size_t)(i / n) * a.slot + (i % n)) * 2, tag, a.max_polls, fail);
  } else {
    for (int i = tid; i < n; i += nt) out[i] = peer_wait_sum(mine, a.world, a.slot, i, tag, a.max_polls, fail);
  }
  if (fail) s_fail = 1;
  __syncthreads();
  if (tid == 0) {
    a.state[0] = epoch;
    if (s_fail && a.state[1] == 0) a.state[1] = epoch;
  }
  __syncthreads();
}

}  // namespace ska
