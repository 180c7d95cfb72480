// ska_peer.cu - the exchange step of the sharded solvers over NVLink peer memory, sm_100a.
//
// The only data-path exchange of this library is small and latency-bound: per LM trial the packed reduced camera system
// (158 .. 1 179 doubles) and 4 trial scalars (csrc/ska_ba.cu), per CG iteration of the regularised LM two dot products
// and a one-frame halo (csrc/ska_ba_reg.cu).  An NCCL collective costs ~20-35 us for such a payload on 8 GPUs - as much
// as all the kernels of a config-3 trial together.  Here every rank PUSHES its payload straight into every peer's
// receive area (plain 8-byte stores to peer-mapped memory: NVLink / NVSwitch), every 64-bit word carrying half a double
// and the number of the exchange, so that a word is its own arrival flag (ska_peer.cuh); the receiver spins on its OWN
// local words and sums / copies the world's payloads in fixed rank order - every rank gets bit-identical results and
// takes identical accept / reject decisions.  One single CTA per exchange, no host involvement, capturable in the trial's
// CUDA graph; the single-CTA consumer kernels (solve, control, CG scalars) run the exchange as their own prologue.
//
// Buffers: receive areas are double-buffered by the parity of the exchange counter.  Every rank runs the same sequence of
// exchanges, so the counters agree without being communicated.  A word that never arrives (a dead peer) makes the kernel
// give up after 2^poll_limit_log2 polls (default 2^24, seconds) and raise an error word the host reads after the solve - it
// never hangs the GPU.
//
// Memory comes from cudaMalloc (exportable with cudaIpcGetMemHandle; one process per GPU) - ska_peer_alloc / export /
// import below; the host side (peer.py) exchanges the 64-byte handles through torch.distributed once at set-up.
#include <cuda_runtime.h>
#include <stdint.h>

#include "ska_internal.h"
#include "ska_peer.cuh"

namespace ska {

int peer_fill(const SkaPeerComm& c, PeerDev& a) {
  if (c.world < 1 || c.world > SKA_MAX_PEERS || c.rank < 0 || c.rank >= c.world) return set_error(SKA_EINVAL, "bad world / rank");
  if (c.slot_doubles < 1 || c.d_state == nullptr) return set_error(SKA_EINVAL, "slot_doubles >= 1 and d_state are required");
  if (c.poll_limit_log2 < 0 || c.poll_limit_log2 > 31) return set_error(SKA_EINVAL, "poll_limit_log2 must be in 0..31");
  a.world = c.world, a.rank = c.rank, a.slot = c.slot_doubles;
  a.max_polls = 1u << (c.poll_limit_log2 ? c.poll_limit_log2 : 24);
  for (int r = 0; r < SKA_MAX_PEERS; ++r) {
    a.recv[r] = r < c.world ? reinterpret_cast<uint64_t*>(c.recv[r]) : nullptr;
    if (r < c.world && a.recv[r] == nullptr) return set_error(SKA_EINVAL, "null peer pointer");
  }
  a.state = c.d_state;
  a.skip = c.d_skip;
  return SKA_OK;
}

namespace {

constexpr int kPeerThreads = 256;

__global__ void __launch_bounds__(kPeerThreads) peer_exchange_kernel(const PeerDev a, const double* in, int n, double* out, int gather) {
  peer_exchange_block(a, in, n, out, gather);
}

int launch_exchange(const SkaPeerComm& c, const double* in, int n, double* out, int gather, cudaStream_t s) {
  PeerDev a;
  const int rc = peer_fill(c, a);
  if (rc != SKA_OK) return rc;
  if (n < 0 || n > c.slot_doubles) return set_error(SKA_EINVAL, "payload larger than the peer slot");
  if (in == nullptr || out == nullptr) return set_error(SKA_EINVAL, "null device pointer");
  peer_exchange_kernel<<<1, kPeerThreads, 0, s>>>(a, in, n, out, gather);
  const cudaError_t ce = cudaGetLastError();
  return ce == cudaSuccess ? SKA_OK : set_error((int)ce, cudaGetErrorString(ce));
}

}  // namespace

int peer_allreduce(const SkaPeerComm& c, double* buf, int n, cudaStream_t s) { return launch_exchange(c, buf, n, buf, 0, s); }
int peer_allgather(const SkaPeerComm& c, const double* in, int n, double* out, cudaStream_t s) { return launch_exchange(c, in, n, out, 1, s); }

int peer_alloc(size_t bytes, void** out) {
  if (out == nullptr || bytes == 0) return set_error(SKA_EINVAL, "bytes > 0 and a result pointer are required");
  void* p = nullptr;
  cudaError_t ce = cudaMalloc(&p, bytes);
  if (ce == cudaSuccess) ce = cudaMemset(p, 0, bytes);
  if (ce != cudaSuccess) return set_error((int)ce, cudaGetErrorString(ce));
  *out = p;
  return SKA_OK;
}
int peer_free(void* p) {
  const cudaError_t ce = cudaFree(p);
  return ce == cudaSuccess ? SKA_OK : set_error((int)ce, cudaGetErrorString(ce));
}
int peer_export(void* p, unsigned char* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  const cudaError_t ce = cudaIpcGetMemHandle(&h, p);
  if (ce != cudaSuccess) return set_error((int)ce, cudaGetErrorString(ce));
  for (int i = 0; i < 64; ++i) handle64[i] = reinterpret_cast<unsigned char*>(&h)[i];
  return SKA_OK;
}
int peer_import(const unsigned char* handle64, void** out) {
  cudaIpcMemHandle_t h;
  for (int i = 0; i < 64; ++i) reinterpret_cast<unsigned char*>(&h)[i] = handle64[i];
  void* p = nullptr;
  const cudaError_t ce = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (ce != cudaSuccess) return set_error((int)ce, cudaGetErrorString(ce));
  *out = p;
  return SKA_OK;
}
int peer_close(void* p) {
  const cudaError_t ce = cudaIpcCloseMemHandle(p);
  return ce == cudaSuccess ? SKA_OK : set_error((int)ce, cudaGetErrorString(ce));
}

}  // namespace ska
