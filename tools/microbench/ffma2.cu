// Microbenchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2, sm_100) issue/throughput on B200.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters, float s) {
  float a[8], b[8];
  float2 p[8];
  for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i; b[i] = a[i] * 0.5f; p[i] = make_float2(a[i], b[i]); }
  const float2 s2 = make_float2(s, s * 0.999f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) { a[i] = fmaf(a[i], s, b[i]); b[i] = fmaf(b[i], s, a[i]); }
        else { p[i] = __ffma2_rn(p[i], s2, p[(i + 1) & 7]); }
      }
    }
  }
  float acc = 0.f;
  for (int i = 0; i < 8; ++i) acc += (MODE == 0) ? (a[i] + b[i]) : (p[i].x + p[i].y);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
double run(int iters) {
  float* d; cudaMalloc(&d, 148 * 8 * 256 * sizeof(float));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 8, 256>>>(d, 10, 1.0001f);
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 256>>>(d, iters, 1.0001f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaFree(d);
  // FMAs: MODE0: 2 per inner op; MODE1: 2 per packed op
  const double fmas = (double)148 * 8 * 256 * iters * 64 * 2;
  return fmas / (ms * 1e-3);
}

int main() {
  const int iters = 20000;
  const double r0 = run<0>(iters), r1 = run<1>(iters);
  printf("FFMA : %.2f TFMA/s (%.2f TFLOP/s)\n", r0 / 1e12, 2 * r0 / 1e12);
  printf("FFMA2: %.2f TFMA/s (%.2f TFLOP/s)\n", r1 / 1e12, 2 * r1 / 1e12);
  return 0;
}
