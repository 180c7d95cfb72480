// ska_internal.h - shared between the translation units of libska.so (not installed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ska.h"
#include "ska_prep.h"

namespace ska {

int set_error(int code, const char* msg);

struct TriArgs {
  const SkaCamera* cams;
  int32_t V;
  double centre[3];
  const double* Rt_frames;
  const float* kpts;
  const float* conf;
  int64_t T;
  int32_t J;
  int32_t layout;
  uint32_t flags;
  float* X;
  float* err;
  float* proj;
  uint8_t* status;
  void* stream;
};

int triangulate_dispatch(const TriArgs& a);

// bundle adjustment (ska_ba.cu / ska_ba_wide.cu)
int ba_red_size(int C);
int ba_max_grid();
int ba_sum(const float* x, int64_t n, double* out, void* workspace, size_t ws_bytes, void* stream);
int ba_linearize(const SkaBaProblem& p, cudaStream_t s);
int ba_linearize_wide(const SkaBaProblem& p, cudaStream_t s);
int ba_backsub(const SkaBaProblem& p, cudaStream_t s);
int ba_solve(int C, uint64_t free_mask, const double* red, double* cams, double* ctrl, double* delta, void* stream);
int ba_control(int C, const double* red, const double* red2, double* cams, double* ctrl, double* hist, void* stream);
int launch_reduce(const double* partials, int rows, int ncol, double* out, cudaStream_t s);

}  // namespace ska
