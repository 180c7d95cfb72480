// ska_internal.h - shared between the translation units of libska.so (not installed).
#pragma once
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

}  // namespace ska
