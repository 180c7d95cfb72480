/* A plain C translation unit against include/ska.h: the boundary is C (no C++ types, no torch), so a maintainer can bind it
 * from any FFI.  Built and run by tests/test_abi.py::test_header_is_plain_c_and_links (no GPU needed: only the entry
 * points that validate their arguments before touching CUDA are called). */
#include <stdio.h>
#include <string.h>

#include "ska.h"

int main(void) {
  if (ska_abi_version() != SKA_ABI_VERSION) {
    printf("abi %d != header %d\n", ska_abi_version(), SKA_ABI_VERSION);
    return 1;
  }
  if (strcmp(ska_build_arch(), "sm_100a") != 0) return 2;
  SkaCamera cams[2];
  memset(cams, 0, sizeof cams);
  /* V = 1 is not a rig: argument error, reported before any CUDA call */
  if (ska_triangulate_reproject_f32(cams, 1, NULL, NULL, NULL, 1, 17, SKA_LAYOUT_VIEW_MAJOR, 0, NULL, NULL, NULL, NULL, NULL) != SKA_EINVAL) return 3;
  if (strlen(ska_last_error()) == 0) return 4;
  SkaBaRegProblem reg;
  memset(&reg, 0, sizeof reg);
  if (ska_ba_reg_linearize_f64(&reg, NULL) != SKA_EINVAL) return 5;
  SkaPeerComm comm;
  memset(&comm, 0, sizeof comm);
  if (ska_peer_allreduce_f64(&comm, NULL, 0, NULL) != SKA_EINVAL) return 6;
  if (ska_peer_region_bytes(2, 16) != 2u * 2u * 16u * 16u) return 7;
  printf("ok abi %d\n", ska_abi_version());
  return 0;
}
