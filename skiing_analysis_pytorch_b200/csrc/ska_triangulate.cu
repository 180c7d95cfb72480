// ska_triangulate.cu - view-count switch for the fused triangulate+reproject kernel
// (instantiations live in ska_tri_v{2..8}.cu, the kernel in ska_triangulate_impl.cuh).
#include "ska_internal.h"

namespace ska {
int tri_dispatch_v2(const TriArgs&);
int tri_dispatch_v3(const TriArgs&);
int tri_dispatch_v4(const TriArgs&);
int tri_dispatch_v5(const TriArgs&);
int tri_dispatch_v6(const TriArgs&);
int tri_dispatch_v7(const TriArgs&);
int tri_dispatch_v8(const TriArgs&);

size_t tri_frames_ws_v2(int64_t);
size_t tri_frames_ws_v3(int64_t);
size_t tri_frames_ws_v4(int64_t);
size_t tri_frames_ws_v5(int64_t);
size_t tri_frames_ws_v6(int64_t);
size_t tri_frames_ws_v7(int64_t);
size_t tri_frames_ws_v8(int64_t);

size_t tri_frames_workspace(int V, int64_t T) {
  switch (V) {
    case 2: return tri_frames_ws_v2(T);
    case 3: return tri_frames_ws_v3(T);
    case 4: return tri_frames_ws_v4(T);
    case 5: return tri_frames_ws_v5(T);
    case 6: return tri_frames_ws_v6(T);
    case 7: return tri_frames_ws_v7(T);
    case 8: return tri_frames_ws_v8(T);
    default: return 0;
  }
}

int triangulate_dispatch(const TriArgs& a) {
  switch (a.V) {
    case 2: return tri_dispatch_v2(a);
    case 3: return tri_dispatch_v3(a);
    case 4: return tri_dispatch_v4(a);
    case 5: return tri_dispatch_v5(a);
    case 6: return tri_dispatch_v6(a);
    case 7: return tri_dispatch_v7(a);
    case 8: return tri_dispatch_v8(a);
    default: return set_error(SKA_EUNSUPPORTED, "V must be in 2..8 for the fused kernel");
  }
}
}  // namespace ska
