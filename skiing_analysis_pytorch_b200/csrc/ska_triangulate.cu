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
