// One translation unit per view count so the instantiations compile in parallel.
#include "ska_triangulate_impl.cuh"
namespace ska {
int tri_dispatch_v3(const TriArgs& a) { return dispatch<3>(a); }
}  // namespace ska
