// One translation unit per view count so the instantiations compile in parallel.
#include "ska_triangulate_impl.cuh"
namespace ska {
int tri_dispatch_v7(const TriArgs& a) { return dispatch<7>(a); }
size_t tri_frames_ws_v7(int64_t T) { return frames_ws_bytes<7>(T); }
}  // namespace ska
