// One translation unit per view count so the instantiations compile in parallel.
#include "ska_triangulate_impl.cuh"
namespace ska {
int tri_dispatch_v6(const TriArgs& a) { return dispatch<6>(a); }
size_t tri_frames_ws_v6(int64_t T) { return frames_ws_bytes<6>(T); }
}  // namespace ska
