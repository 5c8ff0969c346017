// K1 instantiations for callers that hold spikes/abs_diff planes (split_spikes), see dvs_tile_pass.cuh
#include "dvs_tile_pass.cuh"
namespace dvs {
int launch_tile_pass_planes(const TileArgs &A, cudaStream_t st) {
  return launch_tile_pass<SRC_PLANES, false, INH_NONE>(A, st);
}
}  // namespace dvs
