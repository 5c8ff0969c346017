// K1 fast-path instantiations, frame dtype SRC_U8 (see dvs_tile_fast.cuh)
#include "dvs_tile_ranked.cuh"
namespace dvs {
int launch_tile_fast_u8(const TileArgs &A, cudaStream_t st) { return dispatch_fast<SRC_U8>(A, st); }
}  // namespace dvs
