// K1 fast-path instantiations, frame dtype SRC_I16 (see dvs_tile_fast.cuh)
#include "dvs_tile_ranked.cuh"
namespace dvs {
int launch_tile_fast_i16(const TileArgs &A, cudaStream_t st) { return dispatch_fast<SRC_I16>(A, st); }
}  // namespace dvs
