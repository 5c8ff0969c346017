// K1 fast-path instantiations, frame dtype SRC_U8, one-byte state planes ST_U8 (see dvs_tile_fast.cuh)
#include "dvs_tile_ranked.cuh"
namespace dvs {
int launch_tile_fast_u8_s8(const TileArgs &A, cudaStream_t st) { return dispatch_fast<SRC_U8, ST_U8>(A, st); }
}  // namespace dvs
