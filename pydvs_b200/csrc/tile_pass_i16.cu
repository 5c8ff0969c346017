// K1 instantiations: frame dtype SRC_I16, adaptive threshold false (see dvs_tile_pass.cuh)
#include "dvs_tile_pass.cuh"
namespace dvs {
int launch_tile_pass_i16(const TileArgs &A, cudaStream_t st) { return dispatch_inh<SRC_I16, false>(A, st); }
}  // namespace dvs
