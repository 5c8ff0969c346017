// K1 instantiations: frame dtype SRC_I16, adaptive threshold true (see dvs_tile_pass.cuh)
#include "dvs_tile_pass.cuh"
namespace dvs {
int launch_tile_pass_i16_adpt(const TileArgs &A, cudaStream_t st) { return dispatch_inh<SRC_I16, true>(A, st); }
}  // namespace dvs
