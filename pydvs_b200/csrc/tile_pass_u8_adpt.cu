// K1 instantiations: frame dtype SRC_U8, adaptive threshold true (see dvs_tile_pass.cuh)
#include "dvs_tile_pass.cuh"
namespace dvs {
int launch_tile_pass_u8_adpt(const TileArgs &A, cudaStream_t st) { return dispatch_inh<SRC_U8, true>(A, st); }
}  // namespace dvs
