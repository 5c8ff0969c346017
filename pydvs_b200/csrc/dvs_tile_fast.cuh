// dvs_tile_fast.cuh -- K1 fast path: static threshold T >= 2, update_reference_rate, history_weight == 1.
//
// Same contract and bit-exact results as the generic tile pass (dvs_tile_pass.cuh); the difference is
// instruction count.  When every frame / reference pixel of a tile is in [0, 255] (a block-wide vote;
// otherwise the tile falls back to the generic body) nothing can wrap in 16 bits, so
//   * |frame - ref| and the threshold test run on packed u16x2 lanes with DPX min/max
//     (VIMNMX.U16x2): 5 instructions per 2 pixels, no unpacking;
//   * a group of 8 pixels without a spike needs no further work at all: with history_weight == 1 its
//     reference is unchanged (clip(r + 0, 0, 255) == r), so it is not even written back;
//   * a tile without a spike skips the compaction (one block-wide vote) and only zeroes its counters.
// Scalar per-pixel work (sign, reference update, record emission) is done for spiking pixels only.
#pragma once
#include "dvs_tile_pass.cuh"

namespace dvs {

template <int SRC, int J, int KI>
__device__ __noinline__ void tile_body_cold(const TileArgs &A) {
  tile_body<SRC, J, false, KI>(A);
}

DEV uint4 ld16_stream(const void *p) { return __ldcs(reinterpret_cast<const uint4 *>(p)); }

template <int SRC, int J, int KI>
__global__ void __launch_bounds__(256) k_tile_fast(const __grid_constant__ TileArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int16_t *a_sm = reinterpret_cast<int16_t *>(smem_raw);
  const int a_bytes = (KI != INH_NONE) ? ((A.tile_px * 2 + 15) & ~15) : 0;
  uint32_t *wtot = reinterpret_cast<uint32_t *>(smem_raw + a_bytes);
  uint32_t *hist = wtot + 128;
  const int nh = A.enc.mode == DVS_ENC_NONE ? 0 : (A.enc.n_slots + 1);
  uint32_t *misc = hist + 2 * nh;

  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  const int s_idx = tile / A.tiles_per_stream, t_idx = tile - s_idx * A.tiles_per_stream;
  const int row0 = t_idx * A.tile_rows;
  const int rows = min(A.tile_rows, A.H - row0);
  const int npx = rows * A.W;  // multiple of 8 (host-checked)
  const size_t plane_off = (size_t)s_idx * A.H * A.W + (size_t)row0 * A.W;

  // ---- loads: all 2 * J 128-bit requests of the thread are in flight before first use ----------
  uint32_t c_pk[J][4], r_pk[J][4];
  bool have[J];
  uint32_t rng = 0;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int q0 = (j * (int)blockDim.x + tid) * 8;
    have[j] = q0 < npx;
    if (have[j]) {
      if (SRC == SRC_U8) {
        const uint2 v = __ldcs(reinterpret_cast<const uint2 *>(static_cast<const uint8_t *>(A.curr) + plane_off + q0));
        c_pk[j][0] = __byte_perm(v.x, 0, 0x4140); c_pk[j][1] = __byte_perm(v.x, 0, 0x4342);
        c_pk[j][2] = __byte_perm(v.y, 0, 0x4140); c_pk[j][3] = __byte_perm(v.y, 0, 0x4342);
      } else {
        const uint4 v = ld16_stream(static_cast<const int16_t *>(A.curr) + plane_off + q0);
        c_pk[j][0] = v.x; c_pk[j][1] = v.y; c_pk[j][2] = v.z; c_pk[j][3] = v.w;
      }
      const uint4 r = ld16(A.ref + plane_off + q0);
      r_pk[j][0] = r.x; r_pk[j][1] = r.y; r_pk[j][2] = r.z; r_pk[j][3] = r.w;
    }
  }
#pragma unroll
  for (int j = 0; j < J; ++j) {
    if (have[j]) {
      rng |= r_pk[j][0] | r_pk[j][1] | r_pk[j][2] | r_pk[j][3];
      if (SRC != SRC_U8) rng |= c_pk[j][0] | c_pk[j][1] | c_pk[j][2] | c_pk[j][3];
    }
  }
  if (__syncthreads_or((rng & 0xFF00FF00u) != 0)) {  // some pixel outside [0, 255]: exact generic path
    tile_body_cold<SRC, J, KI>(A);
    return;
  }

  for (int i = tid; i < 2 * nh; i += blockDim.x) hist[i] = 0;
  if (tid < 8) misc[tid] = 0;

  // ---- packed threshold: a = |c - r| = max - min per u16 lane; spike iff a > T -------------------
  const int T = A.thr_scalar;                                  // 2 <= T <= 32767 (host-checked)
  const uint32_t kadd = (uint32_t)(0x7fff - T) * 0x00010001u;  // bit 15 of (a + kadd) <=> a > T
  uint32_t a_pk[J][4], sgn[J];
  int any = 0;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    sgn[j] = 0;
    if (!have[j]) continue;
    const int q0 = (j * (int)blockDim.x + tid) * 8;
    uint32_t fl[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const uint32_t mx = __vmaxu2(c_pk[j][w], r_pk[j][w]), mn = __vminu2(c_pk[j][w], r_pk[j][w]);
      const uint32_t a = mx - mn;  // no borrow between lanes: max >= min lane-wise
      fl[w] = (a + kadd) & 0x80008000u;
      a_pk[j][w] = a & ((fl[w] >> 15) * 0xffffu);  // abs_diff * spikes
    }
    if (fl[0] | fl[1] | fl[2] | fl[3]) {
      any = 1;
#pragma unroll
      for (int p = 0; p < 8; ++p)
        if ((fl[p >> 1] >> (15 + 16 * (p & 1))) & 1u)
          sgn[j] |= (get16(c_pk[j], p) > get16(r_pk[j], p) ? 1u : 2u) << (2 * p);
    }
    if (KI != INH_NONE) st16(a_sm + q0, a_pk[j]);
  }
  if (!__syncthreads_or(any)) {  // quiet tile: no records, no events, reference untouched
    for (int c = tid; c < A.C; c += blockDim.x)
      A.tile_counts[((size_t)s_idx * A.C + c) * A.tiles_per_stream + t_idx] = 0;
    return;
  }

  // ---- local inhibition (k == 2), gs:177-221: only groups that still hold a spike ----------------
  if (KI == INH_K2) {
#pragma unroll
    for (int j = 0; j < J; ++j) {
      if (sgn[j] == 0) continue;
      const int q0 = (j * (int)blockDim.x + tid) * 8;
      const int rb = (q0 / A.W) & 1;
      const uint4 o = ld16(a_sm + (rb ? q0 - A.W : q0 + A.W));
      const uint32_t o_pk[4] = {o.x, o.y, o.z, o.w};
      uint32_t keep = 0;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int m0 = get16(a_pk[j], 2 * g), m1 = get16(a_pk[j], 2 * g + 1);
        const int p0 = get16(o_pk, 2 * g), p1 = get16(o_pk, 2 * g + 1);
        const int t00 = rb ? p0 : m0, t01 = rb ? p1 : m1, t10 = rb ? m0 : p0, t11 = rb ? m1 : p1;
        int best = t00, bi = 0;  // row-major scan with strict '>' (argmax gs:125-131)
        if (t01 > best) { best = t01; bi = 1; }
        if (t10 > best) { best = t10; bi = 2; }
        if (t11 > best) { best = t11; bi = 3; }
        if (bi == rb * 2) keep |= 3u << (4 * g);
        if (bi == rb * 2 + 1) keep |= 3u << (4 * g + 2);
      }
      sgn[j] &= keep;
    }
  }

  // ---- reference update (update_reference_rate gs:244-249) for spiking pixels, list membership ----
  uint32_t posm[J], negm[J];
  const unsigned magic = A.upd.div_thr.magic;
  const int max_n = A.upd.max_time_ms;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    posm[j] = 0; negm[j] = 0;
    if (sgn[j] == 0) continue;
    const int q0 = (j * (int)blockDim.x + tid) * 8;
    uint32_t rn_pk[4] = {r_pk[j][0], r_pk[j][1], r_pk[j][2], r_pk[j][3]};
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const int s = sgn_get(sgn[j], p);
      if (s != 0) {
        const int a = get16(a_pk[j], p), r = get16(r_pk[j], p);
        const int upd = min((int)__umulhi((unsigned)a, magic), max_n) * T;  // n * threshold <= 255
        set16(rn_pk, p, s > 0 ? min(r + upd, 255) : max(r - upd, 0));
        bool tp, tn;
        split_px(s, A.polarity, A.uncoded, tp, tn);
        if (tp) posm[j] |= 1u << p;
        if (tn) negm[j] |= 1u << p;
      }
    }
    st16(A.ref + plane_off + q0, rn_pk);
  }
  tile_compact<J>(A, tile, s_idx, t_idx, row0, posm, negm, a_pk, wtot, hist, misc, nh, 0u);
}

template <int SRC, int J, int KI>
static int launch_fast_shape(const TileArgs &A, int bd, size_t smem, cudaStream_t st) {
  auto k = k_tile_fast<SRC, J, KI>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
  }
  k<<<(unsigned)((long long)A.S * A.tiles_per_stream), bd, smem, st>>>(A);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
  return DVS_OK;
}

// Caller guarantees: vec_ok, W % 8 == 0, <= 1024 groups per tile, inh_k in {1, 2}, 2 <= T, rate update,
// history_weight == 1, max_time_ms >= 0, no debug planes.
template <int SRC>
static int dispatch_fast(const TileArgs &A, cudaStream_t st) {
  const int items = A.tile_px / 8;
  const int nh = A.enc.mode == DVS_ENC_NONE ? 0 : (A.enc.n_slots + 1);
  const bool k2 = A.inh_k == 2;
  const size_t smem = (k2 ? ((A.tile_px * 2 + 15) & ~15) : 0) + (128 + 2 * nh + 8) * sizeof(uint32_t);
  auto round32 = [](int x) { return (x + 31) / 32 * 32; };
  if (items <= 256) {
    const int bd = round32(items);
    return k2 ? launch_fast_shape<SRC, 1, INH_K2>(A, bd, smem, st) : launch_fast_shape<SRC, 1, INH_NONE>(A, bd, smem, st);
  }
  const int bd = round32((items + 3) / 4);
  return k2 ? launch_fast_shape<SRC, 4, INH_K2>(A, bd, smem, st) : launch_fast_shape<SRC, 4, INH_NONE>(A, bd, smem, st);
}

}  // namespace dvs
