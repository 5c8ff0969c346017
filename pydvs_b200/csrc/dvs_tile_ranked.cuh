// dvs_tile_ranked.cuh -- K1 for the headline shape: static threshold, 2x2 inhibition on two-row tiles, coded RATE or
// TIME encoder with MERGED lists, history_weight == 1, int16 or uint8 frames (what dispatch_fast_shapes calls the
// FUSED2 shape).  Same contract and bit-exact results as fast2_tile / the generic tile pass; the difference is that
// every candidate knows its FINAL rank in the tile's pos / neg list before the per-candidate phase starts:
//
//   2. one pass over the 2x2 areas a thread owns (frame + reference words of both rows at hand): threshold, first
//      arg-max through a packed key maximum, and the finished work-list entry of the winner (pixel index, reference,
//      abs_diff, sign, "silent" = a coded-rate spike that falls in no slot and is not recorded);  the thread counts its
//      recorded pos and neg winners per row in 8-bit fields;
//   3. ONE warp scan of those packed counts (a warp holds at most 128 per field) + the per-(row, warp) totals through
//      shared memory give each winner its rank among the tile's pos resp. neg spikes in row-major order, and the tile
//      totals -- the work list is laid out [pos in list order][neg in list order]; a silent winner never reaches the
//      list: its only effect, the reference update, is applied by the owner thread while it stores its entries;
//   4. one candidate per lane: reference update, slot histogram, and the record goes STRAIGHT to its place in the tile's
//      list in global memory (lane i of the pos part writes record i: consecutive lanes, consecutive 8-byte records).
//      No ballots, no ranks, no staging in shared memory, no copy-out pass and one barrier less than fast2_tile, which
//      discovers list membership only inside the per-candidate loop.
//
// Shared memory: [wl: fast_cap u32][wtot: 32 u32][hist: 2 * (n_slots + 1) u32]  (~8.5 KB)
#pragma once
#include "dvs_tile_fast.cuh"

namespace dvs {

#ifndef FAST2R_MINB
#define FAST2R_MINB 5
#endif
#ifndef RANKED_LAZY_WIDEN
#define RANKED_LAZY_WIDEN 1   // one-byte planes stay packed in registers (2 per group instead of 4) and a word is widened where it
                              // is used: 40 registers without spills, i.e. 6 CTAs per SM.  K1 ms at 64 x 4K (gpurun sweep): 0.336
                              // against 0.360 (eager, 5 CTAs), 0.344 (eager, 6 CTAs, 8 B of spills), 0.351 (lazy, 5 CTAs);
                              // uint8 frames 0.325 / 0.355, 2048 x VGA 0.679 / 0.735, dense 4K 0.709 / 0.697
#endif
#ifndef FAST2R_MINB_S8
#define FAST2R_MINB_S8 6      // CTAs per SM asked of ptxas for one-byte state (int16 frames)
#endif
#ifndef FAST2R_MINB_S8U8
#define FAST2R_MINB_S8U8 8    // ... with uint8 frames as well (32 registers, 4 B of spills): 256 x 4K 1.190 ms against 1.254 at 6
#endif
#ifndef RANKED_NARROW_I16
#define RANKED_NARROW_I16 0     // int16 frames narrowed to bytes in registers (32 registers with 20 B of spills, 8 CTAs per SM):
                                // 64 x 4K K1 0.347 ms against 0.336 without
#endif
#ifndef RANKED_BD512
#define RANKED_BD512 0        // rows of 257..512 groups as ONE column block per thread in CTAs of 512 threads (one-byte state)
                              // -- 32 registers, 64 warps per SM, and slower: 64 x 4K K1 0.417 ms against 0.336 (16 warps per barrier)
#endif
#ifndef FAST2R_MINB_512
#define FAST2R_MINB_512 4
#endif
#ifndef RANKED_QUIET_SKIP
#define RANKED_QUIET_SKIP 0   // one early-out per 2 x 8 block ahead of the per-area loop: K1 at 64 x 4K 0.405 ms against 0.336
                              // without (packed-bytes build, 20 B of spills; eager build: 0.393 against 0.364)
#endif

template <int SRC, int JH, int ENC, int ST>
DEV void fast2r_tile(const TileArgs &A, const int t_idx, const int s_idx, unsigned char *smem_raw) {
  static_assert(ENC == ENC_RATE_CODED || ENC == ENC_TIME_CODED, "inlined encoders only");
  constexpr int J = 2 * JH;
  constexpr bool LAZY_R = RANKED_LAZY_WIDEN && ST == ST_U8;                  // reference bytes stay packed
  // int16 frames: once the range vote has seen the raw words, a tile that stays here holds only values in [0, 255], so
  // the frame can be narrowed to bytes in registers as well (2 PRMT per group) -- fewer live registers, more CTAs per SM
  constexpr bool NARROW_C = RANKED_NARROW_I16 && RANKED_LAZY_WIDEN && ST == ST_U8 && SRC == SRC_I16;
  constexpr bool LAZY_C = RANKED_LAZY_WIDEN && ST == ST_U8 && (SRC == SRC_U8 || NARROW_C);  // ... and the frame bytes
  using state_t = typename state_elem<ST>::type;
  uint32_t *wl = reinterpret_cast<uint32_t *>(smem_raw);
  uint32_t *wtot = wl + A.fast_cap;
  uint32_t *hist = wtot + 32;
  const int nh = A.enc.n_slots + 1;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;  // J * NW <= 32
  const int tile = s_idx * A.tiles_per_stream + t_idx;
  const int row0 = t_idx * 2;
  const size_t plane_off = ((size_t)s_idx * A.H + row0) * A.W;
  state_t *const ref_t = static_cast<state_t *>(A.ref) + plane_off;
  const int gpr = A.W >> 3;  // groups of 8 pixels per row
  const unsigned char *const cur_thread =
      static_cast<const unsigned char *>(A.curr) + (plane_off + (size_t)tid * 8) * (SRC == SRC_U8 ? 1 : 2);
  const state_t *const ref_thread = ref_t + tid * 8;

  // ---- 1. loads: group j = row (j / JH), column block (j % JH); all requests in flight before first use -------------
  uint32_t c_pk[J][4], r_pk[J][4];
  bool have[JH];
#pragma unroll
  for (int h = 0; h < JH; ++h) have[h] = h * (int)blockDim.x + tid < gpr;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int off = (j / JH) * A.W + (j % JH) * (int)blockDim.x * 8;
    if (have[j % JH]) {
      ld8_raw<(SRC == SRC_U8 ? 1 : 2), true>(cur_thread + (SRC == SRC_U8 ? 1 : 2) * off, c_pk[j]);
      ld8_raw<(ST == ST_U8 ? 1 : 2), false>(ref_thread + off, r_pk[j]);
    }
  }
  if (tid < 2 * nh) hist[tid] = 0;
  for (int i = blockDim.x + tid; i < 2 * nh; i += blockDim.x) hist[i] = 0;
  // bytes -> 16-bit lanes only after every load has been issued (see ld8_raw)
  uint32_t rng0 = 0;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    if (!have[j % JH]) continue;
    if (NARROW_C) {
      rng0 |= (c_pk[j][0] | c_pk[j][1]) | (c_pk[j][2] | c_pk[j][3]);
      c_pk[j][0] = __byte_perm(c_pk[j][0], c_pk[j][1], 0x6420);
      c_pk[j][1] = __byte_perm(c_pk[j][2], c_pk[j][3], 0x6420);
    } else if (!LAZY_C) widen8<(SRC == SRC_U8 ? 1 : 2)>(c_pk[j]);
    if (!LAZY_R) widen8<(ST == ST_U8 ? 1 : 2)>(r_pk[j]);
  }
  auto r_word = [&](int j, int w) -> uint32_t {
    if (LAZY_R) return __byte_perm(r_pk[j][w >> 1], 0, (w & 1) ? 0x4342 : 0x4140);
    return r_pk[j][w];
  };
  auto c_word = [&](int j, int w) -> uint32_t {
    if (LAZY_C) return __byte_perm(c_pk[j][w >> 1], 0, (w & 1) ? 0x4342 : 0x4140);
    return c_pk[j][w];
  };

  // ---- 2. threshold + 2x2 inhibition + finished entries + class counts ----------------------------------------------
  const int T = A.thr_scalar;                                  // 2 <= T <= 32767 (host-checked)
  const uint32_t kadd = (uint32_t)(0x7fff - T) * 0x00010001u;  // bit 15 of (a + kadd) <=> a > T
  // entry: tile-local pixel index (13 bits) | reference << 13 (8) | abs_diff << 21 (8) | negative << 29 | silent << 30
  uint32_t wle[JH][4];
  uint32_t cnt[JH];  // recorded winners of this thread per column block, 8-bit fields: upper row pos, lower row pos, upper neg, lower neg
  uint32_t sil = 0;  // bit h: a silent winner in column block h
  uint32_t rng = rng0;
#pragma unroll
  for (int h = 0; h < JH; ++h) {
    const int ju = h, jl = JH + h;
    cnt[h] = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) wle[h][w] = 0u;
    if (!have[h]) continue;
    if (ST != ST_U8)  // one-byte state cannot leave [0, 255]
      rng |= ((r_pk[ju][0] | r_pk[ju][1] | r_pk[ju][2]) | r_pk[ju][3]) | ((r_pk[jl][0] | r_pk[jl][1] | r_pk[jl][2]) | r_pk[jl][3]);
    if (SRC != SRC_U8 && !NARROW_C)
      rng |= ((c_pk[ju][0] | c_pk[ju][1] | c_pk[ju][2]) | c_pk[ju][3]) | ((c_pk[jl][0] | c_pk[jl][1] | c_pk[jl][2]) | c_pk[jl][3]);
    const uint32_t qu = (uint32_t)(h * (int)blockDim.x * 8 + tid * 8);
#if RANKED_QUIET_SKIP
    {  // most 2 x 8 blocks of a frame hold no spike at all: one branch for the block instead of one per area
      uint32_t any = 0u;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const uint32_t cu = c_word(ju, w), ru = r_word(ju, w), cl = c_word(jl, w), rl = r_word(jl, w);
        any |= ((__vmaxu2(cu, ru) - __vminu2(cu, ru)) + kadd) | ((__vmaxu2(cl, rl) - __vminu2(cl, rl)) + kadd);
      }
      if ((any & 0x80008000u) == 0u) continue;
    }
#endif
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const uint32_t cu = c_word(ju, w), ru = r_word(ju, w), cl = c_word(jl, w), rl = r_word(jl, w);
      const uint32_t au = __vmaxu2(cu, ru) - __vminu2(cu, ru), al = __vmaxu2(cl, rl) - __vminu2(cl, rl);
      const uint32_t fu = (au + kadd) & 0x80008000u, fl2 = (al + kadd) & 0x80008000u;  // thresholded_difference gs:67-78
      if ((fu | fl2) == 0u) continue;
      // local_inhibition gs:177-221 + argmax gs:118-132: keys abs_diff * 8 + (3 - index) * 2 + (frame < reference); the
      // maximum key is the first maximum of abs_diff * spikes in row-major order and carries its sign.  (PRMT with the
      // sign-replicating selector turns bit 15 / 31 of a flag word into a lane mask in one instruction.)
      const uint32_t mu = au & msb_lane_mask(fu), ml = al & msb_lane_mask(fl2);
      const uint32_t nu = (~((cu | 0x80008000u) - ru) >> 15) & 0x00010001u;  // lane bit: frame < reference
      const uint32_t nl = (~((cl | 0x80008000u) - rl) >> 15) & 0x00010001u;
      const uint32_t m2 = __vmaxu2(mu * 8u + 0x00040006u + nu, ml * 8u + 0x00000002u + nl);
      const uint32_t m = max(m2 & 0xffffu, m2 >> 16);
      const uint32_t idx = 3u - ((m >> 1) & 3u), half = idx & 1u, lower = idx >> 1;
      const uint32_t r_w = lower ? rl : ru;
      const uint32_t a_win = m >> 3, neg = m & 1u;
      const bool silent = a_win >= (uint32_t)A.silent_a;  // coded rate, k == 0, records not wanted: no event, no record
      const uint32_t q = qu + (lower ? (uint32_t)A.W : 0u) + 2u * w + half;
      wle[h][w] = q | (((r_w >> (16u * half)) & 0xffu) << 13) | (a_win << 21) | (neg << 29) | (silent ? 1u << 30 : 0u);
      if (silent) sil |= 1u << h;
      else cnt[h] += 1u << (8u * lower + 16u * neg);
    }
  }

  // ---- 3. ranks: one warp scan of the packed counts, (row, warp) totals through shared memory ------------------------
  const bool w_unsafe = __any_sync(kFull, (rng & 0xFF00FF00u) != 0);
  const bool w_silent = __any_sync(kFull, sil != 0u);
  uint32_t inc[JH];
#pragma unroll
  for (int h = 0; h < JH; ++h) inc[h] = cnt[h];
  const bool w_cand = __any_sync(kFull, (cnt[0] | cnt[JH - 1]) != 0u);
  if (w_cand) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
      for (int h = 0; h < JH; ++h) {
        const uint32_t y = __shfl_up_sync(kFull, inc[h], o);
        if (lane >= o) inc[h] += y;
      }
    }
    // group j = row * JH + h; lane j publishes the warp's (pos | neg << 16) total of that group
    uint32_t tot = __shfl_sync(kFull, inc[0], 31);
    if (JH == 2) {
      const uint32_t t1 = __shfl_sync(kFull, inc[JH - 1], 31);
      if (lane & 1) tot = t1;
    }
    if (lane < J) {
      const int row = JH == 2 ? lane >> 1 : lane;
      wtot[lane * NW + warp] = ((tot >> (8 * row)) & 0xffu) | (((tot >> (16 + 8 * row)) & 0xffu) << 16) |
                               (lane == 0 ? (w_unsafe ? 0x80000000u : 0u) | (w_silent ? 0x40000000u : 0u) : 0u);
    }
  } else if (lane < J) {
    wtot[lane * NW + warp] = lane == 0 ? (w_unsafe ? 0x80000000u : 0u) | (w_silent ? 0x40000000u : 0u) : 0u;
  }
  __syncthreads();
  const uint32_t ent_raw = lane < J * NW ? wtot[lane] : 0u;  // every warp reads the <= 32 totals
  const bool unsafe = __any_sync(kFull, (ent_raw >> 31) != 0u);
  const bool has_silent = __any_sync(kFull, ((ent_raw >> 30) & 1u) != 0u);
  const uint32_t ent = ent_raw & 0x3fffffffu;
  const uint32_t totals = __reduce_add_sync(kFull, ent);  // #pos | #neg << 16 of the tile (each <= 2048: no carry)
  const int n_pos = totals & 0xffffu, n_neg = totals >> 16;
  if (totals == 0u && !has_silent && !unsafe) {  // quiet tile: no records, no events, reference untouched
    for (int c = tid; c < A.C; c += blockDim.x)
      A.tile_counts[((size_t)s_idx * A.C + c) * A.tiles_per_stream + t_idx] = 0;
    return;
  }
  if (unsafe) {  // a pixel outside [0, 255]: the generic body replays the tile with exact int16 wrap semantics
    if (tid == 0) A.redo[2 + atomicAdd(A.redo, 1)] = tile;
    return;
  }
  const unsigned magic = A.upd.div_thr.magic;
  const int max_n = A.upd.max_time_ms;
  if (w_cand || w_silent) {  // (on sparse frames most warps of a tile hold no winner at all)
    // pn[j]: next pos rank | next neg work-list position << 16 of this thread in group j
    uint32_t pn[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int h = j % JH, row = j / JH;
      const uint32_t b = __reduce_add_sync(kFull, lane < j * NW + warp ? ent : 0u);  // earlier (group, warp) pairs
      const uint32_t ex = inc[h] - cnt[h];                                            // earlier lanes of this warp
      pn[j] = ((b & 0xffffu) + ((ex >> (8 * row)) & 0xffu)) |
              (((uint32_t)n_pos + (b >> 16) + ((ex >> (16 + 8 * row)) & 0xffu)) << 16);
    }
#pragma unroll
    for (int h = 0; h < JH; ++h) {
      if ((cnt[h] | ((sil >> h) & 1u)) == 0u) continue;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const uint32_t e = wle[h][w];
        if (e == 0u) continue;
        const int q = e & 0x1fff;
        const bool neg = (e & 0x20000000u) != 0u;
        if (e & 0x40000000u) {
          // silent spike: its only effect is the reference update (gs:244-249), done right here by the owner thread
          const int r = (e >> 13) & 0xff, a = (e >> 21) & 0xff;
          const int upd = min((int)__umulhi((unsigned)a, magic), max_n) * T;
          ref_t[q] = (state_t)(neg ? max(r - upd, 0) : min(r + upd, 255));
        } else {
          const bool lower = q >= A.W;
          const uint32_t x = lower ? pn[JH + h] : pn[h];
          wl[neg ? x >> 16 : x & 0xffffu] = e;
          const uint32_t nx = x + (neg ? 0x10000u : 1u);
          if (lower) pn[JH + h] = nx; else pn[h] = nx;
        }
      }
    }
  }
  const int n_rec = n_pos + n_neg;
  if (n_rec == 0) {  // silent spikes only: references are updated, nothing to list
    for (int c = tid; c < A.C; c += blockDim.x)
      A.tile_counts[((size_t)s_idx * A.C + c) * A.tiles_per_stream + t_idx] = 0;
    return;
  }
  __syncthreads();

  // ---- 4. one candidate per lane: reference update, slot histogram, record straight to its place ---------------------
  const int chunks_per_warp = (int)(((unsigned)(((n_rec + 31) >> 5) + NW - 1) * A.inv_nw) >> 16);
  const int per_warp = chunks_per_warp << 5;
  const int lo = min(n_rec, warp * per_warp), hi = min(n_rec, lo + per_warp);
  const size_t region = (size_t)tile * A.rec_stride;
  dvs_record *const pos_dst = A.records + region;
  dvs_record *const neg_dst = A.records + region + (A.rec_stride - n_neg) - n_pos;  // indexed by the work-list position
  unsigned status = 0;
  for (int i = lo + lane; i < hi; i += 32) {
    const uint32_t e = wl[i];
    const int q = e & 0x1fff, r = (e >> 13) & 0xff, a = (e >> 21) & 0xff;
    const bool neg = (e >> 29) & 1u;
    const int y = q >= A.W ? 1 : 0, x = q - (y ? A.W : 0);
    const int upd = min((int)__umulhi((unsigned)a, magic), max_n) * T;  // gs:244-249
    ref_t[q] = (state_t)(neg ? max(r - upd, 0) : min(r + upd, 255));
    int desc;
    if (ENC == ENC_RATE_CODED) {
      desc = max(0, (A.enc.n_slots - 1) - (int)__umulhi((unsigned)a, A.enc.dv.magic));  // k, gs:830-832
      atomicAdd(&hist[(neg ? nh : 0) + desc], 1u);
    } else {
      // make_spike_lists_time gs:903-926 (see fast2_tile)
      const int v = (int)__umulhi((unsigned)a, A.enc.dv.magic);
      int nt = min(v - 1, A.enc.n_slots - 1);
      if (neg) nt = max(nt, 0);
      desc = A.enc.n_slots - nt - 1;
      if (desc >= A.enc.n_slots) { status |= DVS_STATUS_SLOT_RANGE; desc = -1; }
      else atomicAdd(&hist[(neg ? nh : 0) + desc], 1u);
    }
    const uint2 rec = make_uint2((uint32_t)(row0 + y) | ((uint32_t)x << 16), rec_word1(a, desc, true));
    *reinterpret_cast<uint2 *>((i < n_pos ? pos_dst : neg_dst) + i) = rec;
  }
  if (ENC == ENC_TIME_CODED) {
    status = __reduce_or_sync(kFull, status);
    if (status && lane == 0 && A.status) atomicOr(A.status, (int)status);
  }
  __syncthreads();
  write_counters(A, s_idx, t_idx, hist, nh, n_pos, n_neg);
}

template <int SRC, int JH, int ENC, int ST = ST_I16, int BD = 256>
__global__ void __launch_bounds__(BD, BD == 512 ? FAST2R_MINB_512 : ST == ST_U8 ? (SRC == SRC_U8 || RANKED_NARROW_I16 ? FAST2R_MINB_S8U8 : FAST2R_MINB_S8) : FAST2R_MINB)
    k_tile_ranked(const __grid_constant__ TileArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  pdl_trigger();
  fast2r_tile<SRC, JH, ENC, ST>(A, (int)blockIdx.x, (int)blockIdx.y, smem_raw);
}

// Caller guarantees the FUSED2 shape (see dispatch_fast_shapes) and two-row tiles of at most 8192 pixels.
template <int SRC, int JH, int ENC, int ST, int BD = 256>
static int launch_ranked(const TileArgs &A_in, int bd, cudaStream_t st) {
  TileArgs A = A_in;
  A.inv_nw = (65536u + (unsigned)(bd / 32) - 1u) / (unsigned)(bd / 32);
  // coded rate with drop_silent: spikes with k == 0 (abs_diff // T_enc >= n_slots - 1) are updated but not recorded
  A.silent_a = (ENC == ENC_RATE_CODED && A.drop_silent) ? (A.enc.n_slots - 1) * A.enc.threshold : 0x7fffffff;
  const int nh = A.enc.n_slots + 1;
  A.fast_cap = fast_cap_of(A);
  const size_t smem = (size_t)A.fast_cap * 4 + (32 + 2 * nh + 4) * sizeof(uint32_t);
  cudaError_t e;
  k_tile_ranked<SRC, JH, ENC, ST, BD><<<dim3((unsigned)A.tiles_per_stream, (unsigned)A.S), bd, smem, st>>>(A);
  if ((e = cudaGetLastError()) != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
  // generic body over the (normally zero) declined tiles
  auto kr = k_tile_redo<SRC, 4, INH_K2, false>;
  const size_t smem_r = ((A.tile_px * 2 + 15) & ~15) + (128 + 2 * nh + 8) * sizeof(uint32_t);
  if (smem_r > 48 * 1024 && (e = cudaFuncSetAttribute(kr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r)) != cudaSuccess)
    return set_error((int)e, cudaGetErrorString(e));
  const int items = A.tile_px / 8;
  const long long tiles = (long long)A.S * A.tiles_per_stream;
  e = launch_pdl(kr, dim3((unsigned)(tiles < 592 ? tiles : 592)), dim3((unsigned)(((items + 3) / 4 + 31) / 32 * 32)), smem_r, st, A);
  if (e != cudaSuccess || (e = cudaGetLastError()) != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
  return DVS_OK;
}

}  // namespace dvs
