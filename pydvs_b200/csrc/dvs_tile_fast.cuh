// dvs_tile_fast.cuh -- K1 fast path: static threshold T >= 2, update_reference_rate, history_weight == 1.
//
// Same contract and bit-exact results as the generic tile pass (dvs_tile_pass.cuh); the difference is
// instruction count.  When every frame / reference pixel of a tile is in [0, 255] (a block-wide vote;
// otherwise the tile falls back to the generic body) nothing can wrap in 16 bits, so
//   1. |frame - ref|, the threshold test and the sign run on packed u16x2 lanes with DPX min/max
//      (VIMNMX.U16x2), ~8 instructions per 2 pixels with no unpacking; the signed, thresholded
//      difference of every pixel goes to shared memory;
//   2. a tile without a spike (one block-wide vote) only zeroes its counters; with history_weight == 1
//      the reference of a pixel that does not spike is unchanged (clip(r + 0, 0, 255) == r), so quiet
//      pixels are never written back;
//   3. the spiking pixels of the tile are compacted IN ROW-MAJOR ORDER into a shared-memory work list
//      (shuffle scans), and everything that is expensive per spike -- 2x2 inhibition arg-max, reference
//      update, list membership, slot histogram, record emission -- runs one spike per lane with
//      ballot/popc ranks, i.e. cost proportional to the number of spikes, not of pixels.
#pragma once
#include "dvs_tile_pass.cuh"

namespace dvs {

DEV uint4 ld16_stream(const void *p) { return __ldcs(reinterpret_cast<const uint4 *>(p)); }

// shared memory of the fast kernel: [sd: tile_px int16][wl: tile_px uint16][code: tile_px uint8]
//                                   [wtot: 128 u32][hist: 2 * (n_slots + 1) u32][misc: 8 u32]
template <int SRC, int J, int KI>
__global__ void __launch_bounds__(256) k_tile_fast(const __grid_constant__ TileArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int plane_bytes = (A.tile_px * 2 + 15) & ~15;
  int16_t *sd_sm = reinterpret_cast<int16_t *>(smem_raw);
  uint16_t *wl = reinterpret_cast<uint16_t *>(smem_raw + plane_bytes);
  uint8_t *code_sm = smem_raw + 2 * plane_bytes;
  uint32_t *wtot = reinterpret_cast<uint32_t *>(smem_raw + 2 * plane_bytes + ((A.tile_px + 15) & ~15));
  uint32_t *hist = wtot + 128;
  const int nh = A.enc.mode == DVS_ENC_NONE ? 0 : (A.enc.n_slots + 1);
  uint32_t *misc = hist + 2 * nh;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
  const int tile = blockIdx.x;
  const int s_idx = tile / A.tiles_per_stream, t_idx = tile - s_idx * A.tiles_per_stream;
  const int row0 = t_idx * A.tile_rows;
  const int rows = min(A.tile_rows, A.H - row0);
  const int npx = rows * A.W;  // multiple of 8 (host-checked)
  const size_t plane_off = (size_t)s_idx * A.H * A.W + (size_t)row0 * A.W;

  // ---- 1. loads: all 2 * J 128-bit requests of the thread are in flight before first use --------
  uint32_t c_pk[J][4], r_pk[J][4];
  bool have[J];
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
  uint32_t rng = 0;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    if (have[j]) {
      rng |= r_pk[j][0] | r_pk[j][1] | r_pk[j][2] | r_pk[j][3];
      if (SRC != SRC_U8) rng |= c_pk[j][0] | c_pk[j][1] | c_pk[j][2] | c_pk[j][3];
    }
  }
  if (__syncthreads_or((rng & 0xFF00FF00u) != 0)) {  // some pixel outside [0, 255]: exact generic path
    if (tid == 0) A.redo[1 + atomicAdd(A.redo, 1)] = tile;  // k_tile_redo picks it up
    return;
  }

  for (int i = tid; i < 2 * nh; i += blockDim.x) hist[i] = 0;

  // ---- 2. packed threshold + sign: a = max - min per u16 lane, spike iff a > T, negative iff c < r --
  const int T = A.thr_scalar;                                  // 2 <= T <= 32767 (host-checked)
  const uint32_t kadd = (uint32_t)(0x7fff - T) * 0x00010001u;  // bit 15 of (a + kadd) <=> a > T
  uint32_t flags8[J], cnt[J];
  int any = 0;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    flags8[j] = 0; cnt[j] = 0;
    if (!have[j]) continue;
    const int q0 = (j * (int)blockDim.x + tid) * 8;
    uint32_t sd[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const uint32_t c = c_pk[j][w], r = r_pk[j][w];
      const uint32_t mx = __vmaxu2(c, r), mn = __vminu2(c, r);
      const uint32_t a = mx - mn;                                  // no borrow: max >= min lane-wise
      const uint32_t fl = (a + kadd) & 0x80008000u;                // spike flags at bits 15 / 31
      const uint32_t ne = ((mn ^ c) + 0x7fff7fffu) & 0x80008000u;  // lane of min differs from c <=> c > r
      const uint32_t keep = (fl >> 15) * 0xffffu;                  // 0xffff in spiking lanes
      const uint32_t neg = ((fl & ~ne) >> 15) * 0xffffu;           // 0xffff in spiking lanes with c < r
      // abs_diff * spikes, negated (two's complement per lane; a >= 3 there, so no carry) where c < r
      sd[w] = ((a & keep) ^ neg) + (neg & 0x00010001u);
      flags8[j] |= (((fl >> 15) & 1u) | ((fl >> 30) & 2u)) << (2 * w);
    }
    cnt[j] = __popc(flags8[j]);
    any |= (int)flags8[j];
    st16(sd_sm + q0, sd);
  }
  if (!__syncthreads_or(any)) {  // quiet tile: no records, no events, reference untouched
    for (int c = tid; c < A.C; c += blockDim.x)
      A.tile_counts[((size_t)s_idx * A.C + c) * A.tiles_per_stream + t_idx] = 0;
    return;
  }

  // ---- 3. ordered work list of spiking pixels: exclusive scan in (group j, thread) order ---------
  uint32_t incl[J];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    uint32_t x = cnt[j];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(kFull, x, o);
      if (lane >= o) x += y;
    }
    incl[j] = x;
    if (lane == 31) wtot[j * NW + warp] = x;
  }
  __syncthreads();
  if (warp == 0) {
    const uint32_t v = lane < J * NW ? wtot[lane] : 0;  // J * NW <= 32 in this kernel
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(kFull, x, o);
      if (lane >= o) x += y;
    }
    if (lane < J * NW) wtot[lane] = x - v;
    if (lane == 31) misc[0] = x;
  }
  __syncthreads();
  const int n_cand = (int)misc[0];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    if (flags8[j] == 0) continue;
    const int q0 = (j * (int)blockDim.x + tid) * 8;
    uint32_t pos = wtot[j * NW + warp] + incl[j] - cnt[j];
#pragma unroll
    for (int p = 0; p < 8; ++p)
      if ((flags8[j] >> p) & 1u) wl[pos++] = (uint16_t)(q0 + p);
  }
  __syncthreads();

  // ---- 4a. one spiking pixel per lane: inhibition, reference update, list membership, histogram ---
  const int per_warp = (((n_cand + NW - 1) / NW) + 31) & ~31;
  const int lo = min(n_cand, warp * per_warp), hi = min(n_cand, lo + per_warp);
  const unsigned magic = A.upd.div_thr.magic;
  const int max_n = A.upd.max_time_ms;
  unsigned long long hp = 0, hn = 0;
  unsigned status = 0;
  uint32_t w_pos = 0, w_neg = 0;
  for (int i0 = lo; i0 < hi; i0 += 32) {
    const int i = i0 + lane;
    bool tp = false, tn = false;
    if (i < hi) {
      const int q = wl[i];
      const int v = sd_sm[q];
      const int a = v < 0 ? -v : v;
      bool win = true;
      if (KI == INH_K2) {  // local_inhibition gs:177-221: first arg-max of the 2x2 area, row-major, strict '>'
        const int y = (int)__umulhi((unsigned)q, A.inv_w), x = q - y * A.W;
        const int o = (y & ~1) * A.W + (x & ~1);
        const uint32_t top = *reinterpret_cast<const uint32_t *>(sd_sm + o);
        const uint32_t bot = *reinterpret_cast<const uint32_t *>(sd_sm + o + A.W);
        const int t00 = abs((int)(short)(top & 0xffff)), t01 = abs((int)(short)(top >> 16));
        const int t10 = abs((int)(short)(bot & 0xffff)), t11 = abs((int)(short)(bot >> 16));
        int best = t00, bi = 0;
        if (t01 > best) { best = t01; bi = 1; }
        if (t10 > best) { best = t10; bi = 2; }
        if (t11 > best) { best = t11; bi = 3; }
        win = bi == (((y & 1) << 1) | (x & 1));
      }
      if (win) {
        const int s = v < 0 ? -1 : 1;
        int16_t *rp = A.ref + plane_off + q;
        const int r = *rp;
        const int upd = min((int)__umulhi((unsigned)a, magic), max_n) * T;  // update_reference_rate gs:244-249
        *rp = (int16_t)(s > 0 ? min(r + upd, 255) : max(r - upd, 0));
        split_px(s, A.polarity, A.uncoded, tp, tn);
        if (tp || tn) hist_add(A, hist, nh, a, tp, hp, hn, status);
      }
      code_sm[i] = tp ? 1 : (tn ? 2 : 0);
    }
    w_pos += __popc(__ballot_sync(kFull, tp));
    w_neg += __popc(__ballot_sync(kFull, tn));
    // per_warp <= 1024 candidates -> at most 32 records per lane: the 8-bit histogram fields cannot overflow
  }
  if (lane == 0) wtot[64 + warp] = w_pos | (w_neg << 16);
  __syncthreads();
  uint32_t base = 0, totals = 0;
  for (int w = 0; w < NW; ++w) {
    const uint32_t t = wtot[64 + w];
    if (w < warp) base += t;
    totals += t;
  }
  const int n_pos_tile = totals & 0xffff, n_neg_tile = totals >> 16;

  // ---- 4b. records: rank inside the tile = warp base + ballot prefix ------------------------------
  if (totals != 0) {
    const size_t region = (size_t)tile * A.tile_px;
    dvs_record *pos_dst = A.records + region;
    dvs_record *neg_dst = A.records + region + (A.tile_px - n_neg_tile);
    uint32_t op = base & 0xffff, on = base >> 16;
    const unsigned lt = (1u << lane) - 1u;
    for (int i0 = lo; i0 < hi; i0 += 32) {
      const int i = i0 + lane;
      const int code = i < hi ? code_sm[i] : 0;
      const unsigned bp = __ballot_sync(kFull, code == 1), bn = __ballot_sync(kFull, code == 2);
      if (code) {
        const int q = wl[i];
        const int v = sd_sm[q];
        const int y = (int)__umulhi((unsigned)q, A.inv_w), x = q - y * A.W;
        const uint2 rec = make_uint2((uint32_t)(row0 + y) | ((uint32_t)x << 16), (uint32_t)(v < 0 ? -v : v));
        if (code == 1) *reinterpret_cast<uint2 *>(pos_dst + op + __popc(bp & lt)) = rec;
        else *reinterpret_cast<uint2 *>(neg_dst + on + __popc(bn & lt)) = rec;
      }
      op += __popc(bp);
      on += __popc(bn);
    }
  }
  tile_finish(A, s_idx, t_idx, hist, nh, n_pos_tile, n_neg_tile, hp, hn, status);
}

template <int SRC, int J, int KI>
static int launch_fast_shape(const TileArgs &A, int bd, size_t smem, cudaStream_t st) {
  auto k = k_tile_fast<SRC, J, KI>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
  }
  cudaError_t e = cudaMemsetAsync(A.redo, 0, sizeof(int32_t), st);
  if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
  const long long tiles = (long long)A.S * A.tiles_per_stream;
  k<<<(unsigned)tiles, bd, smem, st>>>(A);
  if ((e = cudaGetLastError()) != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
  // generic body over the (normally zero) tiles with pixels outside [0, 255]
  auto kr = k_tile_redo<SRC, J, KI>;
  const int nh = A.enc.mode == DVS_ENC_NONE ? 0 : (A.enc.n_slots + 1);
  const size_t smem_r = ((KI != INH_NONE) ? ((A.tile_px * 2 + 15) & ~15) : 0) + (128 + 2 * nh + 8) * sizeof(uint32_t);
  if (smem_r > 48 * 1024) {
    e = cudaFuncSetAttribute(kr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r);
    if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
  }
  kr<<<(unsigned)(tiles < 592 ? tiles : 592), bd, smem_r, st>>>(A);
  if ((e = cudaGetLastError()) != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
  return DVS_OK;
}

// Caller guarantees: vec_ok, W % 8 == 0, <= 1024 groups per tile, inh_k in {1, 2}, 2 <= T, rate update,
// history_weight == 1, max_time_ms >= 0, no debug planes.
template <int SRC>
static int dispatch_fast(const TileArgs &A, cudaStream_t st) {
  const int items = A.tile_px / 8;
  const int nh = A.enc.mode == DVS_ENC_NONE ? 0 : (A.enc.n_slots + 1);
  const bool k2 = A.inh_k == 2;
  const size_t smem = (size_t)((A.tile_px * 2 + 15) & ~15) * 2 + (size_t)((A.tile_px + 15) & ~15) +
                      (128 + 2 * nh + 8) * sizeof(uint32_t);
  auto round32 = [](int x) { return (x + 31) / 32 * 32; };
  if (items <= 256) {
    const int bd = round32(items);
    return k2 ? launch_fast_shape<SRC, 1, INH_K2>(A, bd, smem, st) : launch_fast_shape<SRC, 1, INH_NONE>(A, bd, smem, st);
  }
  const int bd = round32((items + 3) / 4);
  return k2 ? launch_fast_shape<SRC, 4, INH_K2>(A, bd, smem, st) : launch_fast_shape<SRC, 4, INH_NONE>(A, bd, smem, st);
}

}  // namespace dvs
