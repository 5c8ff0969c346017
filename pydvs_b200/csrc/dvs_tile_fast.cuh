// dvs_tile_fast.cuh -- K1 fast path: static threshold T >= 2 or the adaptive-threshold plane, rate / time-binary
// update, history_weight in [0, 1], inhibition off or 2x2 (two-row tiles).
//
// Same contract and bit-exact results as the generic tile pass (dvs_tile_pass.cuh); the difference is
// instruction count.  When every frame / reference pixel of a tile is in [0, 255] (a block-wide vote;
// otherwise the tile is handed to the generic body through the redo list) nothing can wrap in 16 bits, so
//   1. |frame - ref| and the threshold test run on packed u16x2 lanes with DPX min/max (VIMNMX.U16x2),
//      ~5 instructions per 2 pixels with no unpacking; with two-row tiles each thread owns the same column
//      blocks in both rows, so the 2x2 inhibition arg-max is taken in registers right there;
//   2. with history_weight == 1 the reference of a pixel that does not spike is unchanged
//      (clip(r + 0, 0, 255) == r): quiet pixels are never written back, quiet tiles only zero their counters;
//   3. the surviving spikes of the tile are compacted IN ROW-MAJOR ORDER into a shared-memory work list
//      (shuffle scans over packed counters, one barrier), and everything that is expensive per spike --
//      reference update, list membership, slot histogram (fire-and-forget shared-memory atomics), record --
//      runs one spike per lane with ballot/popc ranks: cost proportional to the number of spikes, spread evenly
//      over the lanes, not to the number of pixels.
#pragma once
#include "dvs_tile_pass.cuh"

namespace dvs {

DEV uint4 ld16_stream(const void *p) { return __ldcs(reinterpret_cast<const uint4 *>(p)); }

// state planes (reference, thresholds) in either layout: element type, and 8 consecutive pixels <-> four words of two
// 16-bit lanes (ST_U8: one 8-byte access, bytes widened / narrowed by PRMT; every lane is in [0, 255] on this path)
template <int ST> struct state_elem { using type = int16_t; };
template <> struct state_elem<ST_U8> { using type = uint8_t; };
template <int ST>
DEV void ld_state8(const void *p, uint32_t (&pk)[4]) {
  if (ST == ST_U8) {
    const uint2 v = *reinterpret_cast<const uint2 *>(p);
    pk[0] = __byte_perm(v.x, 0, 0x4140); pk[1] = __byte_perm(v.x, 0, 0x4342);
    pk[2] = __byte_perm(v.y, 0, 0x4140); pk[3] = __byte_perm(v.y, 0, 0x4342);
  } else {
    const uint4 v = ld16(p);
    pk[0] = v.x; pk[1] = v.y; pk[2] = v.z; pk[3] = v.w;
  }
}
// The same in two steps, so that a thread can have ALL its loads in flight before the first PRMT waits for one of them
// (ptxas keeps an unpack that directly follows its load ahead of the next load: the loads would complete one by one):
// ld8_raw<BYTES> fetches 8 pixels (BYTES = 1: 8 bytes in raw[0..1], BYTES = 2: 16 bytes), widen8<BYTES> turns them
// into the four words of two 16-bit lanes in place.
template <int BYTES, bool STREAM>
DEV void ld8_raw(const void *p, uint32_t (&raw)[4]) {
  if (BYTES == 1) {
    const uint2 v = STREAM ? __ldcs(reinterpret_cast<const uint2 *>(p)) : *reinterpret_cast<const uint2 *>(p);
    raw[0] = v.x; raw[1] = v.y;
  } else {
    const uint4 v = STREAM ? ld16_stream(p) : ld16(p);
    raw[0] = v.x; raw[1] = v.y; raw[2] = v.z; raw[3] = v.w;
  }
}
template <int BYTES>
DEV void widen8(uint32_t (&pk)[4]) {
  if (BYTES == 1) {
    const uint32_t x = pk[0], y = pk[1];
    pk[0] = __byte_perm(x, 0, 0x4140); pk[1] = __byte_perm(x, 0, 0x4342);
    pk[2] = __byte_perm(y, 0, 0x4140); pk[3] = __byte_perm(y, 0, 0x4342);
  }
}
template <int ST>
DEV void st_state8(void *p, const uint32_t (&pk)[4]) {
  if (ST == ST_U8)
    *reinterpret_cast<uint2 *>(p) = make_uint2(__byte_perm(pk[0], pk[1], 0x6420), __byte_perm(pk[2], pk[3], 0x6420));
  else st16(p, pk);
}
// one state pixel as an unsigned value (the int16 layout is read as uint16: in [0, 255] wherever this is used)
template <int ST>
DEV int state_px(const void *plane, int q) {
  if (ST == ST_U8) return (int)static_cast<const uint8_t *>(plane)[q];
  return (int)static_cast<const uint16_t *>(plane)[q];
}

// PRMT with the full selector (bit 3 of a selector nibble replicates the sign bit of the chosen byte over the target
// byte; __byte_perm() ignores that bit, hence the inline PTX)
template <uint32_t SEL>
DEV uint32_t prmt(uint32_t a, uint32_t b = 0u) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(SEL));
  return r;
}
// bit 15 / 31 of x replicated over its 16-bit lane (byte 1 -> bytes 0, 1; byte 3 -> bytes 2, 3)
DEV uint32_t msb_lane_mask(uint32_t x) { return prmt<0xBB99u>(x); }
// two flag words with bits 15 / 31 (nothing else set in bytes 1 and 3 above bit 7 matters) -> flags of the four
// pixels as bits 0-3: the four bytes are gathered with one PRMT, their top bits moved to bits 0, 8, 16, 24, and a
// multiplication lines them up in bits 24-27 (b_k * 2^(8k) * 2^(24 - 7k); all partial products are distinct bits)
DEV uint32_t flags4_of(uint32_t f_lo, uint32_t f_hi) {
  const uint32_t g = prmt<0x7531u>(f_lo, f_hi) & 0x80808080u;
  return ((g >> 7) * 0x01020408u) >> 24;
}
// the inverse for four pixel flags in bits 0-3 of n (n < 16): word with the flag of pixel k in bit 8k + 7, from which
// prmt<0x9988> / prmt<0xBBAA> make the lane masks of the first / second pixel pair
DEV uint32_t bytes_of_flags4(uint32_t n) { return n * 0x10204080u; }

// shared memory of the fast kernel: [sd: tile_px int16][wl: tile_px uint16][code: tile_px uint8]
//                                   [wtot: 128 u32][hist: 2 * (n_slots + 1) u32][misc: 8 u32]
#ifndef FAST2_MINB
#define FAST2_MINB 5
#endif
// tuning switches (tools/build_variant.py builds the alternatives; the defaults are the measured winners)
#ifndef FAST2_THR_BIASED
#define FAST2_THR_BIASED 2   // coded threshold update in the `down`-biased domain (one clip) instead of two clips + select:
                             // 0 never, 1 always, 2 uint8 frames only (K1 ms, 64 x 1080p: int16 0.308 / 0.319, uint8 0.329 / 0.318)
#endif
#ifndef FAST2_INH_KEYS
#define FAST2_INH_KEYS 1     // 2x2 arg-max through packed keys instead of a compare / select chain (int16 frames; with
                             // uint8 frames both switches cost 10 %: K1 0.557 vs 0.504 ms at 64 x 4K)
#endif
#ifndef FAST2_PAIR_XY
#define FAST2_PAIR_XY 1      // two-row tiles: row / column of a candidate by one compare instead of a multiply-high
#endif
#ifndef FAST2_J2
#define FAST2_J2 0           // linear tiles of 257..512 groups with two groups per thread (else four)
#endif
#ifndef FAST2_FUSED2
#define FAST2_FUSED2 1       // fused threshold + 2x2 inhibition + entry build (see FUSED2 below)
#endif
#ifndef FAST2_FUSED2_U8
#define FAST2_FUSED2_U8 1    // ... for uint8 frames as well (K1 at 64 x 4K: 0.424 ms against 0.495 with the index-only list)
#endif
#ifndef FAST2_QONLY_ADAPT
#define FAST2_QONLY_ADAPT 0  // index-only work list for the adaptive kernels as well
#endif
#ifndef FAST2_LAZY_STATE
#define FAST2_LAZY_STATE 1   // adaptive thresholds without inhibition on one-byte state: reference / threshold bytes stay packed
                             // (2 registers per group) and a word is widened where it is used
#endif
#ifndef FAST2_MINB_ADPT_S8
#define FAST2_MINB_ADPT_S8 5
#endif
#ifndef FAST2_THR_EARLY
#define FAST2_THR_EARLY 1    // adaptive thresholds without inhibition: advanced and stored in the threshold pass itself
#endif
constexpr int kFastCap = 2048;  // work-list capacity; denser tiles go to the generic body (redo list)
// Capacity actually reserved for a launch: a tile cannot hold more candidates than pixels (areas with k x k inhibition),
// and small tiles (VGA rows, MNIST-sized streams) get many more resident CTAs per SM out of the smaller footprint.
static inline int fast_cap_of(const TileArgs &A) {
  const int k2 = A.inh_k >= 2 ? A.inh_k * A.inh_k : 1;
  const int bound = (A.tile_px + k2 - 1) / k2;
  return bound >= kFastCap ? kFastCap : (bound + 31) / 32 * 32;
}

// ---- fast kernel, register-only variant ---------------------------------------------------------
// Used when no shared-memory plane is needed: inhibition off (PAIR == false: thread owns JH groups in
// linear order), or 2x2 inhibition with two-row tiles (PAIR == true: thread owns the SAME JH column
// blocks in both rows of the tile, so each 2x2 area lives in one thread's registers and the arg-max
// is taken before the work list is built -- only area winners become candidates).
// shared memory: [wl: kFastCap u32][rec: kFastCap uint2][wtot: 128 u32][hist: 2 * (n_slots + 1) u32]
// ADAPT: per-pixel threshold plane (thresholded_difference_adpt gs:83-113 + the coded
// update_reference_rate_adpt gs:256-297): the plane is read with the frame, compared lane-wise, advanced
// (+up where the pixel spikes after inhibition, -down elsewhere, clipped) on packed lanes and written back
// for every pixel (10 B/px); the division abs_diff // threshold only happens for candidates.
// One tile.  STAGED: the tile's frame / reference rows were brought into shared memory by a bulk async copy
// (st_cur / st_ref); after the first barrier every thread holds its pixels in registers, so the stage buffer
// is dead and doubles as work list + record staging (smem_raw points at it; `ctl` holds wtot / hist).
// ENC: 0 = any encoder / polarity / update the fast path takes (work-list entries carry reference and abs_diff);
//      1 = coded RATE encoder, 2 = coded TIME encoder -- MERGED lists, history_weight == 1, rate (or rate_adpt) update,
//      inlined membership / slot / histogram.  For ENC != 0 the work list holds only the 16-bit tile-local pixel index
//      (plus the old threshold when ADAPT): the lane that takes a candidate re-reads its frame and reference pixel
//      (two 2-byte loads that hit L1 / L2, the tile was just streamed in) instead of the pushing thread extracting
//      them from its packed registers -- the push was the most divergent part of the kernel (a warp runs it as often
//      as its busiest lane has candidates), so the work moves to the phase that is spread evenly over the lanes.
enum { ENC_ANY = 0, ENC_RATE_CODED = 1, ENC_TIME_CODED = 2 };
// a // thr for 0 <= a <= 255 and thr >= 0 (NumPy: x // 0 == 0): (a + 0.5) * rcp(thr) truncates to the exact quotient,
// the nearest integer boundary is >= 0.5 / (a + 1) >= 2^-9 away in relative terms, far above the rcp / product error
DEV int adpt_div(int a, int thr) {
  return thr == 0 ? 0 : (int)(((float)a + 0.5f) * __frcp_rn((float)thr));
}
// QUAD: 4x4 inhibition on four-row tiles -- the PAIR mapping with four rows per column block (JH == 1): a thread holds
// two complete 4x4 areas (columns 0-3 and 4-7 of its block) in registers, the arg-max is a packed maximum over keys
// abs_diff * 16 + (15 - index in the area), and the winners go to the index-only work list (ENC != 0).  Rows of up to
// 4096 pixels: CTAs of up to 512 threads.
template <int SRC, int JH, bool PAIR, int ENC, bool ADAPT, bool STAGED, bool QUAD = false, int ST = ST_I16>
DEV void fast2_tile(const TileArgs &A, const int t_idx, const int s_idx, unsigned char *smem_raw, unsigned char *ctl,
                    const unsigned char *st_cur, const unsigned char *st_ref) {
  static_assert(!QUAD || (PAIR && JH == 1 && ENC != ENC_ANY && !STAGED), "4x4 variant: one column block, inlined encoders");
  static_assert(!STAGED || ST == ST_I16, "the staged rows are int16");
  using state_t = typename state_elem<ST>::type;
  constexpr int J = QUAD ? 4 : (PAIR ? 2 * JH : JH);
  constexpr bool SIMPLE = ENC != ENC_ANY;
  // index-only work list (see above), K1 ms at 64 streams with / without it (gpurun sweep, round 2): uint8 4K frames
  // 0.503 / 0.581, 4K without inhibition 0.598 / 0.625 -- but int16 dense 4K 1.204 / 1.142 (the static 2x2 area walk does
  // not diverge) and adaptive 1080p 0.321 / 0.309 (the old threshold has to be extracted for the entry anyway)
#ifndef FAST2_QONLY_MODE
#define FAST2_QONLY_MODE 2
#endif
  constexpr bool FUSED2_SHAPE = FAST2_FUSED2 && PAIR && !QUAD && !ADAPT && SIMPLE && (SRC == SRC_I16 || FAST2_FUSED2_U8);
  constexpr bool QONLY = QUAD || (!FUSED2_SHAPE && SIMPLE && (FAST2_QONLY_MODE == 1 || (FAST2_QONLY_MODE == 2 && (ADAPT ? FAST2_QONLY_ADAPT != 0 : (!PAIR || SRC == SRC_U8)))));
  // ADAPT with the index-only list: the thresholds are written back AFTER the spike loop (behind its barrier), so the
  // lane that takes a candidate reads the old threshold from the plane like the frame and reference pixels, and the
  // push stores nothing but the index
  constexpr bool TLATE = ADAPT && QONLY && FAST2_QONLY_ADAPT == 2;
  // ADAPT without inhibition: a pixel's spike flag is final as soon as it is thresholded, so the new threshold is
  // computed from the flag word's lane mask (one PRMT) and stored right there; the old thresholds stay in registers
  // for the candidates' divisions, and a tile that turns out to be declined writes them back before it is handed on
  constexpr bool THR_EARLY = FAST2_THR_EARLY && ADAPT && !PAIR && !TLATE;
  // 2x2 tiles, static threshold, inlined encoder, int16 frames (the headline kernel): threshold, inhibition and the
  // work-list entry of each area's winner are computed in one pass over the area words while frame and reference are at
  // hand -- the winner's abs_diff, position and sign all come out of one packed key maximum -- so the push after the
  // scan only stores finished entries
  constexpr bool FUSED2 = FUSED2_SHAPE;
  constexpr bool LAZY_S = FAST2_LAZY_STATE && ST == ST_U8 && ADAPT && !PAIR && !STAGED && !QUAD;
  uint32_t *wl = reinterpret_cast<uint32_t *>(smem_raw);
  uint16_t *wl16 = reinterpret_cast<uint16_t *>(smem_raw);  // ENC != 0: pixel indices only
  const int cap = STAGED ? kFastCap : A.fast_cap;
  uint2 *rec_sm = reinterpret_cast<uint2 *>(smem_raw + cap * 4);
  uint32_t *wtot = reinterpret_cast<uint32_t *>(ctl);
  uint16_t *wl_thr = reinterpret_cast<uint16_t *>(wtot + 128 + 2 * (A.enc.mode == DVS_ENC_NONE ? 0 : A.enc.n_slots + 1) + 8);  // ADAPT
  uint8_t *tr_lut = reinterpret_cast<uint8_t *>(wl_thr + (ADAPT ? cap : 0));  // history_weight != 1: trunc(hw * r), r in [0, 255]
  const bool decay = !SIMPLE && A.upd.hw_is_one == 0;  // (SIMPLE implies plain rate update, no decay) 0 <= history_weight < 1 (host-checked): every pixel's reference moves
  uint32_t *hist = wtot + 128;
  const int nh = A.enc.mode == DVS_ENC_NONE ? 0 : (A.enc.n_slots + 1);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
  const int tile = s_idx * A.tiles_per_stream + t_idx;
  const int row0 = t_idx * A.tile_rows;
  const int npx = min(A.tile_rows, A.H - row0) * A.W;  // multiple of 8 (host-checked)
  const size_t plane_off = ((size_t)s_idx * A.H + row0) * A.W;
  state_t *const ref_t = static_cast<state_t *>(A.ref) + plane_off;
  const unsigned char *const cur_t = static_cast<const unsigned char *>(A.curr) + plane_off * (SRC == SRC_U8 ? 1 : 2);
  const int gpr = A.W >> 3;  // groups of 8 pixels per row

  // group j of this thread starts at tile-local pixel q0(j) = off(j) + 8 * tid, with off(j) uniform over the CTA
  // (PAIR: j = row * JH + column block), so every address below is "per-thread base + uniform offset"
  auto off_of = [&](int j) -> int {
    if (PAIR) return (j / JH) * A.W + (j % JH) * (int)blockDim.x * 8;
    return j * (int)blockDim.x * 8;
  };
  auto in_tile = [&](int j) -> bool {
    if (PAIR) return (j % JH) * (int)blockDim.x + tid < gpr;
    return off_of(j) + tid * 8 < npx;
  };
  auto q0_of = [&](int j) -> int {
    if (PAIR) return in_tile(j) ? off_of(j) + tid * 8 : npx;
    return (j * (int)blockDim.x + tid) * 8;
  };
  const unsigned char *const cur_thread = cur_t + (size_t)tid * 8 * (SRC == SRC_U8 ? 1 : 2);
  state_t *const ref_thread = ref_t + tid * 8;

  // ---- 1. loads: all 2 * J 128-bit requests of the thread are in flight before first use --------
  uint32_t c_pk[J][4], r_pk[J][4], t_pk[ADAPT ? J : 1][4];
  bool have[J];
  state_t *const thr_t = ADAPT ? static_cast<state_t *>(A.thr) + plane_off : nullptr;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int q0 = PAIR ? off_of(j) + tid * 8 : q0_of(j), off = off_of(j);
    have[j] = PAIR ? in_tile(j) : q0 < npx;
    if (have[j]) {
      constexpr int CB = SRC == SRC_U8 ? 1 : 2, SB = ST == ST_U8 ? 1 : 2;
      if (ADAPT) ld8_raw<SB, false>(thr_t + q0, t_pk[j]);
      if (STAGED) ld8_raw<CB, false>(st_cur + CB * q0, c_pk[j]);
      else ld8_raw<CB, true>(PAIR ? cur_thread + CB * off : cur_t + CB * q0, c_pk[j]);
      if (STAGED) ld8_raw<2, false>(st_ref + 2 * q0, r_pk[j]);
      else ld8_raw<SB, false>(PAIR ? ref_thread + off : ref_t + q0, r_pk[j]);
    }
  }
  // bytes -> 16-bit lanes only after every load has been issued (see ld8_raw)
#pragma unroll
  for (int j = 0; j < J; ++j) {
    if (!have[j]) continue;
    if (ADAPT && !LAZY_S) widen8<(ST == ST_U8 ? 1 : 2)>(t_pk[ADAPT ? j : 0]);
    widen8<(SRC == SRC_U8 ? 1 : 2)>(c_pk[j]);
    if (!STAGED && !LAZY_S) widen8<(ST == ST_U8 ? 1 : 2)>(r_pk[j]);
  }
  // word w (two 16-bit lanes) of the reference / threshold group j; LAZY_S widens the packed bytes on the spot
  auto rw = [&](int j, int w) -> uint32_t {
    return LAZY_S ? __byte_perm(r_pk[j][w >> 1], 0, (w & 1) ? 0x4342 : 0x4140) : r_pk[j][w];
  };
  auto tw = [&](int j, int w) -> uint32_t {
    return LAZY_S ? __byte_perm(t_pk[ADAPT ? j : 0][w >> 1], 0, (w & 1) ? 0x4342 : 0x4140) : t_pk[ADAPT ? j : 0][w];
  };
  // pixel p (0..7) of group j as a value
  auto r_px = [&](int j, int p) -> int {
    if (LAZY_S) return (int)((((unsigned long long)r_pk[j][1] << 32 | r_pk[j][0]) >> (8 * p)) & 0xff);
    const unsigned long long lo = ((unsigned long long)r_pk[j][1] << 32) | r_pk[j][0], hi = ((unsigned long long)r_pk[j][3] << 32) | r_pk[j][2];
    return (int)(((p < 4 ? lo : hi) >> (16 * (p & 3))) & 0xffff);
  };
  auto t_px = [&](int j, int p) -> int {
    const int jt = ADAPT ? j : 0;
    if (LAZY_S) return (int)((((unsigned long long)t_pk[jt][1] << 32 | t_pk[jt][0]) >> (8 * p)) & 0xff);
    const unsigned long long lo = ((unsigned long long)t_pk[jt][1] << 32) | t_pk[jt][0], hi = ((unsigned long long)t_pk[jt][3] << 32) | t_pk[jt][2];
    return (int)(((p < 4 ? lo : hi) >> (16 * (p & 3))) & 0xffff);
  };
  if (tid < 2 * nh) hist[tid] = 0;
  for (int i = blockDim.x + tid; i < 2 * nh; i += blockDim.x) hist[i] = 0;
  if (decay)  // DTYPE(history_weight * ref_frame) gs:249 for every possible in-range reference value, exact in fp64
    for (int i = tid; i < 256; i += blockDim.x) tr_lut[i] = (uint8_t)hist_weight(A.upd.hw, false, i);

  // ---- 2. packed threshold: a = max - min per u16 lane (masked to spiking lanes), spike iff a > T --
  const int T = A.thr_scalar;                                  // 2 <= T <= 32767 (host-checked)
  const uint32_t kadd = (uint32_t)(0x7fff - T) * 0x00010001u;  // bit 15 of (a + kadd) <=> a > T
  uint32_t flags8[J];
  uint32_t rng = 0, cnt_lo = 0, cnt_hi = 0;  // counts of groups (0,1) and (2,3), 16-bit fields
  uint32_t tmax = 0;                         // ADAPT: lane-wise maximum of the thresholds of this thread
  // abs_diff of a word of two pixels, kept only where the pixel spikes (abs_diff * spikes, gs:73-75); recomputed
  // where it is needed instead of being held in registers across the phases
  auto masked_abs = [&](int j, int w) -> uint32_t {
    const uint32_t c = c_pk[j][w], r = rw(j, w);
    const uint32_t a = __vmaxu2(c, r) - __vminu2(c, r);
    const uint32_t f = ADAPT ? ~((tw(j, w) | 0x80008000u) - a) & 0x80008000u : (a + kadd) & 0x80008000u;
    return a & ((f >> 15) * 0xffffu);
  };
  // update_reference_rate_adpt gs:291-297 on packed lanes: +up on the lanes of m (pixels that still spike), -down
  // elsewhere, clipped to [min, max]
  auto thr_new = [&](uint32_t th, uint32_t m) -> uint32_t {
    const uint32_t up2 = (uint32_t)A.upd.thr_up * 0x00010001u, dn2 = (uint32_t)A.upd.thr_down * 0x00010001u;
    const uint32_t mn2 = (uint32_t)A.upd.min_thr * 0x00010001u, mx2 = (uint32_t)A.upd.max_thr * 0x00010001u;
    if (!A.upd.uncoded) {  // (the host admits the coded rule only with up + down <= 32768)
      // coded rule in a domain biased by `down` (no lane can borrow): x = th + (spike ? up + down : 0) stands for
      // th + up resp. th - down, clip(x, min + down, max + down) - down is the clipped threshold.  th + up <= 32767
      // (range vote) and up + down <= 32768 keep every lane below 2^16.
      return __vminu2(__vmaxu2(th + (m & (up2 + dn2)), mn2 + dn2), mx2 + dn2) - dn2;
    }
    // uncoded twin gsu:286-295: += up only while th < max, -= down only while th > min, no clip.
    // lane-wise compare of values < 2^15: bit 15 of ((x | 0x8000) - y) is set iff x >= y
    const uint32_t ge_max = msb_lane_mask((th | 0x80008000u) - mx2);  // th >= max
    const uint32_t le_min = msb_lane_mask((mn2 | 0x80008000u) - th);  // min >= th
    return th + ((up2 & ~ge_max) & m) - ((dn2 & ~le_min) & ~m);
  };
  uint32_t wle[FUSED2 ? JH : 1][4];  // FUSED2: finished work-list entry of each area (0 = none)
  if constexpr (FUSED2) {
#pragma unroll
    for (int h = 0; h < JH; ++h) {
      const int ju = h, jl = JH + h;
      flags8[ju] = 0; flags8[jl] = 0;
#pragma unroll
      for (int w = 0; w < 4; ++w) wle[h][w] = 0u;
      if (!have[ju]) continue;
      if (ST != ST_U8) rng |= ((r_pk[ju][0] | r_pk[ju][1] | r_pk[ju][2]) | r_pk[ju][3]) | ((r_pk[jl][0] | r_pk[jl][1] | r_pk[jl][2]) | r_pk[jl][3]);
      rng |= ((c_pk[ju][0] | c_pk[ju][1] | c_pk[ju][2]) | c_pk[ju][3]) | ((c_pk[jl][0] | c_pk[jl][1] | c_pk[jl][2]) | c_pk[jl][3]);
      const int qu = q0_of(ju);
      uint32_t ku = 0, kl = 0;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const uint32_t cu = c_pk[ju][w], ru = r_pk[ju][w], cl = c_pk[jl][w], rl = r_pk[jl][w];
        const uint32_t au = __vmaxu2(cu, ru) - __vminu2(cu, ru), al = __vmaxu2(cl, rl) - __vminu2(cl, rl);
        const uint32_t fu = (au + kadd) & 0x80008000u, fl2 = (al + kadd) & 0x80008000u;  // thresholded_difference gs:67-78
        if ((fu | fl2) == 0u) continue;
        // local_inhibition gs:177-221 + argmax gs:118-132: keys abs_diff * 8 + (3 - index) * 2 + (frame < reference);
        // the maximum key is the first maximum of abs_diff * spikes in row-major order and carries its sign
        const uint32_t mu = au & ((fu >> 15) * 0xffffu), ml = al & ((fl2 >> 15) * 0xffffu);
        const uint32_t nu = (~((cu | 0x80008000u) - ru) >> 15) & 0x00010001u;  // lane bit: frame < reference
        const uint32_t nl = (~((cl | 0x80008000u) - rl) >> 15) & 0x00010001u;
        const uint32_t m2 = __vmaxu2(mu * 8u + 0x00040006u + nu, ml * 8u + 0x00000002u + nl);
        const uint32_t m = max(m2 & 0xffffu, m2 >> 16);
        const uint32_t idx = 3u - ((m >> 1) & 3u), half = idx & 1u;
        const uint32_t r_w = idx < 2u ? ru : rl;
        const uint32_t q = (uint32_t)qu + (idx < 2u ? 0u : (uint32_t)A.W) + 2u * w + half;
        wle[h][w] = q | (((r_w >> (16u * half)) & 0xffu) << 13) | ((m >> 3) << 21) | ((m & 1u) << 29);
        if (idx < 2u) ku |= 1u << (2 * w + half); else kl |= 1u << (2 * w + half);
      }
      flags8[ju] = ku; flags8[jl] = kl;
    }
  } else {
#pragma unroll
  for (int j = 0; j < J; ++j) {
    flags8[j] = 0;
    if (!have[j]) continue;
    uint32_t fl[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const uint32_t c = c_pk[j][w], r = rw(j, w);
      const uint32_t a = __vmaxu2(c, r) - __vminu2(c, r);  // no borrow: max >= min lane-wise
      if (ADAPT) fl[w] = ~((tw(j, w) | 0x80008000u) - a) & 0x80008000u;  // a > threshold lane-wise (both < 2^15)
      else fl[w] = (a + kadd) & 0x80008000u;                 // spike flags at bits 15 / 31
    }
    if (ADAPT && !LAZY_S) tmax = __vmaxu2(tmax, __vmaxu2(__vmaxu2(t_pk[j][0], t_pk[j][1]), __vmaxu2(t_pk[j][2], t_pk[j][3])));
    if (ST != ST_U8) rng |= (r_pk[j][0] | r_pk[j][1] | r_pk[j][2]) | r_pk[j][3];  // one-byte state cannot leave [0, 255]
    if (SRC != SRC_U8) rng |= (c_pk[j][0] | c_pk[j][1] | c_pk[j][2]) | c_pk[j][3];
    if (fl[0] | fl[1] | fl[2] | fl[3]) flags8[j] = flags4_of(fl[0], fl[1]) | (flags4_of(fl[2], fl[3]) << 4);
    if (THR_EARLY) {
      uint32_t tn[4];
#pragma unroll
      for (int w = 0; w < 4; ++w) tn[w] = thr_new(tw(j, w), msb_lane_mask(fl[w]));
      st_state8<ST>(thr_t + q0_of(j), tn);
    }
  }
  }
  if (ADAPT) {
    // a negative threshold (bit 15 as an unsigned lane), or -- coded rule -- one so large that th + up leaves int16
    // and wraps negative in the reference (gs:292): the generic body replays those tiles with exact int16 arithmetic
    // (LAZY_S: bytes, so only th + up can leave int16, and 255 stands in for the largest threshold)
    if (LAZY_S) tmax = 0x00ff00ffu;
    const uint32_t over = A.upd.uncoded ? tmax : (tmax | (tmax + (uint32_t)A.upd.thr_up * 0x00010001u));
    if (over & 0x80008000u) rng |= 0xFF00u;
  }
  if (QUAD) {
    // local_inhibition gs:177-221 for 4x4 areas on registers.  Group j is row j of the tile; words 0-1 / 2-3 of the four
    // groups form the left / right area of the column block.  argmax gs:118-132 keeps the FIRST maximum in row-major
    // order: with keys a * 16 + (15 - index) a plain maximum does that (a <= 255, so a key fits a 16-bit lane).
    if ((flags8[0] | flags8[1]) | (flags8[2] | flags8[3])) {
      uint32_t keep[4] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int sd = 0; sd < 2; ++sd) {
        uint32_t best = 0u;
#pragma unroll
        for (int r4 = 0; r4 < 4; ++r4) {
          if (flags8[r4] == 0) continue;   // no spike in this row of the block: its masked abs_diff is zero
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const uint32_t idx_lo = 15u - (4u * r4 + 2u * h2), idx_hi = idx_lo - 1u;
            best = __vmaxu2(best, masked_abs(r4, 2 * sd + h2) * 16u + (idx_lo | (idx_hi << 16)));
          }
        }
        const uint32_t m = max(best & 0xffffu, best >> 16);
        if ((m >> 4) != 0u) {  // some pixel of the area spikes: only the winner keeps its spike
          const uint32_t idx = 15u - (m & 15u), row = idx >> 2, bit = 1u << (4 * sd + (idx & 3u));
#pragma unroll
          for (int r4 = 0; r4 < 4; ++r4) keep[r4] |= row == (uint32_t)r4 ? bit : 0u;
        }
      }
#pragma unroll
      for (int r4 = 0; r4 < 4; ++r4) flags8[r4] = keep[r4];
    }
  } else if (PAIR && !FUSED2) {
    // local_inhibition gs:177-221 on registers: of each 2x2 area (word w of the upper and of the lower
    // row group) only the first arg-max of abs_diff * spikes keeps its spike (row-major, strict '>')
#pragma unroll
    for (int h = 0; h < JH; ++h) {
      const int ju = h, jl = JH + h;
      if ((flags8[ju] | flags8[jl]) == 0) continue;
      uint32_t ku = 0, kl = 0;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const uint32_t up = flags8[ju] ? masked_abs(ju, w) : 0u, lo = flags8[jl] ? masked_abs(jl, w) : 0u;
        if (FAST2_INH_KEYS && SRC == SRC_I16) {
        // keys abs_diff * 4 + (3 - index in the area): the maximum key is the FIRST maximum in row-major order
        // (abs_diff <= 255 on this path, so a key fits its 16-bit lane)
        const uint32_t m2 = __vmaxu2(up * 4u + 0x00020003u, lo * 4u + 0x00000001u);
        const uint32_t m = max(m2 & 0xffffu, m2 >> 16);
        if (m >> 2) {
          const uint32_t bi = 3u - (m & 3u);
          if (bi < 2) ku |= 1u << (2 * w + bi); else kl |= 1u << (2 * w + bi - 2);
        }
        } else {
        const int t00 = up & 0xffff, t01 = up >> 16, t10 = lo & 0xffff, t11 = lo >> 16;
        int best = t00, bi = 0;
        if (t01 > best) { best = t01; bi = 1; }
        if (t10 > best) { best = t10; bi = 2; }
        if (t11 > best) { best = t11; bi = 3; }
        if (best > 0) {
          if (bi < 2) ku |= 1u << (2 * w + bi); else kl |= 1u << (2 * w + bi - 2);
        }
        }
      }
      flags8[ju] = ku; flags8[jl] = kl;
    }
  }
  // update_reference_rate_adpt gs:291-297 on packed lanes: +up where the pixel still spikes, -down
  // elsewhere, clipped to [min, max]; written back for every pixel.  Deferred until the tile is known to
  // stay in this kernel (a declined tile must reach the generic body with its thresholds untouched).
  auto store_thresholds = [&]() {
#pragma unroll
    for (int j = 0; j < J; ++j) {
      if (!have[j]) continue;
      // lane masks of the four words from the 8 pixel flags: flag k -> bit 8k + 7 of a word (multiplication), then PRMT
      const uint32_t g01 = bytes_of_flags4(flags8[j] & 0xfu), g23 = bytes_of_flags4(flags8[j] >> 4);
      uint32_t tn[4];
      tn[0] = thr_new(tw(j, 0), prmt<0x9988u>(g01));
      tn[1] = thr_new(tw(j, 1), prmt<0xBBAAu>(g01));
      tn[2] = thr_new(tw(j, 2), prmt<0x9988u>(g23));
      tn[3] = thr_new(tw(j, 3), prmt<0xBBAAu>(g23));
      st_state8<ST>(thr_t + q0_of(j), tn);
    }
  };
  // history_weight != 1: the reference of every pixel decays (clip(trunc(hw * r) + 0, 0, 255)); candidates are
  // overwritten with their full update in step 4a.  Deferred like the thresholds (needs the LUT barrier too).
  auto store_decayed_reference = [&]() {
#pragma unroll
    for (int j = 0; j < J; ++j) {
      if (!have[j]) continue;
      uint32_t rn[4];
#pragma unroll
      for (int w = 0; w < 4; ++w)
        rn[w] = (uint32_t)tr_lut[rw(j, w) & 0xff] | ((uint32_t)tr_lut[(rw(j, w) >> 16) & 0xff] << 16);
      st_state8<ST>(ref_t + q0_of(j), rn);
    }
  };
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const uint32_t n = __popc(flags8[j]);
    if (j < 2) cnt_lo += n << (16 * (j & 1)); else cnt_hi += n << (16 * (j & 1));
  }

  // ---- 3. ordered work list: exclusive scan of the candidate counts in (group j, thread) order -----
  const bool w_unsafe = __any_sync(kFull, (rng & 0xFF00FF00u) != 0);
  const bool w_any = __any_sync(kFull, (cnt_lo | cnt_hi) != 0);
  uint32_t in_lo = cnt_lo, in_hi = cnt_hi;
  if (w_any) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y0 = __shfl_up_sync(kFull, in_lo, o), y1 = __shfl_up_sync(kFull, in_hi, o);
      if (lane >= o) { in_lo += y0; in_hi += y1; }
    }
  }
  if (lane == 31) {
    wtot[0 * NW + warp] = (in_lo & 0xffff) | (w_unsafe ? 0x80000000u : 0u);  // bit 31: a pixel outside [0, 255]
    if (J > 1) wtot[1 * NW + warp] = in_lo >> 16;
    if (J > 2) { wtot[2 * NW + warp] = in_hi & 0xffff; wtot[3 * NW + warp] = in_hi >> 16; }
  }
  __syncthreads();
  const uint32_t ent_raw = lane < J * NW ? wtot[lane] : 0u;  // every warp reads the <= 32 totals
  // QUAD: up to 16 warps x 4 rows = 64 (row, warp) totals, two per lane
  const uint32_t ent_raw2 = (QUAD && lane + 32 < J * NW) ? wtot[lane + 32] : 0u;
  const bool unsafe = __any_sync(kFull, ((ent_raw | ent_raw2) >> 31) != 0u);
  const uint32_t ent = ent_raw & 0x7fffffffu, ent2 = ent_raw2 & 0x7fffffffu;
  const int n_cand = (int)__reduce_add_sync(kFull, ent + ent2);  // redux.sync: the tile's candidate count
  if (n_cand == 0 && !unsafe) {  // quiet tile: no records, no events, reference untouched (unless it decays)
    if (ADAPT && !THR_EARLY) store_thresholds();
    if (decay) store_decayed_reference();
    for (int c = tid; c < A.C; c += blockDim.x)
      A.tile_counts[((size_t)s_idx * A.C + c) * A.tiles_per_stream + t_idx] = 0;
    return;
  }
  if (unsafe || n_cand > cap) {  // generic body: exact int16 wrap semantics / dense tiles
    if (THR_EARLY) {  // ... which must find the thresholds as they were
#pragma unroll
      for (int j = 0; j < J; ++j)
        if (have[j]) {
          if (LAZY_S) *reinterpret_cast<uint2 *>(thr_t + q0_of(j)) = make_uint2(t_pk[ADAPT ? j : 0][0], t_pk[ADAPT ? j : 0][1]);
          else st_state8<ST>(thr_t + q0_of(j), t_pk[ADAPT ? j : 0]);
        }
    }
    if (tid == 0) A.redo[2 + atomicAdd(A.redo, 1)] = tile;
    return;
  }
  if (ADAPT && !TLATE && !THR_EARLY) store_thresholds();
  if (decay) store_decayed_reference();
  if (w_any) {  // (a warp without candidates has nothing to push)
    // entry: tile-local pixel index (13 bits) | reference << 13 (8 bits) | abs_diff << 21 (8 bits) | negative << 29
    const uint32_t ex_lo = in_lo - cnt_lo, ex_hi = in_hi - cnt_hi;
    auto entry_of = [&](int q, int c, int r) -> uint32_t {
      // the entry carries the reference the update starts from: trunc(history_weight * r)
      return (uint32_t)q | ((uint32_t)(decay ? (int)tr_lut[r] : r) << 13) | ((uint32_t)(c < r ? r - c : c - r) << 21) |
             (c < r ? 1u << 29 : 0u);
    };
    uint32_t pos_of[J];
#pragma unroll
    for (int j = 0; j < J; ++j)
      pos_of[j] = __reduce_add_sync(kFull, (lane < j * NW + warp ? ent : 0u) + (lane + 32 < j * NW + warp ? ent2 : 0u)) +
                  (((j < 2 ? ex_lo : ex_hi) >> (16 * (j & 1))) & 0xffffu);  // + candidates of earlier (group, warp) pairs
    if (FUSED2) {
      // the entries were finished with the inhibition: store them at their ranks (at most one per area)
#pragma unroll
      for (int h = 0; h < JH; ++h) {
        const int ju = h, jl = JH + h;
        if ((flags8[ju] | flags8[jl]) == 0) continue;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const uint32_t e = wle[h][w];
          if (e == 0u) continue;
          if ((flags8[ju] >> (2 * w)) & 3u) wl[pos_of[ju]++] = e; else wl[pos_of[jl]++] = e;
        }
      }
    } else if (QONLY && PAIR && !QUAD) {
      // index-only entries: one 16-bit store per surviving area
#pragma unroll
      for (int h = 0; h < JH; ++h) {
        const int ju = h, jl = JH + h;
        if ((flags8[ju] | flags8[jl]) == 0) continue;
        const int qu = q0_of(ju), ql = q0_of(jl);
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const uint32_t bu = (flags8[ju] >> (2 * w)) & 3u, bl = (flags8[jl] >> (2 * w)) & 3u;
          if ((bu | bl) == 0) continue;
          const bool up = bu != 0;
          const int half = (int)((up ? bu : bl) >> 1);
          const uint32_t pos = up ? pos_of[ju] : pos_of[jl];
          if (ADAPT && !TLATE) wl_thr[pos] = (uint16_t)((up ? t_pk[ADAPT ? ju : 0][w] : t_pk[ADAPT ? jl : 0][w]) >> (16 * half));
          wl16[pos] = (uint16_t)((up ? qu : ql) + 2 * w + half);
          if (up) pos_of[ju] = pos + 1; else pos_of[jl] = pos + 1;
        }
      }
    } else if (QONLY) {
#pragma unroll
      for (int j = 0; j < J; ++j) {
        if (flags8[j] == 0) continue;
        const int q0 = q0_of(j);
        uint32_t pos = pos_of[j];
        for (uint32_t m = flags8[j]; m; m &= m - 1) {  // candidates only
          const int p = __ffs(m) - 1;
          if (ADAPT && !TLATE) wl_thr[pos] = (uint16_t)t_px(j, p);
          wl16[pos++] = (uint16_t)(q0 + p);
        }
      }
    } else if (PAIR) {
      // after the 2x2 inhibition an area (word w of the upper and of the lower row group) holds at most one
      // candidate: walk the four areas of a column block once, instead of the set bits of each row group
#pragma unroll
      for (int h = 0; h < JH; ++h) {
        const int ju = h, jl = JH + h;
        if ((flags8[ju] | flags8[jl]) == 0) continue;
        const int qu = q0_of(ju), ql = q0_of(jl);
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const uint32_t bu = (flags8[ju] >> (2 * w)) & 3u, bl = (flags8[jl] >> (2 * w)) & 3u;
          if ((bu | bl) == 0) continue;
          const bool up = bu != 0;
          const int half = (int)((up ? bu : bl) >> 1), sh = 16 * half;
          const int c = (int)(((up ? c_pk[ju][w] : c_pk[jl][w]) >> sh) & 0xffffu);
          const int r = (int)(((up ? r_pk[ju][w] : r_pk[jl][w]) >> sh) & 0xffffu);
          const uint32_t pos = up ? pos_of[ju] : pos_of[jl];
          if (ADAPT) wl_thr[pos] = (uint16_t)((up ? t_pk[ADAPT ? ju : 0][w] : t_pk[ADAPT ? jl : 0][w]) >> sh);
          wl[pos] = entry_of((up ? qu : ql) + 2 * w + half, c, r);
          if (up) pos_of[ju] = pos + 1; else pos_of[jl] = pos + 1;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < J; ++j) {
        if (flags8[j] == 0) continue;
        const int q0 = q0_of(j);
        uint32_t pos = pos_of[j];
        const unsigned long long c_lo = ((unsigned long long)c_pk[j][1] << 32) | c_pk[j][0], c_hi = ((unsigned long long)c_pk[j][3] << 32) | c_pk[j][2];
        for (uint32_t m = flags8[j]; m; m &= m - 1) {  // candidates only
          const int p = __ffs(m) - 1;
          const int sh = 16 * (p & 3);
          const int c = (int)(((p < 4 ? c_lo : c_hi) >> sh) & 0xffff), r = r_px(j, p);
          if (ADAPT) wl_thr[pos] = (uint16_t)t_px(j, p);
          wl[pos++] = entry_of(q0 + p, c, r);
        }
      }
    }
  }
  __syncthreads();

  // ---- 4a. one candidate per lane: reference update, list membership, histogram, staged record ------
  // 32-candidate chunks per warp = ceil(ceil(n_cand / 32) / NW), without an integer division by the run-time NW
  // (inv_nw = ceil(2^16 / NW), set by the launcher: exact for the <= 64 + NW chunks a tile can have)
  const int chunks_per_warp = (int)(((unsigned)(((n_cand + 31) >> 5) + NW - 1) * A.inv_nw) >> 16);
  const int per_warp = chunks_per_warp << 5;
  const int lo = min(n_cand, warp * per_warp), hi = min(n_cand, lo + per_warp);
  const unsigned magic = A.upd.div_thr.magic;
  const int max_n = A.upd.max_time_ms;
  const unsigned lt = (1u << lane) - 1u;
  unsigned long long hp = 0, hn = 0;
  unsigned status = 0;
  uint32_t w_pos = 0, w_neg = 0;
  for (int i0 = lo; i0 < hi; i0 += 32) {
    const int i = i0 + lane;
    bool tp = false, tn = false;
    uint2 rec = make_uint2(0u, 0u);
    if (i < hi) {
      int q, r, a;
      bool neg;
      if (QONLY) {  // index-only entry: fetch the pixel again (both values are in [0, 255]: the tile passed the range vote)
        q = wl16[i];
        // the tile's frame pointer is rebuilt from the reference pointer (kernel parameters are free, live registers are not)
        const long long poff = (ref_t - static_cast<const state_t *>(A.ref)) + q;  // pixel index in the planes
        const unsigned char *cur_q = static_cast<const unsigned char *>(A.curr) + (SRC == SRC_U8 ? poff : 2 * poff);
        const int c = SRC == SRC_U8 ? (int)__ldg(cur_q) : (int)__ldg(reinterpret_cast<const uint16_t *>(cur_q));
        r = state_px<ST>(ref_t, q);
        neg = c < r;
        a = neg ? r - c : c - r;
      } else {
        const uint32_t e = wl[i];
        q = e & 0x1fff; r = (e >> 13) & 0xff; a = (e >> 21) & 0xff;
        neg = (e >> 29) & 1u;
      }
      int y, x;
      if (FAST2_PAIR_XY && SRC == SRC_I16 && PAIR && !QUAD) { y = q >= A.W ? 1 : 0; x = q - (y ? A.W : 0); }  // two-row tiles
      else { y = (int)__umulhi((unsigned)q, A.inv_w); x = q - y * A.W; }
      int upd;
      if (ADAPT) {  // gs:287-289: (a // thr) [* min_threshold below]
        const int th_old = TLATE ? state_px<ST>(thr_t, q) : (int)wl_thr[i];
        upd = SIMPLE ? adpt_div(a, th_old) : np_floordiv(a, th_old);
      }
      else if (SIMPLE || A.upd.mode == DVS_UPD_RATE) upd = min((int)__umulhi((unsigned)a, magic), max_n) * T;  // gs:244-249
      else {  // time-binary updates gs:374-376 / gs:421-423: log2 table of abs_diff (raw) or of abs_diff // T
        const int idx = A.upd.mode == DVS_UPD_TIME_BIN_THR ? (int)__umulhi((unsigned)a, magic) : a;
        upd = 0;
        if (idx >= A.upd.lut_len) status |= DVS_STATUS_LUT_INDEX;
        else upd = (int)__ldg(A.upd.lut + idx) * (A.upd.mode == DVS_UPD_TIME_BIN_THR ? T : 1);
      }
      if (ADAPT) {  // gs:289: int16 products and sums wrap before the clip (thresholds below min_threshold can get there)
        const int u = w16((neg ? -upd : upd) * A.upd.min_thr);
        ref_t[q] = (state_t)clip(w16(r + u), 0, 255);
      } else ref_t[q] = (state_t)(neg ? max(r - upd, 0) : min(r + upd, 255));
      int desc = 0;
      if (ENC == ENC_RATE_CODED) {
        tp = !neg; tn = neg;
        desc = max(0, (A.enc.n_slots - 1) - (int)__umulhi((unsigned)a, A.enc.dv.magic));  // k, gs:830-832
        if (desc == 0 && A.drop_silent) { tp = false; tn = false; }  // k == 0: in no slot; the record would only be skipped by K3
        else atomicAdd(&hist[(neg ? nh : 0) + desc], 1u);  // fire-and-forget shared-memory reduction (k <= n_slots - 1)
      } else if (ENC == ENC_TIME_CODED) {
        // make_spike_lists_time gs:903-926: one event in slot num_bins - min(v - 1, num_bins - 1) - 1, the neg list
        // additionally clamps at 0 (gs:919-921); a pos spike with v == 0 lands outside the list of lists
        tp = !neg; tn = neg;
        const int v = (int)__umulhi((unsigned)a, A.enc.dv.magic);
        int nt = min(v - 1, A.enc.n_slots - 1);
        if (neg) nt = max(nt, 0);
        desc = A.enc.n_slots - nt - 1;
        if (desc >= A.enc.n_slots) { status |= DVS_STATUS_SLOT_RANGE; desc = -1; }
        else atomicAdd(&hist[(neg ? nh : 0) + desc], 1u);
      } else {
        split_px(neg ? -1 : 1, A.polarity, A.uncoded, tp, tn);
        if (tp || tn) desc = hist_add(A, hist, nh, a, tp, hp, hn, status, /*allow_regs=*/false);
      }
      rec = make_uint2((uint32_t)(row0 + y) | ((uint32_t)x << 16), rec_word1(a, desc, A.enc.mode != DVS_ENC_NONE));
    }
    const unsigned bp = __ballot_sync(kFull, tp), bn = __ballot_sync(kFull, tn);
    if (tp) rec_sm[lo + w_pos + __popc(bp & lt)] = rec;
    if (tn) rec_sm[hi - 1 - (w_neg + __popc(bn & lt))] = rec;
    w_pos += __popc(bp);
    w_neg += __popc(bn);
  }
  if (lane == 0) wtot[64 + warp] = w_pos | (w_neg << 16);
  if (ENC != ENC_RATE_CODED) {  // only the status bits are left to publish: the histogram went straight to shared memory
    status = __reduce_or_sync(kFull, status);
    if (status && lane == 0 && A.status) atomicOr(A.status, (int)status);
  }
  __syncthreads();
  if (TLATE) store_thresholds();  // every candidate has read its old threshold by now
  uint32_t base = 0, totals = 0;
  {
    const uint32_t t = lane < NW ? wtot[64 + lane] : 0u;  // pos | neg << 16 per warp: the packed sums cannot carry
    base = __reduce_add_sync(kFull, lane < warp ? t : 0u);
    totals = __reduce_add_sync(kFull, t);
  }
  const int n_pos_tile = totals & 0xffff, n_neg_tile = totals >> 16;

  // ---- 4b. coalesced copy of the warp's staged records to their ranks in the tile's lists ----------
  if (w_pos | w_neg) {
    const size_t region = (size_t)tile * A.rec_stride;
    dvs_record *pos_dst = A.records + region + (base & 0xffff);
    dvs_record *neg_dst = A.records + region + (A.rec_stride - n_neg_tile) + (base >> 16);
    for (int i = lane; i < (int)w_pos; i += 32) *reinterpret_cast<uint2 *>(pos_dst + i) = rec_sm[lo + i];
    for (int i = lane; i < (int)w_neg; i += 32) *reinterpret_cast<uint2 *>(neg_dst + i) = rec_sm[hi - 1 - i];
  }
  write_counters(A, s_idx, t_idx, hist, nh, n_pos_tile, n_neg_tile);
}

#ifndef FAST2_MINB_ADPT
#define FAST2_MINB_ADPT 4
#endif
template <int SRC, int JH, bool PAIR, int ENC, bool ADAPT = false, int ST = ST_I16>
__global__ void __launch_bounds__(256, ADAPT ? (FAST2_LAZY_STATE && ST == ST_U8 && !PAIR ? FAST2_MINB_ADPT_S8 : FAST2_MINB_ADPT) : FAST2_MINB) k_tile_fast2(const __grid_constant__ TileArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  pdl_trigger();
  fast2_tile<SRC, JH, PAIR, ENC, ADAPT, false, false, ST>(A, (int)blockIdx.x, (int)blockIdx.y, smem_raw, smem_raw + A.fast_cap * 12,
                                                  nullptr, nullptr);
}

// 4x4 inhibition: one column block per thread over four rows; MAXT = 256 (rows up to 2048 px) or 512 (up to 4096 px)
template <int SRC, int ENC, bool ADAPT, int MAXT, int ST = ST_I16>
__global__ void __launch_bounds__(MAXT, MAXT == 256 ? 4 : 2) k_tile_fast4(const __grid_constant__ TileArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  pdl_trigger();
  fast2_tile<SRC, 1, true, ENC, ADAPT, false, true, ST>(A, (int)blockIdx.x, (int)blockIdx.y, smem_raw, smem_raw + A.fast_cap * 12,
                                                     nullptr, nullptr);
}

// ---- persistent, staged variant (opt-in: dvs_step_params.force_generic == 3) -----------------------
// Measured on 4K x 256 streams: 2.61 ms against 2.08 ms for one tile per CTA (3 CTAs/SM = 24 warps cannot cover
// the barrier / ALU latencies of the compute phases that 5 CTAs/SM = 40 warps do; a single-stage form at 4 CTAs/SM
// took 2.97 ms).  Kept for shapes / future parts where DRAM latency rather than issue slots is the limit.
// One CTA per resident slot (grid = SMs x CTAs/SM) walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...  The rows of
// tile k+1 (frame and reference: two contiguous runs of global memory) are fetched by one elected thread with
// cp.async.bulk into the other half of a two-stage shared-memory ring while tile k is processed; completion is
// signalled through an mbarrier (expect_tx / complete_tx), so no warp ever waits on DRAM with registers reserved
// for in-flight loads.  shared memory: [stage 0][stage 1][wtot 128 u32][hist][tr_lut][2 mbarriers].
DEV uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
DEV void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
DEV void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
DEV void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
DEV void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

#ifndef FAST2S_MINB
#define FAST2S_MINB 3
#endif
template <int SRC, int JH, bool PAIR, int ENC>
__global__ void __launch_bounds__(256, FAST2S_MINB) k_tile_fast2s(const __grid_constant__ TileArgs A, const int stage_bytes) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  pdl_trigger();
  unsigned char *ctl = smem_raw + 2 * stage_bytes;
  const int nh = A.enc.mode == DVS_ENC_NONE ? 0 : (A.enc.n_slots + 1);
  uint64_t *bars = reinterpret_cast<uint64_t *>(ctl + (((128 + 2 * nh + 8) * 4 + 256 + 15) & ~15));
  const long long n_tiles = (long long)A.S * A.tiles_per_stream;
  const int elem = SRC == SRC_U8 ? 1 : 2;
  const int cur_bytes_max = A.tile_px * elem;  // the reference rows follow the frame rows inside a stage

  auto issue = [&](long long tile, int stage) {  // one thread
    const int s_idx = (int)(tile / A.tiles_per_stream), t_idx = (int)(tile - (long long)s_idx * A.tiles_per_stream);
    const int row0 = t_idx * A.tile_rows;
    const int npx = min(A.tile_rows, A.H - row0) * A.W;
    const size_t plane_off = ((size_t)s_idx * A.H + row0) * A.W;
    unsigned char *dst = smem_raw + (size_t)stage * stage_bytes;
    mbar_expect_tx(&bars[stage], (uint32_t)(npx * (elem + 2)));
    bulk_g2s(dst, static_cast<const unsigned char *>(A.curr) + plane_off * elem, (uint32_t)(npx * elem), &bars[stage]);
    bulk_g2s(dst + cur_bytes_max, reinterpret_cast<const unsigned char *>(static_cast<const int16_t *>(A.ref) + plane_off), (uint32_t)(npx * 2), &bars[stage]);
  };

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  long long tile = blockIdx.x;
  if (threadIdx.x == 0 && tile < n_tiles) issue(tile, 0);
  for (int k = 0; tile < n_tiles; ++k, tile += gridDim.x) {
    const int stage = k & 1;
    if (threadIdx.x == 0 && tile + gridDim.x < n_tiles) {
      // the other stage was last touched through the generic proxy (work list / records of tile k - 1)
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(tile + gridDim.x, stage ^ 1);
    }
    mbar_wait(&bars[stage], (uint32_t)((k >> 1) & 1));
    const int s_idx = (int)(tile / A.tiles_per_stream), t_idx = (int)(tile - (long long)s_idx * A.tiles_per_stream);
    unsigned char *st = smem_raw + (size_t)stage * stage_bytes;
    fast2_tile<SRC, JH, PAIR, ENC, false, true>(A, t_idx, s_idx, st, ctl, st, st + cur_bytes_max);
    __syncthreads();  // tile k is finished with its stage, wtot and hist before anything of tile k + 1 touches them
  }
}

template <int SRC, int JH, bool PAIR, int ENC, bool ADAPT = false, int ST = ST_I16>
static int launch_fast2(const TileArgs &A_in, int bd, cudaStream_t st) {
  TileArgs A = A_in;
  A.inv_nw = (65536u + (unsigned)(bd / 32) - 1u) / (unsigned)(bd / 32);
  A.fast_cap = fast_cap_of(A);
  auto k = k_tile_fast2<SRC, JH, PAIR, ENC, ADAPT, ST>;
  const int nh = A.enc.mode == DVS_ENC_NONE ? 0 : (A.enc.n_slots + 1);
  const size_t smem = (size_t)A.fast_cap * 12 + (128 + 2 * nh + 8) * sizeof(uint32_t) + (ADAPT ? A.fast_cap * 2 : 0) + 256;
  cudaError_t e;
  if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
    return set_error((int)e, cudaGetErrorString(e));
  const long long tiles = (long long)A.S * A.tiles_per_stream;
  bool staged = false;
  if constexpr (!ADAPT && PAIR && ENC != ENC_TIME_CODED && ST == ST_I16) {
    // persistent staged kernel: worth it once every resident CTA has several tiles to walk
    const int stage_bytes = (A.tile_px * ((SRC == SRC_U8 ? 1 : 2) + 2) + 127) & ~127;
    const size_t smem_s = 2 * (size_t)stage_bytes + (((128 + 2 * nh + 8) * 4 + 256 + 15) & ~15) + 16;
    static int sm_count = 0;
    if (sm_count == 0) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    }
    const long long resident = (long long)sm_count * FAST2S_MINB;
    if (A.use_staged == 2 && stage_bytes >= kFastCap * 12 && smem_s <= 72 * 1024) {  // opt-in: measured slower than one tile per CTA
      auto ks = k_tile_fast2s<SRC, JH, PAIR, ENC>;
      if ((e = cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s)) != cudaSuccess)
        return set_error((int)e, cudaGetErrorString(e));
      ks<<<(unsigned)resident, bd, smem_s, st>>>(A, stage_bytes);
      staged = true;
    }
  }
  if (!staged) k<<<dim3((unsigned)A.tiles_per_stream, (unsigned)A.S), bd, smem, st>>>(A);
  if ((e = cudaGetLastError()) != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
  // generic body over the (normally zero) declined tiles; its group-to-thread mapping is the linear one
  constexpr int JR = (PAIR ? 2 * JH : JH) > 1 ? 4 : 1;
  constexpr int KR = PAIR ? INH_K2 : INH_NONE;
  auto kr = k_tile_redo<SRC, JR, KR, ADAPT>;
  const size_t smem_r = ((KR != INH_NONE) ? ((A.tile_px * 2 + 15) & ~15) : 0) + (128 + 2 * nh + 8) * sizeof(uint32_t);
  if (smem_r > 48 * 1024 && (e = cudaFuncSetAttribute(kr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r)) != cudaSuccess)
    return set_error((int)e, cudaGetErrorString(e));
  const int items = A.tile_px / 8;
  const int bd_r = JR == 1 ? (items + 31) / 32 * 32 : ((items + 3) / 4 + 31) / 32 * 32;
  e = launch_pdl(kr, dim3((unsigned)(tiles < 592 ? tiles : 592)), dim3((unsigned)bd_r), smem_r, st, A);
  if (e != cudaSuccess || (e = cudaGetLastError()) != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
  return DVS_OK;
}

template <int SRC, int ENC, bool ADAPT, int ST = ST_I16>
static int launch_fast4(const TileArgs &A_in, cudaStream_t st) {
  TileArgs A = A_in;
  A.inv_nw = (65536u + (unsigned)((A.W / 8 + 31) / 32) - 1u) / (unsigned)((A.W / 8 + 31) / 32);
  A.fast_cap = fast_cap_of(A);
  const int nh = A.enc.n_slots + 1;
  const size_t smem = (size_t)A.fast_cap * 12 + (128 + 2 * nh + 8) * sizeof(uint32_t) + (ADAPT ? A.fast_cap * 2 : 0) + 256;
  cudaError_t e;
  const int gpr = A.W / 8, bd = (gpr + 31) / 32 * 32;
  const dim3 grid((unsigned)A.tiles_per_stream, (unsigned)A.S);
  if (bd <= 256) {
    auto k = k_tile_fast4<SRC, ENC, ADAPT, 256, ST>;
    if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
      return set_error((int)e, cudaGetErrorString(e));
    k<<<grid, bd, smem, st>>>(A);
  } else {
    auto k = k_tile_fast4<SRC, ENC, ADAPT, 512, ST>;
    if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
      return set_error((int)e, cudaGetErrorString(e));
    k<<<grid, bd, smem, st>>>(A);
  }
  if ((e = cudaGetLastError()) != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
  // generic body (shared-memory plane, any inhibition width) over the declined tiles
  auto kr = k_tile_redo<SRC, 4, INH_GENERIC, ADAPT, 1024>;
  const size_t smem_r = ((A.tile_px * 2 + 15) & ~15) + (128 + 2 * nh + 8) * sizeof(uint32_t);
  if (smem_r > 48 * 1024 && (e = cudaFuncSetAttribute(kr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r)) != cudaSuccess)
    return set_error((int)e, cudaGetErrorString(e));
  const int items = A.tile_px / 8;
  const long long tiles = (long long)A.S * A.tiles_per_stream;
  e = launch_pdl(kr, dim3((unsigned)(tiles < 296 ? tiles : 296)), dim3((unsigned)(((items + 3) / 4 + 31) / 32 * 32)), smem_r, st, A);
  if (e != cudaSuccess || (e = cudaGetLastError()) != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
  return DVS_OK;
}

// the ranked form of the headline shape (dvs_tile_ranked.cuh; the translation units include that header)
#ifndef FAST2_RANKED
#define FAST2_RANKED 1
#endif
#ifndef RANKED_BD512
#define RANKED_BD512 0
#endif
template <int SRC, int JH, int ENC, int ST, int BD>
static int launch_ranked(const TileArgs &A_in, int bd, cudaStream_t st);

// Caller guarantees: vec_ok (incl. the threshold plane when ADAPT), W % 8 == 0, <= 1024 groups per tile, inhibition
// off or 2x2 on two-row tiles; static: 2 <= T, rate or time-binary update, history_weight in [0, 1], max_time_ms >= 0;
// adaptive: rate_adpt update, history_weight in [0, 1], 0 <= min <= max <= 32767, 0 <= up, down <= 32767; no debug planes.
template <int SRC, bool ADAPT, int ST = ST_I16>
static int dispatch_fast_shapes(const TileArgs &A, cudaStream_t st) {
  const int items = A.tile_px / 8;
  // inlined membership / slot / histogram and index-only work list: MERGED lists of the coded module, coded RATE
  // (<= 8 slots) or TIME encoder, plain rate update (rate_adpt when adaptive) without history decay
  const bool coded = A.uncoded == 0 && A.polarity == DVS_POL_MERGED && A.enc.uncoded == 0 && A.enc.threshold >= 2 &&
                     A.enc.u16_vals == 0 && A.upd.hw_is_one &&
                     (ADAPT ? (A.upd.mode == DVS_UPD_RATE_ADPT && A.upd.uncoded == 0) : A.upd.mode == DVS_UPD_RATE);
  int enc = ENC_ANY;
  if (coded && A.enc.mode == DVS_ENC_RATE && A.enc.n_slots >= 1 && A.enc.n_slots <= 8 && A.fast_hist) enc = ENC_RATE_CODED;
  else if (coded && A.enc.mode == DVS_ENC_TIME && A.enc.n_slots >= 1) enc = ENC_TIME_CODED;
  auto round32 = [](int x) { return (x + 31) / 32 * 32; };
  if (A.inh_k == 4) {  // caller: four-row tiles, W <= 4096, enc != ENC_ANY (dvs_step_fused checks fast4_ok())
    if (enc == ENC_RATE_CODED) return launch_fast4<SRC, ENC_RATE_CODED, ADAPT, ST>(A, st);
    if (enc == ENC_TIME_CODED) return launch_fast4<SRC, ENC_TIME_CODED, ADAPT, ST>(A, st);
    return set_error(DVS_ERR_UNSUPPORTED, "internal: 4x4 fast path needs an inlined encoder");
  }
#define DVS_LAUNCH_ENC(JH_, PAIR_, BD_)                                                                      \
  (enc == ENC_RATE_CODED   ? launch_fast2<SRC, JH_, PAIR_, ENC_RATE_CODED, ADAPT, ST>(A, BD_, st)             \
   : enc == ENC_TIME_CODED ? launch_fast2<SRC, JH_, PAIR_, ENC_TIME_CODED, ADAPT, ST>(A, BD_, st)             \
                           : launch_fast2<SRC, JH_, PAIR_, ENC_ANY, ADAPT, ST>(A, BD_, st))
  if (A.inh_k != 2) {
    if (items <= 256) return DVS_LAUNCH_ENC(1, false, round32(items));
#if FAST2_J2
    if (items <= 512) return DVS_LAUNCH_ENC(2, false, round32((items + 1) / 2));  // two groups per thread: half the registers
#endif
    return DVS_LAUNCH_ENC(4, false, round32((items + 3) / 4));
  }
  const int gpr = A.W / 8;  // two-row tiles: each thread owns the same column block(s) in both rows
  if constexpr (FAST2_RANKED && !ADAPT) {
    if (enc != ENC_ANY && A.use_staged == 0) {
      if (gpr <= 256)
        return enc == ENC_RATE_CODED ? launch_ranked<SRC, 1, ENC_RATE_CODED, ST, 256>(A, round32(gpr), st)
                                     : launch_ranked<SRC, 1, ENC_TIME_CODED, ST, 256>(A, round32(gpr), st);
      if constexpr (RANKED_BD512 && ST == ST_U8) {
        if (gpr <= 512)
          return enc == ENC_RATE_CODED ? launch_ranked<SRC, 1, ENC_RATE_CODED, ST, 512>(A, round32(gpr), st)
                                       : launch_ranked<SRC, 1, ENC_TIME_CODED, ST, 512>(A, round32(gpr), st);
      }
      return enc == ENC_RATE_CODED ? launch_ranked<SRC, 2, ENC_RATE_CODED, ST, 256>(A, round32((gpr + 1) / 2), st)
                                   : launch_ranked<SRC, 2, ENC_TIME_CODED, ST, 256>(A, round32((gpr + 1) / 2), st);
    }
  }
  if (gpr <= 256) return DVS_LAUNCH_ENC(1, true, round32(gpr));
  return DVS_LAUNCH_ENC(2, true, round32((gpr + 1) / 2));
#undef DVS_LAUNCH_ENC
}

template <int SRC, int ST = ST_I16>
static int dispatch_fast(const TileArgs &A, cudaStream_t st) { return dispatch_fast_shapes<SRC, false, ST>(A, st); }

template <int SRC, int ST = ST_I16>
static int dispatch_fast_adpt(const TileArgs &A, cudaStream_t st) { return dispatch_fast_shapes<SRC, true, ST>(A, st); }

}  // namespace dvs
