// dvs_tile_pass.cuh -- K1, the fused per-tile pass (template + launcher).  Instantiated per
// (frame dtype, adaptive) in tile_pass_*.cu so that the translation units compile in parallel.
#pragma once
#include <limits.h>

#include <utility>

#include "dvs_device.cuh"

#ifndef GENERIC_J1
#define GENERIC_J1 1   // generic body with one group per thread and CTAs of up to 1024 threads: a quarter of the code of the
                       // four-group form, whose problem is instruction-cache misses (profiles/r2_k1_generic_*): dense 4K
                       // x 16 streams 1.94 -> 1.53 ms, moving bar 1.27 -> 1.23 ms
#endif

namespace dvs {

// ================================================================================================
// K1: tile pass
// ================================================================================================
enum { SRC_I16 = 0, SRC_U8 = 1, SRC_PLANES = 2 };
enum { INH_NONE = 0, INH_K2 = 2, INH_GENERIC = -1 };
// layout of the resident state planes (reference, thresholds): int16 as the reference's API holds them, or one byte per
// pixel (dvs_step_params.state_dtype == DVS_STATE_U8: every update rule clips the reference to [0, 255] and the coded
// threshold rule to [min, max], so with 0 <= min <= max <= 255 nothing is lost and the pass moves half the state bytes)
enum { ST_I16 = 0, ST_U8 = 1 };

struct TileArgs {
  int S, H, W, tile_rows, tiles_per_stream, tile_px, C;
  int rec_stride;          // records reserved per tile region (tile_px, or the inhibition bound tile_px / k^2)
  int adaptive, polarity, uncoded, inh_k, vec_ok, fast_hist, drop_silent, use_staged;
  int thr_scalar;
  int state_u8;            // ref / thr planes hold one byte per pixel (ST_U8)
  unsigned inv_w;          // ceil(2^32 / W): q / W == umulhi(q, inv_w) for q < 32768
  unsigned inv_nw;         // fast kernels: ceil(2^16 / warps per CTA)
  int fast_cap;            // fast kernels: work-list capacity of a tile (<= kFastCap; the most candidates a tile can hold)
  int silent_a;            // ranked kernel: abs_diff from which a coded-rate spike falls in no slot (and is not recorded)
  UpdateArgs upd;
  EncArgs enc;
  const void *curr;        // frames (SRC_I16 / SRC_U8) or spikes plane (SRC_PLANES)
  const int16_t *abs_in;   // SRC_PLANES
  void *ref, *thr;         // state planes, int16 or uint8 (state_u8)
  int16_t *diff_out, *abs_out, *spk_out;
  dvs_record *records;
  uint32_t *tile_counts;
  int32_t *status;
  int32_t *global_max;     // [S][2] (pos list, neg list), SRC_PLANES only; INT_MIN = empty
  int32_t *redo;           // [2 + tiles]: count, CTA ticket, tiles the fast kernel hands to the generic body
};

DEV uint4 ld16(const void *p) { return *reinterpret_cast<const uint4 *>(p); }
DEV void st16(void *p, const uint32_t (&v)[4]) {
  *reinterpret_cast<uint4 *>(p) = make_uint4(v[0], v[1], v[2], v[3]);
}

// Loads 8 consecutive int16 pixels (nvalid of them in range) as 4 packed words.
DEV void load8_i16(const int16_t *base, size_t off, int nvalid, bool vec, uint32_t (&pk)[4]) {
  if (vec && nvalid == 8) {
    uint4 v = ld16(base + off);
    pk[0] = v.x; pk[1] = v.y; pk[2] = v.z; pk[3] = v.w;
  } else {
    pk[0] = pk[1] = pk[2] = pk[3] = 0;
#pragma unroll
    for (int p = 0; p < 8; ++p)
      if (p < nvalid) set16(pk, p, base[off + p]);
  }
}
DEV void load8_u8(const uint8_t *base, size_t off, int nvalid, bool vec, uint32_t (&pk)[4]) {
  if (vec && nvalid == 8) {
    uint2 v = *reinterpret_cast<const uint2 *>(base + off);
    pk[0] = __byte_perm(v.x, 0, 0x4140); pk[1] = __byte_perm(v.x, 0, 0x4342);
    pk[2] = __byte_perm(v.y, 0, 0x4140); pk[3] = __byte_perm(v.y, 0, 0x4342);
  } else {
    pk[0] = pk[1] = pk[2] = pk[3] = 0;
#pragma unroll
    for (int p = 0; p < 8; ++p)
      if (p < nvalid) set16(pk, p, base[off + p]);
  }
}
DEV void store8_i16(int16_t *base, size_t off, int nvalid, bool vec, const uint32_t (&pk)[4]) {
  if (vec && nvalid == 8) st16(base + off, pk);
  else {
#pragma unroll
    for (int p = 0; p < 8; ++p)
      if (p < nvalid) base[off + p] = (int16_t)get16(pk, p);
  }
}

DEV void store8_u8(uint8_t *base, size_t off, int nvalid, bool vec, const uint32_t (&pk)[4]) {
  if (vec && nvalid == 8)
    *reinterpret_cast<uint2 *>(base + off) = make_uint2(__byte_perm(pk[0], pk[1], 0x6420), __byte_perm(pk[2], pk[3], 0x6420));
  else {
#pragma unroll
    for (int p = 0; p < 8; ++p)
      if (p < nvalid) base[off + p] = (uint8_t)get16(pk, p);
  }
}
// state planes in either layout (generic body: a uniform run-time branch; the fast kernels are templated on ST_*)
DEV void load8_state(const void *base, bool u8, size_t off, int nvalid, bool vec, uint32_t (&pk)[4]) {
  if (u8) load8_u8(static_cast<const uint8_t *>(base), off, nvalid, vec, pk);
  else load8_i16(static_cast<const int16_t *>(base), off, nvalid, vec, pk);
}
DEV void store8_state(void *base, bool u8, size_t off, int nvalid, bool vec, const uint32_t (&pk)[4]) {
  if (u8) store8_u8(static_cast<uint8_t *>(base), off, nvalid, vec, pk);
  else store8_i16(static_cast<int16_t *>(base), off, nvalid, vec, pk);
}

// sign codes, 2 bits per pixel: 0 -> 0, 1 -> +1, 2 -> -1
DEV int sgn_get(uint32_t bits, int p) {
  const uint32_t c = (bits >> (2 * p)) & 3u;
  return c == 1 ? 1 : (c == 2 ? -1 : 0);
}
DEV uint32_t sgn_code(int s) { return s > 0 ? 1u : (s < 0 ? 2u : 0u); }

// ---- per-tile slot histogram helpers --------------------------------------------------------------
// One record's contribution to the tile's per-slot event counts.  fast_hist (RATE / TIME with
// n_slots <= 8): 8-bit fields of two 64-bit registers per thread (callers add < 256 records per
// thread between flushes); otherwise shared-memory atomics.
DEV void hist_add_desc(const TileArgs &A, uint32_t *hist, int nh, int desc, bool tp, unsigned long long &hp,
                       unsigned long long &hn, unsigned &status, bool allow_regs = true) {
  if (A.fast_hist && allow_regs) {  // desc in [-1, 8]
    if (desc >= 0 && desc < 8) {
      if (tp) hp += 1ull << (8 * desc); else hn += 1ull << (8 * desc);
    }
  } else if (A.enc.mode == DVS_ENC_RATE || A.enc.mode == DVS_ENC_TIME) {
    if (desc >= 0 && desc < nh) atomicAdd(&hist[(tp ? 0 : nh) + desc], 1u);
  } else {
    for (int jj = 0; jj < A.enc.n_slots; ++jj) {
      unsigned inv;
      const unsigned sm = bin_slotmask(A.enc, jj, inv);
      if (desc & inv) status |= DVS_STATUS_SLOT_RANGE;
      const int m = __popc((unsigned)desc & sm);
      if (m) atomicAdd(&hist[(tp ? 0 : nh) + jj], (unsigned)m);
    }
  }
}

// Returns the record's encoder descriptor (enc_desc) so that callers can store it with the record.
DEV int hist_add(const TileArgs &A, uint32_t *hist, int nh, int a, bool tp, unsigned long long &hp,
                 unsigned long long &hn, unsigned &status, bool allow_regs = true) {
  if (A.enc.mode == DVS_ENC_NONE) return 0;
  const int desc = enc_desc(A.enc, a, tp, status);
  hist_add_desc(A, hist, nh, desc, tp, hp, hn, status, allow_regs);
  return desc;
}

// Records carry their encoder descriptor in the upper half of the second word as desc + 2 (0 = absent), so
// that the expand kernel does not have to redo the division / table look-up.
DEV uint32_t rec_word1(int a, int desc, bool has_desc) {
  return (uint32_t)(a & 0xffff) | (has_desc ? (uint32_t)(desc + 2) << 16 : 0u);
}

// Flush the register histograms into the shared-memory histogram (one redux.sync per pair of bins) and
// publish the status bits.  Must be reached by every thread of the warp.
DEV void hist_flush(const TileArgs &A, uint32_t *hist, int nh, unsigned long long hp, unsigned long long hn,
                    unsigned status) {
  const int lane = threadIdx.x & 31;
  if (A.enc.mode != DVS_ENC_NONE && A.fast_hist && __any_sync(kFull, (hp | hn) != 0)) {
#pragma unroll
    for (int w = 0; w < 8; ++w) {  // 16 bins (2 polarities x 8) as 8 words of two 16-bit fields
      const unsigned long long h = (w < 4) ? hp : hn;
      const int b = (w & 3) * 2;
      uint32_t v = (uint32_t)((h >> (8 * b)) & 0xff) | ((uint32_t)((h >> (8 * b + 8)) & 0xff) << 16);
      v = __reduce_add_sync(kFull, v);
      if (lane == 0 && v) {
        const int base = (w < 4 ? 0 : nh) + b;
        if (v & 0xffff) atomicAdd(&hist[base], v & 0xffff);
        if ((v >> 16) && b + 1 < nh) atomicAdd(&hist[base + 1], v >> 16);
      }
    }
  }
  status = __reduce_or_sync(kFull, status);
  if (status && lane == 0 && A.status) atomicOr(A.status, (int)status);
}

// Write the tile's counters component-major: tile_counts[s][c][t].  Needs a barrier after the last
// hist_flush / shared-memory atomic.
DEV void write_counters(const TileArgs &A, int s_idx, int t_idx, const uint32_t *hist, int nh, int n_pos_tile,
                        int n_neg_tile) {
  for (int c = threadIdx.x; c < A.C; c += blockDim.x) {
    uint32_t v;
    if (c == 0) v = n_pos_tile;
    else if (c == 1) v = n_neg_tile;
    else {
      const int slot = (c - 2) >> 1, pol = (c - 2) & 1;
      const uint32_t *h = hist + (pol ? nh : 0);
      if (A.enc.mode == DVS_ENC_RATE) {  // slot j holds every pixel with k > j
        v = 0;
        for (int kk = slot + 1; kk <= A.enc.n_slots; ++kk) v += h[kk];
      } else v = h[slot];
    }
    A.tile_counts[((size_t)s_idx * A.C + c) * A.tiles_per_stream + t_idx] = v;
  }
}

DEV void tile_finish(const TileArgs &A, int s_idx, int t_idx, uint32_t *hist, int nh, int n_pos_tile,
                     int n_neg_tile, unsigned long long hp, unsigned long long hn, unsigned status) {
  hist_flush(A, hist, nh, hp, hn, status);
  __syncthreads();
  write_counters(A, s_idx, t_idx, hist, nh, n_pos_tile, n_neg_tile);
}

// ---- ordered compaction shared by the generic and the fast tile kernels ------------------------
// posm/negm: 8-bit list-membership masks per group; a_pk: abs_diff of the group's pixels.
// Block-wide exclusive scan in (group j, thread) order == row-major pixel order, records written to the
// tile's region (pos list from the front, neg list ending at the back), per-slot event counts of the
// tile accumulated in `hist`, then the per-tile counters written component-major.
template <int J>
DEV void tile_compact(const TileArgs &A, int tile, int s_idx, int t_idx, int row0, const uint32_t (&posm)[J],
                      const uint32_t (&negm)[J], const uint32_t (&a_pk)[J][4], uint32_t *wtot, uint32_t *hist,
                      uint32_t *misc, int nh, unsigned status) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
  uint32_t cnt[J];
#pragma unroll
  for (int j = 0; j < J; ++j) cnt[j] = __popc(posm[j]) | (__popc(negm[j]) << 16);
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
    const int n_ent = J * NW;  // <= 128
    uint32_t carry = 0;
    for (int e0 = 0; e0 < n_ent; e0 += 32) {
      const int e = e0 + lane;
      const uint32_t v = e < n_ent ? wtot[e] : 0;
      uint32_t x = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(kFull, x, o);
        if (lane >= o) x += y;
      }
      if (e < n_ent) wtot[e] = carry + x - v;
      carry += __shfl_sync(kFull, x, 31);
    }
    if (lane == 0) misc[0] = carry;  // tile totals: #pos | #neg << 16
  }
  __syncthreads();
  const uint32_t totals = misc[0];
  const int n_pos_tile = totals & 0xffff, n_neg_tile = totals >> 16;
  unsigned long long hp = 0, hn = 0;  // 8 x 8-bit register histograms (fast_hist)

  if (totals != 0) {
    const size_t region = (size_t)tile * A.rec_stride;
    dvs_record *pos_dst = A.records + region;
    dvs_record *neg_dst = A.records + region + (A.rec_stride - n_neg_tile);
#pragma unroll
    for (int j = 0; j < J; ++j) {
      if ((posm[j] | negm[j]) == 0) continue;
      const int q0 = (j * (int)blockDim.x + tid) * 8;
      const uint32_t excl = wtot[j * NW + warp] + incl[j] - cnt[j];
      uint32_t op = excl & 0xffff, on = excl >> 16;
      const int yq = q0 / A.W;
      int y = row0 + yq, x = q0 - yq * A.W;
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const bool tp = (posm[j] >> p) & 1u, tn = (negm[j] >> p) & 1u;
        if (tp || tn) {
          const int a = get16(a_pk[j], p);
          const int desc = hist_add(A, hist, nh, a, tp, hp, hn, status);
          const uint2 rec = make_uint2((uint32_t)y | ((uint32_t)x << 16), rec_word1(a, desc, A.enc.mode != DVS_ENC_NONE));
          if (tp) *reinterpret_cast<uint2 *>(pos_dst + op++) = rec;
          else *reinterpret_cast<uint2 *>(neg_dst + on++) = rec;
        }
        if (++x == A.W) { x = 0; ++y; }
      }
    }
  }
  tile_finish(A, s_idx, t_idx, hist, nh, n_pos_tile, n_neg_tile, hp, hn, status);
}

// The generic tile pass: every mode, every geometry, exact int16 wrap semantics for arbitrary inputs.
template <int SRC, int J, bool ADAPT, int KI>
DEV void tile_body(const TileArgs &A, const int tile) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: [a_sm: tile_px int16 (KI != 0)] [wtot: 128 u32] [hist: 2*(n_slots+1) u32] [misc: 8 u32]
  int16_t *a_sm = reinterpret_cast<int16_t *>(smem_raw);
  const int a_bytes = (KI != INH_NONE) ? ((A.tile_px * 2 + 15) & ~15) : 0;
  uint32_t *wtot = reinterpret_cast<uint32_t *>(smem_raw + a_bytes);
  uint32_t *hist = wtot + 128;
  const int nh = A.enc.mode == DVS_ENC_NONE ? 0 : (A.enc.n_slots + 1);
  uint32_t *misc = hist + 2 * nh;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
  const int s_idx = tile / A.tiles_per_stream, t_idx = tile - s_idx * A.tiles_per_stream;
  const int row0 = t_idx * A.tile_rows;
  const int rows = min(A.tile_rows, A.H - row0);
  const int npx = rows * A.W;
  const size_t plane_off = (size_t)s_idx * A.H * A.W + (size_t)row0 * A.W;
  const bool vec = A.vec_ok != 0;

  for (int i = tid; i < 2 * nh; i += blockDim.x) hist[i] = 0;
  if (tid < 8) misc[tid] = 0;

  uint32_t a_pk[J][4], r_pk[J][4], t_pk[ADAPT ? J : 1][4], c_pk[J][4];
  uint32_t sgn[J];
  int nval[J];
  unsigned status = 0;

  // ---- phase A: loads (all issued before first use), threshold + sign split ------------------
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int q0 = (j * (int)blockDim.x + tid) * 8;
    nval[j] = min(8, max(0, npx - q0));
    if (nval[j] > 0) {
      if (SRC == SRC_I16) load8_i16(static_cast<const int16_t *>(A.curr), plane_off + q0, nval[j], vec, c_pk[j]);
      else if (SRC == SRC_U8) load8_u8(static_cast<const uint8_t *>(A.curr), plane_off + q0, nval[j], vec, c_pk[j]);
      else load8_i16(static_cast<const int16_t *>(A.curr), plane_off + q0, nval[j], vec, c_pk[j]);
      if (SRC == SRC_PLANES) load8_i16(A.abs_in, plane_off + q0, nval[j], vec, a_pk[j]);
      else load8_state(A.ref, A.state_u8 != 0, plane_off + q0, nval[j], vec, r_pk[j]);
      if (ADAPT) load8_state(A.thr, A.state_u8 != 0, plane_off + q0, nval[j], vec, t_pk[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < J; ++j) {
    sgn[j] = 0;
    if (nval[j] <= 0) continue;
    const int q0 = (j * (int)blockDim.x + tid) * 8;
    if (SRC == SRC_PLANES) {
#pragma unroll
      for (int p = 0; p < 8; ++p)
        if (p < nval[j]) sgn[j] |= sgn_code(get16(c_pk[j], p)) << (2 * p);
    } else {
      uint32_t d_pk[4] = {0, 0, 0, 0};
      a_pk[j][0] = a_pk[j][1] = a_pk[j][2] = a_pk[j][3] = 0;
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        if (p < nval[j]) {
          int T, negT;
          if (ADAPT) { T = get16(t_pk[j], p); negT = w16(-T); }  // gs:110
          else { T = A.thr_scalar; negT = -T; }                  // gs:74
          int d, a, s;
          threshold_px(get16(c_pk[j], p), get16(r_pk[j], p), T, negT, d, a, s);
          set16(d_pk, p, d);
          set16(a_pk[j], p, a);
          sgn[j] |= sgn_code(s) << (2 * p);
        }
      }
      if (A.diff_out) store8_i16(A.diff_out, plane_off + q0, nval[j], vec, d_pk);
      if (A.abs_out) store8_i16(A.abs_out, plane_off + q0, nval[j], vec, a_pk[j]);
    }
    if (KI != INH_NONE) {
      if (nval[j] == 8) st16(a_sm + q0, a_pk[j]);
      else {
#pragma unroll
        for (int p = 0; p < 8; ++p)
          if (p < nval[j]) a_sm[q0 + p] = (int16_t)get16(a_pk[j], p);
      }
    }
  }

  // ---- phase B: local inhibition, gs:177-221: keep only the first arg-max of each k x k area ----
  if (KI != INH_NONE) {
    __syncthreads();
    const int k = A.inh_k;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      if (sgn[j] == 0) continue;  // nothing to inhibit in this group (also covers nval == 0)
      const int q0 = (j * (int)blockDim.x + tid) * 8;
      if (KI == INH_K2 && vec) {
        // the group's 8 pixels sit in one row (W % 8 == 0): four 2x2 areas with the partner row
        const int ry = q0 / A.W;
        const int rb = ry & 1;
        uint4 o = ld16(a_sm + (rb ? q0 - A.W : q0 + A.W));
        const uint32_t o_pk[4] = {o.x, o.y, o.z, o.w};
        uint32_t keep = 0;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int m0 = get16(a_pk[j], 2 * g), m1 = get16(a_pk[j], 2 * g + 1);
          const int p0 = get16(o_pk, 2 * g), p1 = get16(o_pk, 2 * g + 1);
          // row-major scan of the area with strict '>' (argmax gs:125-131)
          const int t00 = rb ? p0 : m0, t01 = rb ? p1 : m1, t10 = rb ? m0 : p0, t11 = rb ? m1 : p1;
          int best = t00, bi = 0;
          if (t01 > best) { best = t01; bi = 1; }
          if (t10 > best) { best = t10; bi = 2; }
          if (t11 > best) { best = t11; bi = 3; }
          const int mine0 = rb * 2, mine1 = rb * 2 + 1;
          if (bi == mine0) keep |= 3u << (4 * g);
          if (bi == mine1) keep |= 3u << (4 * g + 2);
        }
        sgn[j] &= keep;
      } else {
        const int row_in_tile0 = q0 / A.W;
        int y = row_in_tile0, x = q0 - row_in_tile0 * A.W;
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          if (p < nval[j] && ((sgn[j] >> (2 * p)) & 3u)) {
            const int ty0 = (y / k) * k, tx0 = (x / k) * k;  // tile rows start at multiples of k
            int best = a_sm[ty0 * A.W + tx0], br = 0, bc = 0;
            for (int rr = 0; rr < k; ++rr)
              for (int cc = 0; cc < k; ++cc) {
                const int v = a_sm[(ty0 + rr) * A.W + tx0 + cc];
                if (v > best) { best = v; br = rr; bc = cc; }
              }
            if (br != y - ty0 || bc != x - tx0) sgn[j] &= ~(3u << (2 * p));
          }
          if (++x == A.W) { x = 0; ++y; }
        }
      }
    }
  }

  // ---- phase C: reference / threshold update, list membership, counts -------------------------
  uint32_t posm[J], negm[J];  // 8-bit membership masks
  int mx_pos = INT_MIN, mx_neg = INT_MIN;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    posm[j] = 0; negm[j] = 0;
    if (nval[j] <= 0) continue;
    const int q0 = (j * (int)blockDim.x + tid) * 8;
    uint32_t rn_pk[4] = {0, 0, 0, 0}, tn_pk[4] = {0, 0, 0, 0}, s_pk[4] = {0, 0, 0, 0};
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      if (p < nval[j]) {
        const int s = sgn_get(sgn[j], p);
        const int a = get16(a_pk[j], p);
        if (SRC != SRC_PLANES) {
          int th_state = ADAPT ? get16(t_pk[j], p) : 0, th_ret = th_state;
          const int rn = update_px(A.upd, a, s, get16(r_pk[j], p), th_state, th_ret, status);
          set16(rn_pk, p, rn);
          if (ADAPT) set16(tn_pk, p, th_ret);
          set16(s_pk, p, s);
        }
        bool tp, tn;
        split_px(s, A.polarity, A.uncoded, tp, tn);
        if (tp) { posm[j] |= 1u << p; mx_pos = max(mx_pos, a); }
        if (tn) { negm[j] |= 1u << p; mx_neg = max(mx_neg, a); }
      }
    }
    if (SRC != SRC_PLANES) {
      if (A.upd.mode != DVS_UPD_NONE) store8_state(A.ref, A.state_u8 != 0, plane_off + q0, nval[j], vec, rn_pk);
      if (ADAPT && A.upd.mode == DVS_UPD_RATE_ADPT) store8_state(A.thr, A.state_u8 != 0, plane_off + q0, nval[j], vec, tn_pk);
      if (A.spk_out) store8_i16(A.spk_out, plane_off + q0, nval[j], vec, s_pk);
    }
  }

  if (SRC == SRC_PLANES && A.global_max) {
    mx_pos = __reduce_max_sync(kFull, mx_pos);
    mx_neg = __reduce_max_sync(kFull, mx_neg);
    if (lane == 0) {
      if (mx_pos != INT_MIN) atomicMax(&A.global_max[2 * s_idx], mx_pos);
      if (mx_neg != INT_MIN) atomicMax(&A.global_max[2 * s_idx + 1], mx_neg);
    }
  }
  tile_compact<J>(A, tile, s_idx, t_idx, row0, posm, negm, a_pk, wtot, hist, misc, nh, status);
}

template <int SRC, int J, int BDMAX, bool ADAPT, int KI>
__global__ void __launch_bounds__(BDMAX) k_tile_pass(const __grid_constant__ TileArgs A) {
  pdl_trigger();
  tile_body<SRC, J, ADAPT, KI>(A, blockIdx.x);
}

// Generic body over the tiles the fast kernel declined (pixels outside [0, 255]): A.redo[0] = count,
// A.redo[2..] = tile indices.  Launched after every fast pass; normally finds count == 0.  The last CTA to finish
// (A.redo[1] counts them) zeroes both words again, so the list is ready for the next step without a memset launch.
template <int SRC, int J, int KI, bool ADAPT = false, int BDMAX = 256>
__global__ void __launch_bounds__(BDMAX) k_tile_redo(const __grid_constant__ TileArgs A) {
  pdl_trigger();
  pdl_wait();  // the fast pass has completed: its redo list and everything it wrote are visible
  const int n = A.redo[0];
  for (int i = blockIdx.x; i < n; i += gridDim.x) {
    tile_body<SRC, J, ADAPT, KI>(A, A.redo[2 + i]);
    __syncthreads();  // shared memory is reused by the next tile
  }
  if (n != 0 && threadIdx.x == 0) {  // (n == 0: nothing to reset, and every CTA sees the same n)
    __threadfence();
    if (atomicAdd(&A.redo[1], 1) == (int)gridDim.x - 1) {
      A.redo[1] = 0;
      A.redo[0] = 0;
    }
  }
}


// ---- host side ---------------------------------------------------------------------------------
int set_error(int code, const char *msg);  // dvs_b200.cu: stores the thread-local message, returns code

// Launch with the programmatic-stream-serialization attribute (see pdl_wait() in dvs_device.cuh); the kernel must call
// pdl_wait() before it touches anything the previous kernel of the stream wrote.
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = DVS_PDL;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// Tile shapes: <= 256 groups of 8 px -> one group per thread; <= 1024 groups -> four per thread in
// CTAs of <= 256 threads; larger tiles (W * inh_k > 8192) only exist in the generic-inhibition
// variant, which also covers inh_k == 1 (every pixel is its own 1x1 area).
template <int SRC, int J, int BDMAX, bool ADAPT, int KI>
static int launch_shape(const TileArgs &A, int bd, size_t smem, cudaStream_t st) {
  auto k = k_tile_pass<SRC, J, BDMAX, ADAPT, KI>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
  }
  k<<<(unsigned)((long long)A.S * A.tiles_per_stream), bd, smem, st>>>(A);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
  return DVS_OK;
}

template <int SRC, bool ADAPT, int KI>
static int launch_tile_pass(const TileArgs &A, cudaStream_t st) {
  const int items = (A.tile_px + 7) / 8;
  const int nh = A.enc.mode == DVS_ENC_NONE ? 0 : (A.enc.n_slots + 1);
  const size_t smem = ((KI != INH_NONE) ? ((A.tile_px * 2 + 15) & ~15) : 0) + (128 + 2 * nh + 8) * sizeof(uint32_t);
  if ((long long)A.S * A.tiles_per_stream > 0x7fffffffll) return set_error(DVS_ERR_UNSUPPORTED, "too many tiles");
  auto round32 = [](int x) { return (x + 31) / 32 * 32; };
  if (items <= 256) return launch_shape<SRC, 1, 256, ADAPT, KI>(A, round32(items), smem, st);
#if GENERIC_J1
  if (items <= 1024) return launch_shape<SRC, 1, 1024, ADAPT, KI>(A, round32(items), smem, st);
#endif
  if (items <= 1024) return launch_shape<SRC, 4, 256, ADAPT, KI>(A, round32((items + 3) / 4), smem, st);
  if constexpr (KI == INH_GENERIC || SRC == SRC_PLANES) {
    const int bd = round32((items + 3) / 4);
    if (bd > 1024) return set_error(DVS_ERR_UNSUPPORTED, "tile larger than 32768 pixels");
    return launch_shape<SRC, 4, 1024, ADAPT, KI>(A, bd, smem, st);
  } else {
    return set_error(DVS_ERR_UNSUPPORTED, "internal: wide tile routed to a narrow variant");
  }
}

template <int SRC, bool ADAPT>
static int dispatch_inh(const TileArgs &A, cudaStream_t st) {
  const bool wide = (A.tile_px + 7) / 8 > 1024;
  if (!wide && A.inh_k <= 1) return launch_tile_pass<SRC, ADAPT, INH_NONE>(A, st);
  if (!wide && A.inh_k == 2 && A.vec_ok && A.W % 8 == 0) return launch_tile_pass<SRC, ADAPT, INH_K2>(A, st);
  return launch_tile_pass<SRC, ADAPT, INH_GENERIC>(A, st);
}

// one per translation unit tile_pass_*.cu
int launch_tile_pass_i16(const TileArgs &A, cudaStream_t st);
int launch_tile_pass_i16_adpt(const TileArgs &A, cudaStream_t st);
int launch_tile_pass_u8(const TileArgs &A, cudaStream_t st);
int launch_tile_pass_u8_adpt(const TileArgs &A, cudaStream_t st);
int launch_tile_pass_planes(const TileArgs &A, cudaStream_t st);
// fast path (dvs_tile_fast.cuh): static threshold, rate update, history_weight == 1
int launch_tile_fast_i16(const TileArgs &A, cudaStream_t st);
int launch_tile_fast_u8(const TileArgs &A, cudaStream_t st);
// fast path with a per-pixel adaptive threshold plane (coded update_reference_rate_adpt, history_weight == 1)
int launch_tile_fast_i16_adpt(const TileArgs &A, cudaStream_t st);
int launch_tile_fast_u8_adpt(const TileArgs &A, cudaStream_t st);
// ... with one-byte state planes (ST_U8)
int launch_tile_fast_i16_s8(const TileArgs &A, cudaStream_t st);
int launch_tile_fast_u8_s8(const TileArgs &A, cudaStream_t st);
int launch_tile_fast_i16_adpt_s8(const TileArgs &A, cudaStream_t st);
int launch_tile_fast_u8_adpt_s8(const TileArgs &A, cudaStream_t st);

}  // namespace dvs
