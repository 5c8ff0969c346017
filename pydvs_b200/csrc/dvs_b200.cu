// dvs_b200.cu -- C-ABI of libpydvs_b200 (include/pydvs_b200.h) and every kernel except the fused tile pass.
//
// Data path of one frame over S streams (DESIGN.md has the full account):
//   K1 tile pass   csrc/dvs_tile_fast.cuh (specialised) / csrc/dvs_tile_pass.cuh (generic), instantiated in
//                  tile_*.cu: threshold + sign split, k x k inhibition, reference (+ adaptive threshold)
//                  update in place, ordered compaction of the spiking pixels into the tile's record lists,
//                  per-tile per-slot event counts.
//   K2 k_scan_*    exclusive scans of the per-tile counters -> CSR row pointers (this file).
//   K3 k_expand    ballot/popc expansion of records into (x, y, t, p) events at their final,
//                  reference-ordered positions (this file).
//   plus the unfused 1:1 kernels behind the drop-in module functions and the float32 DVSOperator mode.
// Everything is HBM-bound integer work: no tensor cores, no TMEM.
#include <cuda_runtime.h>

#include <algorithm>
#include <limits.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "dvs_device.cuh"
#include "dvs_tile_pass.cuh"
#include "pydvs_b200.h"

namespace dvs {

// ================================================================================================
// K1': (rows; cols; vals) lists -> tiles of one stream (make_spike_lists_* signatures)
// ================================================================================================
struct ListArgs {
  const int16_t *pos, *neg;
  long long n_pos, n_neg, pos_stride, neg_stride;
  int chunk, n_pos_tiles, tiles, C;
  EncArgs enc;
  dvs_record *records;
  uint32_t *tile_counts;
  int32_t *status;
};

__global__ void __launch_bounds__(256) k_lists_to_tiles(const __grid_constant__ ListArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t *hist = reinterpret_cast<uint32_t *>(smem_raw);  // n_slots + 1
  const int nh = A.enc.n_slots + 1;
  const int tile = blockIdx.x, tid = threadIdx.x;
  const bool is_pos = tile < A.n_pos_tiles;
  const long long first = (long long)(is_pos ? tile : tile - A.n_pos_tiles) * A.chunk;
  const long long total = is_pos ? A.n_pos : A.n_neg;
  const int n = (int)max(0ll, min((long long)A.chunk, total - first));
  const int16_t *src = is_pos ? A.pos : A.neg;
  const long long stride = is_pos ? A.pos_stride : A.neg_stride;
  for (int i = tid; i < nh; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  unsigned status = 0;
  dvs_record *dst = A.records + (size_t)tile * A.chunk + (is_pos ? 0 : A.chunk - n);
  for (int i = tid; i < n; i += blockDim.x) {
    const int row = (uint16_t)src[first + i], col = (uint16_t)src[stride + first + i];
    const int val = src[2 * stride + first + i];
    *reinterpret_cast<uint2 *>(dst + i) =
        make_uint2((uint32_t)row | ((uint32_t)col << 16), (uint32_t)(val & 0xffff));
    const int desc = enc_desc(A.enc, val, is_pos, status);
    if (A.enc.mode == DVS_ENC_RATE || A.enc.mode == DVS_ENC_TIME) {
      if (desc >= 0) atomicAdd(&hist[desc], 1u);
    } else if (A.enc.mode != DVS_ENC_NONE) {
      for (int jj = 0; jj < A.enc.n_slots; ++jj) {
        unsigned inv;
        const unsigned sm = bin_slotmask(A.enc, jj, inv);
        if (desc & inv) status |= DVS_STATUS_SLOT_RANGE;
        const int m = __popc((unsigned)desc & sm);
        if (m) atomicAdd(&hist[jj], (unsigned)m);
      }
    }
  }
  if (status && A.status) atomicOr(A.status, (int)status);
  __syncthreads();
  for (int c = tid; c < A.C; c += blockDim.x) {
    uint32_t v = 0;
    if (c < 2) v = (c == (is_pos ? 0 : 1)) ? n : 0;
    else {
      const int slot = (c - 2) >> 1, pol = (c - 2) & 1;
      if (pol == (is_pos ? 0 : 1)) {
        if (A.enc.mode == DVS_ENC_RATE) {
          for (int kk = slot + 1; kk <= A.enc.n_slots; ++kk) v += hist[kk];
        } else v = hist[slot];
      }
    }
    A.tile_counts[(size_t)c * A.tiles + tile] = v;
  }
}

// ================================================================================================
// K2: scans
// ================================================================================================
// CSR row pointers over (stream, slot, polarity) from the stream totals, by ONE CTA: the S * 2 * n_slots segments are
// cut into one contiguous run per thread (segment k = words 2.. of row k / per of `totals`), so every thread of the
// CTA has work even for a few streams and its loads are in flight together.  The totals may have been written by other
// CTAs of the same launch (the tile scan's last block): they are read past L1 (ld.cg).
DEV void seg_scan_block(int S, int n_slots, const uint32_t *totals, long long *seg_offsets) {
  __shared__ long long wsum[32];
  __shared__ long long carry_sm;
  const int C = 2 + 2 * n_slots, per = 2 * n_slots;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = (int)blockDim.x >> 5;
  const int n_seg = S * per;  // <= 8192 (host-checked)
  const int chunk = (n_seg + (int)blockDim.x - 1) / (int)blockDim.x;
  const int lo = min(n_seg, tid * chunk), hi = min(n_seg, lo + chunk);
  const int row0 = lo / per, col0 = lo - row0 * per;
  long long sum = 0;
  {
    const uint32_t *p = totals + (long long)row0 * C + 2 + col0;
    int c = col0;
#pragma unroll 8
    for (int k = lo; k < hi; ++k) {
      sum += __ldcg(p);
      ++p;
      if (++c == per) { c = 0; p += 2; }
    }
  }
  long long x = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long y = __shfl_up_sync(kFull, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) wsum[warp] = x;
  __syncthreads();
  if (warp == 0) {
    long long w = lane < n_warps ? wsum[lane] : 0, wx = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long y = __shfl_up_sync(kFull, wx, o);
      if (lane >= o) wx += y;
    }
    wsum[lane] = wx - w;
    if (lane == 31) carry_sm = wx;
  }
  __syncthreads();
  long long run = wsum[warp] + x - sum;
  {
    const uint32_t *p = totals + (long long)row0 * C + 2 + col0;
    int c = col0;
#pragma unroll 8
    for (int k = lo; k < hi; ++k) {
      seg_offsets[k] = run;
      run += __ldcg(p);
      ++p;
      if (++c == per) { c = 0; p += 2; }
    }
  }
  if (tid == 0) seg_offsets[n_seg] = carry_sm;
}

// one warp per (stream, component): exclusive scan over the stream's tiles (contiguous memory).  With a `ticket`
// (one zero word, left zero) the last CTA to finish also builds the CSR row pointers (seg_scan_block), which saves
// the separate single-CTA launch: at 32 streams per GPU the two scans were 14 us of a 250 us step.
__global__ void __launch_bounds__(256) k_scan_tiles(int n_rows, int tps, const uint32_t *__restrict__ counts,
                                                    uint32_t *__restrict__ excl, uint32_t *totals, int S, int n_slots,
                                                    long long *seg_offsets, unsigned *ticket) {
  const int gw = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  pdl_trigger();
  pdl_wait();
  if (gw < n_rows) {
    const size_t base = (size_t)gw * tps;
    uint32_t carry = 0;
    // 256 tiles per pass: the eight loads of a lane are independent and in flight together, and the loads of the next
    // pass are issued before the current one is scanned (the counters of a large step have left L2 by now: a warp
    // that waits for each pass pays one DRAM latency per 256 tiles, five of them for a 4K stream)
    uint32_t v[8], nx[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int t = u * 32 + lane;
      v[u] = t < tps ? counts[base + t] : 0;
    }
    for (int t0 = 0; t0 < tps; t0 += 256) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int t = t0 + 256 + u * 32 + lane;
        nx[u] = t < tps ? counts[base + t] : 0;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int t = t0 + u * 32 + lane;
        uint32_t x = v[u];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t y = __shfl_up_sync(kFull, x, o);
          if (lane >= o) x += y;
        }
        if (t < tps) excl[base + t] = carry + x - v[u];
        carry += __shfl_sync(kFull, x, 31);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = nx[u];
    }
    if (lane == 0) totals[gw] = carry;
  }
  if (ticket == nullptr) return;
  __shared__ int is_last;
  __syncthreads();  // every warp of this CTA has written its totals
  if (threadIdx.x == 0) {
    __threadfence();
    is_last = atomicAdd(ticket, 1u) == gridDim.x - 1u;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (threadIdx.x == 0) *ticket = 0u;
  seg_scan_block(S, n_slots, totals, seg_offsets);
}

// The same scan with one CTA (256 threads) per (stream, component) row, for rows of more than 256 tiles: all loads of
// a row (up to 2048 tiles per round) are in flight at once and the 8 x 8 (pass, warp) totals meet in shared memory
// behind a single barrier, instead of one warp walking the row pass by pass (4K streams: 1080 tiles per row; at 32
// streams per GPU the warp-per-row form took 9.6 us, as long as 5 % of the step).
__global__ void __launch_bounds__(256) k_scan_tiles_wide(int tps, const uint32_t *__restrict__ counts,
                                                         uint32_t *__restrict__ excl, uint32_t *totals, int S, int n_slots,
                                                         long long *seg_offsets, unsigned *ticket) {
  __shared__ uint32_t wt[64];
  __shared__ int is_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t base = (size_t)blockIdx.x * tps;
  uint32_t carry = 0;
  pdl_trigger();
  pdl_wait();
  for (int t0 = 0; t0 < tps; t0 += 2048) {
    uint32_t v[8], x[8];
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const int t = t0 + p * 256 + tid;
      v[p] = t < tps ? counts[base + t] : 0u;
    }
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      x[p] = v[p];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(kFull, x[p], o);
        if (lane >= o) x[p] += y;
      }
      if (lane == 31) wt[p * 8 + warp] = x[p];
    }
    __syncthreads();
    // exclusive scan of the 64 (pass, warp) totals, redundantly by every warp: lane l holds entries l and l + 32
    const uint32_t w0 = wt[lane], w1 = wt[lane + 32];
    uint32_t s0 = w0, s1 = w1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y0 = __shfl_up_sync(kFull, s0, o), y1 = __shfl_up_sync(kFull, s1, o);
      if (lane >= o) { s0 += y0; s1 += y1; }
    }
    const uint32_t half = __shfl_sync(kFull, s0, 31);
    const uint32_t e0 = s0 - w0, e1 = s1 - w1 + half;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const int k = p * 8 + warp;  // uniform over the warp
      const uint32_t off = __shfl_sync(kFull, p < 4 ? e0 : e1, k & 31);
      const int t = t0 + p * 256 + tid;
      if (t < tps) excl[base + t] = carry + off + x[p] - v[p];
    }
    carry += half + __shfl_sync(kFull, s1, 31);
    __syncthreads();  // wt is rewritten by the next round
  }
  if (tid == 0) totals[blockIdx.x] = carry;
  if (ticket == nullptr) return;
  if (tid == 0) {
    __threadfence();
    is_last = atomicAdd(ticket, 1u) == gridDim.x - 1u;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (tid == 0) *ticket = 0u;
  seg_scan_block(S, n_slots, totals, seg_offsets);
}

// single CTA: the row pointers as a launch of their own (dvs_scan_tiles without a ticket)
__global__ void __launch_bounds__(1024) k_scan_segments(int S, int n_slots, const uint32_t *totals, long long *seg_offsets) {
  seg_scan_block(S, n_slots, totals, seg_offsets);
}

// Many streams (S * 2 * n_slots > 8192 segments): a single CTA would have to pull megabytes through one SM.
// Three small launches instead, using seg_offsets itself as scratch: (1) one thread per stream sums its row
// into seg_offsets[s * per]; (2) one CTA scans those S values in place (exclusive); (3) one thread per stream
// expands its base into the stream's 2 * n_slots row pointers.
__global__ void __launch_bounds__(128) k_seg_stream_sums(int S, int n_slots, const uint32_t *__restrict__ totals,
                                                         long long *__restrict__ seg_offsets) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const int C = 2 + 2 * n_slots, per = 2 * n_slots;
  const uint32_t *row = totals + (long long)s * C + 2;
  long long sum = 0;
#pragma unroll 4
  for (int j = 0; j < per; ++j) sum += row[j];
  seg_offsets[(long long)s * per] = sum;
}

// exclusive scan, in place, of the per-stream event totals parked in seg_offsets[s * per], by one CTA (the values may
// have been written by other CTAs of the same launch: read past L1)
DEV void seg_scan_streams_block(int S, int per, long long *seg_offsets) {
  __shared__ long long wsum[32];
  __shared__ long long carry_sm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = (int)blockDim.x >> 5;
  const int chunk = (S + (int)blockDim.x - 1) / (int)blockDim.x;
  const int lo = min(S, tid * chunk), hi = min(S, lo + chunk);
  long long sum = 0;
#pragma unroll 16
  for (int s = lo; s < hi; ++s) sum += __ldcg(seg_offsets + (long long)s * per);
  long long x = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long y = __shfl_up_sync(kFull, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) wsum[warp] = x;
  __syncthreads();
  if (warp == 0) {
    long long w = lane < n_warps ? wsum[lane] : 0, wx = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long y = __shfl_up_sync(kFull, wx, o);
      if (lane >= o) wx += y;
    }
    wsum[lane] = wx - w;
    if (lane == 31) carry_sm = wx;
  }
  __syncthreads();
  long long run = wsum[warp] + x - sum;
#pragma unroll 16
  for (int s = lo; s < hi; ++s) {
    const long long v = __ldcg(seg_offsets + (long long)s * per);
    seg_offsets[(long long)s * per] = run;
    run += v;
  }
  if (tid == 0) seg_offsets[(long long)S * per] = carry_sm;
}

__global__ void __launch_bounds__(1024) k_seg_scan_streams(int S, int per, long long *seg_offsets) {
  seg_scan_streams_block(S, per, seg_offsets);
}

// Streams of a few tiles each (tps <= 8: e.g. 4096 MNIST-sized streams of one tile): one warp per STREAM, a lane per
// component walks the row's tiles; the warp also sums the stream's events.  mode 1: the last CTA builds all row
// pointers (few segments); mode 2: the stream sums are parked in seg_offsets[s * per] and the last CTA scans them in
// place (k_seg_fill completes the rows); mode 3: sums parked only; mode 0: no row pointers wanted.
__global__ void __launch_bounds__(256) k_scan_small(int S, int tps, int n_slots, const uint32_t *__restrict__ counts,
                                                    uint32_t *__restrict__ excl, uint32_t *totals, long long *seg_offsets,
                                                    unsigned *ticket, int mode) {
  const int s = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  const int C = 2 + 2 * n_slots;
  pdl_trigger();
  pdl_wait();
  if (s < S) {
    unsigned long long events = 0;
    for (int c = lane; c < C; c += 32) {
      const size_t base = ((size_t)s * C + c) * tps;
      uint32_t run = 0;
      for (int t = 0; t < tps; ++t) {
        const uint32_t v = counts[base + t];
        excl[base + t] = run;
        run += v;
      }
      totals[(size_t)s * C + c] = run;
      if (c >= 2) events += run;
    }
    if (mode >= 2) {
#pragma unroll
      for (int o = 16; o; o >>= 1) events += __shfl_xor_sync(kFull, events, o);
      if (lane == 0) seg_offsets[(long long)s * 2 * n_slots] = (long long)events;
    }
  }
  if (mode == 0 || mode == 3) return;
  __shared__ int is_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    is_last = atomicAdd(ticket, 1u) == gridDim.x - 1u;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (threadIdx.x == 0) *ticket = 0u;
  if (mode == 1) seg_scan_block(S, n_slots, totals, seg_offsets);
  else seg_scan_streams_block(S, 2 * n_slots, seg_offsets);
}

__global__ void __launch_bounds__(128) k_seg_fill(int S, int n_slots, const uint32_t *__restrict__ totals,
                                                  long long *__restrict__ seg_offsets) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const int C = 2 + 2 * n_slots, per = 2 * n_slots;
  const uint32_t *row = totals + (long long)s * C + 2;
  long long *out = seg_offsets + (long long)s * per;
  long long run = out[0];
#pragma unroll 4
  for (int j = 0; j < per; ++j) {
    out[j] = run;
    run += row[j];
  }
}

// ================================================================================================
// K3: expand records into events, one warp per (tile, polarity)
// ================================================================================================
struct ExpandArgs {
  int S, tps, tile_px, C, tpw;
  EncArgs enc;
  const dvs_record *records;
  const uint32_t *tile_counts, *tile_excl;
  const long long *seg_offsets;
  dvs_event *events;
  long long event_cap;
  int32_t *status;
};

// grid = (ceil(2 * tiles_per_stream / 256), S).  Task index within a stream = pol * tps + t, so that the
// 256 record counts a block looks at are two contiguous runs of tile_counts; every thread loads one
// count (coalesced), and each warp then walks the non-empty tasks of its 32 lanes with the whole warp.
// NS8 (n_slots <= 8, no time-binary multiplicities): every lane also pre-loads its own task's 8 segment
// positions (coalesced over tiles) and the first records of the NEXT task are prefetched while the
// current one is expanded, so the dependent-load chain per task is off the critical path.
#ifndef EXPAND_SMALL_TASK
#define EXPAND_SMALL_TASK 4
#endif
#ifndef EXPAND_MINB
#define EXPAND_MINB 5   // 48 registers (no spills on the NS8 path); 4 / 6 / 8 measured within 5 us of each other at 256 x 4K
#endif
template <bool NS8>
__global__ void __launch_bounds__(256, EXPAND_MINB) k_expand(const __grid_constant__ ExpandArgs A) {
  pdl_wait();  // the scan (and through it the tile pass) has completed
  const int lane = threadIdx.x & 31;
  const int s = blockIdx.y;
  // A.tpw tasks per warp (32 when there are plenty of tiles, fewer to keep >= ~16k warps in flight)
  const int warp_g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int my_task = lane < A.tpw ? warp_g * A.tpw + lane : 2 * A.tps;
  const int my_pol = my_task >= A.tps, my_t = my_task - my_pol * A.tps;
  int my_n = 0;
  if (my_task < 2 * A.tps) {
    my_n = (int)A.tile_counts[((size_t)s * A.C + my_pol) * A.tps + my_t];
    // rate coding: slot 0 holds every record that emits at all -> tiles whose records all have k == 0
    // (large changes emit nothing, Appendix A.8) are skipped without touching their records
    if (my_n > 0 && A.enc.mode == DVS_ENC_RATE && A.tile_counts[((size_t)s * A.C + 2 + my_pol) * A.tps + my_t] == 0) my_n = 0;
  }
  unsigned todo = __ballot_sync(kFull, my_n > 0);
  if (todo == 0) return;
  const long long n_seg = (long long)A.S * A.enc.n_slots * 2;
  if (A.seg_offsets[n_seg] > A.event_cap) {
    if (lane == 0 && A.status) atomicOr(A.status, (int)DVS_STATUS_EVENT_OVERFLOW);
    return;
  }
  const bool bin_mode = A.enc.mode == DVS_ENC_TIME_BIN || A.enc.mode == DVS_ENC_TIME_BIN_THR;
  const unsigned lt = (1u << lane) - 1u;
  unsigned status = 0;
  const dvs_record *my_rec = A.records + ((size_t)s * A.tps + my_t) * A.tile_px + (my_pol ? A.tile_px - my_n : 0);

  if (NS8) {
    // this lane's task: position of its first event in each slot's segment, relative to the stream's events
    const long long stream_base = A.seg_offsets[(long long)s * A.enc.n_slots * 2];
    dvs_event *const ev_s = A.events + stream_base;
    uint32_t my_base[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      my_base[j] = 0;
      if (my_n > 0 && j < A.enc.n_slots)
        my_base[j] = (uint32_t)(A.seg_offsets[((long long)s * A.enc.n_slots + j) * 2 + my_pol] - stream_base) +
                     A.tile_excl[((size_t)s * A.C + 2 + 2 * j + my_pol) * A.tps + my_t];
    }
    // Tasks with a handful of records (the usual case on sparse frames: one or two per tile) are expanded by
    // their own lane, 32 tasks at a time, instead of occupying the whole warp one after the other.
    constexpr int kSmallTask = EXPAND_SMALL_TASK;
    if (__any_sync(kFull, my_n > 0 && my_n <= kSmallTask)) {
      if (my_n > 0 && my_n <= kSmallTask) {
        const uint32_t pb = (my_pol ? 0xffffu : 1u) << 16;
        for (int i = 0; i < my_n; ++i) {
          const uint2 r = *reinterpret_cast<const uint2 *>(my_rec + i);
          const int desc = (r.y >> 16) ? (int)(r.y >> 16) - 2 : enc_desc(A.enc, (int)(short)(r.y & 0xffff), my_pol == 0, status);
          const uint32_t xy = (r.x >> 16) | (r.x << 16);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (A.enc.mode == DVS_ENC_RATE ? desc > j : desc == j) {
              *reinterpret_cast<uint2 *>(ev_s + my_base[j]) = make_uint2(xy, (uint32_t)j | pb);
              my_base[j] += 1;
            }
          }
        }
      }
      todo = __ballot_sync(kFull, my_n > kSmallTask);
      if (todo == 0) {
        status = __reduce_or_sync(kFull, status);
        if (status && lane == 0 && A.status) atomicOr(A.status, (int)status);
        return;
      }
    }
    // Tasks that fit a group of G = 8 (then 16) lanes: 32 / G tasks are expanded per pass (group g walks the tasks
    // owned by its own lanes), the ballot of a slot is cut into the groups' bit fields.
    auto grouped = [&](const int G, const int n_lo, const int n_hi) {
      const int g0 = lane & ~(G - 1), sub = lane & (G - 1);
      const unsigned sub_lt = (1u << sub) - 1u, gm = G == 8 ? 0xffu : 0xffffu;
#pragma unroll 1
      for (int k = 0; k < G; ++k) {
        const int src_g = g0 + k;
        const int n_g = __shfl_sync(kFull, my_n, src_g);
        const bool act = n_g > n_lo && n_g <= n_hi;  // uniform within the group
        if (!__any_sync(kFull, act)) continue;
        const dvs_record *rec_g = reinterpret_cast<const dvs_record *>(__shfl_sync(kFull, (unsigned long long)my_rec, src_g));
        const int pol_g = __shfl_sync(kFull, my_pol, src_g);
        uint2 r = make_uint2(0u, 0u);
        int desc = (A.enc.mode == DVS_ENC_TIME) ? -1 : 0;
        if (act && sub < n_g) {
          r = *reinterpret_cast<const uint2 *>(rec_g + sub);
          desc = (r.y >> 16) ? (int)(r.y >> 16) - 2 : enc_desc(A.enc, (int)(short)(r.y & 0xffff), pol_g == 0, status);
        }
        const uint32_t xy = (r.x >> 16) | (r.x << 16);
        const uint32_t pb = (pol_g ? 0xffffu : 1u) << 16;
        const int dmax = __reduce_max_sync(kFull, desc);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (A.enc.mode == DVS_ENC_RATE ? j >= dmax : j > dmax) break;
          const uint32_t bj = __shfl_sync(kFull, my_base[j], src_g);
          const bool m = (A.enc.mode == DVS_ENC_RATE) ? desc > j : desc == j;
          const unsigned gb = (__ballot_sync(kFull, m) >> g0) & gm;
          if (m) *reinterpret_cast<uint2 *>(ev_s + bj + __popc(gb & sub_lt)) = make_uint2(xy, (uint32_t)j | pb);
        }
      }
    };
    constexpr int kGroupTask = 16;
    if (__any_sync(kFull, my_n > kSmallTask && my_n <= 8)) grouped(8, kSmallTask, 8);
    if (__any_sync(kFull, my_n > 8 && my_n <= kGroupTask)) grouped(16, 8, kGroupTask);
    if (__any_sync(kFull, my_n > kSmallTask && my_n <= kGroupTask)) {
      todo = __ballot_sync(kFull, my_n > kGroupTask);
      if (todo == 0) {
        status = __reduce_or_sync(kFull, status);
        if (status && lane == 0 && A.status) atomicOr(A.status, (int)status);
        return;
      }
    }
    int src = __ffs(todo) - 1;
    todo &= todo - 1;
    const dvs_record *rec = reinterpret_cast<const dvs_record *>(__shfl_sync(kFull, (unsigned long long)my_rec, src));
    int n = __shfl_sync(kFull, my_n, src);
    uint2 r_first = lane < n ? *reinterpret_cast<const uint2 *>(rec + lane) : make_uint2(0, 0);
    while (true) {
      // prefetch the first chunk of the next task
      const int nsrc = todo ? __ffs(todo) - 1 : -1;
      const dvs_record *nrec = rec;
      int nn = 0;
      uint2 r_next = make_uint2(0, 0);
      if (nsrc >= 0) {
        todo &= todo - 1;
        nrec = reinterpret_cast<const dvs_record *>(__shfl_sync(kFull, (unsigned long long)my_rec, nsrc));
        nn = __shfl_sync(kFull, my_n, nsrc);
        if (lane < nn) r_next = *reinterpret_cast<const uint2 *>(nrec + lane);
      }
      const int pol = __shfl_sync(kFull, my_pol, src);
      uint32_t base[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) base[j] = __shfl_sync(kFull, my_base[j], src);
      const uint32_t pbits = (pol ? 0xffffu : 1u) << 16;
      // four 32-record chunks per iteration: their loads are issued back to back (4x the bytes in flight)
      for (int i0 = 0; i0 < n; i0 += 128) {
        uint2 rr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + 32 * u + lane;
          rr[u] = (u == 0 && i0 == 0) ? r_first : (i < n ? *reinterpret_cast<const uint2 *>(rec + i) : make_uint2(0, 0));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (i0 + 32 * u >= n) break;
          const int i = i0 + 32 * u + lane;
          const uint2 r = rr[u];
          int desc = (A.enc.mode == DVS_ENC_TIME) ? -1 : 0;
          if (i < n) desc = (r.y >> 16) ? (int)(r.y >> 16) - 2 : enc_desc(A.enc, (int)(short)(r.y & 0xffff), pol == 0, status);
          const uint32_t xy = (r.x >> 16) | (r.x << 16);  // record is (row, col); event is (x = col, y = row)
          const int dmax = __reduce_max_sync(kFull, desc);
          if (A.enc.mode == DVS_ENC_RATE && dmax <= 0) continue;  // nothing in this chunk emits
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (A.enc.mode == DVS_ENC_RATE ? j >= dmax : j > dmax) break;  // no record of the chunk reaches slot j
            const bool m = (A.enc.mode == DVS_ENC_RATE) ? desc > j : desc == j;
            const unsigned bal = __ballot_sync(kFull, m);
            if (m) *reinterpret_cast<uint2 *>(ev_s + base[j] + __popc(bal & lt)) = make_uint2(xy, (uint32_t)j | pbits);
            base[j] += __popc(bal);
          }
        }
      }
      if (nsrc < 0) break;
      src = nsrc; rec = nrec; n = nn; r_first = r_next;
    }
    status = __reduce_or_sync(kFull, status);
    if (status && lane == 0 && A.status) atomicOr(A.status, (int)status);
    return;
  }

  while (todo) {
    const int src_lane = __ffs(todo) - 1;
    todo &= todo - 1;
    const int n = __shfl_sync(kFull, my_n, src_lane);
    const int pol = __shfl_sync(kFull, my_pol, src_lane), t = __shfl_sync(kFull, my_t, src_lane);
    const dvs_record *rec = reinterpret_cast<const dvs_record *>(__shfl_sync(kFull, (unsigned long long)my_rec, src_lane));
    for (int sb = 0; sb < A.enc.n_slots; sb += 32) {
      const int jmine = sb + lane;
      long long base = 0;
      unsigned slotmask = 0;
      if (jmine < A.enc.n_slots) {
        base = A.seg_offsets[((long long)s * A.enc.n_slots + jmine) * 2 + pol] +
               A.tile_excl[((size_t)s * A.C + 2 + 2 * jmine + pol) * A.tps + t];
        if (bin_mode) { unsigned inv; slotmask = bin_slotmask(A.enc, jmine, inv); }
      }
      const int nj = min(32, A.enc.n_slots - sb);
      for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        uint2 r = make_uint2(0, 0);
        int desc = (A.enc.mode == DVS_ENC_TIME) ? -1 : 0;
        if (i < n) {
          r = *reinterpret_cast<const uint2 *>(rec + i);
          desc = enc_desc(A.enc, (int)(short)(r.y & 0xffff), pol == 0, status);
        }
        int jend = nj;
        if (A.enc.mode == DVS_ENC_RATE) jend = min(nj, max(0, __reduce_max_sync(kFull, desc) - sb));
        const uint32_t xy = (r.x >> 16) | (r.x << 16);
        for (int jj = 0; jj < jend; ++jj) {
          const unsigned sm = __shfl_sync(kFull, slotmask, jj);
          const int m = enc_mult(A.enc, desc, sb + jj, sm);
          const long long b = __shfl_sync(kFull, base, jj);
          int tot, ex;
          if (!bin_mode || !__any_sync(kFull, m > 1)) {  // (time-binary: multiplicities above 1 only when a value wraps)
            const unsigned bal = __ballot_sync(kFull, m != 0);
            tot = __popc(bal);
            ex = __popc(bal & lt);
          } else {
            int x = m;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const int y = __shfl_up_sync(kFull, x, o);
              if (lane >= o) x += y;
            }
            tot = __shfl_sync(kFull, x, 31);
            ex = x - m;
          }
          if (m) {
            const uint2 ev = make_uint2(xy, (uint32_t)(sb + jj) | ((pol ? 0xffffu : 1u) << 16));
            for (int q = 0; q < m; ++q) *reinterpret_cast<uint2 *>(A.events + b + ex + q) = ev;
          }
          if (lane == jj) base += tot;
        }
      }
    }
  }
  status = __reduce_or_sync(kFull, status);
  if (status && lane == 0 && A.status) atomicOr(A.status, (int)status);
}

// ================================================================================================
// rows of W pixels -> rows of Wp pixels (Wp % 8 == 0, pad columns zero): the frame copy in front of the tile pass
// when the planes are kept with padded rows (DVSStreams, rows that are not a multiple of 8 pixels with inhibition).
// One thread per 8 destination pixels, four pair loads (W is even: every row starts on a pair boundary).
// ================================================================================================
template <typename PAIR, typename VEC>
__global__ void __launch_bounds__(256) k_pad_rows(const PAIR *__restrict__ src, VEC *__restrict__ dst, long long n_groups,
                                                  int pairs_per_row, int groups_per_row) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n_groups) return;
  const long long r = i / groups_per_row;
  const int g = (int)(i - r * groups_per_row);
  const PAIR *s = src + r * pairs_per_row + g * 4;
  uint32_t v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = g * 4 + k < pairs_per_row ? (uint32_t)s[k] : 0u;
  if constexpr (sizeof(PAIR) == 4) dst[i] = make_uint4(v[0], v[1], v[2], v[3]);
  else dst[i] = make_uint2(v[0] | (v[1] << 16), v[2] | (v[3] << 16));
}

// ================================================================================================
// gather tile-local record lists of one stream into contiguous (rows; cols; vals) arrays
// ================================================================================================
__global__ void __launch_bounds__(256) k_gather_split(int s, int tps, int tile_px, int C,
                                                      const dvs_record *__restrict__ records,
                                                      const uint32_t *__restrict__ counts,
                                                      const uint32_t *__restrict__ excl,
                                                      int16_t *pos_out, long long pos_stride,
                                                      int16_t *neg_out, long long neg_stride) {
  const int lane = threadIdx.x & 31;
  const long long warps_total = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long task = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; task < 2ll * tps;
       task += warps_total) {
    const int t = (int)(task >> 1), pol = (int)(task & 1);
    const size_t ci = ((size_t)s * C + pol) * tps + t;
    const int n = (int)counts[ci];
    if (n == 0) continue;
    const long long off = excl[ci];
    const size_t tile = (size_t)s * tps + t;
    const dvs_record *rec = records + tile * tile_px + (pol ? tile_px - n : 0);
    int16_t *out = pol ? neg_out : pos_out;
    const long long stride = pol ? neg_stride : pos_stride;
    for (int i = lane; i < n; i += 32) {
      const uint2 r = *reinterpret_cast<const uint2 *>(rec + i);
      out[off + i] = (int16_t)(r.x & 0xffff);
      out[stride + off + i] = (int16_t)(r.x >> 16);
      out[2 * stride + off + i] = (int16_t)(r.y & 0xffff);
    }
  }
}

// ================================================================================================
// unfused function-for-function kernels
// ================================================================================================
__global__ void __launch_bounds__(256) k_thresholded_difference(const int16_t *__restrict__ curr,
                                                                const int16_t *__restrict__ ref,
                                                                const int16_t *__restrict__ thr_plane,
                                                                int thr_scalar, int16_t *__restrict__ diff,
                                                                int16_t *__restrict__ abs_diff,
                                                                int16_t *__restrict__ spikes, long long n,
                                                                int vec) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long q0 = g * 8;
  if (q0 >= n) return;
  const int nv = (int)min(8ll, n - q0);
  uint32_t c[4], r[4], t[4] = {0, 0, 0, 0}, d_pk[4] = {0, 0, 0, 0}, a_pk[4] = {0, 0, 0, 0}, s_pk[4] = {0, 0, 0, 0};
  load8_i16(curr, q0, nv, vec, c);
  load8_i16(ref, q0, nv, vec, r);
  if (thr_plane) load8_i16(thr_plane, q0, nv, vec, t);
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    if (p < nv) {
      int T, negT;
      if (thr_plane) { T = get16(t, p); negT = w16(-T); }
      else { T = thr_scalar; negT = -T; }
      int d, a, s;
      threshold_px(get16(c, p), get16(r, p), T, negT, d, a, s);
      set16(d_pk, p, d); set16(a_pk, p, a); set16(s_pk, p, s);
    }
  }
  store8_i16(diff, q0, nv, vec, d_pk);
  store8_i16(abs_diff, q0, nv, vec, a_pk);
  store8_i16(spikes, q0, nv, vec, s_pk);
}

// one thread per k x k area; in place on spikes
__global__ void __launch_bounds__(256) k_local_inhibition(int16_t *__restrict__ spikes,
                                                          const int16_t *__restrict__ abs_diff, int S, int H,
                                                          int W, int k) {
  const int aw = W / k, ah = H / k;
  const long long n_areas = (long long)S * aw * ah;
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_areas) return;
  const int ax = (int)(g % aw);
  const long long rest = g / aw;
  const int ay = (int)(rest % ah), s = (int)(rest / ah);
  const size_t o = (size_t)s * H * W + (size_t)ay * k * W + (size_t)ax * k;
  int best = abs_diff[o], br = 0, bc = 0;
  for (int rr = 0; rr < k; ++rr)
    for (int cc = 0; cc < k; ++cc) {
      const int v = abs_diff[o + (size_t)rr * W + cc];
      if (v > best) { best = v; br = rr; bc = cc; }
    }
  for (int rr = 0; rr < k; ++rr)
    for (int cc = 0; cc < k; ++cc)
      if (rr != br || cc != bc) spikes[o + (size_t)rr * W + cc] = 0;
}

// k == 2 on 16-byte aligned rows with W % 8 == 0: a thread owns 8 columns of a row pair (four 2x2 areas), reads
// abs_diff and spikes with 16-byte loads and only writes the rows back when some spike was actually removed.
__global__ void __launch_bounds__(256) k_local_inhibition2_vec(int16_t *__restrict__ spikes,
                                                               const int16_t *__restrict__ abs_diff, long long n_pairs,
                                                               int W) {
  const int gpr = W >> 3;
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_pairs * gpr) return;
  const long long pair = g / gpr;  // row pairs are contiguous over streams because H is even
  const int cg = (int)(g - pair * gpr);
  const size_t o = (size_t)pair * 2 * W + (size_t)cg * 8;
  const uint4 su = ld16(spikes + o), sl = ld16(spikes + o + W);
  if ((su.x | su.y | su.z | su.w | sl.x | sl.y | sl.z | sl.w) == 0u) return;  // nothing to inhibit
  const uint4 au = ld16(abs_diff + o), al = ld16(abs_diff + o + W);
  const uint32_t a_u[4] = {au.x, au.y, au.z, au.w}, a_l[4] = {al.x, al.y, al.z, al.w};
  uint32_t s_u[4] = {su.x, su.y, su.z, su.w}, s_l[4] = {sl.x, sl.y, sl.z, sl.w};
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const int t00 = (int16_t)(a_u[w] & 0xffffu), t01 = (int16_t)(a_u[w] >> 16);
    const int t10 = (int16_t)(a_l[w] & 0xffffu), t11 = (int16_t)(a_l[w] >> 16);
    int best = t00, bi = 0;  // argmax gs:118-132: first maximum in row-major order, strict '>'
    if (t01 > best) { best = t01; bi = 1; }
    if (t10 > best) { best = t10; bi = 2; }
    if (t11 > best) { best = t11; bi = 3; }
    s_u[w] &= bi == 0 ? 0x0000ffffu : (bi == 1 ? 0xffff0000u : 0u);
    s_l[w] &= bi == 2 ? 0x0000ffffu : (bi == 3 ? 0xffff0000u : 0u);
  }
  if ((s_u[0] ^ su.x) | (s_u[1] ^ su.y) | (s_u[2] ^ su.z) | (s_u[3] ^ su.w)) st16(spikes + o, s_u);
  if ((s_l[0] ^ sl.x) | (s_l[1] ^ sl.y) | (s_l[2] ^ sl.z) | (s_l[3] ^ sl.w)) st16(spikes + o + W, s_l);
}

__global__ void __launch_bounds__(256) k_update_reference(const __grid_constant__ UpdateArgs U,
                                                          const int16_t *__restrict__ abs_diff,
                                                          const int16_t *__restrict__ spikes,
                                                          const int16_t *__restrict__ ref_in,
                                                          int16_t *__restrict__ ref_out, int16_t *thr,
                                                          int16_t *thr_out, long long n, int vec,
                                                          int32_t *status_out) {
  const long long q0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (q0 >= n) return;
  const int nv = (int)min(8ll, n - q0);
  const bool adpt = U.mode == DVS_UPD_RATE_ADPT;
  uint32_t a[4], s[4], r[4], t[4] = {0, 0, 0, 0}, rn[4] = {0, 0, 0, 0}, ts[4] = {0, 0, 0, 0}, tr[4] = {0, 0, 0, 0};
  load8_i16(abs_diff, q0, nv, vec, a);
  load8_i16(spikes, q0, nv, vec, s);
  load8_i16(ref_in, q0, nv, vec, r);
  if (adpt) load8_i16(thr, q0, nv, vec, t);
  unsigned status = 0;
  if (U.mode == DVS_UPD_RATE && U.hw_is_one && U.div_thr.T >= 2 && nv == 8) {
    // update_reference_rate gs:244-249 with history_weight == 1 and a scalar threshold >= 2, word by word (two
    // pixels): quiet words reduce to clip(ref, 0, 255) on packed signed lanes; a spiking pixel with abs_diff >= 0
    // takes the magic-number division; anything else goes through the general rule.  Every int16 wrap is kept.
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      if (s[w] == 0u) {
        rn[w] = __vmins2(__vmaxs2(r[w], 0u), 0x00ff00ffu);
        continue;
      }
      uint32_t out = 0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int av = (int)(short)(a[w] >> (16 * h)), sv = (int)(short)(s[w] >> (16 * h)), rv = (int)(short)(r[w] >> (16 * h));
        int v;
        if (av >= 0) {
          const int nq = min((int)__umulhi((unsigned)av, U.div_thr.magic), U.max_time_ms);  // clip(q, 0, max): q >= 0 here
          v = clip(w16(rv + w16(w16(nq * sv) * U.threshold)), 0, 255);
        } else {
          int th_state = 0, th_ret = 0;
          v = update_px(U, av, sv, rv, th_state, th_ret, status);
        }
        out |= ((uint32_t)v & 0xffffu) << (16 * h);
      }
      rn[w] = out;
    }
    store8_i16(ref_out, q0, nv, vec, rn);
    return;
  }
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    if (p < nv) {
      int th_state = get16(t, p), th_ret = th_state;
      const int v = update_px(U, get16(a, p), get16(s, p), get16(r, p), th_state, th_ret, status);
      set16(rn, p, v); set16(ts, p, th_state); set16(tr, p, th_ret);
    }
  }
  store8_i16(ref_out, q0, nv, vec, rn);
  if (adpt) {
    store8_i16(thr, q0, nv, vec, ts);
    if (thr_out != thr) store8_i16(thr_out, q0, nv, vec, tr);
  }
  if (status && status_out) atomicOr(status_out, (int)status);
}

// noise variants gs:428-473: mark sampled pixels, then one elementwise pass
__global__ void __launch_bounds__(256) k_mark_samples(const int16_t *__restrict__ sample, long long n_sample,
                                                      int W, int H, uint8_t *__restrict__ flags,
                                                      int32_t *status_out) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_sample) return;
  const int v = sample[g];
  const int q = np_floordiv(v, W);
  int r = w16(q), c = w16(v - q * W);  // sample // width, sample % width (int16 arrays)
  if (r < 0) r += H;
  if (c < 0) c += W;
  if (r < 0 || r >= H || c < 0 || c >= W) { if (status_out) atomicOr(status_out, (int)DVS_STATUS_INDEX); return; }
  flags[(size_t)r * W + c] = 1;
}

__global__ void __launch_bounds__(256) k_update_noise(int thresh_variant, const int16_t *__restrict__ abs_diff,
                                                      const int16_t *__restrict__ spikes,
                                                      const int16_t *__restrict__ ref_in,
                                                      int16_t *__restrict__ ref_out, long long n, FastDiv dv,
                                                      int max_time_ms, double hw, int hw_is_one,
                                                      const uint8_t *__restrict__ lut, int lut_len,
                                                      const uint8_t *__restrict__ flags, int32_t *status_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int a = abs_diff[i];
  int mult;
  if (thresh_variant) {
    mult = w16(clip(w16(div_T(a, dv)), 0, max_time_ms));  // gs:463
    if (flags[i]) mult = 0;                               // gs:469
  } else {
    int idx = a;
    if (idx < 0) idx += lut_len;
    if (idx < 0 || idx >= lut_len) { if (status_out) atomicOr(status_out, (int)DVS_STATUS_LUT_INDEX); mult = 0; }
    else mult = lut[idx];                                 // gs:440
    if (flags[i]) mult = w16(mult ^ 0xFF);                // gs:445
  }
  // gs:447 / gs:471: no clip; the thresh variant omits `* threshold`
  ref_out[i] = (int16_t)w16(hist_weight(hw, hw_is_one != 0, ref_in[i]) + w16(spikes[i] * mult));
}

// DVSOperator::operator() cpp/dvs_op.cpp:12-31; explicit _rn intrinsics forbid FMA contraction
__global__ void __launch_bounds__(256) k_step_f32(const float *__restrict__ src, float *__restrict__ diff,
                                                  float *__restrict__ ref, float *__restrict__ thr, float relax,
                                                  float up, float down, long long n, int vec) {
  const long long q0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (q0 >= n) return;
  float sv[4], rv[4], tv[4], dv[4];
  const int nv = (int)min(4ll, n - q0);
  if (vec && nv == 4) {
    const float4 a = *reinterpret_cast<const float4 *>(src + q0), b = *reinterpret_cast<const float4 *>(ref + q0),
                 c = *reinterpret_cast<const float4 *>(thr + q0);
    sv[0] = a.x; sv[1] = a.y; sv[2] = a.z; sv[3] = a.w;
    rv[0] = b.x; rv[1] = b.y; rv[2] = b.z; rv[3] = b.w;
    tv[0] = c.x; tv[1] = c.y; tv[2] = c.z; tv[3] = c.w;
  } else {
    for (int p = 0; p < 4; ++p) {
      sv[p] = p < nv ? src[q0 + p] : 0.f; rv[p] = p < nv ? ref[q0 + p] : 0.f; tv[p] = p < nv ? thr[q0 + p] : 0.f;
    }
  }
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    float d = __fsub_rn(sv[p], rv[p]);
    const bool test = (d < -tv[p]) || (d > tv[p]);
    d = __fmul_rn(d, test ? 1.0f : 0.0f);
    dv[p] = d;
    rv[p] = __fadd_rn(__fmul_rn(relax, rv[p]), d);
    tv[p] = test ? __fmul_rn(tv[p], up) : __fmul_rn(tv[p], down);
  }
  if (vec && nv == 4) {
    if (diff) *reinterpret_cast<float4 *>(diff + q0) = make_float4(dv[0], dv[1], dv[2], dv[3]);
    *reinterpret_cast<float4 *>(ref + q0) = make_float4(rv[0], rv[1], rv[2], rv[3]);
    *reinterpret_cast<float4 *>(thr + q0) = make_float4(tv[0], tv[1], tv[2], tv[3]);
  } else {
    for (int p = 0; p < nv; ++p) {
      if (diff) diff[q0 + p] = dv[p];
      ref[q0 + p] = rv[p]; thr[q0 + p] = tv[p];
    }
  }
}

__global__ void __launch_bounds__(256) k_keys_from_events(const dvs_event *__restrict__ ev, long long n, int uncoded,
                                                          int flag_shift, int data_shift, int data_mask,
                                                          int16_t *__restrict__ keys) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint2 e = *reinterpret_cast<const uint2 *>(ev + i);
  const int x = e.x & 0xffff, y = e.x >> 16, p = (short)(e.y >> 16);
  keys[i] = (int16_t)spike_key(y, x, p > 0, uncoded, flag_shift, data_shift, data_mask);
}

// the same with the event count read on the DEVICE (the last CSR row pointer of the step), so that a sink can queue the
// key pass behind a step without the host ever learning the count first
__global__ void __launch_bounds__(256) k_keys_from_events_devn(const dvs_event *__restrict__ ev, const long long *__restrict__ n_dev,
                                                               long long cap, int uncoded, int flag_shift, int data_shift,
                                                               int data_mask, int16_t *__restrict__ keys) {
  const long long n = min(*n_dev, cap);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {   // (the grid is sized by the SMs)
    const uint2 e = *reinterpret_cast<const uint2 *>(ev + i);
    const int x = e.x & 0xffff, y = e.x >> 16, p = (short)(e.y >> 16);
    keys[i] = (int16_t)spike_key(y, x, p > 0, uncoded, flag_shift, data_shift, data_mask);
  }
}

// batched move_image (+ mask_image): one thread per output pixel, gs:1184-1215 / gs:1124-1127
__global__ void __launch_bounds__(256) k_move_mask(const int16_t *__restrict__ images, int n_images,
                                                   const int32_t *__restrict__ img_index, const int32_t *__restrict__ dx,
                                                   const int32_t *__restrict__ dy, int bg, const double *__restrict__ mask,
                                                   int16_t *__restrict__ out, int S, int H, int W) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long P = (long long)H * W;
  if (g >= (long long)S * P) return;
  const int s = (int)(g / P);
  const int q = (int)(g - (long long)s * P);
  const int y = q / W, x = q - y * W;
  const int sx = x - dx[s], sy = y + dy[s];  // content moves right for dx > 0 and up for dy > 0
  int v = bg;
  const int img = img_index[s];
  if (sx >= 0 && sx < W && sy >= 0 && sy < H && img >= 0 && img < n_images) v = images[(long long)img * P + (long long)sy * W + sx];
  if (mask) v = w16(__double2int_rz((double)v * mask[q]));  // DTYPE(original * mask)
  out[g] = (int16_t)v;
}

}  // namespace dvs

// ================================================================================================
// C-ABI
// ================================================================================================
using namespace dvs;

static thread_local char g_err[512] = "";
static thread_local int g_last_path = -1;  // which tile kernel the last dvs_step_fused of this thread launched

static int fail(int code, const char *msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
namespace dvs {
int set_error(int code, const char *msg) { return fail(code, msg); }
}  // namespace dvs
static int check_launch(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return DVS_OK;
}
static FastDiv make_div(int T) {
  FastDiv f;
  f.T = T;
  f.magic = (T >= 2) ? (unsigned)((0x100000000ull + (unsigned)T - 1) / (unsigned)T) : 0u;
  return f;
}
static EncArgs make_enc(const dvs_enc_params *e) {
  EncArgs E;
  memset(&E, 0, sizeof(E));
  if (!e) { E.mode = DVS_ENC_NONE; return E; }
  E.mode = e->mode; E.uncoded = e->uncoded; E.threshold = e->threshold;
  E.n_slots = e->mode == DVS_ENC_NONE ? 0 : e->n_slots;
  E.u16_vals = e->u16_vals; E.lut_len = e->lut_len; E.lut = e->lut;
  E.dv = make_div(e->threshold);
  return E;
}
static int check_enc(const dvs_enc_params *e) {
  if (!e || e->mode == DVS_ENC_NONE) return DVS_OK;
  if (e->mode < 0 || e->mode > DVS_ENC_NONE) return fail(DVS_ERR_INVALID_ARGUMENT, "encoder mode");
  if (e->n_slots < 0 || e->n_slots > kMaxSlots) return fail(DVS_ERR_UNSUPPORTED, "n_slots outside [0, 1024]");
  if ((e->mode == DVS_ENC_TIME_BIN || e->mode == DVS_ENC_TIME_BIN_THR) && (!e->lut || e->lut_len <= 0))
    return fail(DVS_ERR_INVALID_ARGUMENT, "time-binary encoders need a log2 table");
  if (e->mode != DVS_ENC_TIME_BIN && e->threshold == 0)
    return fail(DVS_ERR_INVALID_ARGUMENT, "encoder threshold is zero (reference: ZeroDivisionError)");
  return DVS_OK;
}
static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

int dvs_abi_version(void) { return DVS_ABI_VERSION; }
int dvs_last_step_path(void) { return g_last_path; }
const char *dvs_last_error_string(void) { return g_err; }

int dvs_get_limits(dvs_limits *out) {
  if (!out) return fail(DVS_ERR_INVALID_ARGUMENT, "null limits");
  out->max_tile_px = kMaxTilePx; out->max_slots = kMaxSlots; out->max_width = 65535; out->sm_count = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) {
    int sm = 0;
    if (cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess) out->sm_count = sm;
  }
  cudaGetLastError();
  return DVS_OK;
}

int dvs_choose_tile_rows(int32_t S, int32_t H, int32_t W, int32_t inh_k) {
  const int k = inh_k > 1 ? inh_k : 1;
  if (S <= 0 || H <= 0 || W <= 0 || (long long)W * k > kMaxTilePx) return -1;
  if (k == 2 && W % 8 == 0 && H % 2 == 0 && W >= 64) return 2;  // 2x2 areas stay inside one thread (k_tile_fast2)
  if (k == 4 && W % 8 == 0 && H % 4 == 0 && W >= 64 && W <= 4096) return 4;  // 4x4 areas likewise (k_tile_fast4)
  const int target_px = 8192;
  int rows = (target_px / W) / k * k;
  if (rows < k) rows = k;
  const int h_up = (H + k - 1) / k * k;
  if (rows > h_up) rows = h_up;
  // keep at least ~4 CTAs per SM in flight when the problem allows it
  const long long want = 4ll * 148;
  while (rows > k && (long long)S * ((H + rows - 1) / rows) < want) {
    int nr = (rows / 2) / k * k;
    if (nr < k) nr = k;
    if (nr == rows) break;
    rows = nr;
  }
  if (k == 1 && W % 8 != 0) {
    // rows that do not end on a 16-byte boundary: tiles of a multiple of 8 / gcd(W, 8) rows still start aligned, which
    // is all the linear (no inhibition) fast kernel needs
    const int g = (W % 4 == 0) ? 2 : (W % 2 == 0 ? 4 : 8);
    if (rows >= g) rows = rows / g * g;
    else if (H >= g && (long long)g * W <= kMaxTilePx) rows = g;
  }
  return rows;
}

static int fill_tiling(TileArgs &A, const dvs_tiling &t, int n_slots_or_none, int enc_mode) {
  if (t.S <= 0 || t.H <= 0 || t.W <= 0 || t.tile_rows <= 0) return fail(DVS_ERR_INVALID_ARGUMENT, "tiling");
  if (t.W > 65535 || t.H > 65535) return fail(DVS_ERR_UNSUPPORTED, "W, H must fit uint16");
  if ((long long)t.tile_rows * t.W > kMaxTilePx) return fail(DVS_ERR_UNSUPPORTED, "tile_rows * W > 32768");
  A.S = t.S; A.H = t.H; A.W = t.W; A.tile_rows = t.tile_rows;
  A.tiles_per_stream = (t.H + t.tile_rows - 1) / t.tile_rows;
  A.tile_px = t.tile_rows * t.W;
  A.rec_stride = A.tile_px;
  A.inv_w = t.W >= 2 ? (unsigned)((0x100000000ull + (unsigned)t.W - 1) / (unsigned)t.W) : 0u;
  A.C = 2 + 2 * (enc_mode == DVS_ENC_NONE ? 0 : n_slots_or_none);
  return DVS_OK;
}

int dvs_step_fused(const dvs_step_params *p, void *stream) {
  if (!p || !p->curr || !p->ref || !p->records || !p->tile_counts) return fail(DVS_ERR_INVALID_ARGUMENT, "null pointer");
  if (p->adaptive && !p->thr) return fail(DVS_ERR_INVALID_ARGUMENT, "adaptive mode needs a threshold plane");
  if (p->update_mode < 0 || p->update_mode > DVS_UPD_NONE) return fail(DVS_ERR_INVALID_ARGUMENT, "update mode");
  if (p->update_mode == DVS_UPD_RATE_ADPT && !p->adaptive) return fail(DVS_ERR_INVALID_ARGUMENT, "rate_adpt needs adaptive");
  if ((p->update_mode == DVS_UPD_TIME_BIN_RAW || p->update_mode == DVS_UPD_TIME_BIN_THR) &&
      (!p->enc.lut || p->enc.lut_len <= 0))
    return fail(DVS_ERR_INVALID_ARGUMENT, "time-binary update needs a log2 table");
  int rc = check_enc(&p->enc);
  if (rc) return rc;
  TileArgs A;
  memset(&A, 0, sizeof(A));
  rc = fill_tiling(A, p->tiling, p->enc.n_slots, p->enc.mode);
  if (rc) return rc;
  const int k = p->inh_k > 1 ? p->inh_k : 1;
  if (k > 1 && (A.W % k || A.H % k || A.tile_rows % k))
    return fail(DVS_ERR_INVALID_ARGUMENT, "inhibition needs W, H and tile_rows divisible by inh_k (reference: IndexError)");
  A.adaptive = p->adaptive; A.polarity = p->polarity; A.uncoded = p->uncoded; A.inh_k = k;
  if (p->record_stride != 0) {  // the caller sized the tile regions by the inhibition bound: one record per k x k area
    const int bound = k > 1 ? (A.tile_rows / k) * (A.W / k) : A.tile_px;
    if (p->record_stride < bound || p->record_stride > A.tile_px)
      return fail(DVS_ERR_INVALID_ARGUMENT, "record_stride below the per-tile record bound (or above tile_rows * W)");
    A.rec_stride = p->record_stride;
  }
  A.thr_scalar = p->threshold;
  A.upd.mode = p->update_mode; A.upd.uncoded = p->uncoded; A.upd.threshold = p->threshold;
  A.upd.min_thr = p->min_thr; A.upd.max_thr = p->max_thr; A.upd.thr_down = p->thr_down; A.upd.thr_up = p->thr_up;
  A.upd.max_time_ms = p->max_time_ms; A.upd.lut = p->enc.lut; A.upd.lut_len = p->enc.lut_len;
  A.upd.hw = p->history_weight; A.upd.hw_is_one = p->history_weight == 1.0; A.upd.div_thr = make_div(p->threshold);
  A.enc = make_enc(&p->enc);
  A.fast_hist = (A.enc.mode == DVS_ENC_RATE || A.enc.mode == DVS_ENC_TIME) && A.enc.n_slots <= 8;
  A.drop_silent = p->drop_silent != 0;
  // force_generic == 3: the persistent staged kernel whenever the tile shape allows it (opt-in, see dvs_tile_fast.cuh)
  A.use_staged = p->force_generic == 3 ? 2 : 0;
  A.curr = p->curr; A.ref = p->ref; A.thr = p->thr;
  if (p->state_dtype != DVS_STATE_I16 && p->state_dtype != DVS_STATE_U8) return fail(DVS_ERR_INVALID_ARGUMENT, "state_dtype");
  A.state_u8 = p->state_dtype == DVS_STATE_U8;
  if (A.state_u8 && p->adaptive &&
      (p->uncoded || p->update_mode != DVS_UPD_RATE_ADPT || p->min_thr < 0 || p->max_thr > 255 || p->min_thr > p->max_thr))
    // (the uncoded rule keeps the unclipped threshold, gsu:286-295, which may leave [0, 255])
    return fail(DVS_ERR_UNSUPPORTED, "one-byte state needs the coded threshold rule with 0 <= min_thr <= max_thr <= 255");
  A.diff_out = p->diff_out; A.abs_out = p->abs_diff_out; A.spk_out = p->spikes_out;
  A.records = p->records; A.tile_counts = p->tile_counts; A.status = p->status;
  const bool u8 = p->frame_dtype == DVS_FRAME_U8;
  const long long P = (long long)A.H * A.W;
  A.vec_ok = (P % 8 == 0) && (A.tile_px % 8 == 0) && aligned16(p->ref) && (!p->thr || aligned16(p->thr)) &&
             (u8 ? ((reinterpret_cast<uintptr_t>(p->curr) & 7) == 0) : aligned16(p->curr)) &&
             (!p->diff_out || aligned16(p->diff_out)) && (!p->abs_diff_out || aligned16(p->abs_diff_out)) &&
             (!p->spikes_out || aligned16(p->spikes_out));
  cudaStream_t st = as_stream(stream);
  // fast path: static threshold, rate update, no history decay, vector-friendly geometry
  // (the register-only kernels also take 0 <= history_weight < 1 and the time-binary updates)
  const bool plain = p->update_mode == DVS_UPD_RATE && p->history_weight == 1.0;
  // 4x4 inhibition in registers: four-row tiles, one column block per thread (<= 512 threads) and an inlined encoder
  const bool coded_lists = p->uncoded == 0 && p->polarity == DVS_POL_MERGED && p->enc.uncoded == 0 && p->enc.threshold >= 2 &&
                           p->enc.u16_vals == 0 && p->history_weight == 1.0 &&
                           ((p->enc.mode == DVS_ENC_RATE && p->enc.n_slots >= 1 && p->enc.n_slots <= 8) ||
                            (p->enc.mode == DVS_ENC_TIME && p->enc.n_slots >= 1));
  const bool quad = k == 4 && A.tile_rows == 4 && A.H % 4 == 0 && A.W % 8 == 0 && A.W / 8 <= 512 && coded_lists &&
                    (p->adaptive ? p->update_mode == DVS_UPD_RATE_ADPT : p->update_mode == DVS_UPD_RATE);
  const int max_groups = quad ? 2048 : 1024;
  const bool regs_only = k == 1 || (k == 2 && A.tile_rows == 2 && A.H % 2 == 0 && A.W % 8 == 0) || quad;
  const bool upd_ok = regs_only && (plain || (p->history_weight >= 0.0 && p->history_weight <= 1.0 &&
                                             (p->update_mode == DVS_UPD_RATE || p->update_mode == DVS_UPD_TIME_BIN_RAW ||
                                              (p->update_mode == DVS_UPD_TIME_BIN_THR && p->threshold <= 128))));
  const bool fast = !p->adaptive && upd_ok &&
                    p->threshold >= 2 && p->threshold <= 32767 && p->max_time_ms >= 0 && A.vec_ok &&
                    A.tile_px / 8 <= max_groups && !p->diff_out && !p->abs_diff_out &&
                    !p->spikes_out && p->force_generic != 1 && p->redo && A.S <= 65535;
  A.redo = p->redo;
  g_last_path = DVS_PATH_GENERIC;
  if (fast) {
    g_last_path = quad ? DVS_PATH_FAST_4X4 : DVS_PATH_FAST;
    if (A.state_u8) return u8 ? launch_tile_fast_u8_s8(A, st) : launch_tile_fast_i16_s8(A, st);
    return u8 ? launch_tile_fast_u8(A, st) : launch_tile_fast_i16(A, st);
  }
  // adaptive fast path: coded rate_adpt rule, no history decay, inhibition off or 2x2 on two-row tiles
  // (uncoded rule: th - down must not borrow below zero lane-wise, so thresholds start and stay >= down is not
  //  guaranteed -> require min_thr >= 0 and let the range vote catch negative planes; th > min >= 0 => th - down
  //  can go negative only if down > th - 0, which the generic body handles: restrict to down <= min_thr + 1)
  const bool unc_ok = p->uncoded ? (p->thr_down <= p->min_thr + 1) : ((long long)p->thr_up + p->thr_down <= 32768);
  const bool fast_adpt = p->adaptive && p->update_mode == DVS_UPD_RATE_ADPT && unc_ok &&
                         p->history_weight >= 0.0 && p->history_weight <= 1.0 &&
                         p->min_thr >= 0 && p->min_thr <= p->max_thr && p->max_thr <= 32767 && p->thr_up >= 0 &&
                         p->thr_up <= 32767 && p->thr_down >= 0 && p->thr_down <= 32767 && A.vec_ok &&
                         A.tile_px / 8 <= max_groups && regs_only &&
                         !p->diff_out && !p->abs_diff_out && !p->spikes_out && p->force_generic != 1 && p->redo && A.S <= 65535;
  if (fast_adpt) {
    g_last_path = quad ? DVS_PATH_FAST_4X4 : DVS_PATH_FAST;
    if (A.state_u8) return u8 ? launch_tile_fast_u8_adpt_s8(A, st) : launch_tile_fast_i16_adpt_s8(A, st);
    return u8 ? launch_tile_fast_u8_adpt(A, st) : launch_tile_fast_i16_adpt(A, st);
  }
  if (u8) return p->adaptive ? launch_tile_pass_u8_adpt(A, st) : launch_tile_pass_u8(A, st);
  return p->adaptive ? launch_tile_pass_i16_adpt(A, st) : launch_tile_pass_i16(A, st);
}

int dvs_split_tiles(const dvs_tiling *tiling, const int16_t *spikes, const int16_t *abs_diff, int32_t polarity,
                    const dvs_enc_params *enc, dvs_record *records, uint32_t *tile_counts, int32_t *global_max,
                    int32_t *status, void *stream) {
  if (!tiling || !spikes || !abs_diff || !records || !tile_counts) return fail(DVS_ERR_INVALID_ARGUMENT, "null pointer");
  int rc = check_enc(enc);
  if (rc) return rc;
  TileArgs A;
  memset(&A, 0, sizeof(A));
  A.enc = make_enc(enc);
  rc = fill_tiling(A, *tiling, A.enc.n_slots, A.enc.mode);
  if (rc) return rc;
  A.polarity = polarity; A.uncoded = enc ? enc->uncoded : 0; A.inh_k = 1;
  A.upd.mode = DVS_UPD_NONE;
  A.fast_hist = (A.enc.mode == DVS_ENC_RATE || A.enc.mode == DVS_ENC_TIME) && A.enc.n_slots <= 8;
  A.curr = spikes; A.abs_in = abs_diff; A.records = records; A.tile_counts = tile_counts;
  A.status = status; A.global_max = global_max;
  const long long P = (long long)A.H * A.W;
  A.vec_ok = (P % 8 == 0) && (A.tile_px % 8 == 0) && aligned16(spikes) && aligned16(abs_diff);
  return launch_tile_pass_planes(A, as_stream(stream));
}

int dvs_lists_to_tiles(const int16_t *pos, int64_t n_pos, int64_t pos_stride, const int16_t *neg, int64_t n_neg,
                       int64_t neg_stride, int32_t chunk, const dvs_enc_params *enc, dvs_record *records,
                       uint32_t *tile_counts, int32_t *status, void *stream) {
  if (!enc || !records || !tile_counts || chunk <= 0 || chunk > kMaxTilePx || n_pos < 0 || n_neg < 0)
    return fail(DVS_ERR_INVALID_ARGUMENT, "lists_to_tiles arguments");
  if ((n_pos > 0 && !pos) || (n_neg > 0 && !neg)) return fail(DVS_ERR_INVALID_ARGUMENT, "null list");
  int rc = check_enc(enc);
  if (rc) return rc;
  ListArgs A;
  memset(&A, 0, sizeof(A));
  A.pos = pos; A.neg = neg; A.n_pos = n_pos; A.n_neg = n_neg; A.pos_stride = pos_stride; A.neg_stride = neg_stride;
  A.chunk = chunk;
  A.n_pos_tiles = (int)((n_pos + chunk - 1) / chunk);
  const int n_neg_tiles = (int)((n_neg + chunk - 1) / chunk);
  A.tiles = A.n_pos_tiles + n_neg_tiles;
  if (A.tiles == 0) A.tiles = 1;  // one empty tile keeps the scan / expand shapes valid
  A.enc = make_enc(enc);
  A.C = 2 + 2 * A.enc.n_slots;
  A.records = records; A.tile_counts = tile_counts; A.status = status;
  const size_t smem = (size_t)(A.enc.n_slots + 1) * sizeof(uint32_t);
  k_lists_to_tiles<<<A.tiles, 256, smem, as_stream(stream)>>>(A);
  return check_launch("k_lists_to_tiles");
}

static int scan_tiles_impl(int32_t S, int32_t tiles_per_stream, int32_t n_slots, const uint32_t *tile_counts,
                           uint32_t *tile_excl, uint32_t *stream_totals, int64_t *seg_offsets, uint32_t *ticket, void *stream) {
  if (S <= 0 || tiles_per_stream <= 0 || n_slots < 0 || n_slots > kMaxSlots || !tile_counts || !tile_excl ||
      !stream_totals)
    return fail(DVS_ERR_INVALID_ARGUMENT, "scan_tiles arguments");
  const int C = 2 + 2 * n_slots;
  const long long n_rows = (long long)S * C;
  if (n_rows > 0x7fffffffll / 32) return fail(DVS_ERR_UNSUPPORTED, "S * C too large");
  const unsigned grid = (unsigned)((n_rows * 32 + 255) / 256);
  long long *seg = reinterpret_cast<long long *>(seg_offsets);
  const bool small = (long long)S * 2 * n_slots <= 8192;
  const bool fused = ticket && seg && n_slots > 0 && small;  // the last CTA of the tile scan builds the row pointers
  // (const / non-const copies of the arguments so that cudaLaunchKernelEx sees the kernels' exact parameter types)
  const uint32_t *counts_c = tile_counts;
  long long *seg_k = fused ? seg : nullptr;
  unsigned *ticket_k = fused ? ticket : nullptr;
  int tps_k = tiles_per_stream, S_k = S, ns_k = n_slots, rows_k = (int)n_rows;
  cudaError_t le;
  if (tiles_per_stream <= 8 && S <= 0x7fffffff / 32) {
    // many small streams: a warp per stream, which also delivers the per-stream event totals
    const bool want = seg && n_slots > 0;
    int mode = !want ? 0 : (small ? (ticket ? 1 : 0) : (ticket ? 2 : 3));
    seg_k = mode ? seg : nullptr;
    ticket_k = ticket;
    le = launch_pdl(k_scan_small, dim3((unsigned)(((long long)S * 32 + 255) / 256)), dim3(256), 0, as_stream(stream), S_k, tps_k,
                    ns_k, counts_c, tile_excl, stream_totals, seg_k, ticket_k, mode);
    if (le != cudaSuccess) return fail((int)le, cudaGetErrorString(le));
    int rc = check_launch("k_scan_small");
    if (rc || mode == 1) return rc;
    if (mode >= 2) {
      const unsigned blocks = (unsigned)((S + 127) / 128);
      if (mode == 3) k_seg_scan_streams<<<1, 1024, 0, as_stream(stream)>>>(S, 2 * n_slots, seg);
      k_seg_fill<<<blocks, 128, 0, as_stream(stream)>>>(S, n_slots, stream_totals, seg);
      return check_launch("k_seg_fill");
    }
    // mode 0 with row pointers wanted (few segments, no ticket): the separate launch below
  } else {
    if (tiles_per_stream > 256 && n_rows <= 0x7fffffffll / 256)
      le = launch_pdl(k_scan_tiles_wide, dim3((unsigned)n_rows), dim3(256), 0, as_stream(stream), tps_k, counts_c, tile_excl,
                      stream_totals, S_k, ns_k, seg_k, ticket_k);
    else
      le = launch_pdl(k_scan_tiles, dim3(grid), dim3(256), 0, as_stream(stream), rows_k, tps_k, counts_c, tile_excl, stream_totals,
                      S_k, ns_k, seg_k, ticket_k);
    if (le != cudaSuccess) return fail((int)le, cudaGetErrorString(le));
    int rc = check_launch("k_scan_tiles");
    if (rc || fused) return rc;
  }
  int rc = DVS_OK;
  if (seg_offsets) {
    if (n_slots == 0) {
      cudaError_t e = cudaMemsetAsync(seg_offsets, 0, sizeof(int64_t), as_stream(stream));
      if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
      return DVS_OK;
    }
    if (small) {
      k_scan_segments<<<1, 1024, 0, as_stream(stream)>>>(S, n_slots, stream_totals, seg);
    } else {
      const unsigned blocks = (unsigned)((S + 127) / 128);
      k_seg_stream_sums<<<blocks, 128, 0, as_stream(stream)>>>(S, n_slots, stream_totals, seg);
      k_seg_scan_streams<<<1, 1024, 0, as_stream(stream)>>>(S, 2 * n_slots, seg);
      k_seg_fill<<<blocks, 128, 0, as_stream(stream)>>>(S, n_slots, stream_totals, seg);
    }
    rc = check_launch("k_scan_segments");
  }
  return rc;
}

int dvs_scan_tiles(int32_t S, int32_t tiles_per_stream, int32_t n_slots, const uint32_t *tile_counts,
                   uint32_t *tile_excl, uint32_t *stream_totals, int64_t *seg_offsets, void *stream) {
  return scan_tiles_impl(S, tiles_per_stream, n_slots, tile_counts, tile_excl, stream_totals, seg_offsets, nullptr, stream);
}

int dvs_scan_tiles_ticket(int32_t S, int32_t tiles_per_stream, int32_t n_slots, const uint32_t *tile_counts,
                          uint32_t *tile_excl, uint32_t *stream_totals, int64_t *seg_offsets, uint32_t *ticket,
                          void *stream) {
  return scan_tiles_impl(S, tiles_per_stream, n_slots, tile_counts, tile_excl, stream_totals, seg_offsets, ticket, stream);
}

int dvs_aer_expand(int32_t S, int32_t tiles_per_stream, int32_t tile_px, const dvs_enc_params *enc,
                   const dvs_record *records, const uint32_t *tile_counts, const uint32_t *tile_excl,
                   const int64_t *seg_offsets, dvs_event *events, int64_t event_cap, int32_t *status, void *stream) {
  if (S <= 0 || tiles_per_stream <= 0 || tile_px <= 0 || !enc || enc->mode == DVS_ENC_NONE || !records ||
      !tile_counts || !tile_excl || !seg_offsets || (!events && event_cap > 0))
    return fail(DVS_ERR_INVALID_ARGUMENT, "aer_expand arguments");
  int rc = check_enc(enc);
  if (rc) return rc;
  ExpandArgs A;
  memset(&A, 0, sizeof(A));
  A.S = S; A.tps = tiles_per_stream; A.tile_px = tile_px; A.enc = make_enc(enc); A.C = 2 + 2 * A.enc.n_slots;
  A.records = records; A.tile_counts = tile_counts; A.tile_excl = tile_excl;
  A.seg_offsets = reinterpret_cast<const long long *>(seg_offsets);
  A.events = events; A.event_cap = event_cap; A.status = status;
  if (S > 65535) return fail(DVS_ERR_UNSUPPORTED, "more than 65535 streams per launch");
  int tpw = 32;
  while (tpw > 1 && (long long)S * ((2 * tiles_per_stream + tpw - 1) / tpw) < 16384) tpw >>= 1;
  A.tpw = tpw;
  // streams of one or two tiles (thousands of MNIST-sized streams) have fewer than eight warps' worth of tasks: smaller CTAs
  // instead of idle warps that hold residency slots
  const int warps = std::min(8, (2 * tiles_per_stream + tpw - 1) / tpw);
  const int tasks_per_block = warps * tpw;
  const dim3 grid((unsigned)((2 * tiles_per_stream + tasks_per_block - 1) / tasks_per_block), (unsigned)S);
  const bool ns8 = (A.enc.mode == DVS_ENC_RATE || A.enc.mode == DVS_ENC_TIME) && A.enc.n_slots <= 8;
  const cudaError_t le = ns8 ? launch_pdl(k_expand<true>, grid, dim3(32 * warps), 0, as_stream(stream), A)
                             : launch_pdl(k_expand<false>, grid, dim3(32 * warps), 0, as_stream(stream), A);
  if (le != cudaSuccess) return fail((int)le, cudaGetErrorString(le));
  return check_launch("k_expand");
}

int dvs_gather_split(int32_t s, int32_t tiles_per_stream, int32_t tile_px, int32_t n_slots, const dvs_record *records,
                     const uint32_t *tile_counts, const uint32_t *tile_excl, int16_t *pos_out, int64_t pos_stride,
                     int16_t *neg_out, int64_t neg_stride, void *stream) {
  if (s < 0 || tiles_per_stream <= 0 || tile_px <= 0 || !records || !tile_counts || !tile_excl)
    return fail(DVS_ERR_INVALID_ARGUMENT, "gather_split arguments");
  const int C = 2 + 2 * n_slots;
  long long grid = (2ll * tiles_per_stream + 7) / 8;
  if (grid > 148ll * 16) grid = 148ll * 16;
  k_gather_split<<<(unsigned)grid, 256, 0, as_stream(stream)>>>(s, tiles_per_stream, tile_px, C, records, tile_counts,
                                                                tile_excl, pos_out, pos_stride, neg_out, neg_stride);
  return check_launch("k_gather_split");
}

int dvs_thresholded_difference_i16(const int16_t *curr, const int16_t *ref, const int16_t *thr_plane, int32_t thr_scalar,
                                   int16_t *diff, int16_t *abs_diff, int16_t *spikes, int64_t n_px, void *stream) {
  if (!curr || !ref || !diff || !abs_diff || !spikes || n_px < 0) return fail(DVS_ERR_INVALID_ARGUMENT, "null pointer");
  if (n_px == 0) return DVS_OK;
  const int vec = aligned16(curr) && aligned16(ref) && aligned16(diff) && aligned16(abs_diff) && aligned16(spikes) &&
                  (!thr_plane || aligned16(thr_plane));
  const long long groups = (n_px + 7) / 8;
  k_thresholded_difference<<<(unsigned)((groups + 255) / 256), 256, 0, as_stream(stream)>>>(
      curr, ref, thr_plane, thr_scalar, diff, abs_diff, spikes, n_px, vec);
  return check_launch("k_thresholded_difference");
}

int dvs_local_inhibition_i16(int16_t *spikes, const int16_t *abs_diff, int32_t S, int32_t H, int32_t W, int32_t k,
                             void *stream) {
  if (!spikes || !abs_diff || S <= 0 || H <= 0 || W <= 0 || k <= 0) return fail(DVS_ERR_INVALID_ARGUMENT, "inhibition arguments");
  if (W % k || H % k) return fail(DVS_ERR_INVALID_ARGUMENT, "inhibition needs W and H divisible by inh_width (reference: IndexError)");
  if (k == 1) return DVS_OK;
  const long long n_areas = (long long)S * (W / k) * (H / k);
  if (k == 2 && W % 8 == 0 && aligned16(spikes) && aligned16(abs_diff)) {
    const long long n_pairs = (long long)S * (H / 2), n_thr = n_pairs * (W / 8);
    k_local_inhibition2_vec<<<(unsigned)((n_thr + 255) / 256), 256, 0, as_stream(stream)>>>(spikes, abs_diff, n_pairs, W);
    return check_launch("k_local_inhibition2_vec");
  }
  k_local_inhibition<<<(unsigned)((n_areas + 255) / 256), 256, 0, as_stream(stream)>>>(spikes, abs_diff, S, H, W, k);
  return check_launch("k_local_inhibition");
}

int dvs_update_reference_i16(int32_t mode, int32_t uncoded, const int16_t *abs_diff, const int16_t *spikes,
                             const int16_t *ref_in, int16_t *ref_out, int16_t *thr, int16_t *thr_out, int32_t threshold,
                             int32_t min_thr, int32_t max_thr, int32_t thr_down, int32_t thr_up, int32_t max_time_ms,
                             double history_weight, const uint8_t *lut, int32_t lut_len, int64_t n_px, int32_t *status,
                             void *stream) {
  if (!abs_diff || !spikes || !ref_in || !ref_out || n_px < 0) return fail(DVS_ERR_INVALID_ARGUMENT, "null pointer");
  if (mode < 0 || mode >= DVS_UPD_NONE) return fail(DVS_ERR_INVALID_ARGUMENT, "update mode");
  if (mode == DVS_UPD_RATE_ADPT && (!thr || !thr_out)) return fail(DVS_ERR_INVALID_ARGUMENT, "rate_adpt needs threshold planes");
  if ((mode == DVS_UPD_TIME_BIN_RAW || mode == DVS_UPD_TIME_BIN_THR) && (!lut || lut_len <= 0))
    return fail(DVS_ERR_INVALID_ARGUMENT, "time-binary update needs a log2 table");
  if (n_px == 0) return DVS_OK;
  UpdateArgs U;
  memset(&U, 0, sizeof(U));
  U.mode = mode; U.uncoded = uncoded; U.threshold = threshold; U.min_thr = min_thr; U.max_thr = max_thr;
  U.thr_down = thr_down; U.thr_up = thr_up; U.max_time_ms = max_time_ms; U.lut = lut; U.lut_len = lut_len;
  U.hw = history_weight; U.hw_is_one = history_weight == 1.0; U.div_thr = make_div(threshold);
  const int vec = aligned16(abs_diff) && aligned16(spikes) && aligned16(ref_in) && aligned16(ref_out) &&
                  (!thr || aligned16(thr)) && (!thr_out || aligned16(thr_out));
  const long long groups = (n_px + 7) / 8;
  k_update_reference<<<(unsigned)((groups + 255) / 256), 256, 0, as_stream(stream)>>>(U, abs_diff, spikes, ref_in, ref_out,
                                                                                     thr, thr_out, n_px, vec, status);
  return check_launch("k_update_reference");
}

int dvs_update_reference_noise_i16(int32_t thresh_variant, const int16_t *abs_diff, const int16_t *spikes,
                                   const int16_t *ref_in, int16_t *ref_out, int32_t W, int32_t H, int32_t threshold,
                                   int32_t max_time_ms, double history_weight, const uint8_t *lut, int32_t lut_len,
                                   const int16_t *sample, int64_t n_sample, uint8_t *flags, int32_t *status, void *stream) {
  if (!abs_diff || !spikes || !ref_in || !ref_out || !flags || W <= 0 || H <= 0 || n_sample < 0 || (n_sample > 0 && !sample))
    return fail(DVS_ERR_INVALID_ARGUMENT, "noise update arguments");
  if (!thresh_variant && (!lut || lut_len <= 0)) return fail(DVS_ERR_INVALID_ARGUMENT, "raw noise update needs a log2 table");
  const long long n = (long long)W * H;
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(flags, 0, (size_t)n, st);
  if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
  if (n_sample > 0) {
    k_mark_samples<<<(unsigned)((n_sample + 255) / 256), 256, 0, st>>>(sample, n_sample, W, H, flags, status);
    int rc = check_launch("k_mark_samples");
    if (rc) return rc;
  }
  k_update_noise<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(thresh_variant, abs_diff, spikes, ref_in, ref_out, n,
                                                             make_div(threshold), max_time_ms, history_weight,
                                                             history_weight == 1.0, lut, lut_len, flags, status);
  return check_launch("k_update_noise");
}

int dvs_step_f32(const float *src, float *diff, float *ref, float *thr, float relax, float up, float down, int64_t n_px,
                 void *stream) {
  if (!src || !ref || !thr || n_px < 0) return fail(DVS_ERR_INVALID_ARGUMENT, "null pointer");
  if (n_px == 0) return DVS_OK;
  const int vec = aligned16(src) && aligned16(ref) && aligned16(thr) && (!diff || aligned16(diff));
  const long long groups = (n_px + 3) / 4;
  k_step_f32<<<(unsigned)((groups + 255) / 256), 256, 0, as_stream(stream)>>>(src, diff, ref, thr, relax, up, down, n_px, vec);
  return check_launch("k_step_f32");
}

int dvs_move_mask_i16(const int16_t *images, int32_t n_images, const int32_t *img_index, const int32_t *dx,
                      const int32_t *dy, int32_t bg, const double *mask, int16_t *out, int32_t S, int32_t H, int32_t W,
                      void *stream) {
  if (!images || !img_index || !dx || !dy || !out || n_images <= 0 || S <= 0 || H <= 0 || W <= 0)
    return fail(DVS_ERR_INVALID_ARGUMENT, "move_mask arguments");
  const long long n = (long long)S * H * W;
  k_move_mask<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(images, n_images, img_index, dx, dy, bg, mask, out,
                                                                       S, H, W);
  return check_launch("k_move_mask");
}

int dvs_pad_rows(const void *src, void *dst, int64_t rows, int32_t W, int32_t Wp, int32_t elem_bytes, void *stream) {
  if (!src || !dst || rows < 0 || W <= 0 || (W & 1) || Wp < W || Wp - W >= 8 || (Wp & 7) || (elem_bytes != 1 && elem_bytes != 2))
    return fail(DVS_ERR_INVALID_ARGUMENT, "pad_rows: W even, Wp = W rounded up to a multiple of 8, 1- or 2-byte pixels");
  if ((reinterpret_cast<uintptr_t>(src) & (2 * elem_bytes - 1)) || (reinterpret_cast<uintptr_t>(dst) & (8 * elem_bytes - 1)))
    return fail(DVS_ERR_INVALID_ARGUMENT, "pad_rows: unaligned buffer");
  const long long n_groups = (long long)rows * (Wp / 8);
  if (n_groups == 0) return DVS_OK;
  const unsigned grid = (unsigned)((n_groups + 255) / 256);
  if (elem_bytes == 2)
    k_pad_rows<uint32_t, uint4><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const uint32_t *>(src), static_cast<uint4 *>(dst),
                                                                  n_groups, W / 2, Wp / 8);
  else
    k_pad_rows<uint16_t, uint2><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const uint16_t *>(src), static_cast<uint2 *>(dst),
                                                                  n_groups, W / 2, Wp / 8);
  return check_launch("k_pad_rows");
}

int dvs_keys_from_events(const dvs_event *events, int64_t n, int32_t uncoded, int32_t flag_shift, int32_t data_shift,
                         int32_t data_mask, int16_t *keys, void *stream) {
  if (n < 0 || (n > 0 && (!events || !keys))) return fail(DVS_ERR_INVALID_ARGUMENT, "keys_from_events arguments");
  if (n == 0) return DVS_OK;
  k_keys_from_events<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(events, n, uncoded, flag_shift, data_shift,
                                                                                data_mask, keys);
  return check_launch("k_keys_from_events");
}

int dvs_keys_from_events_devn(const dvs_event *events, const int64_t *n_dev, int64_t cap, int32_t uncoded, int32_t flag_shift,
                              int32_t data_shift, int32_t data_mask, int16_t *keys, void *stream) {
  if (cap < 0 || !n_dev || (cap > 0 && (!events || !keys))) return fail(DVS_ERR_INVALID_ARGUMENT, "keys_from_events_devn arguments");
  if (cap == 0) return DVS_OK;
  const long long blocks = (cap + 255) / 256;
  k_keys_from_events_devn<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, as_stream(stream)>>>(
      events, reinterpret_cast<const long long *>(n_dev), cap, uncoded, flag_shift, data_shift, data_mask, keys);
  return check_launch("k_keys_from_events_devn");
}

}  // extern "C"
