// dvs_device.cuh -- device-side arithmetic shared by every kernel of libpydvs_b200.
//
// Integer semantics follow NumPy int16 arrays with Python-int scalars: each
// value the reference holds in an int16 array is narrowed with w16(); integer division is floor
// division with x // 0 == 0.  Reference citations: "gs" = pydvs/generate_spikes.pyx, "gsu" =
// pydvs/generate_spikes_uncoded.pyx.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pydvs_b200.h"

#define DEV __device__ __forceinline__

namespace dvs {

constexpr int kMaxTilePx = 32768;  // 16-bit packed per-tile counters
constexpr int kMaxSlots = 1024;
constexpr unsigned kFull = 0xffffffffu;

// Programmatic dependent launch (sm_90+): the small kernels that follow K1 in a step (redo, scan, expand) are
// launched with the programmatic-stream-serialization attribute, so their CTAs are set up while the previous kernel
// drains; pdl_wait() then blocks until that kernel has completed and its writes are visible (a no-op for a normal
// launch).  pdl_trigger() lets the NEXT kernel start that early set-up.  Stream order is unchanged.
#ifndef DVS_PDL
#define DVS_PDL 1
#endif
DEV void pdl_wait() {
#if DVS_PDL
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
DEV void pdl_trigger() {
#if DVS_PDL
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

DEV int w16(int x) { return (int)(short)x; }

// NumPy floor_divide for int16 operands, exact for |a|, |b| <= 32768: the correctly rounded fp32
// quotient differs from a/b by < 2^-9/|b| while a non-integer a/b is >= 1/|b| away from an integer.
DEV int np_floordiv(int a, int b) {
  if (b == 0) return 0;
  return (int)floorf(__fdiv_rn((float)a, (float)b));
}

// Division by a launch-invariant scalar threshold.  magic = ceil(2^32 / T) is exact for 0 <= a < 2^16
// and 2 <= T < 2^15 (a * (magic*T - 2^32) < 2^31).
struct FastDiv {
  int T;
  unsigned magic;
};
DEV int div_T(int a, const FastDiv &f) {
  if (f.T >= 2 && a >= 0) return (int)__umulhi((unsigned)a, f.magic);
  if (f.T == 1) return a;
  return np_floordiv(a, f.T);
}

DEV int clip(int x, int lo, int hi) {  // np.clip == minimum(maximum(x, lo), hi)
  return min(max(x, lo), hi);
}

// DTYPE(history_weight * ref_frame), gs:249: float64 product, C cast toward zero to int16.
DEV int hist_weight(double hw, bool hw_is_one, int r) {
  if (hw_is_one) return r;
  return w16(__double2int_rz(hw * (double)r));
}

// ---- thresholded_difference[_adpt], gs:67-78 / gs:104-113 --------------------------------------
DEV void threshold_px(int c, int r, int T, int negT, int &d, int &a, int &s) {
  d = w16(c - r);
  a = w16(d < 0 ? -d : d);  // np.abs wraps at -32768
  s = (a > T) ? 1 : 0;
  a = s ? a : 0;            // abs_diff * spikes
  if (d < negT) s = -1;
}

// ---- update_reference_*, gs:225-425 (Appendix A.3) ---------------------------------------------
struct UpdateArgs {
  int mode, uncoded;
  int threshold, min_thr, max_thr, thr_down, thr_up, max_time_ms;
  int lut_len;
  const uint8_t *lut;
  double hw;
  int hw_is_one;
  FastDiv div_thr;
};

// Returns the new reference value; for DVS_UPD_RATE_ADPT also advances the threshold: th_state is
// what the reference leaves in the caller's array (gs:292-295), th_ret what it returns (gs:297).
DEV int update_px(const UpdateArgs &U, int a, int s, int r, int &th_state, int &th_ret,
                  unsigned &status) {
  const int base = hist_weight(U.hw, U.hw_is_one != 0, r);
  int upd;
  switch (U.mode) {
    case DVS_UPD_RATE: {  // gs:244-249 (== gs:330-332)
      int n = clip(w16(div_T(a, U.div_thr)), 0, U.max_time_ms);
      upd = w16(w16(n * s) * U.threshold);
      break;
    }
    case DVS_UPD_RATE_ADPT: {  // gs:287-297, gsu:281-295
      int th = th_state;
      int n = w16(np_floordiv(a, th));
      upd = w16(w16(n * s) * U.min_thr);
      if (!U.uncoded) {
        th = w16(th + w16((s != 0 ? 1 : 0) * U.thr_up));
        th = w16(th - w16((s == 0 ? 1 : 0) * U.thr_down));
        th_ret = w16(clip(th, U.min_thr, U.max_thr));
      } else {
        th = w16(th + w16(((s != 0 && th < U.max_thr) ? 1 : 0) * U.thr_up));
        th = w16(th - w16(((s == 0 && th > U.min_thr) ? 1 : 0) * U.thr_down));
        th_ret = th;
      }
      th_state = th;
      // gs:289 adds before clipping; no max_time_ms clip in this rule
      return w16(clip(w16(base + upd), 0, 255));
    }
    case DVS_UPD_TIME_BIN_RAW:
    case DVS_UPD_TIME_BIN_THR: {  // gs:374-376, gs:421-423
      int idx = (U.mode == DVS_UPD_TIME_BIN_THR) ? w16(div_T(a, U.div_thr)) : a;
      if (idx < 0) idx += U.lut_len;  // NumPy negative index
      int mult = 0;
      if (idx < 0 || idx >= U.lut_len) status |= DVS_STATUS_LUT_INDEX;
      else mult = U.lut[idx];
      upd = w16(s * mult);
      if (U.mode == DVS_UPD_TIME_BIN_THR) upd = w16(upd * U.threshold);
      break;
    }
    default:
      return r;
  }
  return w16(clip(w16(base + upd), 0, 255));
}

// ---- slot encoders, Appendix A.8 ---------------------------------------------------------------
struct EncArgs {
  int mode, uncoded, threshold, n_slots, u16_vals, lut_len;
  const uint8_t *lut;
  FastDiv dv;
};

// Python floor division of C ints (cdivision=False); threshold == 0 is rejected on the host.
DEV int enc_div(int val, const EncArgs &E) { return div_T(val, E.dv); }

// Slot of bit i (MSB-first index 0..7) of a time-binary code, or -1 when the reference would raise.
DEV int bin_slot(const EncArgs &E, int i) {
  int t;
  if (E.mode == DVS_ENC_TIME_BIN) t = i;  // gs:991-992
  else {
    int ii = min(i, E.n_slots);                      // gs:1075-1076
    t = E.uncoded ? E.n_slots - ii : ii - 1;         // gsu:978 / gs:1077
    if (t < 0) t += E.n_slots;                       // list[-1]
  }
  return (t >= 0 && t < E.n_slots) ? t : -1;
}

// Per-record descriptor: RATE -> k (slots [0,k)), TIME -> slot (or -1), TIME_BIN* -> 8-bit code.
DEV int enc_desc(const EncArgs &E, int val_raw, bool is_pos, unsigned &status) {
  const int val = E.u16_vals ? (val_raw & 0xffff) : val_raw;
  switch (E.mode) {
    case DVS_ENC_RATE: {
      int v = w16(enc_div(val, E));  // cdef DTYPE_t val, gs:830
      int k;
      if (!E.uncoded) k = max(0, w16((E.n_slots - 1) - v));  // gs:831-832
      else k = max(0, min(E.n_slots - 1, v));                // gsu:752
      if (k > E.n_slots) { k = E.n_slots; status |= DVS_STATUS_SLOT_RANGE; }
      return k;
    }
    case DVS_ENC_TIME: {
      int nt = min(enc_div(val, E) - 1, E.n_slots - 1);  // gs:910
      if (!E.uncoded && !is_pos) nt = max(nt, 0);        // gs:919-921 (gsu:829: no max)
      int t = E.n_slots - nt - 1;                        // gs:924
      if (t < 0 || t >= E.n_slots) { status |= DVS_STATUS_SLOT_RANGE; return -1; }
      return t;
    }
    case DVS_ENC_TIME_BIN:
    case DVS_ENC_TIME_BIN_THR: {
      int idx = (E.mode == DVS_ENC_TIME_BIN_THR) ? enc_div(val, E) : val;  // gs:1068 / gs:987
      if (idx < 0) idx += E.lut_len;
      if (idx < 0 || idx >= E.lut_len) { status |= DVS_STATUS_LUT_INDEX; return 0; }
      return E.lut[idx];
    }
    default:
      return 0;
  }
}

// How many times the record appears in slot j.  slotmask (time-binary modes): bit (0x80 >> i) set
// when bit i maps to slot j.
DEV int enc_mult(const EncArgs &E, int desc, int j, unsigned slotmask_j) {
  switch (E.mode) {
    case DVS_ENC_RATE: return desc > j;
    case DVS_ENC_TIME: return desc == j;
    case DVS_ENC_TIME_BIN:
    case DVS_ENC_TIME_BIN_THR: return __popc((unsigned)desc & slotmask_j);
    default: return 0;
  }
}

DEV unsigned bin_slotmask(const EncArgs &E, int j, unsigned &invalid_bits) {
  unsigned m = 0;
  invalid_bits = 0;
  for (int i = 0; i < 8; ++i) {
    int t = bin_slot(E, i);
    if (t == j) m |= 0x80u >> i;
    if (t < 0) invalid_bits |= 0x80u >> i;
  }
  return m;
}

// ---- split_spikes list membership, gs:673-713 / gsu:644-648 ------------------------------------
DEV void split_px(int s, int polarity, int uncoded, bool &to_pos, bool &to_neg) {
  if (uncoded || polarity == DVS_POL_MERGED) { to_pos = s > 0; to_neg = s < 0; }
  else if (polarity == DVS_POL_RECTIFIED) { to_pos = s != 0; to_neg = false; }
  else if (polarity == DVS_POL_UP) { to_pos = s > 0; to_neg = false; }
  else if (polarity == DVS_POL_DOWN) { to_pos = s < 0; to_neg = false; }
  else { to_pos = false; to_neg = false; }
}

// ---- keys, gs:763-778 / gsu:694-705 ------------------------------------------------------------
DEV int spike_key(int row, int col, int is_pos, int uncoded, int flag_shift, int data_shift,
                  int data_mask) {
  unsigned long long d = 0;
  if (!uncoded) {
    if (is_pos) d |= 1ull;
    d |= (unsigned long long)(long long)(row << 1);
    d |= (unsigned long long)((long long)col << ((data_shift + 1) & 63));
  } else {
    if (is_pos) d |= 1ull << (flag_shift & 63);
    d |= (unsigned long long)((long long)(row & data_mask) << (data_shift & 63));
    d |= (unsigned long long)(col & data_mask);
  }
  return (int)(short)(unsigned short)d;
}

// ---- packed int16x2 helpers --------------------------------------------------------------------
DEV int get16(const uint32_t (&pk)[4], int p) { return (int)(short)(pk[p >> 1] >> ((p & 1) * 16)); }
DEV void set16(uint32_t (&pk)[4], int p, int v) {
  const uint32_t sh = (p & 1) * 16;
  pk[p >> 1] = (pk[p >> 1] & ~(0xffffu << sh)) | (((uint32_t)v & 0xffffu) << sh);
}

}  // namespace dvs
