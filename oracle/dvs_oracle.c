/*
 * dvs_oracle.c -- CPU restatement of pyDVS's spike-generation path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * link or call this file.  The product (pydvs_b200/) never does: it has no CPU path at all.
 *
 * Every function restates, as plain scalar C loops over flat row-major planes, the arithmetic of
 * one function of the reference's Cython modules (citations are /root/reference/pydvs/<file>:<lines>;
 * "gs" = generate_spikes.pyx, "gsu" = generate_spikes_uncoded.pyx, "cpp" = cpp/dvs_op.cpp).
 * The reference computes with NumPy int16 arrays and Python-int scalars, so every intermediate that
 * NumPy would hold in an int16 array is narrowed with w16() here (two's-complement wrap), and
 * integer division follows NumPy: floor division, x // 0 == 0.
 *
 * Parity status: PINNED -- this file is checked in tests/ against (a) the unmodified reference
 * compiled into oracle/_ref by oracle/build_ref.py (live differential tests, when present) and
 * (b) golden vectors generated from that compiled reference, committed under tests/golden/.
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -o oracle/libdvs_oracle.so oracle/dvs_oracle.c
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_OK 0
#define ORC_ERR_INDEX 1     /* the reference raises IndexError here                */
#define ORC_ERR_CAPACITY 2  /* caller-provided output buffer too small             */
#define ORC_ERR_ZERODIV 3   /* the reference raises ZeroDivisionError here         */
#define ORC_ERR_UNDEFINED 4 /* the reference indexes a list out of range (UB, B.7) */

enum { POL_UP = 0, POL_DOWN = 1, POL_MERGED = 2, POL_RECTIFIED = 3 };

static inline int16_t w16(int32_t x) { return (int16_t)(uint16_t)(uint32_t)x; }

/* NumPy integer floor_divide on int16 operands (result narrowed by the caller). */
static inline int32_t np_floordiv(int32_t a, int32_t b) {
  if (b == 0) return 0;
  int32_t q = a / b, r = a % b;
  if (r != 0 && ((r < 0) != (b < 0))) q -= 1;
  return q;
}

static inline int32_t np_clip(int32_t x, int32_t lo, int32_t hi) {
  /* np.clip == minimum(maximum(x, lo), hi), also when lo > hi */
  int32_t y = x < lo ? lo : x;
  return y > hi ? hi : y;
}

/* int16(trunc(float64(hw) * x)) -- gs:249 `DTYPE(history_weight*ref_frame)` */
static inline int16_t hist(double hw, int16_t r) { return (int16_t)(hw * (double)r); }

/* ---- A.1 thresholded_difference / _adpt : gs:49-78, gs:83-113 (gsu:43-106 identical) ---------- */
void orc_thresholded_difference(const int16_t *curr, const int16_t *ref, const int16_t *thr_plane,
                                int thr_scalar, int16_t *diff, int16_t *abs_diff, int16_t *spikes,
                                long n) {
  for (long i = 0; i < n; ++i) {
    int16_t d = w16((int32_t)curr[i] - (int32_t)ref[i]);   /* gs:67  */
    int16_t a = w16(d < 0 ? -(int32_t)d : (int32_t)d);     /* gs:68 np.abs wraps at -32768 */
    int32_t t, negt;
    if (thr_plane) { t = thr_plane[i]; negt = w16(-t); }   /* gs:110: -threshold is an int16 array */
    else           { t = thr_scalar;   negt = -t; }        /* gs:74: Python int */
    int16_t s = (a > t) ? 1 : 0;                           /* gs:70  */
    a = w16((int32_t)a * s);                               /* gs:72  */
    if (d < negt) s = -1;                                  /* gs:74-75 */
    diff[i] = d; abs_diff[i] = a; spikes[i] = s;
  }
}

/* ---- A.2 local_inhibition : gs:177-221 with argmax gs:118-132, transform_coords gs:136-151 ----- */
int orc_local_inhibition(int16_t *spikes, const int16_t *abs_diff, int W, int H, int k) {
  if (k <= 0) return ORC_ERR_ZERODIV;
  long new_w = (long)k * k;
  long new_h = ((long)H * W) / new_w;      /* gs:202-204 */
  long ratio = W / k;
  if (new_h > 0 && ratio == 0) return ORC_ERR_ZERODIV;
  int16_t *keep = (int16_t *)calloc((size_t)H * W + 1, sizeof(int16_t));
  if (!keep) return ORC_ERR_CAPACITY;
  for (long u = 0; u < new_h; ++u) {
    long r0 = k * (u / ratio), c0 = k * (u % ratio);
    if (r0 + k > H || c0 + k > W) { free(keep); return ORC_ERR_INDEX; } /* fancy index out of range */
    long best = 0; int16_t bv = abs_diff[r0 * W + c0];
    for (long v = 1; v < new_w; ++v) {                      /* gs:127-130 strict > keeps the first */
      int16_t x = abs_diff[(r0 + v / k) * W + (c0 + v % k)];
      if (x > bv) { bv = x; best = v; }
    }
    long p = (r0 + best / k) * W + (c0 + best % k);
    keep[p] = spikes[p];                                    /* gs:216 */
  }
  memcpy(spikes, keep, (size_t)H * W * sizeof(int16_t));    /* gs:217-218 */
  free(keep);
  return ORC_OK;
}

/* ---- A.3 reference updates --------------------------------------------------------------------- */
/* update_reference_rate gs:225-251 == update_reference_time_thresh gs:301-334 */
void orc_update_reference_rate(const int16_t *abs_diff, const int16_t *spikes, const int16_t *ref,
                               int thr, int max_time_ms, double hw, int16_t *ref_out, long n) {
  for (long i = 0; i < n; ++i) {
    int16_t ns = w16(np_floordiv(abs_diff[i], thr));                  /* gs:244 */
    ns = w16(np_clip(ns, 0, max_time_ms));                            /* gs:247 */
    int16_t upd = w16((int32_t)w16((int32_t)ns * spikes[i]) * thr);   /* gs:249 */
    ref_out[i] = w16(np_clip(w16((int32_t)hist(hw, ref[i]) + upd), 0, 255));
  }
}

/* update_reference_rate_adpt: coded gs:256-297, uncoded gsu:251-295.  thr is updated IN PLACE to the
 * unclipped value (gs:292-295); thr_out receives what the function returns (coded: clipped copy,
 * uncoded: the same mutated values). */
void orc_update_reference_rate_adpt(const int16_t *abs_diff, const int16_t *spikes, const int16_t *ref,
                                    int16_t *thr, int min_thr, int max_thr, int down, int up,
                                    double hw, int uncoded, int16_t *ref_out, int16_t *thr_out,
                                    long n) {
  for (long i = 0; i < n; ++i) {
    int16_t th = thr[i];
    int16_t ns = w16(np_floordiv(abs_diff[i], th));                        /* gs:287 */
    int16_t upd = w16((int32_t)w16((int32_t)ns * spikes[i]) * min_thr);    /* gs:289 */
    int16_t r = w16((int32_t)hist(hw, ref[i]) + upd);
    if (!uncoded) {
      th = w16(th + w16((spikes[i] != 0 ? 1 : 0) * up));                   /* gs:291-292 */
      th = w16(th - w16((spikes[i] == 0 ? 1 : 0) * down));                 /* gs:294-295 */
      thr_out[i] = w16(np_clip(th, min_thr, max_thr));                     /* gs:297 */
    } else {
      th = w16(th + w16(((spikes[i] != 0 && th < max_thr) ? 1 : 0) * up));   /* gsu:286-289 */
      th = w16(th - w16(((spikes[i] == 0 && th > min_thr) ? 1 : 0) * down)); /* gsu:291-293 */
      thr_out[i] = th;                                                       /* gsu:295 */
    }
    thr[i] = th;
    ref_out[i] = w16(np_clip(r, 0, 255));
  }
}

/* update_reference_time_binary_raw gs:339-378 (thresh=0) / _binary_thresh gs:382-425 (thresh=1) */
int orc_update_reference_time_binary(const int16_t *abs_diff, const int16_t *spikes,
                                     const int16_t *ref, int thr, double hw, const uint8_t *lut,
                                     int lut_len, int thresh, int16_t *ref_out, long n) {
  for (long i = 0; i < n; ++i) {
    int32_t idx = thresh ? w16(np_floordiv(abs_diff[i], thr)) : abs_diff[i];  /* gs:374 / gs:421 */
    if (idx < 0) idx += lut_len;                       /* NumPy negative index wrap */
    if (idx < 0 || idx >= lut_len) return ORC_ERR_INDEX;
    int16_t mult = lut[idx];
    int16_t upd = w16((int32_t)spikes[i] * mult);
    if (thresh) upd = w16((int32_t)upd * thr);         /* gs:423 spikes*mult*threshold */
    ref_out[i] = w16(np_clip(w16((int32_t)hist(hw, ref[i]) + upd), 0, 255));
  }
  return ORC_OK;
}

/* noise variants gs:428-449 (raw) and gs:453-473 (thresh) with the random pixel indices INJECTED
 * (the reference draws them from the unseeded global NumPy RNG, Appendix B.9).  `sample` holds the
 * int16-cast flat indices exactly as gs:442/465 produce them; rows = sample // W, cols = sample % W
 * in int16 floor arithmetic, negative results wrap like NumPy indices. */
int orc_update_reference_noise(const int16_t *abs_diff, const int16_t *spikes, const int16_t *ref,
                               int W, int H, int thr, int max_time_ms, double hw, const uint8_t *lut,
                               int lut_len, int thresh_variant, const int16_t *sample, long n_sample,
                               int16_t *ref_out) {
  long n = (long)W * H;
  int16_t *mult = (int16_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int16_t));
  if (!mult) return ORC_ERR_CAPACITY;
  for (long i = 0; i < n; ++i) {
    if (thresh_variant) {
      mult[i] = w16(np_clip(w16(np_floordiv(abs_diff[i], thr)), 0, max_time_ms)); /* gs:463 */
    } else {
      int32_t idx = abs_diff[i];
      if (idx < 0) idx += lut_len;
      if (idx < 0 || idx >= lut_len) { free(mult); return ORC_ERR_INDEX; }
      mult[i] = lut[idx];                                                          /* gs:440 */
    }
  }
  for (long j = 0; j < n_sample; ++j) {
    int32_t r = w16(np_floordiv(sample[j], W));
    int32_t c = sample[j] - np_floordiv(sample[j], W) * W;
    c = w16(c);
    if (r < 0) r += H;
    if (c < 0) c += W;
    if (r < 0 || r >= H || c < 0 || c >= W) { free(mult); return ORC_ERR_INDEX; }
  }
  if (thresh_variant) {
    for (long j = 0; j < n_sample; ++j) {                                          /* gs:469 */
      int32_t r = w16(np_floordiv(sample[j], W)), c = w16(sample[j] - np_floordiv(sample[j], W) * W);
      if (r < 0) r += H;
      if (c < 0) c += W;
      mult[(long)r * W + c] = 0;
    }
  } else {
    /* gs:445 `mult[rows,cols] = mult[rows,cols] ^ 0xFF`: gather-then-scatter, so a pixel listed
     * twice is still flipped only once relative to its ORIGINAL value. */
    int16_t *flipped = (int16_t *)malloc((size_t)(n_sample > 0 ? n_sample : 1) * sizeof(int16_t));
    if (!flipped) { free(mult); return ORC_ERR_CAPACITY; }
    for (long j = 0; j < n_sample; ++j) {
      int32_t r = w16(np_floordiv(sample[j], W)), c = w16(sample[j] - np_floordiv(sample[j], W) * W);
      if (r < 0) r += H;
      if (c < 0) c += W;
      flipped[j] = w16(mult[(long)r * W + c] ^ 0xFF);
    }
    for (long j = 0; j < n_sample; ++j) {
      int32_t r = w16(np_floordiv(sample[j], W)), c = w16(sample[j] - np_floordiv(sample[j], W) * W);
      if (r < 0) r += H;
      if (c < 0) c += W;
      mult[(long)r * W + c] = flipped[j];
    }
    free(flipped);
  }
  for (long i = 0; i < n; ++i)  /* gs:447 / gs:471: no clip, and the thresh variant omits *T */
    ref_out[i] = w16((int32_t)hist(hw, ref[i]) + w16((int32_t)spikes[i] * mult[i]));
  free(mult);
  return ORC_OK;
}

/* ---- A.4 log2 LUT : encode_time_n_bits_single gs:477-500, generate_log2_table gs:1105-1118 ------ */
static uint8_t enc_n_bits(int value, int num_bits) {
  if (value == 0) return 0;
  int hi = 31 - __builtin_clz((unsigned)value);     /* DTYPE(np.log2(v)) for an exact-int v */
  uint8_t mult = (uint8_t)(1u << hi);               /* DTYPE_U8(np.power(2, hi)): wraps above 255 */
  if (num_bits == 1) return mult;
  for (int it = 0; it < num_bits - 1; ++it) {       /* gs:493-498 */
    int rest = value - mult;
    if (rest > 0) {
      int b = 31 - __builtin_clz((unsigned)rest);
      mult |= (uint8_t)(1u << b);
    } else break;
  }
  return mult;
}

void orc_generate_log2_table(int max_active_bits, int bit_resolution, uint8_t *table) {
  int n = 1 << bit_resolution;
  for (int r = 0; r < max_active_bits; ++r)
    for (int v = 0; v < n; ++v) table[(long)r * n + v] = enc_n_bits(v, r);
}

/* ---- A.5 split_spikes : coded gs:651-719, uncoded gsu:623-672 ---------------------------------- */
/* neg/pos are [3][cap] row-major int16 (the uncoded twin's uint16 has the same bits).  Returns the
 * counts through n_neg/n_pos and global_max. */
int orc_split_spikes(const int16_t *spikes, const int16_t *abs_diff, int W, int H, int polarity,
                     int uncoded, int16_t *neg, int16_t *pos, long cap, long *n_neg, long *n_pos,
                     int *global_max) {
  long nn = 0, np_ = 0;
  int have_n = 0, have_p = 0;
  int16_t mx_n = 0, mx_p = 0;
  for (long r = 0; r < H; ++r)
    for (long c = 0; c < W; ++c) {
      int16_t s = spikes[r * W + c], a = abs_diff[r * W + c];
      int to_pos, to_neg;
      if (uncoded) { to_pos = s > 0; to_neg = s < 0; }              /* gsu:644-648 always both */
      else switch (polarity) {
        case POL_RECTIFIED: to_pos = s != 0; to_neg = 0; break;     /* gs:673-680 */
        case POL_UP:        to_pos = s > 0;  to_neg = 0; break;     /* gs:682-689 */
        case POL_DOWN:      to_pos = s < 0;  to_neg = 0; break;     /* gs:692-699 pos <- negatives */
        case POL_MERGED:    to_pos = s > 0;  to_neg = s < 0; break; /* gs:701-713 */
        default:            to_pos = 0; to_neg = 0; break;          /* no branch taken */
      }
      if (to_pos) {
        if (np_ >= cap) return ORC_ERR_CAPACITY;
        pos[0 * cap + np_] = (int16_t)r; pos[1 * cap + np_] = (int16_t)c; pos[2 * cap + np_] = a;
        if (!have_p || a > mx_p) { mx_p = a; have_p = 1; }
        ++np_;
      }
      if (to_neg) {
        if (nn >= cap) return ORC_ERR_CAPACITY;
        neg[0 * cap + nn] = (int16_t)r; neg[1 * cap + nn] = (int16_t)c; neg[2 * cap + nn] = a;
        if (!have_n || a > mx_n) { mx_n = a; have_n = 1; }
        ++nn;
      }
    }
  int g = 0;
  if (!uncoded) {
    if (have_n && have_p) g = mx_n > mx_p ? mx_n : mx_p;
    else if (have_n) g = mx_n;
    else if (have_p) g = mx_p;
  } else {                                                           /* gsu:650-665 */
    if (polarity == POL_MERGED) {
      if (have_n && have_p) g = mx_n > mx_p ? mx_n : mx_p;
      else if (have_n) g = mx_n;
      else if (have_p) g = mx_p;
    } else if (polarity == POL_UP) { if (have_p) g = mx_p; }
    else { if (have_n) g = mx_n; }
  }
  *n_neg = nn; *n_pos = np_; *global_max = g;
  return ORC_OK;
}

/* ---- A.6 keys : coded gs:744-778, uncoded gsu:676-705 ------------------------------------------ */
int16_t orc_spike_to_key(int row, int col, int flag_shift, int data_shift, int data_mask, int is_pos,
                         int uncoded) {
  uint64_t d = 0;
  if (!uncoded) {
    if (is_pos) d |= 1;                                             /* gs:766-768 */
    d |= (uint64_t)(int64_t)((int32_t)row << 1);                    /* gs:772 */
    d |= (uint64_t)((int64_t)col << ((data_shift + 1) & 63));       /* gs:776 */
  } else {
    if (is_pos) d |= (uint64_t)1 << (flag_shift & 63);              /* gsu:698 */
    d |= (uint64_t)((int64_t)(row & data_mask) << (data_shift & 63)); /* gsu:701 */
    d |= (uint64_t)(col & data_mask);                               /* gsu:703 */
  }
  return (int16_t)(uint16_t)d;                                      /* DTYPE_U16_t d -> DTYPE_t */
}

/* ---- A.7/A.8 encoders.  Output = CSR: slot_counts[n_slots] and `out` filled slot-major, inside a
 * slot all pos entries (input order) then all neg entries.  For key modes `out` holds one int32 per
 * event (the int16 key widened); for the uncoded rate/time encoders it holds (row, col, +1|-1)
 * triples (gsu:748-757, gsu:820-827).  rows/cols/vals are read as int16, or as uint16 when
 * `u16_in` (the twin's DTYPE_U16_t signatures gsu:710-711, 773-774). ---------------------------- */
enum { ENC_RATE = 0, ENC_TIME = 1, ENC_TIME_BIN = 2, ENC_TIME_BIN_THR = 3 };

typedef struct {
  int mode, uncoded, u16_in;
  int threshold;      /* rate: threshold;  time / time_bin_thr: min_threshold */
  int n_slots;        /* rate: max_time_ms; others: num_bins */
  int flag_shift, data_shift, data_mask;
  const uint8_t *lut; int lut_len;
} orc_enc_params;

static inline int32_t rd(const int16_t *p, int u16_in) {
  return u16_in ? (int32_t)(uint16_t)*p : (int32_t)*p;
}

/* Python floor division between C ints with cdivision=False (gs:830, gs:910, gs:1068). */
static inline int py_floordiv(int a, int b, int *err) {
  if (b == 0) { *err = ORC_ERR_ZERODIV; return 0; }
  int q = a / b, r = a % b;
  if (r != 0 && ((r < 0) != (b < 0))) q -= 1;
  return q;
}

/* Visits the slots one pixel contributes to, in the order the reference appends them. */
static int enc_slots(const orc_enc_params *p, int32_t val, int is_pos, int *slots, int *n_out) {
  int err = 0, n = 0;
  switch (p->mode) {
    case ENC_RATE: {
      int v = (int16_t)py_floordiv(val, p->threshold, &err);         /* cdef DTYPE_t val */
      if (err) return err;
      int k;
      if (!p->uncoded) {
        v = (int16_t)((uint32_t)(p->n_slots - 1) - (uint32_t)v);     /* gs:831 unsigned, narrowed */
        k = v > 0 ? v : 0;                                           /* gs:832 */
      } else {
        /* gsu:752 min(max_spikes-1, val) */
        k = (p->n_slots - 1) < v ? (p->n_slots - 1) : v;
        if (k < 0) k = 0;
      }
      for (int j = 0; j < k; ++j) slots[n++] = j;                    /* gs:847-848 */
      break;
    }
    case ENC_TIME: {
      int nt = py_floordiv(val, p->threshold, &err) - 1;             /* gs:910 */
      if (err) return err;
      if (nt > p->n_slots - 1) nt = p->n_slots - 1;
      if (!p->uncoded && !is_pos && nt < 0) nt = 0;                  /* gs:919-921; gsu:829 has no max */
      int t = p->n_slots - nt - 1;                                   /* gs:924 */
      if (t < 0) t += p->n_slots;                                    /* Python wraparound stays on */
      if (t < 0 || t >= p->n_slots) return ORC_ERR_UNDEFINED;        /* B.7: UB in the reference */
      slots[n++] = t;
      break;
    }
    case ENC_TIME_BIN:
    case ENC_TIME_BIN_THR: {
      int idx = val;
      if (p->mode == ENC_TIME_BIN_THR) {
        idx = py_floordiv(val, p->threshold, &err);                  /* gs:1068 */
        if (err) return err;
      }
      if (idx < 0) idx += p->lut_len;
      if (idx < 0 || idx >= p->lut_len) return ORC_ERR_INDEX;
      uint8_t code = p->lut[idx];
      for (int i = 0; i < 8; ++i) {                                  /* np.unpackbits: MSB first */
        if (!(code & (0x80u >> i))) continue;
        int t;
        if (p->mode == ENC_TIME_BIN) t = i;                          /* gs:991-992 */
        else {
          int ii = i > p->n_slots ? p->n_slots : i;                  /* gs:1075-1076 */
          t = p->uncoded ? p->n_slots - ii : ii - 1;                 /* gsu:978 / gs:1077 */
        }
        if (t < 0) t += p->n_slots;                                  /* list[-1] wraps */
        if (t < 0 || t >= p->n_slots) return ORC_ERR_INDEX;          /* Python list IndexError */
        slots[n++] = t;
      }
      break;
    }
    default: return ORC_ERR_INDEX;
  }
  *n_out = n;
  return ORC_OK;
}

int orc_make_spike_lists(const orc_enc_params *p, const int16_t *pos, long n_pos, long pos_stride,
                         const int16_t *neg, long n_neg, long neg_stride, long *slot_counts,
                         int32_t *out, long out_cap, long *n_events) {
  int triples = p->uncoded && (p->mode == ENC_RATE || p->mode == ENC_TIME);
  int width = triples ? 3 : 1;
  int max_per_px = p->n_slots > 8 ? p->n_slots : 8;
  int *slots = (int *)malloc(sizeof(int) * (size_t)(max_per_px + 1));
  long *fill = (long *)calloc((size_t)p->n_slots + 1, sizeof(long));
  if (!slots || !fill) { free(slots); free(fill); return ORC_ERR_CAPACITY; }
  for (int j = 0; j < p->n_slots; ++j) slot_counts[j] = 0;
  int rc = ORC_OK;
  for (int pass = 0; pass < 2 && rc == ORC_OK; ++pass) {
    if (pass == 1) {
      long acc = 0;
      for (int j = 0; j < p->n_slots; ++j) { fill[j] = acc; acc += slot_counts[j]; }
      *n_events = acc;
      if (acc > out_cap) { rc = ORC_ERR_CAPACITY; break; }
    }
    for (long i = 0; i < n_pos + n_neg && rc == ORC_OK; ++i) {
      int is_pos = i < n_pos;
      const int16_t *src = is_pos ? pos : neg;
      long stride = is_pos ? pos_stride : neg_stride, j = is_pos ? i : i - n_pos;
      int32_t row = rd(src + 0 * stride + j, p->u16_in), col = rd(src + 1 * stride + j, p->u16_in);
      int32_t val = rd(src + 2 * stride + j, p->u16_in);
      int n = 0;
      rc = enc_slots(p, val, is_pos, slots, &n);
      if (rc != ORC_OK) break;
      for (int q = 0; q < n; ++q) {
        if (pass == 0) { slot_counts[slots[q]]++; continue; }
        int32_t *dst = out + width * fill[slots[q]]++;
        if (triples) { dst[0] = row; dst[1] = col; dst[2] = is_pos ? 1 : -1; }
        else dst[0] = orc_spike_to_key((int16_t)row, (int16_t)col, p->flag_shift, p->data_shift,
                                       p->data_mask, is_pos, p->uncoded);
      }
    }
  }
  free(slots); free(fill);
  return rc;
}

/* ---- A.10 native float mode : cpp/dvs_op.cpp:4-33 (one pass, in place; `ev` untouched) --------- */
/* Compile with -ffp-contract=off: the x86-64 reference build does not fuse relax*ref + diff. */
void orc_dvs_op_f32(const float *src, float *diff, float *ref, float *thr, float relax, float up,
                    float down, long n) {
  for (long i = 0; i < n; ++i) {
    float d = src[i] - ref[i];                       /* cpp:13 */
    int test = (d < -thr[i]) || (d > thr[i]);        /* cpp:14 */
    d = d * (float)test;                             /* cpp:15 */
    diff[i] = d;
    ref[i] = (relax * ref[i]) + d;                   /* cpp:16 */
    thr[i] = test ? thr[i] * up : thr[i] * down;     /* cpp:17-21 */
  }
}

/* ---- whole-frame pipeline, the reference's call sequence stand_alone.py:160-185 for one stream --
 * thresholded_difference[_adpt] -> local_inhibition -> update_reference_* -> split_spikes ->
 * make_spike_lists_*.  Used for multi-frame differential tests and as the "port" CPU baseline. */
typedef struct {
  int W, H;
  int adaptive;        /* 0: scalar threshold; 1: per-pixel threshold plane + rate_adpt update */
  int update_mode;     /* 0 rate/time_thresh, 2 time_binary_raw, 3 time_binary_thresh */
  int threshold, min_thr, max_thr, down, up, max_time_ms;
  double hw;
  int inh_k;           /* <= 1: off */
  int polarity;
  int uncoded;
  orc_enc_params enc;
} orc_pipe_params;

int orc_pipeline_frame(const orc_pipe_params *pp, const int16_t *curr, int16_t *ref, int16_t *thr,
                       int16_t *scratch /* 5 planes + 6*P */, long *slot_counts, int32_t *out,
                       long out_cap, long *n_events) {
  long P = (long)pp->W * pp->H;
  int16_t *diff = scratch, *abs_diff = scratch + P, *spikes = scratch + 2 * P;
  int16_t *ref_new = scratch + 3 * P, *thr_new = scratch + 4 * P;
  int16_t *neg = scratch + 5 * P, *pos = scratch + 8 * P;
  orc_thresholded_difference(curr, ref, pp->adaptive ? thr : NULL, pp->threshold, diff, abs_diff,
                             spikes, P);
  int rc;
  if (pp->inh_k > 1) {
    rc = orc_local_inhibition(spikes, abs_diff, pp->W, pp->H, pp->inh_k);
    if (rc) return rc;
  }
  if (pp->adaptive) {
    orc_update_reference_rate_adpt(abs_diff, spikes, ref, thr, pp->min_thr, pp->max_thr, pp->down,
                                   pp->up, pp->hw, pp->uncoded, ref_new, thr_new, P);
    memcpy(thr, thr_new, (size_t)P * sizeof(int16_t));   /* callers do thr[:] = returned */
  } else if (pp->update_mode == 0) {
    orc_update_reference_rate(abs_diff, spikes, ref, pp->threshold, pp->max_time_ms, pp->hw, ref_new, P);
  } else {
    rc = orc_update_reference_time_binary(abs_diff, spikes, ref, pp->threshold, pp->hw, pp->enc.lut,
                                          pp->enc.lut_len, pp->update_mode == 3, ref_new, P);
    if (rc) return rc;
  }
  memcpy(ref, ref_new, (size_t)P * sizeof(int16_t));
  long nn = 0, np_ = 0; int gmax = 0;
  rc = orc_split_spikes(spikes, abs_diff, pp->W, pp->H, pp->polarity, pp->uncoded, neg, pos, P, &nn,
                        &np_, &gmax);
  if (rc) return rc;
  return orc_make_spike_lists(&pp->enc, pos, np_, P, neg, nn, P, slot_counts, out, out_cap, n_events);
}

/* ================================================================================================
 * float NVS variants, pydvs/numba_nvs.py:6-147 (A.11).
 * Arithmetic as numba types it [probed by running the reference under numba 0.65]:
 *  - thresholded_difference (nvs:6-20) works on float32 arrays only -> float32 arithmetic; `//` is
 *    NumPy's floor_divide (npy_divmodf: fmod, (a-mod)/b, sign fix-up, floor with the >0.5 snap);
 *  - every expression that multiplies an array by one of the Python-float scalars (leaks, base,
 *    increments, target) is evaluated in float64 and rounded to float32 once, on assignment;
 *  - `reference += spikes` (nvs:29) and `reference * leak_active + spikes` (nvs:88,123: all float32
 *    arrays) stay in float32.
 * The noisy variants draw `uniform(0,1) <= p` from numba's private, unseeded RNG (nvs:85,119,133):
 * not reproducible, so the draws are injected as byte masks (1 = leak this pixel).  p >= 1 / p < 0
 * pin the two branches against the reference.
 * ================================================================================================ */
static float np_floor_dividef(float a, float b) {
  if (b == 0.0f) return a / b;                       /* npy_divmodf: b == 0 -> a / b (inf or nan) */
  float mod = fmodf(a, b);
  float div = (a - mod) / b;
  if (mod != 0.0f) {
    if ((b < 0) != (mod < 0)) div -= 1.0f;
  }
  if (div != 0.0f) {
    float fl = floorf(div);
    if (div - fl > 0.5f) fl += 1.0f;
    return fl;
  }
  return copysignf(0.0f, a / b);
}

/* nvs:6-20.  spikes <- discretised difference, returns activity in `active` (may be NULL). */
static inline uint8_t nvs_td(float frame, float ref, float thr, float *spike) {
  float d = frame - ref;                              /* nvs:13 */
  uint8_t act = (uint8_t)(fabsf(d) >= thr);           /* nvs:14 (>=, not >); nan compares false */
  d = d * (float)act;                                 /* nvs:15 */
  *spike = np_floor_dividef(d, thr) * thr;            /* nvs:18 */
  return act;
}

enum { NVS_0 = 0, NVS_LEAKY = 1, NVS_LEAKY_ADAPTIVE = 2, NVS_NOISY_LEAKY_ADAPTIVE = 3 };

static inline float nvs_thr_update(float thr, uint8_t act, double base, double leak, double incr) {
  double tb = (double)thr - base;                     /* nvs:52-55, left to right, float64 */
  return (float)((base + tb * leak) + (tb * incr) * (double)act);
}

static inline float nvs_leak_active(uint8_t m, double ref_leak) {
  /* nvs:85-86: ones(float32) += (draw <= p) * (reference_leak - 1): float64 sum, stored as float32 */
  return (float)(1.0 + (double)m * (ref_leak - 1.0));
}

void orc_nvs_f32(int mode, const float *frame, float *ref, float *thr, float *spikes, uint8_t *active,
                 const uint8_t *leak_mask, double ref_leak, double thr_base, double thr_incr,
                 double thr_leak, long n) {
  for (long i = 0; i < n; ++i) {
    float sp;
    uint8_t act = nvs_td(frame[i], ref[i], thr[i], &sp);
    spikes[i] = sp;
    if (active) active[i] = act;
    if (mode == NVS_0) {                              /* nvs:23-29 */
      ref[i] = ref[i] + sp;
    } else if (mode == NVS_LEAKY) {                   /* nvs:32-40 */
      ref[i] = (float)((double)ref[i] * ref_leak + (double)sp);
    } else if (mode == NVS_LEAKY_ADAPTIVE) {          /* nvs:43-60 */
      thr[i] = nvs_thr_update(thr[i], act, thr_base, thr_leak, thr_incr);
      ref[i] = (float)((double)ref[i] * ref_leak + (double)sp);
    } else {                                          /* nvs:64-90 */
      thr[i] = nvs_thr_update(thr[i], act, thr_base, thr_leak, thr_incr);
      float la = nvs_leak_active(leak_mask ? leak_mask[i] : 0, ref_leak);
      ref[i] = ref[i] * la + sp;                      /* float32 arrays only */
    }
  }
}

/* nvs_bi_noisy_leaky_adaptive nvs:93-147: an ON cell (ref/thr) and an OFF cell (ref_off/thr_off);
 * the OFF reference relaxes towards target_off starting from the *updated ON reference* (nvs:135). */
void orc_nvs_bi_f32(const float *frame, float *ref, float *thr, float *spikes, float *ref_off,
                    float *thr_off, float *spikes_off, const uint8_t *leak_mask,
                    const uint8_t *leak_mask_off, double ref_leak, double thr_base, double thr_incr,
                    double thr_leak, double target_off, long n) {
  for (long i = 0; i < n; ++i) {
    float sp, spo;
    uint8_t act = nvs_td(frame[i], ref[i], thr[i], &sp);               /* nvs:107 */
    uint8_t acto = nvs_td(frame[i], ref_off[i], thr_off[i], &spo);     /* nvs:108-109 */
    spikes[i] = sp;
    if (spikes_off) spikes_off[i] = spo;
    thr[i] = nvs_thr_update(thr[i], act, thr_base, thr_leak, thr_incr);  /* nvs:111-114 */
    float la = nvs_leak_active(leak_mask ? leak_mask[i] : 0, ref_leak);  /* nvs:117-120 */
    ref[i] = ref[i] * la + sp;                                           /* nvs:121-123 */
    thr_off[i] = nvs_thr_update(thr_off[i], acto, thr_base, thr_leak, thr_incr);  /* nvs:125-128 */
    float lo = nvs_leak_active(leak_mask_off ? leak_mask_off[i] : 0, ref_leak);   /* nvs:130-132 */
    ref_off[i] = (float)((target_off + ((double)ref[i] - target_off) * (double)lo) + (double)spo);  /* nvs:133-135 */
  }
}

/* ================================================================================================
 * receiver side, pydvs/decode_spikes.pyx:46-197.
 * Keys are int16 arrays read into a C uint16 argument: a negative key raises OverflowError in the
 * reference (-> rc 5).  The decoder uses the uncoded bit layout with flag bit 0 meaning +1 (B.12).
 * ================================================================================================ */
static inline void key_to_spike(uint16_t key, int flag_shift, int data_shift, int data_mask,
                                int *row, int *col, int *sign) {
  *row = (int16_t)((key >> data_shift) & data_mask);  /* dec:58 */
  *col = (int16_t)(key & data_mask);                  /* dec:59 */
  *sign = (((key >> flag_shift) & 1) == 0) ? 1 : -1;  /* dec:62-65 */
}

/* update_ref_spike_list_rate dec:71-94: ref_out = int16(ref*hw); then += threshold*sign per key.
 * Out-of-image (row, col) is undefined behaviour in the reference (boundscheck off) -> rc 1. */
int orc_decode_rate(const int16_t *ref, int16_t *ref_out, int H, int W, const int16_t *keys, long n,
                    int threshold, double hw, int flag_shift, int data_shift, int data_mask) {
  for (long i = 0; i < (long)H * W; ++i) ref_out[i] = hist(hw, ref[i]);  /* dec:85 */
  for (long i = 0; i < n; ++i) {
    if (keys[i] < 0) return 5;
    int r, c, s;
    key_to_spike((uint16_t)keys[i], flag_shift, data_shift, data_mask, &r, &c, &s);
    if (r >= H || c >= W) return 1;
    ref_out[(long)r * W + c] = w16((int32_t)ref_out[(long)r * W + c] + (int32_t)(int16_t)threshold * s);  /* dec:90 */
  }
  return 0;
}

/* float64 -> C int16 as Cython converts a NumPy float64 object: int(x) then range check.
 * nan -> ValueError (rc 6), out of range / inf -> OverflowError (rc 5). */
static int f64_to_i16(double v, int16_t *out) {
  if (isnan(v)) return 6;
  if (isinf(v)) return 5;
  double t = trunc(v);
  if (t < -32768.0 || t > 32767.0) return 5;
  *out = (int16_t)t;
  return 0;
}

static inline double py_mod(double a, double b) {     /* np.mod on floats: sign follows the divisor */
  if (b == 0.0) return NAN;
  double m = fmod(a, b);
  if (m != 0.0) { if ((b < 0) != (m < 0)) m += b; } else m = copysign(0.0, b);
  return m;
}

/* update_ref_spike_list_time dec:98-145 (binary == 0) and _time_bin dec:149-197 (binary == 1).
 * In place on time_last / val_last (the reference mutates and returns them), ref_out is new. */
int orc_decode_time(int binary, const int16_t *ref, int16_t *ref_out, int16_t *time_last,
                    int16_t *val_last, int H, int W, const int16_t *keys, long n, int num_bins,
                    double bin_width, int time_period, uint32_t time_now, int threshold, double hw,
                    int flag_shift, int data_shift, int data_mask) {
  for (long i = 0; i < (long)H * W; ++i) ref_out[i] = hist(hw, ref[i]);  /* dec:121 / dec:172 */
  const int16_t thr = (int16_t)threshold, nb = (int16_t)num_bins, per = (int16_t)time_period;
  for (long i = 0; i < n; ++i) {
    if (keys[i] < 0) return 5;
    int r, c, s;
    key_to_spike((uint16_t)keys[i], flag_shift, data_shift, data_mask, &r, &c, &s);
    if (r >= H || c >= W) return 1;
    long q = (long)r * W + c;
    if (val_last[q] == -1) {                          /* dec:128-133 / dec:178-184 */
      int32_t first = binary ? ((int32_t)thr + 1) * s : ((int32_t)thr + 1);
      ref_out[q] = w16((int32_t)ref_out[q] + first);
      val_last[q] = w16(first);
      time_last[q] = (int16_t)(uint16_t)time_now;
      continue;
    }
    double time_diff, v;
    if (!binary) {                                    /* dec:136-137 */
      time_diff = (double)time_now - ((double)time_last[q] - ((double)nb - (double)val_last[q] * bin_width));
      v = floor((double)nb - py_mod(time_diff, (double)per) / bin_width);
    } else {                                          /* dec:186-190 */
      double lg = log2((double)val_last[q]);          /* nan for negatives, -inf for 0 */
      time_diff = (double)time_now - ((double)time_last[q] - 2.0 * (((double)nb - lg * bin_width) + 1.0));
      v = pow(2.0, floor((double)nb - py_mod(time_diff, (double)per) / bin_width));
    }
    int16_t val_now;
    int rc = f64_to_i16(v, &val_now);
    if (rc) return rc;
    int32_t upd = (int32_t)thr * s * (int32_t)val_now;  /* C int arithmetic, stored to int16 */
    ref_out[q] = w16((int32_t)ref_out[q] + upd);      /* dec:141 / dec:192 */
    val_last[q] = w16(upd);
    time_last[q] = (int16_t)(uint16_t)time_now;
  }
  return 0;
}
