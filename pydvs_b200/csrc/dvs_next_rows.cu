// Kernels for the rows either side of the spike-generation path:
//   rank 4: the float32 NVS variants of pydvs/numba_nvs.py ("nvs") as one fused element-wise pass;
//   rank 3: the receiver of pydvs/decode_spikes.pyx ("dec"): history decay + key scatter.
// Both are HBM-bound: every plane is read and written once with 16-byte accesses; the scatter part of
// the decoder touches only the pixels named by the keys.
#include <cuda_runtime.h>
#include <stdint.h>

#include "pydvs_b200.h"

namespace dvs {
int set_error(int code, const char *msg);  // dvs_b200.cu
}

namespace {

// ---- numba_nvs.py ------------------------------------------------------------------------------
// NumPy floor_divide for float32 (npy_divmodf: mod = fmod(a, b); div = (a - mod) / b; sign fix-up; floor with a
// "> 0.5" snap).  For a normal finite divisor and |a / b| < 2^21 that procedure returns exactly the mathematical
// floor of the real quotient (its intermediate error is below 2^-2, which the snap absorbs), so the common case
// takes one IEEE division, a floor and an exact-sign remainder test; everything else replays the NumPy steps.
__device__ __noinline__ float np_floor_dividef_slow(float a, float b) {
  if (b == 0.0f) return __fdiv_rn(a, b);
  const float mod = fmodf(a, b);
  float div = __fdiv_rn(__fsub_rn(a, mod), b);
  if (mod != 0.0f && ((b < 0.0f) != (mod < 0.0f))) div = __fsub_rn(div, 1.0f);
  if (div != 0.0f) {
    float fl = floorf(div);
    if (__fsub_rn(div, fl) > 0.5f) fl = __fadd_rn(fl, 1.0f);
    return fl;
  }
  return copysignf(0.0f, __fdiv_rn(a, b));
}
__device__ __forceinline__ float np_floor_dividef(float a, float b) {
  const float q = __fdiv_rn(a, b);
  const float ab = fabsf(b);
  if (fabsf(q) < 2097152.0f && ab >= 1.17549435e-38f && ab <= 3.402823466e+38f) {
    if (a == 0.0f) return q;                    // +-0 / b: NumPy returns copysign(0, a / b)
    float fl = floorf(q);
    const float rem = fmaf(-fl, b, a);          // a - fl * b with one rounding: its sign is exact
    if (b > 0.0f ? rem < 0.0f : rem > 0.0f) fl = __fsub_rn(fl, 1.0f);  // q was rounded up onto an integer
    return fl;
  }
  return np_floor_dividef_slow(a, b);
}

// thresholded_difference nvs:6-20 for one pixel: float32 only.
__device__ __forceinline__ unsigned nvs_td(float frame, float ref, float thr, float &spike) {
  float d = __fsub_rn(frame, ref);
  const unsigned act = fabsf(d) >= thr ? 1u : 0u;
  d = __fmul_rn(d, (float)act);
  spike = __fmul_rn(np_floor_dividef(d, thr), thr);
  return act;
}

// thresholds as leaky integrators nvs:52-55: float64 (Python-float scalars), one rounding to float32.
__device__ __forceinline__ float nvs_thr_update(float thr, unsigned act, double base, double leak, double incr) {
  const double tb = __dsub_rn((double)thr, base);
  return (float)__dadd_rn(__dadd_rn(base, __dmul_rn(tb, leak)), __dmul_rn(__dmul_rn(tb, incr), (double)act));
}

__device__ __forceinline__ float nvs_leak_active(unsigned m, double ref_leak) {  // nvs:85-86
  return (float)__dadd_rn(1.0, __dmul_rn((double)m, __dsub_rn(ref_leak, 1.0)));
}

struct NvsArgs {
  const float *frame;
  float *ref, *thr, *spikes;
  uint8_t *active;
  const uint8_t *leak_mask;
  float *ref_off, *thr_off, *spikes_off;
  const uint8_t *leak_mask_off;
  double ref_leak, thr_base, thr_incr, thr_leak, target_off;
  long long n;
};

template <int MODE>  // 0 nvs_0, 1 nvs_leaky, 2 nvs_leaky_adaptive, 3 nvs_noisy_leaky_adaptive, 4 nvs_bi_noisy_leaky_adaptive
__device__ __forceinline__ void nvs_px(const NvsArgs &A, float f, float &r, float &t, float &sp, unsigned &act, unsigned m,
                                       float &ro, float &to, float &spo, unsigned mo) {
  act = nvs_td(f, r, t, sp);
  if (MODE == 0) {
    r = __fadd_rn(r, sp);  // nvs:29
    return;
  }
  if (MODE == 1 || MODE == 2) {
    if (MODE == 2) t = nvs_thr_update(t, act, A.thr_base, A.thr_leak, A.thr_incr);
    r = (float)__dadd_rn(__dmul_rn((double)r, A.ref_leak), (double)sp);  // nvs:40, 58-60
    return;
  }
  unsigned acto = 0;
  if (MODE == 4) acto = nvs_td(f, ro, to, spo);  // nvs:108-109
  t = nvs_thr_update(t, act, A.thr_base, A.thr_leak, A.thr_incr);
  r = __fadd_rn(__fmul_rn(r, nvs_leak_active(m, A.ref_leak)), sp);  // nvs:88-90 / 121-123: float32 arrays
  if (MODE == 4) {
    to = nvs_thr_update(to, acto, A.thr_base, A.thr_leak, A.thr_incr);  // nvs:125-128
    const double lo = (double)nvs_leak_active(mo, A.ref_leak);
    ro = (float)__dadd_rn(__dadd_rn(A.target_off, __dmul_rn(__dsub_rn((double)r, A.target_off), lo)), (double)spo);  // nvs:133-135
  }
}

__device__ __forceinline__ void ld4(const float *p, long long q, int nv, bool vec, float v[4]) {
  if (vec) {
    const float4 a = __ldcs(reinterpret_cast<const float4 *>(p + q));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = i < nv ? p[q + i] : 0.f;
  }
}
__device__ __forceinline__ void st4(float *p, long long q, int nv, bool vec, const float v[4]) {
  if (vec) {
    __stcs(reinterpret_cast<float4 *>(p + q), make_float4(v[0], v[1], v[2], v[3]));
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i < nv) p[q + i] = v[i];
  }
}
__device__ __forceinline__ unsigned ldm4(const uint8_t *p, long long q, int nv, bool vec) {
  if (!p) return 0u;
  if (vec) return *reinterpret_cast<const unsigned *>(p + q);
  unsigned w = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i < nv) w |= (unsigned)p[q + i] << (8 * i);
  return w;
}

// Two 4-pixel groups per thread, one block-stride apart: all loads of both groups are issued before any
// arithmetic so that each thread keeps 6-10 16-byte requests in flight.
constexpr int kNvsGroups = 1;

template <int MODE>
__global__ void __launch_bounds__(256) k_nvs(NvsArgs A, int vec_ok) {
  const long long base = ((long long)blockIdx.x * (blockDim.x * kNvsGroups) + threadIdx.x) * 4;
  float f[kNvsGroups][4], r[kNvsGroups][4], t[kNvsGroups][4], ro[kNvsGroups][4], to[kNvsGroups][4];
  unsigned m[kNvsGroups], mo[kNvsGroups];
  int nv[kNvsGroups];
  bool vec[kNvsGroups];
#pragma unroll
  for (int u = 0; u < kNvsGroups; ++u) {
    const long long q = base + (long long)u * blockDim.x * 4;
    nv[u] = q < A.n ? (int)min(4ll, A.n - q) : 0;
    vec[u] = vec_ok && nv[u] == 4;
    if (nv[u] == 0) continue;
    ld4(A.frame, q, nv[u], vec[u], f[u]);
    ld4(A.ref, q, nv[u], vec[u], r[u]);
    ld4(A.thr, q, nv[u], vec[u], t[u]);
    if (MODE == 4) {
      ld4(A.ref_off, q, nv[u], vec[u], ro[u]);
      ld4(A.thr_off, q, nv[u], vec[u], to[u]);
    }
    m[u] = MODE >= 3 ? ldm4(A.leak_mask, q, nv[u], vec[u]) : 0u;
    mo[u] = MODE == 4 ? ldm4(A.leak_mask_off, q, nv[u], vec[u]) : 0u;
  }
#pragma unroll
  for (int u = 0; u < kNvsGroups; ++u) {
    if (nv[u] == 0) continue;
    const long long q = base + (long long)u * blockDim.x * 4;
    float sp[4], spo[4] = {0.f, 0.f, 0.f, 0.f};
    unsigned actw = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      unsigned act;
      nvs_px<MODE>(A, f[u][i], r[u][i], t[u][i], sp[i], act, ((m[u] >> (8 * i)) & 0xffu) ? 1u : 0u, ro[u][i], to[u][i],
                   spo[i], ((mo[u] >> (8 * i)) & 0xffu) ? 1u : 0u);
      actw |= act << (8 * i);
    }
    st4(A.spikes, q, nv[u], vec[u], sp);
    st4(A.ref, q, nv[u], vec[u], r[u]);
    if (MODE >= 2) st4(A.thr, q, nv[u], vec[u], t[u]);
    if (MODE == 4) {
      st4(A.ref_off, q, nv[u], vec[u], ro[u]);
      st4(A.thr_off, q, nv[u], vec[u], to[u]);
      if (A.spikes_off) st4(A.spikes_off, q, nv[u], vec[u], spo);
    }
    if (A.active) {
      if (vec[u]) *reinterpret_cast<unsigned *>(A.active + q) = actw;
      else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i < nv[u]) A.active[q + i] = (uint8_t)((actw >> (8 * i)) & 1u);
      }
    }
  }
}

// ---- decode_spikes.pyx -------------------------------------------------------------------------
// ref_out = int16(ref * history_weight) (dec:85/121/172): float64 product, C cast towards zero.
__global__ void __launch_bounds__(256) k_decay_i16(const int16_t *__restrict__ ref, int16_t *__restrict__ out, double hw,
                                                   long long n, int vec_ok) {
  const long long q = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (q >= n) return;
  if (vec_ok && n - q >= 8) {
    const uint4 v = __ldcs(reinterpret_cast<const uint4 *>(ref + q));
    unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int lo = (int16_t)(w[i] & 0xffffu), hi = (int16_t)(w[i] >> 16);
      const int a = (int)__dmul_rn(hw, (double)lo), b = (int)__dmul_rn(hw, (double)hi);
      w[i] = ((unsigned)a & 0xffffu) | ((unsigned)b << 16);
    }
    __stcs(reinterpret_cast<uint4 *>(out + q), make_uint4(w[0], w[1], w[2], w[3]));
  } else {
    for (long long i = q; i < min(n, q + 8); ++i) out[i] = (int16_t)(int)__dmul_rn(hw, (double)ref[i]);
  }
}

// int16 += v with two's-complement wrap, safe against concurrent updates of the neighbouring half-word.
__device__ __forceinline__ void atomic_add_i16(int16_t *p, int v) {
  unsigned short *a = reinterpret_cast<unsigned short *>(p);
  unsigned short old = *a, assumed;
  do {
    assumed = old;
    old = atomicCAS(a, assumed, (unsigned short)(assumed + (unsigned)v));
  } while (old != assumed);
}

struct KeyArgs {
  const int16_t *keys;
  const long long *stream_offsets;  // [S + 1] positions in keys
  long long n;
  int S, H, W, flag_shift, data_shift, data_mask;
};

// The reference raises at the first offending key; `first_error` keeps the smallest (index << 2 | kind).
__device__ __forceinline__ void report(int32_t *first_error, long long i, int kind) {
  if (first_error) atomicMin(first_error, (int)((i << 2) | kind));
}

// key_to_spike dec:46-67 (+ which stream the key belongs to).  false: the reference would have raised.
__device__ __forceinline__ bool decode_key(const KeyArgs &K, long long i, long long &q, int &sign, int32_t *first_error) {
  const int k16 = K.keys[i];
  if (k16 < 0) {  // int16 -> C uint16 argument: OverflowError
    report(first_error, i, DVS_DECODE_KEY_NEGATIVE);
    return false;
  }
  const unsigned key = (unsigned)k16;
  const int row = (int)((key >> K.data_shift) & (unsigned)K.data_mask), col = (int)(key & (unsigned)K.data_mask);
  sign = ((key >> K.flag_shift) & 1u) == 0u ? 1 : -1;
  if (row >= K.H || col >= K.W) {  // undefined behaviour in the reference (boundscheck off)
    report(first_error, i, DVS_DECODE_INDEX);
    return false;
  }
  int lo = 0, hi = K.S;  // last s with stream_offsets[s] <= i
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (K.stream_offsets[mid] <= i) lo = mid; else hi = mid;
  }
  q = ((long long)lo * K.H + row) * K.W + col;
  return true;
}

__global__ void __launch_bounds__(256) k_decode_rate(KeyArgs K, int16_t *__restrict__ ref, int threshold, int32_t *first_error) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K.n) return;
  long long q; int sign;
  if (!decode_key(K, i, q, sign, first_error)) return;
  atomic_add_i16(ref + q, (int)(int16_t)threshold * sign);  // dec:90
}

constexpr int kClaimFree = 0x7f7f7f7f;  // what a byte-wise memset(0x7f) leaves

// One round of the sequential time decoders: of the keys not yet applied, the first (lowest index) of
// every pixel is applied; later duplicates wait for the next round, so the per-pixel order of dec:125 /
// dec:176 is kept.  claim[] must hold kClaimFree everywhere on entry and is restored on exit.
__global__ void __launch_bounds__(256) k_decode_claim(KeyArgs K, const uint8_t *__restrict__ done, int *__restrict__ claim,
                                                      int32_t *first_error) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K.n || done[i]) return;
  long long q; int sign;
  if (!decode_key(K, i, q, sign, first_error)) return;
  atomicMin(claim + q, (int)i);
}

__device__ __forceinline__ double np_mod(double a, double b) {  // np.mod on float64
  if (b == 0.0) return nan("");
  double m = fmod(a, b);
  if (m != 0.0) { if ((b < 0.0) != (m < 0.0)) m = __dadd_rn(m, b); } else m = copysign(0.0, b);
  return m;
}

__global__ void __launch_bounds__(256) k_decode_time(KeyArgs K, int binary, uint8_t *__restrict__ done, int *__restrict__ claim,
                                                     int16_t *__restrict__ ref, int16_t *__restrict__ time_last,
                                                     int16_t *__restrict__ val_last, const double *__restrict__ log2_lut,
                                                     int num_bins, double bin_width, int time_period, unsigned time_now,
                                                     int threshold, int32_t *first_error, int *__restrict__ remaining) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K.n || done[i]) return;
  long long q; int sign;
  if (!decode_key(K, i, q, sign, nullptr)) { done[i] = 1; return; }
  if (claim[q] != (int)i) { atomicAdd(remaining, 1); return; }
  done[i] = 1;
  claim[q] = kClaimFree;
  const int thr = (int16_t)threshold;
  const double nb = (double)(int16_t)num_bins, per = (double)(int16_t)time_period;
  const int vl = val_last[q];
  int upd;
  if (vl == -1) {  // dec:128-133 / dec:178-184
    upd = binary ? (thr + 1) * sign : (thr + 1);
  } else {
    double v;
    if (!binary) {  // dec:136-137
      const double td = __dsub_rn((double)time_now,
                                  __dsub_rn((double)time_last[q], __dsub_rn(nb, __dmul_rn((double)vl, bin_width))));
      v = floor(__dsub_rn(nb, __ddiv_rn(np_mod(td, per), bin_width)));
    } else {  // dec:186-190; np.log2 of an int16 comes from the host-built table (exactly NumPy's values)
      const double lg = vl > 0 ? log2_lut[vl] : (vl == 0 ? -INFINITY : nan(""));
      const double inner = __dadd_rn(__dsub_rn(nb, __dmul_rn(lg, bin_width)), 1.0);
      const double td = __dsub_rn((double)time_now, __dsub_rn((double)time_last[q], __dmul_rn(2.0, inner)));
      const double e = floor(__dsub_rn(nb, __ddiv_rn(np_mod(td, per), bin_width)));
      v = isnan(e) ? e : (e > 1100.0 ? INFINITY : (e < -1100.0 ? 0.0 : ldexp(1.0, (int)e)));
    }
    if (isnan(v)) { report(first_error, i, DVS_DECODE_VALUE_NAN); return; }  // int(nan): ValueError
    if (isinf(v) || trunc(v) < -32768.0 || trunc(v) > 32767.0) {             // does not fit the C int16: OverflowError
      report(first_error, i, DVS_DECODE_VALUE_RANGE);
      return;
    }
    upd = thr * sign * (int)(int16_t)(int)trunc(v);
  }
  ref[q] = (int16_t)(unsigned short)((unsigned)(int)ref[q] + (unsigned)upd);
  val_last[q] = (int16_t)(unsigned short)(unsigned)upd;
  time_last[q] = (int16_t)(unsigned short)time_now;
}

// ---- frame animators gs:1128-1182, batched over streams -----------------------------------------
// mode 0: traverse_image gs:1128-1152 -- n = int16(frame_number * speed) (0 -> 1); n <= W: the last n columns of
//         the image enter from the left (NumPy slice rules, so a negative n shifts left by |n|); W < n < 2W:
//         with n' = 2W - n + 1 the first n' columns leave on the right; otherwise background.
// mode 1: fade_image gs:1156-1182 -- int16(original * alpha), alpha = f / half before half - 1,
//         (2 half - f) / half after half + 1, else 1 (float64, truncation; bg_gray is overwritten).
__global__ void __launch_bounds__(256) k_animate(int mode, const int16_t *__restrict__ images, int n_images,
                                                 const int32_t *__restrict__ img_index,
                                                 const int32_t *__restrict__ frame_number, double speed, int half_frame,
                                                 int bg, int16_t *__restrict__ out, int S, int H, int W) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long P = (long long)H * W;
  if (g >= (long long)S * P) return;
  const int s = (int)(g / P);
  const int q = (int)(g - (long long)s * P);
  const int y = q / W, x = q - y * W;
  const int img = img_index[s];
  const bool have = img >= 0 && img < n_images;
  const int16_t *src = images + (long long)(have ? img : 0) * P + (long long)y * W;
  const int fn = (int16_t)frame_number[s];
  int v = bg;
  if (mode == 0) {
    int n = (int16_t)__double2int_rz(__dmul_rn((double)fn, speed));
    if (n == 0) n = 1;
    int sx = -1;
    if (n <= W) {
      if (n > 0) { if (x < n) sx = W - n + x; }
      else if (x < W + n) sx = x - n;
    } else if (n < 2 * W) {
      const int m = 2 * W - n + 1;
      if (x >= W - m) sx = x - (W - m);
    }
    if (sx >= 0 && have) v = src[sx];
  } else {
    double alpha = 1.0;
    if (fn < half_frame - 1) alpha = __ddiv_rn((double)fn, (double)half_frame);
    else if (fn > half_frame + 1) alpha = __ddiv_rn((double)(int16_t)(half_frame * 2 - fn), (double)half_frame);
    v = have ? (int)(int16_t)__double2int_rz(__dmul_rn((double)src[x], alpha)) : 0;
  }
  out[g] = (int16_t)v;
}

cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }
int launch_status(const char *what) {
  (void)what;
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? DVS_OK : dvs::set_error((int)e, cudaGetErrorString(e));
}
bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// attention_image gs:1246-1298, the saccade search, for S cameras at once: one CTA per stream.
//   tiny_x  = cv2.resize(x, (new_w, new_w), INTER_AREA)  -> for int16 at exactly 2x: (sum of the 2x2 block + 2) >> 2
//   diff    = np.abs(tiny_prev - tiny_ref)               -> int16 arithmetic (wraps like NumPy)
//   top_n   = np.argsort(diff.reshape(-1))[-5:]          -> the five largest (value, index) pairs in ascending order;
//                                                           NumPy's default sort leaves the order of EQUAL values
//                                                           unspecified -- here ties are ordered by index (stable sort)
//   max_idx = top_n[pick]; row, col = (max_idx // new_w) * 2, (max_idx % new_w) * 2
//   cx, cy  = new_w - col, new_w - row  (reset to 0, 0 when outside [-new_w, new_w])
// Selection: five rounds of a block-wide arg-max over 64-bit keys (value + 32768) << 32 | index, each round bounded
// above by the previous winner.
__device__ __forceinline__ int att_w16(int x) { return (int)(short)x; }
__global__ void __launch_bounds__(256) k_attention_targets(const int16_t *__restrict__ prev, const int16_t *__restrict__ ref,
                                                           const int32_t *__restrict__ picks, int32_t *__restrict__ cx,
                                                           int32_t *__restrict__ cy, int Wd) {
  __shared__ unsigned long long warp_best[8];
  __shared__ unsigned long long chosen[5];
  const int s = blockIdx.x, new_w = Wd >> 1, cells = new_w * new_w;
  const int16_t *p = prev + (size_t)s * Wd * Wd, *r = ref + (size_t)s * Wd * Wd;
  auto key_of = [&](int cell) -> unsigned long long {
    const int row = (cell / new_w) * 2, col = (cell % new_w) * 2;
    const int16_t *pp = p + (size_t)row * Wd + col, *rr = r + (size_t)row * Wd + col;
    const int tp = att_w16(((int)pp[0] + pp[1] + pp[Wd] + pp[Wd + 1] + 2) >> 2);
    const int tr = att_w16(((int)rr[0] + rr[1] + rr[Wd] + rr[Wd + 1] + 2) >> 2);
    const int d = att_w16(tp - tr);
    const int a = att_w16(d < 0 ? -d : d);
    return ((unsigned long long)(unsigned)(a + 32768) << 32) | (unsigned)cell;
  };
  unsigned long long bound = ~0ull;
  for (int round = 0; round < 5; ++round) {
    unsigned long long best = 0ull;  // keys are > 0 only when a cell qualifies: (a + 32768) << 32 >= 0, cell 0 with a = -32768 is key 0
    bool have = false;
    for (int c = threadIdx.x; c < cells; c += blockDim.x) {
      const unsigned long long k = key_of(c);
      if (k < bound && (!have || k > best)) { best = k; have = true; }
    }
    // encode "nothing" as bound itself cannot win: use k + 1 so that 0 means none
    unsigned long long v = have ? best + 1 : 0ull;
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long y = __shfl_xor_sync(0xffffffffu, v, o);
      if (y > v) v = y;
    }
    if ((threadIdx.x & 31) == 0) warp_best[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long m = 0ull;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = warp_best[w] > m ? warp_best[w] : m;
      chosen[round] = m;  // 0 = fewer than 5 cells
    }
    __syncthreads();
    if (chosen[round] == 0ull) break;
    bound = chosen[round] - 1;
  }
  if (threadIdx.x == 0) {
    // top_n is ascending: top_n[4] is round 0's winner.  Fewer than five cells: NumPy's [-5:] keeps them all.
    int n_top = 0;
    while (n_top < 5 && n_top < cells) ++n_top;
    const int pick = picks[s];
    int idx_in_top = pick;                       // index into the ascending list of n_top entries
    if (idx_in_top < 0) idx_in_top += n_top;     // Python negative index
    int new_cx = 0, new_cy = 0;
    if (idx_in_top >= 0 && idx_in_top < n_top) {
      const unsigned long long k = chosen[n_top - 1 - idx_in_top] - 1;
      const int max_idx = (int)(unsigned)(k & 0xffffffffull);
      const int row = (max_idx / new_w) * 2, col = (max_idx % new_w) * 2;
      new_cx = new_w - col;
      new_cy = new_w - row;
      if (new_cx > new_w || new_cx < -new_w || new_cy > new_w || new_cy < -new_w) { new_cx = 0; new_cy = 0; }
    }
    cx[s] = new_cx;
    cy[s] = new_cy;
  }
}

}  // namespace

extern "C" {

int dvs_attention_targets_i16(const int16_t *prev, const int16_t *ref, const int32_t *picks, int32_t *cx, int32_t *cy,
                              int32_t S, int32_t W, void *stream) {
  if (!prev || !ref || !picks || !cx || !cy || S <= 0 || W < 2 || (W & 1))
    return dvs::set_error(DVS_ERR_INVALID_ARGUMENT, "attention_targets: square frames of even width, one pick per stream");
  k_attention_targets<<<(unsigned)S, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(prev, ref, picks, cx, cy, W);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return dvs::set_error((int)e, cudaGetErrorString(e));
  return DVS_OK;
}

int dvs_nvs_step_f32(const dvs_nvs_params *p, void *stream) {
  if (!p || !p->frame || !p->ref || !p->thr || !p->spikes || p->n_px < 0)
    return dvs::set_error(DVS_ERR_INVALID_ARGUMENT, "nvs_step arguments");
  if (p->mode < DVS_NVS_0 || p->mode > DVS_NVS_BI_NOISY_LEAKY_ADAPTIVE)
    return dvs::set_error(DVS_ERR_INVALID_ARGUMENT, "nvs mode");
  if (p->mode == DVS_NVS_BI_NOISY_LEAKY_ADAPTIVE && (!p->ref_off || !p->thr_off))
    return dvs::set_error(DVS_ERR_INVALID_ARGUMENT, "the two-cell variant needs ref_off and thr_off");
  if (p->n_px == 0) return DVS_OK;
  NvsArgs A;
  A.frame = p->frame; A.ref = p->ref; A.thr = p->thr; A.spikes = p->spikes; A.active = p->active;
  A.leak_mask = p->leak_mask; A.ref_off = p->ref_off; A.thr_off = p->thr_off; A.spikes_off = p->spikes_off;
  A.leak_mask_off = p->leak_mask_off; A.ref_leak = p->reference_leak; A.thr_base = p->threshold_base;
  A.thr_incr = p->threshold_mult_incr; A.thr_leak = p->threshold_leak; A.target_off = p->target_off; A.n = p->n_px;
  int vec = aligned16(A.frame) && aligned16(A.ref) && aligned16(A.thr) && aligned16(A.spikes) &&
            (!A.ref_off || aligned16(A.ref_off)) && (!A.thr_off || aligned16(A.thr_off)) &&
            (!A.spikes_off || aligned16(A.spikes_off)) && (!A.active || (reinterpret_cast<uintptr_t>(A.active) & 3u) == 0) &&
            (!A.leak_mask || (reinterpret_cast<uintptr_t>(A.leak_mask) & 3u) == 0) &&
            (!A.leak_mask_off || (reinterpret_cast<uintptr_t>(A.leak_mask_off) & 3u) == 0);
  const unsigned blocks = (unsigned)(((A.n + 3) / 4 + 256 * kNvsGroups - 1) / (256 * kNvsGroups));
  cudaStream_t st = as_stream(stream);
  switch (p->mode) {
    case DVS_NVS_0: k_nvs<0><<<blocks, 256, 0, st>>>(A, vec); break;
    case DVS_NVS_LEAKY: k_nvs<1><<<blocks, 256, 0, st>>>(A, vec); break;
    case DVS_NVS_LEAKY_ADAPTIVE: k_nvs<2><<<blocks, 256, 0, st>>>(A, vec); break;
    case DVS_NVS_NOISY_LEAKY_ADAPTIVE: k_nvs<3><<<blocks, 256, 0, st>>>(A, vec); break;
    default: k_nvs<4><<<blocks, 256, 0, st>>>(A, vec); break;
  }
  return launch_status("k_nvs");
}

int dvs_animate_i16(int32_t mode, const int16_t *images, int32_t n_images, const int32_t *img_index,
                    const int32_t *frame_number, double speed, int32_t half_frame, int32_t bg, int16_t *out, int32_t S,
                    int32_t H, int32_t W, void *stream) {
  if (!images || !img_index || !frame_number || !out || n_images <= 0 || S <= 0 || H <= 0 || W <= 0 ||
      (mode != DVS_ANIM_TRAVERSE && mode != DVS_ANIM_FADE))
    return dvs::set_error(DVS_ERR_INVALID_ARGUMENT, "animate arguments");
  const long long n = (long long)S * H * W;
  k_animate<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(mode, images, n_images, img_index, frame_number, speed,
                                                                     half_frame, bg, out, S, H, W);
  return launch_status("k_animate");
}

int dvs_decay_i16(const int16_t *ref, int16_t *ref_out, double history_weight, int64_t n_px, void *stream) {
  if (n_px < 0 || (n_px > 0 && (!ref || !ref_out))) return dvs::set_error(DVS_ERR_INVALID_ARGUMENT, "decay arguments");
  if (n_px == 0) return DVS_OK;
  const int vec = aligned16(ref) && aligned16(ref_out);
  k_decay_i16<<<(unsigned)(((n_px + 7) / 8 + 255) / 256), 256, 0, as_stream(stream)>>>(ref, ref_out, history_weight, n_px, vec);
  return launch_status("k_decay_i16");
}

static int check_keys(const dvs_key_list *k) {
  if (!k || k->n < 0 || k->S <= 0 || k->H <= 0 || k->W <= 0 || (k->n > 0 && (!k->keys || !k->stream_offsets)))
    return dvs::set_error(DVS_ERR_INVALID_ARGUMENT, "key list");
  if (k->n > 0x1fffffffll) return dvs::set_error(DVS_ERR_UNSUPPORTED, "more than 2^29 keys in one call");
  return DVS_OK;
}
static KeyArgs make_keys(const dvs_key_list *k) {
  KeyArgs K;
  K.keys = k->keys; K.stream_offsets = reinterpret_cast<const long long *>(k->stream_offsets); K.n = k->n;
  K.S = k->S; K.H = k->H; K.W = k->W; K.flag_shift = k->flag_shift; K.data_shift = k->data_shift; K.data_mask = k->data_mask;
  return K;
}

int dvs_decode_rate_i16(const dvs_key_list *keys, int16_t *ref, int32_t threshold, int32_t *first_error, void *stream) {
  int rc = check_keys(keys);
  if (rc) return rc;
  if (!ref) return dvs::set_error(DVS_ERR_INVALID_ARGUMENT, "decode_rate: null reference");
  if (keys->n == 0) return DVS_OK;
  k_decode_rate<<<(unsigned)((keys->n + 255) / 256), 256, 0, as_stream(stream)>>>(make_keys(keys), ref, threshold, first_error);
  return launch_status("k_decode_rate");
}

int dvs_decode_time_round_i16(const dvs_key_list *keys, int32_t binary, int16_t *ref, int16_t *time_last, int16_t *val_last,
                              const double *log2_lut, int32_t num_bins, double bin_width, int32_t time_period,
                              uint32_t time_now, int32_t threshold, uint8_t *done, int32_t *claim, int32_t *remaining,
                              int32_t *first_error, void *stream) {
  int rc = check_keys(keys);
  if (rc) return rc;
  if (!ref || !time_last || !val_last || !done || !claim || !remaining || (binary && !log2_lut))
    return dvs::set_error(DVS_ERR_INVALID_ARGUMENT, "decode_time: null pointer");
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(remaining, 0, sizeof(int32_t), st);
  if (e != cudaSuccess) return dvs::set_error((int)e, cudaGetErrorString(e));
  if (keys->n == 0) return DVS_OK;
  const unsigned blocks = (unsigned)((keys->n + 255) / 256);
  const KeyArgs K = make_keys(keys);
  k_decode_claim<<<blocks, 256, 0, st>>>(K, done, claim, first_error);
  k_decode_time<<<blocks, 256, 0, st>>>(K, binary, done, claim, ref, time_last, val_last, log2_lut, num_bins, bin_width,
                                        time_period, time_now, threshold, first_error, remaining);
  return launch_status("k_decode_time");
}

}  // extern "C"
