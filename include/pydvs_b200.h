/*
 * pydvs_b200.h -- C-ABI of libpydvs_b200.so: hand-written sm_100a CUDA kernels for pyDVS's
 * spike-generation hot path.
 *
 * The reference has no FFI for this path: its interface is the set of module-level Python callables
 * of `pydvs.generate_spikes` / `pydvs.generate_spikes_uncoded` (Cython).  Each entry point below names
 * the reference function(s) it replaces as file:line under /root/reference/pydvs ("gs" =
 * generate_spikes.pyx, "gsu" = generate_spikes_uncoded.pyx, "cpp" = cpp/dvs_op.cpp).  The Python
 * modules pydvs_b200.generate_spikes[_uncoded] bind these with ctypes and keep the reference's
 * function names and positional signatures; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host; planes are row-major
 *     [S][H][W] (S independent camera streams; the reference's 2-D arrays are S == 1);
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); nothing here
 *     allocates, synchronises or touches the host: launches are asynchronous on `stream`;
 *   - return value: 0 on success, otherwise a DVS_ERR_* / cudaError_t code; the message is
 *     available from dvs_last_error_string() (thread-local);
 *   - no globals: calls on different streams / threads are independent;
 *   - integer arithmetic reproduces NumPy int16 semantics exactly (two's-complement wrap of every
 *     int16 intermediate, floor division, x // 0 == 0) -- SURVEY.md Appendix A;
 *   - data-dependent faults the reference reports as exceptions are reported through a device
 *     status word (`status`, may be NULL): bits DVS_STATUS_*.
 */
#ifndef PYDVS_B200_H
#define PYDVS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DVS_ABI_VERSION 4

/* error codes (>= 1000 so they cannot collide with cudaError_t) */
#define DVS_OK 0
#define DVS_ERR_INVALID_ARGUMENT 1001 /* bad geometry / mode / NULL pointer                   */
#define DVS_ERR_UNSUPPORTED 1002      /* outside the documented limits (see dvs_limits)       */
#define DVS_ERR_NO_DEVICE 1003

/* status word bits (device int32, OR-ed by kernels) */
#define DVS_STATUS_LUT_INDEX 1u       /* log2-table index out of range (reference: IndexError)   */
#define DVS_STATUS_SLOT_RANGE 2u      /* time slot outside [0, n_slots): dropped (Appendix B.7)  */
#define DVS_STATUS_EVENT_OVERFLOW 4u  /* event buffer too small: nothing was written by expand   */
#define DVS_STATUS_INDEX 8u           /* injected noise index out of range (reference: IndexError) */

/* polarity modes, gs:36-39 */
enum { DVS_POL_UP = 0, DVS_POL_DOWN = 1, DVS_POL_MERGED = 2, DVS_POL_RECTIFIED = 3 };

/* reference-update rules */
enum {
  DVS_UPD_RATE = 0,         /* update_reference_rate gs:225-251 == update_reference_time_thresh gs:301-334 */
  DVS_UPD_RATE_ADPT = 1,    /* update_reference_rate_adpt gs:256-297 (uncoded flag: gsu:251-295)           */
  DVS_UPD_TIME_BIN_RAW = 2, /* update_reference_time_binary_raw gs:339-378                                */
  DVS_UPD_TIME_BIN_THR = 3, /* update_reference_time_binary_thresh gs:382-425                             */
  DVS_UPD_NONE = 4          /* leave the reference untouched                                              */
};

/* AER slot encoders */
enum {
  DVS_ENC_RATE = 0,         /* make_spike_lists_rate gs:783-852 (gsu:710-768)          */
  DVS_ENC_TIME = 1,         /* make_spike_lists_time gs:857-928 (gsu:773-835)          */
  DVS_ENC_TIME_BIN = 2,     /* make_spike_lists_time_bin gs:933-1010 (gsu:840-913)     */
  DVS_ENC_TIME_BIN_THR = 3, /* make_spike_lists_time_bin_thr gs:1015-1099 (gsu:918-998) */
  DVS_ENC_NONE = 4          /* records only (split_spikes), no slot counting           */
};

enum { DVS_FRAME_I16 = 0, DVS_FRAME_U8 = 1 };
enum { DVS_STATE_I16 = 0, DVS_STATE_U8 = 1 };  /* dvs_step_params.state_dtype */

/* One spiking pixel, the element of split_spikes' (rows; cols; vals) lists, gs:716-719. */
typedef struct dvs_record {
  uint16_t row, col;
  int16_t val;  /* abs_diff at the pixel */
  int16_t pad;
} dvs_record;

/* One AER event: (x, y, polarity, t).  p is +1 for entries of the reference's "pos" list and -1
 * for its "neg" list (i.e. the is_pos_spike flag of gs:744-778, including the DOWN/RECTIFIED quirk
 * that those modes mark everything positive); t is the time-slot index inside the frame interval. */
typedef struct dvs_event {
  uint16_t x, y;
  uint16_t t;
  int16_t p;
} dvs_event;

typedef struct dvs_limits {
  int32_t max_tile_px;   /* pixels per tile                     */
  int32_t max_slots;     /* n_slots                             */
  int32_t max_width;     /* W, H must fit uint16                */
  int32_t sm_count;      /* of the current device (0 if none)   */
} dvs_limits;

/* Slot-encoder parameters (the per-pixel arithmetic inside make_spike_lists_*, Appendix A.8). */
typedef struct dvs_enc_params {
  int32_t mode;          /* DVS_ENC_*                                                        */
  int32_t uncoded;       /* 0: generate_spikes.pyx rules, 1: generate_spikes_uncoded.pyx     */
  int32_t threshold;     /* rate: `threshold`; time / time_bin_thr: `min_threshold`          */
  int32_t n_slots;       /* rate: max_time_ms; others: num_bins                              */
  int32_t u16_vals;      /* record vals are uint16 (uncoded rate/time signatures gsu:710,773) */
  int32_t lut_len;
  const uint8_t *lut;    /* log2 table row (time_bin modes), device                          */
} dvs_enc_params;

/* Tile geometry shared by the fused step, the scan and the expand kernels.  A tile is `tile_rows`
 * full-width rows of one stream; tiles_per_stream = ceil(H / tile_rows); tile t of stream s has
 * global index s * tiles_per_stream + t and owns records[global * tile_rows * W ...). */
typedef struct dvs_tiling {
  int32_t S, H, W;
  int32_t tile_rows;
} dvs_tiling;

/* Counter layout (all uint32, component-major so that scans run over contiguous memory):
 *   tile_counts[s][c][t], c in [0, C), C = 2 + 2 * n_slots:
 *     c = 0: #pos records of the tile, c = 1: #neg records,
 *     c = 2 + 2 * slot + pol (pol 0 = pos list, 1 = neg list): #events of that segment.
 *   tile_excl has the same shape and holds the exclusive prefix over t.
 *   stream_totals[s][c] holds the sum over t.
 *   seg_offsets[(s * n_slots + slot) * 2 + pol] (int64, S*n_slots*2 + 1 entries) is the CSR row
 *   pointer of that segment in the global event buffer: events are ordered stream-major, then
 *   slot-major, then all pos entries row-major, then all neg entries row-major (Appendix A.7). */

typedef struct dvs_step_params {
  dvs_tiling tiling;
  int32_t frame_dtype;      /* DVS_FRAME_*: dtype of `curr`                                   */
  int32_t adaptive;         /* 1: per-pixel threshold plane `thr` (thresholded_difference_adpt) */
  int32_t update_mode;      /* DVS_UPD_*                                                      */
  int32_t uncoded;          /* rate_adpt rule + encoder rules of the uncoded twin             */
  int32_t threshold;        /* scalar threshold (static mode)                                 */
  int32_t min_thr, max_thr, thr_down, thr_up; /* adaptive-threshold parameters gs:260-263     */
  int32_t max_time_ms;
  int32_t inh_k;            /* local_inhibition area width; <= 1: off                         */
  int32_t polarity;         /* DVS_POL_*                                                      */
  int32_t force_generic;    /* testing / profiling: 1 = never take the specialised fast kernel, 3 = its persistent
                             * staged form (bulk async copies + mbarrier) whenever the tile shape allows it   */
  int32_t drop_silent;      /* 1: records that emit no event (coded rate, k == 0) may be left out of the
                             * tile lists: identical events, but the lists no longer are split_spikes' output */
  int32_t record_stride;    /* records reserved per tile region; 0 = tile_rows * W.  With k x k inhibition a tile
                             * holds at most (tile_rows / k) * (W / k) records, so callers may pass that bound
                             * (the same number is the `tile_px` argument of dvs_aer_expand / dvs_gather_split) */
  int32_t state_dtype;      /* DVS_STATE_*: layout of the resident planes `ref` / `thr`.  DVS_STATE_U8 keeps one byte per
                             * pixel: every update_reference_* rule ends in clip(.., 0, 255) (gs:251, 297, 378, 425) and
                             * the coded threshold rule in clip(.., min, max) (gs:297), so for planes that start in
                             * [0, 255] (and 0 <= min_thr <= max_thr <= 255) the values are the reference's own, and the
                             * pass moves half the state bytes.  Identical events, records and state values.          */
  double history_weight;
  dvs_enc_params enc;
  const void *curr;         /* [S][H][W] frames, int16 or uint8                               */
  void *ref;                /* [S][H][W] reference state (int16_t or uint8_t per state_dtype), updated in place */
  void *thr;                /* [S][H][W] threshold state (adaptive; same layout), updated in place            */
  int16_t *diff_out;        /* optional planes (NULL = skip): what thresholded_difference and */
  int16_t *abs_diff_out;    /*   local_inhibition would have returned for this frame          */
  int16_t *spikes_out;
  dvs_record *records;      /* S * tiles_per_stream * record_stride records (tile-local lists)  */
  uint32_t *tile_counts;    /* [S][C][tiles_per_stream]                                       */
  int32_t *status;          /* device status word or NULL                                     */
  int32_t *redo;            /* optional scratch, 2 + S * tiles_per_stream int32, the first two  */
                            /*   words ZERO before the first step (every step leaves them zero):*/
                            /*   enables the specialised fast kernel (tiles it declines are     */
                            /*   listed here and replayed by the generic body)                  */
} dvs_step_params;

/* ---- library ---------------------------------------------------------------------------------- */
int dvs_abi_version(void);
/* Which tile kernel the calling thread's last dvs_step_fused launched (tests and benchmarks report it). */
enum { DVS_PATH_GENERIC = 0, DVS_PATH_FAST = 1, DVS_PATH_FAST_4X4 = 2 };
int dvs_last_step_path(void);
const char *dvs_last_error_string(void);
int dvs_get_limits(dvs_limits *out_host);
/* Host helper: tile_rows for a geometry (multiple of max(inh_k,1), tile_rows*W <= max_tile_px). */
int dvs_choose_tile_rows(int32_t S, int32_t H, int32_t W, int32_t inh_k);
/* Host helper for callers whose frames live in HOST memory: the reference hands 8-bit gray values over as int16 planes
 * (stand_alone.py:160-166 `.astype(int16)`), and a pinned host->device copy is what bounds such a step, so the feed may
 * narrow a frame to bytes on `threads` host threads (0 = all), copy half the bytes and run the step with
 * DVS_FRAME_U8.  Exact: *out_of_range is 1 when any value lies outside [0, 255] (copy the int16 plane then).
 * `dst` 32-byte aligned enables the AVX2 / streaming-store path.  No CUDA call. */
int dvs_host_narrow_i16_u8(const int16_t *src, uint8_t *dst, int64_t n, int32_t threads, int32_t *out_of_range);

/* ---- unfused, function-for-function kernels (drop-in module API) ------------------------------ */

/* thresholded_difference gs:49-78 (thr_plane == NULL) / thresholded_difference_adpt gs:83-113. */
int dvs_thresholded_difference_i16(const int16_t *curr, const int16_t *ref, const int16_t *thr_plane,
                                   int32_t thr_scalar, int16_t *diff, int16_t *abs_diff,
                                   int16_t *spikes, int64_t n_px, void *stream);

/* local_inhibition gs:177-221 (argmax gs:118-132, transform_coords gs:136-151): in place on spikes.
 * Requires W % k == 0 and H % k == 0 (the reference raises IndexError otherwise). */
int dvs_local_inhibition_i16(int16_t *spikes, const int16_t *abs_diff, int32_t S, int32_t H, int32_t W,
                             int32_t k, void *stream);

/* update_reference_* gs:225-425.  mode = DVS_UPD_*; for DVS_UPD_RATE_ADPT `thr` is updated in place
 * to the unclipped value (gs:292-295) and thr_out receives what the reference returns. */
int dvs_update_reference_i16(int32_t mode, int32_t uncoded, const int16_t *abs_diff,
                             const int16_t *spikes, const int16_t *ref_in, int16_t *ref_out,
                             int16_t *thr, int16_t *thr_out, int32_t threshold, int32_t min_thr,
                             int32_t max_thr, int32_t thr_down, int32_t thr_up, int32_t max_time_ms,
                             double history_weight, const uint8_t *lut, int32_t lut_len, int64_t n_px,
                             int32_t *status, void *stream);

/* update_reference_time_binary_raw_noise gs:428-449 (thresh_variant 0) and
 * update_reference_time_thresh_noise gs:453-473 (1) with the random flat indices supplied by the
 * caller (int16-cast exactly as gs:442 / gs:465 produce them).  `flags` is a P-byte scratch plane. */
int dvs_update_reference_noise_i16(int32_t thresh_variant, const int16_t *abs_diff,
                                   const int16_t *spikes, const int16_t *ref_in, int16_t *ref_out,
                                   int32_t W, int32_t H, int32_t threshold, int32_t max_time_ms,
                                   double history_weight, const uint8_t *lut, int32_t lut_len,
                                   const int16_t *sample, int64_t n_sample, uint8_t *flags,
                                   int32_t *status, void *stream);

/* DVSOperator::operator() cpp:4-33: float32, in place on diff/ref/thr, no FMA contraction. */
int dvs_step_f32(const float *src, float *diff, float *ref, float *thr, float relax, float up,
                 float down, int64_t n_px, void *stream);

/* ---- ordered compaction pipeline --------------------------------------------------------------- */

/* K1, fused per-pixel pass: thresholded_difference[_adpt] -> local_inhibition -> update_reference_*
 * -> split_spikes (tile-local record lists) -> per-tile slot counts of make_spike_lists_*.
 * One pass over curr/ref(/thr); state updated in place. */
int dvs_step_fused(const dvs_step_params *p, void *stream);

/* K1 variant for callers that already hold spikes/abs_diff planes (split_spikes gs:651-719). */
int dvs_split_tiles(const dvs_tiling *tiling, const int16_t *spikes, const int16_t *abs_diff,
                    int32_t polarity, const dvs_enc_params *enc, dvs_record *records,
                    uint32_t *tile_counts, int32_t *global_max, int32_t *status, void *stream);

/* K1 variant for callers that hold (rows; cols; vals) lists (the make_spike_lists_* signatures):
 * chunks of `chunk` entries become tiles of a single stream: first ceil(n_pos/chunk) all-pos tiles,
 * then ceil(n_neg/chunk) all-neg tiles.  pos/neg are [3][*] with the given row strides. */
int dvs_lists_to_tiles(const int16_t *pos, int64_t n_pos, int64_t pos_stride, const int16_t *neg,
                       int64_t n_neg, int64_t neg_stride, int32_t chunk, const dvs_enc_params *enc,
                       dvs_record *records, uint32_t *tile_counts, int32_t *status, void *stream);

/* K2: exclusive scans of tile_counts over tiles (per stream, per component), stream totals, and the
 * global CSR row pointers seg_offsets (last entry = total number of events). */
int dvs_scan_tiles(int32_t S, int32_t tiles_per_stream, int32_t n_slots, const uint32_t *tile_counts,
                   uint32_t *tile_excl, uint32_t *stream_totals, int64_t *seg_offsets, void *stream);
/* The same in ONE launch: `ticket` is a device word that is zero before the first call (and left zero); the last
 * CTA of the tile scan builds seg_offsets.  Falls back to the separate launches when there are more than 8192
 * (stream, slot, polarity) segments.  Replaces the list concatenation order of make_spike_lists_* (gs:783-1099)
 * exactly like dvs_scan_tiles. */
int dvs_scan_tiles_ticket(int32_t S, int32_t tiles_per_stream, int32_t n_slots, const uint32_t *tile_counts,
                          uint32_t *tile_excl, uint32_t *stream_totals, int64_t *seg_offsets, uint32_t *ticket,
                          void *stream);

/* K3: expand tile-local records into AER events at their final CSR positions (make_spike_lists_*).
 * tile_px = records per tile region (tile_rows * W, or `chunk`).  If the total exceeds event_cap
 * nothing is written and DVS_STATUS_EVENT_OVERFLOW is raised; the call can be repeated. */
int dvs_aer_expand(int32_t S, int32_t tiles_per_stream, int32_t tile_px, const dvs_enc_params *enc,
                   const dvs_record *records, const uint32_t *tile_counts, const uint32_t *tile_excl,
                   const int64_t *seg_offsets, dvs_event *events, int64_t event_cap, int32_t *status,
                   void *stream);

/* Gather tile-local record lists into the contiguous (rows; cols; vals) arrays split_spikes returns.
 * pos_out / neg_out are [3][pos_stride] / [3][neg_stride] int16 for ONE stream `s`. */
int dvs_gather_split(int32_t s, int32_t tiles_per_stream, int32_t tile_px, int32_t n_slots,
                     const dvs_record *records, const uint32_t *tile_counts, const uint32_t *tile_excl,
                     int16_t *pos_out, int64_t pos_stride, int16_t *neg_out, int64_t neg_stride,
                     void *stream);

/* Frame rows of W pixels -> rows of Wp pixels (Wp = W rounded up to a multiple of 8, pad columns zero): lets a caller
 * keep its planes with 16-byte aligned rows, so that geometries such as 346 x 260 (DAVIS346) with local_inhibition
 * (gs:177-221) run the register-only tile kernels; pad pixels (frame 0, reference 0) never exceed a threshold >= 0. */
int dvs_pad_rows(const void *src, void *dst, int64_t rows, int32_t W, int32_t Wp, int32_t elem_bytes, void *stream);

/* The reference's 16-bit AER key for each event: coded spike_to_key gs:744-778 (uncoded = 0) or
 * gsu:676-705 (uncoded = 1), including the truncation to uint16 and the int16 return type. */
int dvs_keys_from_events(const dvs_event *events, int64_t n, int32_t uncoded, int32_t flag_shift,
                         int32_t data_shift, int32_t data_mask, int16_t *keys, void *stream);
/* The same with the number of events read on the device (`n_dev`: e.g. the last CSR row pointer of the step, clamped to
 * `cap`): a sink can queue the key pass and its copies behind a step without a host round trip for the count. */
int dvs_keys_from_events_devn(const dvs_event *events, const int64_t *n_dev, int64_t cap, int32_t uncoded,
                              int32_t flag_shift, int32_t data_shift, int32_t data_mask, int16_t *keys, void *stream);

/* ---- frame sources (the step before the hot path) ----------------- */

/* Batched move_image gs:1184-1215 (what usaccade_image gs:1218-1242 / attention_image apply) followed by the
 * optional mask_image gs:1124-1127: stream s shows images[img_index[s]] shifted right by dx[s] and up by
 * dy[s]; vacated pixels = bg; with a mask (float64 [H][W]) out = int16(trunc(moved * mask)).
 * images: [n_images][H][W] int16, out: [S][H][W] int16. */
int dvs_move_mask_i16(const int16_t *images, int32_t n_images, const int32_t *img_index, const int32_t *dx,
                      const int32_t *dy, int32_t bg, const double *mask, int16_t *out, int32_t S, int32_t H,
                      int32_t W, void *stream);

/* The saccade search of attention_image gs:1246-1298 for S square cameras of width W (even): tiny = 2x2 area
 * means (cv2.resize INTER_AREA on int16: (sum + 2) >> 2) of the previous frame and of the DVS reference plane,
 * diff = |tiny_prev - tiny_ref| in int16, top_n = the five largest cells in ascending (value, index) order,
 * max_idx = top_n[picks[s]] (the reference draws picks from np.random.randint(5)); cx[s], cy[s] receive the new
 * image centre (new_w - col, new_w - row; 0, 0 when outside [-new_w, new_w]).  Equal diff values: the reference's
 * np.argsort leaves their order unspecified; here they are ordered by cell index. */
int dvs_attention_targets_i16(const int16_t *prev, const int16_t *ref, const int32_t *picks, int32_t *cx,
                              int32_t *cy, int32_t S, int32_t W, void *stream);

/* ---- rows either side of the path ---------------------------------------------- */

/* Float32 NVS variants, pydvs/numba_nvs.py ("nvs").  One fused pass, in place on ref/thr (and the OFF
 * cell), `spikes` overwritten, exactly the precision numba uses: float32 for thresholded_difference
 * (nvs:6-20, NumPy floor_divide) and for array-only expressions, float64 with one rounding wherever a
 * Python-float scalar enters (nvs:40, 52-60, 133-135).  The unseeded draws `uniform(0,1) <= p` of the
 * noisy variants (nvs:85, 119, 133) are supplied by the caller as byte masks (non-zero = leak). */
enum {
  DVS_NVS_0 = 0,                      /* nvs_0 nvs:23-29                          */
  DVS_NVS_LEAKY = 1,                  /* nvs_leaky nvs:32-40                      */
  DVS_NVS_LEAKY_ADAPTIVE = 2,         /* nvs_leaky_adaptive nvs:43-60             */
  DVS_NVS_NOISY_LEAKY_ADAPTIVE = 3,   /* nvs_noisy_leaky_adaptive nvs:64-90       */
  DVS_NVS_BI_NOISY_LEAKY_ADAPTIVE = 4 /* nvs_bi_noisy_leaky_adaptive nvs:93-147   */
};
typedef struct dvs_nvs_params {
  int32_t mode;              /* DVS_NVS_*                                                        */
  int32_t reserved0;
  int64_t n_px;              /* elements in every plane (any shape, any number of streams)       */
  const float *frame;
  float *ref, *thr, *spikes;
  uint8_t *active;           /* optional: what thresholded_difference returns (0/1 per pixel)   */
  const uint8_t *leak_mask;  /* modes 3, 4; NULL = no pixel leaks                                */
  float *ref_off, *thr_off;  /* mode 4: the OFF cell                                             */
  float *spikes_off;         /* mode 4, optional (a local in the reference)                      */
  const uint8_t *leak_mask_off;
  double reference_leak, threshold_base, threshold_mult_incr, threshold_leak, target_off;
} dvs_nvs_params;
int dvs_nvs_step_f32(const dvs_nvs_params *p, void *stream);

/* Batched frame animators (the step before the path): stream s shows images[img_index[s]] at its own
 * frame_number[s].  DVS_ANIM_TRAVERSE = traverse_image gs:1128-1152 (`speed` pixels per frame, vacated
 * pixels = bg); DVS_ANIM_FADE = fade_image gs:1156-1182 (`half_frame`; float64 alpha, truncation). */
enum { DVS_ANIM_TRAVERSE = 0, DVS_ANIM_FADE = 1 };
int dvs_animate_i16(int32_t mode, const int16_t *images, int32_t n_images, const int32_t *img_index,
                    const int32_t *frame_number, double speed, int32_t half_frame, int32_t bg, int16_t *out,
                    int32_t S, int32_t H, int32_t W, void *stream);

/* Receiver, pydvs/decode_spikes.pyx ("dec").  ref_out = int16(ref * history_weight), the first
 * statement of every update_ref_spike_list_* (dec:85, 121, 172). */
int dvs_decay_i16(const int16_t *ref, int16_t *ref_out, double history_weight, int64_t n_px, void *stream);

/* A batch of 16-bit keys (key_to_spike dec:46-67: row = key >> data_shift & mask, col = key & mask,
 * sign = +1 when the flag bit is 0) for S streams of H x W pixels: the keys of stream s are
 * keys[stream_offsets[s] .. stream_offsets[s+1]).  The reference call is S == 1. */
typedef struct dvs_key_list {
  const int16_t *keys;
  const int64_t *stream_offsets; /* device, S + 1 entries */
  int64_t n;
  int32_t S, H, W;
  int32_t flag_shift, data_shift, data_mask;
} dvs_key_list;

/* What the reference raises at an offending key; *first_error (device int32, initialise to
 * DVS_DECODE_NO_ERROR) receives the smallest (key index << 2 | kind). */
enum {
  DVS_DECODE_KEY_NEGATIVE = 0, /* int16 key < 0 passed to a C uint16: OverflowError                 */
  DVS_DECODE_INDEX = 1,        /* (row, col) outside the plane: undefined in the reference; skipped */
  DVS_DECODE_VALUE_NAN = 2,    /* val_now is nan: ValueError                                        */
  DVS_DECODE_VALUE_RANGE = 3   /* val_now does not fit int16: OverflowError                         */
};
#define DVS_DECODE_NO_ERROR 0x7f7f7f7f

/* update_ref_spike_list_rate dec:71-94 after the decay: ref[row, col] += threshold * sign per key
 * (int16 wrap; duplicates accumulate). */
int dvs_decode_rate_i16(const dvs_key_list *keys, int16_t *ref, int32_t threshold, int32_t *first_error,
                        void *stream);

/* update_ref_spike_list_time dec:98-145 (binary == 0) / _time_bin dec:149-197 (binary == 1) after the
 * decay, in place on ref / time_last / val_last.  The reference walks the keys sequentially; one call
 * applies, for every pixel, the first key not yet applied (`done`, n bytes, zero before the first
 * round) and counts the keys still waiting in *remaining: call again until it reads 0 (one round when
 * no pixel repeats, as in every list the encoders produce).  `claim` is a S*H*W int32 scratch plane
 * holding DVS_DECODE_NO_ERROR everywhere (memset 0x7f) on entry and on exit.  `log2_lut` (binary
 * only): 32768 float64, log2_lut[v] = np.log2(v) for v >= 1, built by the host so that the values are
 * NumPy's own. */
int dvs_decode_time_round_i16(const dvs_key_list *keys, int32_t binary, int16_t *ref, int16_t *time_last,
                              int16_t *val_last, const double *log2_lut, int32_t num_bins, double bin_width,
                              int32_t time_period, uint32_t time_now, int32_t threshold, uint8_t *done,
                              int32_t *claim, int32_t *remaining, int32_t *first_error, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PYDVS_B200_H */
