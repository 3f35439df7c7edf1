/*
 * bpm_b200.h -- C ABI of libbpm_b200.so: the data-parallel numeric front end of
 * pixeru/bpm_analysis, written for NVIDIA B200 (sm_100a).
 *
 * The reference has no FFI of its own: its hot path is Python calling
 * numpy / scipy.signal / pandas (bpm_analysis.py:3-6).  Each entry point below
 * replaces the library calls made at the cited reference lines; the Python
 * mirror of the reference's functions (bpm_analysis_b200/frontend.py) binds
 * them with ctypes.  INTEGRATION.md shows the binding a maintainer adds.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - the library never allocates: callers pass a workspace sized by the
 *     matching *_workspace_bytes() query; nothing is kept between calls;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*) and the call
 *     returns without synchronising; results are valid after the stream drains;
 *   - return value: 0 on success, a negative BPM_ERR_* code otherwise; never
 *     throws, never exits;
 *   - a call processes a ragged BATCH of recordings described by an array of
 *     BpmItem; every per-sample buffer holds the recordings back to back at
 *     `m_off`, every per-recording scalar / count is an array of n_items;
 *   - variable-length results (trough / peak lists) are written as int64 at
 *     `m_off` of the recording (capacity m) with the length in a per-recording
 *     int64 count.
 */
#ifndef BPM_B200_H_
#define BPM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BPM_ABI_VERSION 2

enum {
  BPM_OK = 0,
  BPM_ERR_ARG = -1,        /* bad argument (null pointer, negative size, unknown dtype) */
  BPM_ERR_WORKSPACE = -2,  /* workspace too small */
  BPM_ERR_CUDA = -3,       /* a CUDA runtime call or launch failed */
  BPM_ERR_TOO_SHORT = -4   /* a recording is not longer than the 15-sample filter padding */
};

/* sample formats scipy.io.wavfile.read can hand to preprocess_audio (bpm_analysis.py:1014) */
enum { BPM_PCM_I16 = 0, BPM_PCM_I32 = 1, BPM_PCM_U8 = 2, BPM_PCM_F32 = 3, BPM_PCM_F64 = 4 };

typedef struct {
  int64_t in_off;  /* first frame of this recording in the PCM buffer (frames, not bytes) */
  int64_t n_in;    /* frames in this recording */
  int64_t m_off;   /* first element of this recording in every envelope-rate buffer */
  int64_t m;       /* envelope-rate samples: ceil(ceil(n_in / stride) / block) */
} BpmItem;

/* Layout (float64 words) of the filter design image built by
 * bpm_analysis_b200/design.py::BlockFilterDesign.packed():
 *   [0] block  [1] lookback_tiles  [2] D  [3] reserved
 *   [4..16)  sos[2][6]      [16..20) zi[4]      [20..24) C[4]
 *   [24..40) Ad[4][4]       [40..56) P[4][4]    [56..312) pow_chunk[16][4][4]
 *   [312 .. 312+4*block)               wf[block][4]
 *   [312+4*block .. 312+4*(2*block+1))  q[block+1][4]
 *   [312+4*(2*block+1) .. +8*(block+1)) wq8[block+1][8] = (wf[l] | q[l]), wf[block] = 0
 *   [312+16*block+12 .. +512)           pow_lane[32][4][4] = Ad^(8 l), l = 0..31
 * total words = 312 + 16*block + 12 + 512                                         */
#define BPM_DESIGN_HEADER_WORDS 312
#define BPM_DESIGN_LANE_WORDS 512

int bpm_abi_version(void);
const char* bpm_error_string(int code);

/* ---- K0+K1+K2(+K2b): preprocess_audio, bpm_analysis.py:1015-1054 -----------------
 * pcm -> zero-phase Butterworth band-pass (scipy filtfilt semantics: odd padding 15,
 * steady-state zi) evaluated at the kept samples -> |.| -> centred rolling mean.
 *   stride/block: (ds, 1) = the reference's decimate-then-filter order;
 *                 (1, ds) = filter at the original rate, keep every ds-th output.
 *   design / design_host: the same packed design image (layout above) in device and in HOST
 *   memory; the (stride = ds, block = 1) order passes its coefficients to the kernels as launch
 *   parameters (constant bank) and reads only design_host, which may then be the only one given.
 *   filtered, envelope: out, float64[total_m];  absmax: out, float64[n_items] = max|filtered|.
 *   filtered may be NULL when block == 1 and env_window <= 65 (the envelope is then formed inside
 *   the backward pass and the band-passed signal never reaches HBM). */
size_t bpm_frontend_workspace_bytes(int64_t total_m, int n_items);
int bpm_frontend(const void* pcm, int pcm_dtype, int channels,
                 const BpmItem* items, const BpmItem* items_host, int n_items,
                 int64_t stride, const double* design, const double* design_host, int64_t design_words,
                 int env_window, double* filtered, double* envelope, double* absmax,
                 void* workspace, size_t workspace_bytes, void* stream);

/* K0 alone: np.mean(axis=1) + audio_data[::stride] as float64, bpm_analysis.py:1016, :1033.
 * frames_out: float64, recording i at m_off, length m = ceil(n_in / stride).  `pcm` may be
 * mapped pinned HOST memory (zero-copy ingest).  Feeding frames_out to bpm_frontend /
 * bpm_stage_a as BPM_PCM_F64 with stride 1 gives bit-identical results to passing the PCM
 * itself with this stride; a host pipeline uses the split to overlap the PCIe-bound ingest
 * of the next recording with the compute of the current one. */
int bpm_gather_frames(const void* pcm, int pcm_dtype, int channels, const BpmItem* items,
                      const BpmItem* items_host, int n_items, int64_t stride, double* frames_out,
                      void* stream);

/* K0 by the COPY ENGINE: audio_data[::stride] (bpm_analysis.py:1033) as one strided 2-D copy per
 * recording (cudaMemcpy2DAsync: rows = kept frames, width = one frame, source pitch = stride
 * frames), in the PCM's own dtype with all channels of each kept frame.  `pcm` is host memory
 * (pinned for an asynchronous copy) or device memory; frames_out (device) holds recording i at
 * frame m_off, m = ceil(n_in / stride) frames of `channels` samples.  Feeding frames_out to
 * bpm_frontend / bpm_stage_a with the SAME pcm_dtype / channels and stride 1 is bit-identical to
 * passing the PCM itself with this stride.  No kernel runs: the SMs stay free for the compute of
 * the previous recording, and only the kept frames cross PCIe (measured on B200 for C2, 1.09 M
 * frames of 2 bytes every 318 bytes: 1.52 ms, against 1.82 ms for bpm_gather_frames reading mapped
 * pinned memory and 6.24 ms for copying the whole recording).  items_host only. */
int bpm_copy_frames(const void* pcm, int pcm_dtype, int channels, const BpmItem* items_host, int n_items,
                    int64_t stride, void* frames_out, void* stream);

/* K2b: np.int16(y / max|y| * 32767), bpm_analysis.py:1049 (truncating cast). */
int bpm_debug_wav(const double* filtered, const double* absmax, const BpmItem* items,
                  const BpmItem* items_host, int n_items, int16_t* out, void* stream);

/* ---- K3: np.quantile(x, q) (method 'linear'), bpm_analysis.py:225,1067,1075,1114 ---
 * out[i] = quantile of recording i.  Exact order statistics (radix select on the
 * float64 bit pattern) and numpy's two-branch lerp. */
size_t bpm_quantile_workspace_bytes(int n_items);
int bpm_quantile(const double* x, const BpmItem* items, const BpmItem* items_host, int n_items,
                 double q, double* out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- K4: scipy.signal.find_peaks(sign*x, height=, prominence=, distance=) ----------
 * bpm_analysis.py:227 (sign=+1, height = noise floor) and :1070 (sign=-1, no height).
 * Order of conditions as scipy applies them: plateau-aware local maxima -> height ->
 * distance (greedy, highest first) -> prominence (wlen=None).
 *   height: float64[total_m] or NULL; prominence: float64[n_items] (device) or NULL;
 *   distance: samples (>= 1); out_idx: int64[total_m] (list of recording i at m_off);
 *   out_count: int64[n_items]. */
size_t bpm_find_peaks_workspace_bytes(int64_t total_m, int n_items);
int bpm_find_peaks(const double* x, int sign, const double* height, const double* prominence,
                   int distance, const BpmItem* items, const BpmItem* items_host, int n_items,
                   int64_t* out_idx, int64_t* out_count,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ---- K5+K6: Series(idx->val).reindex(arange(m)).interpolate()
 *             .rolling(window, min_periods=3, center=True).quantile(q).bfill().ffill()
 * bpm_analysis.py:1081-1086 / :1103-1106.  The interpolated series is never
 * materialised.  Output is all-NaN for a recording whose knot count is < 1. */
size_t bpm_rolling_floor_workspace_bytes(int64_t total_m, int n_items);
int bpm_rolling_floor(const double* envelope, const int64_t* knots, const int64_t* knot_count,
                      const BpmItem* items, const BpmItem* items_host, int n_items,
                      int window, double q, double* floor_out,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ---- a2 whole: _calculate_dynamic_noise_floor, bpm_analysis.py:1064-1117 ------------
 * K3 + K4(-env) + draft floor + K7 sanitisation + final floor, with the reference's
 * fall-backs (<5 troughs: constant floor and ALL troughs returned; <=2 kept: draft
 * floor; all-NaN: q(0.1)) resolved on the device.
 *   floor_out: float64[total_m]; troughs_out: int64[total_m]; trough_count: int64[n_items].
 *   trough_total, floor_mode (optional, int64[n_items]): troughs before sanitisation and the
 *   branch taken -- 0 sanitised floor, 1 draft floor (:1107-1110), 2 static floor (:1073-1077) --
 *   which is what the reference's log lines :1074, :1099, :1109 report. */
size_t bpm_noise_floor_workspace_bytes(int64_t total_m, int n_items);
int bpm_noise_floor(const double* envelope, const BpmItem* items, const BpmItem* items_host,
                    int n_items, int distance, double trough_prom_q, double floor_q, int window,
                    double rejection_multiplier, double* floor_out, int64_t* troughs_out,
                    int64_t* trough_count, int64_t* trough_total, int64_t* floor_mode,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ---- K7 on its own: trough sanitisation, bpm_analysis.py:1090-1097 -------------------
 * kept = [t for t in troughs if not isnan(draft[t]) and envelope[t] <= mult * draft[t]].
 * Needed separately when a long recording is split into halo-overlapped time chunks and
 * the draft floor of a chunk lives on the rank that owns it (bpm_analysis_b200/stream.py).
 *   troughs / kept_out: int64 lists at m_off; trough_count / kept_count: int64[n_items]. */
size_t bpm_sanitize_troughs_workspace_bytes(int64_t total_m, int n_items);
int bpm_sanitize_troughs(const double* envelope, const double* draft_floor, const int64_t* troughs,
                         const int64_t* trough_count, const BpmItem* items, const BpmItem* items_host,
                         int n_items, double rejection_multiplier, int64_t* kept_out,
                         int64_t* kept_count, void* workspace, size_t workspace_bytes, void* stream);

/* ---- a3: PeakClassifier._find_raw_peaks, bpm_analysis.py:223-229 ------------------- */
size_t bpm_raw_peaks_workspace_bytes(int64_t total_m, int n_items);
int bpm_raw_peaks(const double* envelope, const double* floor, const BpmItem* items,
                  const BpmItem* items_host, int n_items, int distance, double prom_q,
                  int64_t* peaks_out, int64_t* peak_count,
                  void* workspace, size_t workspace_bytes, void* stream);

/* ---- K8: numeric part of PeakClassifier._initialize_state, bpm_analysis.py:93-100 ---
 * strength[p] = max(0, env - floor) at peaks; deviation[p-1]; smoothed deviation
 * (centred rolling mean, window max(5, int((P-1)*smoothing_factor)), min_periods=1).
 * Outputs are laid out like the peak list (at m_off, length P resp. P-1); the unused tail of
 * `deviation` (entries P .. 2P-1 of the recording's m) is used as scratch for a prefix sum. */
int bpm_peak_metrics(const double* envelope, const double* floor, const int64_t* peaks,
                     const int64_t* peak_count, const BpmItem* items, const BpmItem* items_host,
                     int n_items, double smoothing_factor, double* strength, double* deviation,
                     double* smoothed, void* stream);

/* ---- K8b: surrounding-trough noise per raw peak (optional extra output, PARITY UNPINNED) ----
 * The reference's documentation describes it (Documentation/Changelog.md:454,
 * "BPM Detection logic explained.md":262, :276-278) and its config keys survive
 * (config.py:30-31: trough_veto_multiplier, trough_noise_multiplier), but no function of
 * bpm_analysis.py computes it any more; the oracle is oracle/ref_port.py::peak_trough_noise.
 *   prev_amp / next_amp: envelope at the sanitised trough just before / after each peak (NaN if
 *   none); ratio = min(prev, next) / floor[peak]; flags bit 0: ratio > trough_noise_multiplier,
 *   bit 1: look-ahead veto  m*(env[p]-next) < (env[p_next]-next).  Laid out like the peak list. */
int bpm_peak_trough_noise(const double* envelope, const double* floor, const int64_t* peaks,
                          const int64_t* peak_count, const int64_t* troughs, const int64_t* trough_count,
                          const BpmItem* items, const BpmItem* items_host, int n_items,
                          double trough_noise_multiplier, double trough_veto_multiplier,
                          double* prev_amp, double* next_amp, double* ratio, unsigned char* flags,
                          void* stream);

/* ---- beat-list reductions (K9, K10, K12).  A "beat list" is int64 indices at the
 * envelope rate; lists of a batch are back to back, described by BpmItem with
 * m = number of beats and m_off = first beat (in_off / n_in unused).            */

/* K9 calculate_bpm_series, bpm_analysis.py:1463-1484.  Per beat list with B beats:
 *   n_valid = number of intervals with dt > 1e-6; inst / smoothed / times_sec: float64,
 *   stamp_us: int64 microseconds since the series epoch (what datetime.timedelta
 *   quantises to); all four laid out at m_off, length n_valid.  window_us = smoothing
 *   window in microseconds; the centred window is (t - w/2, t + w/2]. */
int bpm_bpm_series(const int64_t* beats, const BpmItem* lists, const BpmItem* lists_host,
                   int n_lists, int rate, int64_t window_us, double* inst, double* smoothed,
                   double* times_sec, int64_t* stamp_us, int64_t* n_valid, void* stream);

/* K10 find_peak_recovery_rate / find_peak_exertion_rate, bpm_analysis.py:1552-1595.
 *   sign = -1: steepest decline searched from the series maximum onward (recovery);
 *   sign = +1: steepest incline over the whole series (exertion).
 *   sign =  0: both in one launch, result float64[n_lists][2][4] = {exertion, recovery}.
 *   result: float64[n_lists][4] = {found(0/1), start index, end index, slope}; indices
 *   are positions in the series of that list. */
int bpm_steepest_slope(const double* smoothed, const int64_t* stamp_us, const int64_t* n_valid,
                       const BpmItem* lists, const BpmItem* lists_host, int n_lists, int sign,
                       double window_sec, double* result, void* workspace, size_t workspace_bytes,
                       void* stream);
size_t bpm_steepest_slope_workspace_bytes(int64_t total_beats, int n_lists);

/* K12 calculate_windowed_hrv, bpm_analysis.py:1414-1461.  rows[i] = number of windows
 * of list i; out: float64[.][4] = {time, rmssdc, sdnn, bpm}, rows of list i start at
 * row m_off (capacity m). */
int bpm_windowed_hrv(const int64_t* beats, const BpmItem* lists, const BpmItem* lists_host,
                     int n_lists, int rate, int window_beats, int step_beats,
                     double* out, int64_t* rows, void* stream);

/* ---- ONE long recording as halo-overlapped time chunks over several GPUs ------------------------
 * (north star: "long single recordings split into halo-overlapped time chunks per GPU, with NCCL
 * used only to gather per-chunk peak lists and window stats"; host side: stream.ShardedFrontEnd.)
 * A rank runs every stage on its chunk + halo; the only stream-wide quantities are the np.quantile
 * thresholds (:1067, :225) -- resolved by bpm_key_histogram / bpm_key_collect with the histograms
 * summed over the ranks -- and the deviation-smoothing window (:99), which depends on the total peak
 * count, so bpm_deviation_series runs on the gathered strength list.
 *
 * bpm_key_histogram: hist[b] += number of samples whose order-preserving key k has
 *   (k >> (shift + bits)) == prefix  (every sample when shift + bits == 64)  and  ((k >> shift) & (2^bits - 1)) == b;
 *   hist: uint64[2^bits], accumulated (zero it first), bits <= 11.
 * bpm_key_collect: keys with (k >> up_shift) == prefix -> out_keys (first `cap` of them), count_min[0] += their
 *   number, count_min[1] = min(.., smallest key above the bucket), count_min[2] / [3] = smallest / largest key
 *   in the bucket  (preset count_min = {0, ~0, ~0, 0}).
 * The key of a double v: bits(v) with the sign bit set for v >= 0, all bits flipped for v < 0.
 * `state` (device, uint64[3] = {prefix, rank inside the bucket, bucket size}; may be NULL): when given,
 * the prefix is read from it instead of the argument, so that a whole descent -- histogram, all-reduce,
 * bpm_key_pick, ..., bpm_key_collect, all-gather, bpm_key_finish -- is enqueued without the host ever
 * waiting for the device.  Start it as {0, k, n} with k = floor((n - 1) q).
 * bpm_key_pick: the bin of the summed histogram that holds element `rank` becomes the next digit.
 * bpm_key_finish: rows = per rank {count, smallest key above, smallest, largest key in the bucket, keys[cap]}
 *   (the bpm_key_collect outputs, all-gathered, `cap + 4` words each); orders the bucket and writes numpy's
 *   linear-method quantile for the fraction gamma = (n - 1) q - k.  A bucket of ONE repeated value (digital
 *   silence) is resolved whatever its size; *status = 1 if the bucket holds more than 4096 distinct-looking
 *   keys (resolve another digit first). */
int bpm_key_histogram(const double* x, int64_t n, int shift, int bits, uint64_t prefix, const uint64_t* state,
                      uint64_t* hist, void* stream);
int bpm_key_collect(const double* x, int64_t n, int up_shift, uint64_t prefix, const uint64_t* state, int64_t cap,
                    uint64_t* out_keys, uint64_t* count_min, void* stream);
int bpm_key_pick(const uint64_t* hist, int bits, uint64_t* state, void* stream);
int bpm_key_finish(const uint64_t* rows, int world, int64_t cap, const uint64_t* state, double gamma, double* out,
                   int64_t* status, void* stream);

/* find_peaks on a chunk [ext_lo, ext_hi) of a stream (one recording per call).  core_lo / core_hi: the part
 * of the chunk (chunk-relative indices) whose peaks the caller keeps; open_left / open_right: that end of
 * the chunk is artificial.  *edge_hits (device, uint64, accumulated) counts every decision about a core
 * peak that could depend on samples beyond an open end -- a distance chain longer than the kernel
 * follows, a prominence walk that stopped at the end, a flat run from the end into the core.
 * anchors (device, int64[2], preset {-1, INT64_MAX}): [0] = the largest position <= core_lo, [1] = the
 * smallest position >= core_hi - 1 of a candidate that outranks every candidate within `distance` of it.
 * Such a candidate is kept whatever lies beyond it and removes all candidates within the distance, so
 * no decision on one side of it depends on the other side.  The core's peaks are exactly those of the
 * unchunked evaluation when *edge_hits == 0 and, towards each open end, an anchor lies between the core
 * and the zone where the chunk's candidates may differ from the stream's (ShardedFrontEnd checks this
 * and otherwise falls back to the replicated evaluation). */
int bpm_find_peaks_chunk(const double* x, int sign, const double* height, const double* prominence, int distance,
                         const BpmItem* items, const BpmItem* items_host, int64_t core_lo, int64_t core_hi,
                         int open_left, int open_right, int64_t* out_idx, int64_t* out_count, uint64_t* edge_hits,
                         int64_t* anchors, void* workspace, size_t workspace_bytes, void* stream);

/* _calculate_dynamic_noise_floor (:1064-1117) on a chunk with the stream-wide trough prominence threshold
 * GIVEN (device double[1]) and without the count-based fall-backs (:1073, :1102, :1113: decided on the
 * stream's totals by the caller).  all_troughs_out / all_count: the troughs before sanitisation. */
size_t bpm_noise_floor_chunk_workspace_bytes(int64_t m);
int bpm_noise_floor_chunk(const double* envelope, const BpmItem* items, const BpmItem* items_host, int distance,
                          const double* trough_prominence, double floor_q, int window, double rejection_multiplier,
                          int64_t core_lo, int64_t core_hi, int open_left, int open_right, double* floor_out,
                          int64_t* troughs_out, int64_t* trough_count, int64_t* all_troughs_out, int64_t* all_count,
                          uint64_t* edge_hits, int64_t* anchors, void* workspace, size_t workspace_bytes, void* stream);

/* The chunk's proof obligations evaluated on the device (one thread; the lists stay where they are):
 * counts = {kept, all, peaks}, flags = {edge hits, trough anchors l/r, peak anchors l/r} as filled by the
 * two calls above, quantile_status (may be NULL) = the bpm_key_finish status words of the thresholds.
 * [trough_lo, trough_hi) is the core given to bpm_noise_floor_chunk, [core_lo, core_hi) the chunk's own.
 * out int64[8] = {bad, all troughs in core, kept troughs in core, peaks in core, index of the first kept
 * trough in core, index of the first peak in core, proven floor range lo, hi}.  bad = 0 means: every
 * core output equals the unchunked evaluation's. */
int bpm_chunk_proof(const int64_t* all_troughs, const int64_t* kept_troughs, const int64_t* peaks, const int64_t* counts,
                    const int64_t* flags, const int64_t* quantile_status, int64_t n, int64_t core_lo, int64_t core_hi,
                    int64_t trough_lo, int64_t trough_hi, int at_start, int at_end, int64_t filter_halo, int distance,
                    int window, int64_t* out, void* stream);

/* The one list exchange of a chunked stream.  bpm_chunk_pack: a rank's contribution
 * [kept troughs of its core | peaks of its core | their strengths], indices moved to stream coordinates
 * (+ chunk_origin), the parts padded to cap_troughs / cap_peaks (the longest among the ranks); `proof` is
 * the rank's bpm_chunk_proof output.  bpm_chunk_unpack: rows = the all-gathered contributions, table =
 * the all-gathered proofs (world x 8) -> the stream's lists in rank order. */
int bpm_chunk_pack(const int64_t* kept, const int64_t* peaks, const double* strength, const int64_t* proof,
                   int64_t chunk_origin, int64_t cap_troughs, int64_t cap_peaks, int64_t* out, void* stream);
int bpm_chunk_unpack(const int64_t* rows, const int64_t* table, int world, int64_t cap_troughs, int64_t cap_peaks,
                     int64_t* troughs, int64_t* peaks, double* strength, void* stream);

/* strength[k] = max(0, env[p_k] - floor[p_k]) (:93-95) alone: what a chunk contributes to the stream's list */
int bpm_peak_strength(const double* envelope, const double* floor_, const int64_t* peaks, const int64_t* peak_count,
                      const BpmItem* items, const BpmItem* items_host, int n_items, double* strength, void* stream);

/* deviation / smoothed deviation (:96-100) from a given strength list; the descriptors give the LIST
 * lengths (m = number of peaks); `deviation` needs room for 2 P values per recording. */
int bpm_deviation_series(const double* strength, const int64_t* peak_count, const BpmItem* items,
                         const BpmItem* items_host, int n_items, double smoothing_factor, double* deviation,
                         double* smoothed, void* stream);

/* ---- float32 outputs (north star: "1e-4 (float32 mode)") ---------------------------------------
 * dst[i] = (float) src[i].  Every stage computes in float64 -- a float32 filter recurrence at these
 * pole radii is off by 1e-3 -- and only the signals handed back to the host (envelope, noise
 * floor, per-peak series) are rounded: half the read-back bytes, 6e-8 relative error. */
int bpm_cast_f32(const double* src, float* dst, int64_t n, void* stream);

/* ---- a1..a4 chained in one call (what analyze_wav_file does at :1731-1732 + :1635) --- */
typedef struct {
  int64_t stride, block;          /* filter placement, see bpm_frontend */
  int32_t pcm_dtype, channels;
  int32_t env_window;             /* rate // 10                 (:1053) */
  int32_t distance;               /* int(min_peak_distance_sec * rate)  (:226, :1066) */
  int32_t noise_window;           /* int(noise_window_sec * rate)       (:1083) */
  int32_t want_debug_wav;
  double trough_prom_q, peak_prom_q, floor_q, rejection_multiplier, smoothing_factor;
} BpmStageAConfig;

typedef struct {
  double* filtered; /* may be NULL, see bpm_frontend */
  double* envelope;  double* absmax;   int16_t* debug_wav; /* may be NULL */
  double* floor;      int64_t* troughs;  int64_t* trough_count;
  int64_t* peaks;     int64_t* peak_count;
  double* strength;   double* deviation; double* smoothed_dev;
  int64_t* trough_total; int64_t* floor_mode;   /* may be NULL, see bpm_noise_floor */
} BpmStageAOutputs;

size_t bpm_stage_a_workspace_bytes(int64_t total_m, int n_items);
int bpm_stage_a(const void* pcm, const BpmItem* items, const BpmItem* items_host, int n_items,
                const double* design, const double* design_host, int64_t design_words,
                const BpmStageAConfig* cfg_host,
                const BpmStageAOutputs* out_host, void* workspace, size_t workspace_bytes,
                void* stream);

/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t bpm_launch_count(void);

/* Diagnostic per-kernel timing (bench.py's roofline line).  Between begin and end an event
 * is recorded after every launch on `stream`; end synchronises and writes one line per
 * kernel, "name launches total_ms", into the HOST buffer `text_host`. */
int bpm_profile_begin(void* stream);
int bpm_profile_end(char* text_host, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* BPM_B200_H_ */
