/*
 * bpm_host.h -- C ABI of libbpm_host.so: compiled HOST code for the sequential stage that
 * consumes the B200 front end's outputs.
 *
 * The reference runs its S1/S2 classifier as a Python loop over the raw peaks with
 * loop-carried state (bpm_analysis.py:113-329 and the helpers at :1120-1255), twice per
 * file (:1635, :1740); once the front end is on the GPU it is what is left of the run time.
 * It does not data-parallelise (every decision feeds the next through the long-term BPM
 * belief, the candidate list and the rejection counter), so it is compiled host code, not
 * a kernel; batches are classified one recording per host thread by the caller.
 *
 * All pointers are HOST pointers.  The library has no global state and is re-entrant.
 * Return value: 0, or BPM_HOST_ERR_*.  It never throws and never exits.
 */
#ifndef BPM_HOST_H_
#define BPM_HOST_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BPM_HOST_ABI_VERSION 2

enum { BPM_HOST_OK = 0, BPM_HOST_ERR_ARG = -1, BPM_HOST_ERR_NOMEM = -2 };

/* PeakType of the reference (bpm_analysis.py:26-36) as far as the classifier assigns it */
enum {
  BPM_PEAK_UNSET = 0,
  BPM_PEAK_S1_PAIRED = 1,          /* "S1 (Paired)" */
  BPM_PEAK_S2_PAIRED = 2,          /* "S2 (Paired)" */
  BPM_PEAK_LONE_S1 = 3,            /* "Lone S1" */
  BPM_PEAK_LONE_S1_CASCADE = 4,    /* "Lone S1 (Corrected by Cascade Reset)" */
  BPM_PEAK_LONE_S1_LAST = 5,       /* "Lone S1 (Last Peak)" */
  BPM_PEAK_NOISE = 6               /* "Noise" */
};

/* The parameters classify_peaks reads from DEFAULT_PARAMS (config.py), with the reference's
 * .get() defaults applied by the caller, plus the constructor arguments of PeakClassifier
 * (bpm_analysis.py:71-83).  Integers of the Python dict are passed as doubles: every use is
 * a float expression.  */
typedef struct {
  double pairing_confidence_threshold;      /* :267 */
  double contractility_bpm_low;             /* :237, :1136, :1171 */
  double contractility_bpm_high;
  double s1_s2_interval_cap_sec;            /* :246 */
  double s1_s2_interval_rr_fraction;
  double interval_penalty_start_factor;     /* :250-252 */
  double interval_penalty_full_factor;
  double interval_max_penalty;
  double kickstart_check_threshold;         /* :143 */
  double stability_confidence_floor;        /* :1155-1156 */
  double stability_confidence_ceiling;
  double s2_s1_ratio_low_bpm;               /* :1172-1174 */
  double s2_s1_ratio_high_bpm;
  double penalty_amount_min;                /* :1179-1180 */
  double penalty_amount_max;
  double s1_s2_boost_ratio;                 /* :1188 */
  double boost_amount_min;                  /* :1190-1191 */
  double boost_amount_max;
  double lone_s1_confidence_threshold;      /* :310 */
  double lone_s1_forward_check_pct;         /* :319 */
  double lone_s1_rhythm_weight;             /* :326, :1236 */
  double lone_s1_amplitude_weight;
  double cascade_reset_trigger_count;       /* :293 */
  double min_bpm;                           /* :1258 */
  double max_bpm;
  double start_bpm;                         /* state['long_term_bpm'] at entry (:103) */
  double peak_bpm_time_sec;                 /* used when has_recovery_window != 0 (:1168-1169) */
  double recovery_end_time_sec;
  int32_t has_recovery_window;
  int32_t enable_interval_penalty;          /* :249 */
  int32_t stability_history_window;         /* :135, :180 */
  int32_t flags;                            /* BPM_CLASSIFY_* */
} BpmClassifierParams;

/* flags: skip the per-peak debug strings (text_bytes = n_peaks NULs).  Decisions, beats, types,
 * the BPM trace and the events are unchanged -- e.g. for the preliminary pass, whose strings the
 * reference throws away (bpm_analysis.py:1635), or for batch services that only want beats. */
#define BPM_CLASSIFY_NO_TEXT 1

/* things the reference reports through logging.info while classifying */
enum { BPM_EVENT_KICKSTART = 1, BPM_EVENT_CASCADE_RESET = 2 };
typedef struct {
  int32_t kind;
  int32_t a;       /* KICKSTART: matches            CASCADE_RESET: position of the forced peak */
  int32_t b;       /* KICKSTART: recent lone S1s     CASCADE_RESET: 0 */
  int32_t pad;
} BpmClassifierEvent;

/* Result of one classification; owned by the library, released with bpm_classification_free. */
typedef struct {
  int64_t n_peaks;                 /* raw peaks classified */
  int64_t n_beats;                 /* len(candidate_beats) */
  int64_t n_history;               /* len(long_term_bpm_history) */
  int64_t n_events;
  int64_t text_bytes;
  const int64_t* beat_positions;   /* [n_beats] positions INTO raw_peaks, ascending */
  const int32_t* peak_types;       /* [n_peaks] BPM_PEAK_* */
  const int64_t* text_offsets;     /* [n_peaks + 1] byte offsets into text */
  const char* text;                /* UTF-8, one NUL-terminated entry per raw peak, back to back:
                                      beat_debug_info[raw_peaks[i]] = text + off[i]  (off[i+1] - off[i] - 1 bytes) */
  const double* history_times;     /* [n_history] candidate_beats[-1] / sample_rate */
  const double* history_bpm;       /* [n_history] long_term_bpm after each decision */
  const BpmClassifierEvent* events;
  double final_long_term_bpm;
  int64_t final_consecutive_rr_rejections;
} BpmClassification;

int bpm_host_abi_version(void);

/* The fixed-point formatter behind the debug strings: writes format(v, ".{prec}f") as Python
 * prints it (correctly rounded, "nan" / "inf" spelled like Python), NUL-terminated; returns the
 * length or BPM_HOST_ERR_ARG.  Exported so that tests can pin it against Python directly. */
int bpm_host_format_fixed(double v, int prec, char* out, size_t capacity);

/* PeakClassifier.classify_peaks (bpm_analysis.py:113-131) for n_peaks >= 2 raw peaks.
 *   envelope, noise_floor: float64[m] (audio_envelope, dynamic_noise_floor.values)
 *   raw_peaks: int64[n_peaks], strictly ascending, each in [0, m)  (state['all_peaks'])
 *   dev_times, dev_values: float64[n_dev], the smoothed deviation series (index seconds ascending, values)
 *   sample_rate: the envelope rate (a Python int in the reference; every use is a true division)
 * Decisions, candidate beats, the long-term BPM trace and every per-peak debug string are those
 * of the reference, bit for bit and byte for byte.  */
int bpm_classify_peaks(const double* envelope, const double* noise_floor, int64_t m,
                       const int64_t* raw_peaks, int64_t n_peaks,
                       const double* dev_times, const double* dev_values, int64_t n_dev,
                       double sample_rate, const BpmClassifierParams* params,
                       BpmClassification** out);
void bpm_classification_free(BpmClassification* c);

/* A batch of independent recordings, one per host thread (the loop inside a recording is
 * sequential; recordings are not).  Job i is classified exactly as bpm_classify_peaks would;
 * status[i] receives its return code and out[i] its result (NULL on error), each released with
 * bpm_classification_free.  n_threads <= 0: one thread per hardware thread, at most n_jobs. */
typedef struct {
  const double* envelope;
  const double* noise_floor;
  int64_t m;
  const int64_t* raw_peaks;
  int64_t n_peaks;
  const double* dev_times;
  const double* dev_values;
  int64_t n_dev;
  double sample_rate;
  const BpmClassifierParams* params;
} BpmClassifyJob;
int bpm_classify_peaks_batch(const BpmClassifyJob* jobs, int64_t n_jobs, int n_threads,
                             BpmClassification** out, int* status);

/* ---- correction passes over the classified beat list ----------
 * The loops of correct_peaks_by_rhythm (bpm_analysis.py:1257-1306) and
 * _fix_rhythmic_discontinuities (:1309-1412).  The few vectorised numpy calls in front of them
 * (np.diff, np.median, np.percentile -> the thresholds) stay numpy calls in the Python mirror;
 * what is compiled is the per-beat Python loop, including the scan of EVERY raw peak the
 * reference repeats for each long gap (:1355-1356).  Decisions are reported as events; the
 * mirror applies them to the debug-string dict and writes the reference's log lines. */
enum {
  BPM_CORR_REPLACED = 1,        /* a = current peak, b = the accepted peak it replaced   (:1291) */
  BPM_CORR_DISCARDED = 2,       /* a = current peak                                      (:1295) */
  BPM_CORR_LONG_INTERVAL = 3,   /* a = S1 peak that starts the long interval             (:1353) */
  BPM_CORR_RELABELLED = 4,      /* a, b = POSITIONS in raw_peaks of the new S1 / S2      (:1373-1381) */
  BPM_CORR_SHORT_INTERVAL = 5,  /* a, b = the two beats, x = their interval in seconds   (:1398) */
  BPM_CORR_REMOVED = 6          /* a = the weaker beat                                   (:1405-1411) */
};
typedef struct {
  int32_t kind;
  int32_t pad;
  int64_t a;
  int64_t b;
  double x;
} BpmCorrectionEvent;

/* correct_peaks_by_rhythm for len(peaks) >= 5: peaks closer than threshold_sec to the last accepted
 * one conflict; the higher envelope amplitude wins.  out: capacity n; events: capacity n. */
int bpm_correct_peaks_by_rhythm(const int64_t* peaks, int64_t n, const double* envelope, int64_t m,
                                double sample_rate, double threshold_sec,
                                int64_t* out, int64_t* n_out, BpmCorrectionEvent* events, int64_t* n_events);

/* _fix_rhythmic_discontinuities after its thresholds are known (len(s1_peaks) >= 6).
 *   raw_is_noise[i] != 0  <=>  "Noise" in debug_info.get(raw_peaks[i], "")   (evaluated by the caller on
 *   the strings, as the reference does: a re-labelled peak keeps "Noise" inside its ORIGINAL_REASON)
 *   out: capacity n_s1 + n_raw; events: capacity 4 * (n_s1 + n_raw) + 8. */
int bpm_fix_rhythmic_discontinuities(const int64_t* s1_peaks, int64_t n_s1, const int64_t* raw_peaks, int64_t n_raw,
                                     const uint8_t* raw_is_noise, const double* envelope, const double* noise_floor,
                                     int64_t m, double sample_rate, double short_threshold_sec,
                                     double long_threshold_sec, double waiver_strength_ratio,
                                     double waiver_max_s2_s1_ratio, int64_t* out, int64_t* n_out,
                                     BpmCorrectionEvent* events, int64_t events_capacity, int64_t* n_events,
                                     int64_t* corrections_made);

/* ---- ingest -------------------------------------------------------
 * K0 on the host cores: out[j] = frame j * stride of pcm, j = 0 .. ceil(n_frames / stride) - 1
 * (audio_data[::downsample_factor], bpm_analysis.py:1033; a frame = all channels of one sample,
 * frame_bytes = channels * sample size, kept in the PCM's own dtype).  `out` is normally a pinned
 * staging buffer that one cudaMemcpyAsync then moves to the device, where bpm_frontend / bpm_stage_a
 * read it with stride 1 -- bit-identical to passing the whole recording with this stride.
 * n_threads <= 0: bpm_host_threads() threads = $BPM_HOST_THREADS, else this process's share of the cores
 * it may run on (its affinity mask divided by $LOCAL_WORLD_SIZE, which torchrun sets: the ranks of a
 * box split the cores instead of oversubscribing them), at most the pool's 32. */
int bpm_host_threads(void);
int bpm_host_gather_frames(const void* pcm, int64_t frame_bytes, int64_t n_frames, int64_t stride, void* out,
                           int n_threads);

/* The same for 24-bit PCM (3-byte little-endian samples, `channels` per frame): out[j * channels + c] =
 * sample c of frame j * stride as int32 with the 24 bits in the upper three bytes -- the array
 * scipy.io.wavfile.read gives the reference for such a file (bpm_analysis.py:1014), which cannot be
 * memory-mapped and which scipy expands in full.  stride 1 expands the whole recording. */
int bpm_host_gather_s24(const void* pcm, int64_t channels, int64_t n_frames, int64_t stride, int32_t* out,
                        int n_threads);

/* ---- on-disk outputs ----------------------------------------------
 * One text table in one call: for every i < n (rows with NaN b[i] left out when skip_nan_b != 0)
 *     head  format(a[i], ".{prec_a}f")  mid  format(b[i], ".{prec_b}f")  tail
 * with bpm_host_format_fixed's digits (= Python's).  Replaces the per-row f-strings of the
 * `_bpm_plot.csv` writer (bpm_analysis.py:466-469: head "", mid ",", tail "\r\n", 3 / 3 digits) and
 * of the summary's heartbeat table (:979-981: "| ", " | ", " |\n", 2 / 1 digits).
 * Returns the number of bytes the table takes; if that exceeds `capacity` the contents of `out`
 * are unspecified and the call is to be repeated with a buffer of that size.  No NUL is appended.
 * Negative: BPM_HOST_ERR_ARG. */
int64_t bpm_host_format_rows(const double* a, const double* b, int64_t n, int prec_a, int prec_b, const char* head,
                             const char* mid, const char* tail, int skip_nan_b, char* out, int64_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* BPM_HOST_H_ */
