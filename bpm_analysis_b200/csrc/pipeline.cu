// extern "C" surface of libbpm_b200.so (include/bpm_b200.h) and the chained stages
// a2 (_calculate_dynamic_noise_floor, bpm_analysis.py:1064-1117), a3 (_find_raw_peaks,
// :223-229) and a1..a4 in one call.  All data-dependent fall-backs of the reference are
// resolved by per-recording control words on the device, so a stage is enqueued without
// any host synchronisation.
#include "common.cuh"

namespace bpm {

// filter.cu
size_t frontend_workspace_bytes(int64_t total_m, int n_items);
int frontend_run(const void* pcm, int pcm_dtype, int channels, const BpmItem* items, const BpmItem* items_host,
                 int n_items, int64_t stride, const double* design, const double* design_host, int64_t design_words,
                 int block, int env_window, double* filtered, double* envelope, double* absmax, Workspace& ws,
                 cudaStream_t st);
int debug_wav_run(const double* filtered, const double* absmax, const BpmItem* items, const BpmItem* items_host,
                  int n_items, int16_t* out, cudaStream_t st);
int gather_frames_run(const void* pcm, int pcm_dtype, int channels, const BpmItem* items, const BpmItem* items_host,
                      int n_items, int64_t stride, double* out, cudaStream_t st);
// select.cu
size_t quantile_workspace_bytes(int n_items);
int quantile_run(const double* x, const BpmItem* items, const BatchShape& sh, double q, const int* cond,
                 double* out, Workspace& ws, cudaStream_t st);
int quantile_multi_run(const double* x, const BpmItem* items, const BatchShape& sh, int nq, const double* q,
                       const int* const* cond, double* const* out, Workspace& ws, cudaStream_t st);
// peaks.cu
size_t find_peaks_workspace_bytes(int64_t total_m, int n_items);
int find_peaks_run(const double* x, int sign, const double* height, const double* prominence, int distance,
                   const BpmItem* items, const BatchShape& sh, int64_t* out_idx, int64_t* out_count,
                   Workspace& ws, cudaStream_t st, cudaEvent_t prominence_ready = nullptr,
                   const ChunkInfo* chunk = nullptr);
// floor.cu
size_t rolling_floor_workspace_bytes(int64_t total_m, int n_items);
bool rolling_floor_sparse_ok(int window);
int rolling_floor_run(const double* env, const int64_t* knots, const int64_t* knot_count, const BpmItem* items,
                      const BatchShape& sh, int window, double q, const int64_t* mode_n_all,
                      const int64_t* mode_n_kept, const int64_t* alt_knots, const int64_t* alt_count, const double* cval,
                      const double* nan_fill, int64_t* total_out, int64_t* mode_out, double* out, double* sparse_out,
                      Workspace& ws, cudaStream_t st);
size_t sanitize_workspace_bytes(int64_t total_m, int n);
int sanitize_run(const double* env, const double* draft, const int64_t* troughs, const int64_t* trough_count,
                 int keep_few, const BpmItem* items, const BatchShape& sh, double mult, int draft_by_knot,
                 int64_t* kept_out, int64_t* kept_count, Workspace& ws, cudaStream_t st);
// metrics.cu
int peak_metrics_run(const double* env, const double* floor_, const int64_t* peaks, const int64_t* peak_count,
                     const BpmItem* items, const BatchShape& sh, double factor, double* strength,
                     double* deviation, double* smoothed, cudaStream_t st);
int peak_trough_noise_run(const double* env, const double* floor_, const int64_t* peaks, const int64_t* peak_count,
                          const int64_t* troughs, const int64_t* trough_count, const BpmItem* items,
                          const BatchShape& sh, double noise_mult, double veto_mult, double* prev_amp,
                          double* next_amp, double* ratio, unsigned char* flags, cudaStream_t st);
int bpm_series_run(const int64_t* beats, const BpmItem* lists, const BatchShape& sh, int rate, int64_t window_us,
                   double* inst, double* smoothed, double* times_sec, int64_t* stamp_us, int64_t* n_valid,
                   cudaStream_t st);
int steepest_run(const double* smoothed, const int64_t* stamp_us, const int64_t* n_valid, const BpmItem* lists,
                 int n_lists, int64_t max_len, int sign, double window_sec, double* result, cudaStream_t st);
int cast_f32_run(const double* src, float* dst, int64_t n, cudaStream_t st);
int deviation_series_run(const double* strength, const int64_t* peak_count, const BpmItem* items, const BatchShape& sh,
                         double factor, double* deviation, double* smoothed, cudaStream_t st, bool list_sized = false);
int peak_strength_run(const double* env, const double* floor_, const int64_t* peaks, const int64_t* peak_count,
                      const BpmItem* items, const BatchShape& sh, double* strength, cudaStream_t st);
int hrv_run(const int64_t* beats, const BpmItem* lists, const BatchShape& sh, int rate, int win, int step,
            double* out, int64_t* rows, cudaStream_t st);

// ------------------------------------------------------------------ a2
struct NoiseFloorScratch {
  double* q_tp;        // [n] trough prominence threshold
  double* q_nf;        // [n] static floor for the "<5 troughs" path
  double* q_fb;        // [n] q(0.1) for the all-NaN path (aliases q_tp when trough_prom_q == 0.1)
  int64_t* all_troughs;
  int64_t* n_all;
  double* draft;
  char* sanitize_ws;
  size_t sanitize_ws_bytes;
  char* select_ws;          // the quantile passes run concurrently with the trough search: own scratch
  size_t select_ws_bytes;
};

static int carve_noise_floor(Workspace& ws, int64_t total_m, int n, NoiseFloorScratch* s) {
  s->q_tp = ws.take<double>(n);
  s->q_nf = ws.take<double>(n);
  s->q_fb = ws.take<double>(n);
  s->all_troughs = ws.take<int64_t>(total_m);
  s->n_all = ws.take<int64_t>(n);
  s->draft = ws.take<double>(total_m);
  s->sanitize_ws_bytes = sanitize_workspace_bytes(total_m, n);
  s->sanitize_ws = ws.take<char>(s->sanitize_ws_bytes);
  s->select_ws_bytes = quantile_workspace_bytes(n);
  s->select_ws = ws.take<char>(s->select_ws_bytes);
  return ws.overflow ? BPM_ERR_WORKSPACE : BPM_OK;
}

// Sub-steps run one after another on one stream, so their scratch can share memory: each
// takes a nested Workspace over the same tail region.
static Workspace sub_ws(Workspace& parent, size_t bytes) {
  if (parent.measuring()) {
    // account for the largest sub-step once
    return Workspace(nullptr, 0);
  }
  char* p = parent.base ? parent.base + parent.used : nullptr;
  size_t cap = parent.cap > parent.used ? parent.cap - parent.used : 0;
  (void)bytes;
  return Workspace(p, cap);
}

static size_t noise_floor_sub_bytes(int64_t total_m, int n) {
  size_t a = quantile_workspace_bytes(n), b = find_peaks_workspace_bytes(total_m, n),
         c = rolling_floor_workspace_bytes(total_m, n);
  size_t mx = a > b ? a : b;
  return mx > c ? mx : c;
}

size_t noise_floor_workspace_bytes(int64_t total_m, int n) {
  Workspace ws(nullptr, 0);
  NoiseFloorScratch s;
  carve_noise_floor(ws, total_m, n, &s);
  return ws.used + noise_floor_sub_bytes(total_m, n);
}

int noise_floor_run(const double* env, const BpmItem* items, const BatchShape& sh, int distance,
                    double trough_prom_q, double floor_q, int window, double mult, double* floor_out,
                    int64_t* troughs_out, int64_t* trough_count, int64_t* total_out /* optional [n] */,
                    int64_t* mode_out /* optional [n] */, double* q_tp_out /* optional [n] */,
                    Workspace& ws, cudaStream_t st) {
  if (!env || !items || !floor_out || !troughs_out || !trough_count) return BPM_ERR_ARG;
  if (distance < 1 || window < MIN_PERIODS) return BPM_ERR_ARG;
  const int n = sh.n_items;
  NoiseFloorScratch s;
  BPM_TRY(carve_noise_floor(ws, sh.total_m, n, &s));
  double* q_tp = q_tp_out ? q_tp_out : s.q_tp;
  const double* q_fb = q_tp;
  {
    // every quantile of the envelope this stage can need, resolved in ONE set of radix passes:
    // q(trough_prominence) (:1067), q(noise_floor_quantile) for the "<5 troughs" path (:1075 -- computed
    // unconditionally, it costs no extra launch) and q(0.1) for the all-NaN path (:1114).  They run on
    // the auxiliary stream, next to the local-maximum / distance steps of the trough search, which
    // only need the prominence threshold at their very end (own workspace: the two overlap in time).
    double qs[3] = {trough_prom_q, floor_q, 0.1};
    double* outs[3] = {q_tp, s.q_nf, s.q_fb};
    int nq = 2;
    if (trough_prom_q != 0.1) { nq = 3; q_fb = s.q_fb; }
    ForkJoin fj;
    BPM_TRY(fj.begin(st));
    Workspace wq(ws.measuring() ? nullptr : s.select_ws, s.select_ws_bytes);
    BPM_TRY(quantile_multi_run(env, items, sh, nq, qs, nullptr, outs, wq, fj.aux));
    BPM_TRY(fj.end_aux());
    Workspace w = sub_ws(ws, 0);
    BPM_TRY(find_peaks_run(env, -1, nullptr, q_tp, distance, items, sh, s.all_troughs, s.n_all, w, st,
                           fj.join_event()));                                                        // :1070
    if (fj.active && cudaGetLastError() != cudaSuccess) return BPM_ERR_CUDA;
  }
  // draft floor from all troughs (:1081-1086).  It is only ever read AT the troughs (:1093), so it
  // is computed there only (one value per trough) whenever the block-cooperative kernel applies.
  const bool sparse = rolling_floor_sparse_ok(window);
  {
    Workspace w = sub_ws(ws, 0);
    BPM_TRY(rolling_floor_run(env, s.all_troughs, s.n_all, items, sh, window, floor_q, s.n_all, nullptr, nullptr, nullptr,
                              s.q_nf, nullptr, nullptr, nullptr, sparse ? nullptr : s.draft, sparse ? s.draft : nullptr,
                              w, st));
  }
  {
    Workspace wz(ws.measuring() ? nullptr : s.sanitize_ws, s.sanitize_ws_bytes);
    BPM_TRY(sanitize_run(env, s.draft, s.all_troughs, s.n_all, 1, items, sh, mult, sparse ? 1 : 0, troughs_out,
                         trough_count, wz, st));                                                     // :1090-1097
  }
  {
    // final floor from the kept troughs (:1102-1106); when <= 2 are kept the reference reuses the
    // draft (:1107-1110) = the same rolling quantile over ALL troughs, recomputed here (mode 1
    // makes the knot table take the other list); constant on the "<5" path (:1073-1077), q(0.1)
    // when everything is NaN (:1113-1115)
    Workspace w = sub_ws(ws, 0);
    BPM_TRY(rolling_floor_run(env, troughs_out, trough_count, items, sh, window, floor_q, s.n_all, trough_count,
                              s.all_troughs, s.n_all, s.q_nf, q_fb, total_out, mode_out, floor_out, nullptr, w, st));
  }
  return BPM_OK;
}

// ------------------------------------------------------------------ a2 on one time chunk of a longer stream
// The stream-wide prominence threshold is GIVEN (it is an order statistic of the whole envelope), and
// the count-based fall-backs of the reference (<5 troughs, <=2 kept, all-NaN) are NOT applied here:
// they are decisions on the stream's totals, which the caller takes after gathering the counts.
int noise_floor_chunk_run(const double* env, const BpmItem* items, const BatchShape& sh, int distance,
                          const double* q_tp, double floor_q, int window, double mult, const ChunkInfo& ci,
                          double* floor_out, int64_t* troughs_out, int64_t* trough_count, int64_t* all_troughs_out,
                          int64_t* all_count, Workspace& ws, cudaStream_t st) {
  if (!env || !items || !q_tp || !floor_out || !troughs_out || !trough_count || !all_troughs_out || !all_count)
    return BPM_ERR_ARG;
  if (distance < 1 || window < MIN_PERIODS || sh.n_items != 1) return BPM_ERR_ARG;
  double* draft = ws.take<double>(sh.total_m);
  const size_t zb = sanitize_workspace_bytes(sh.total_m, 1);
  char* sanitize_ws = ws.take<char>(zb);
  if (ws.overflow) return BPM_ERR_WORKSPACE;
  {
    Workspace w = sub_ws(ws, 0);
    BPM_TRY(find_peaks_run(env, -1, nullptr, q_tp, distance, items, sh, all_troughs_out, all_count, w, st, nullptr, &ci));
  }
  const bool sparse = rolling_floor_sparse_ok(window);
  {
    Workspace w = sub_ws(ws, 0);
    BPM_TRY(rolling_floor_run(env, all_troughs_out, all_count, items, sh, window, floor_q, nullptr, nullptr, nullptr, nullptr,
                              nullptr, nullptr, nullptr, nullptr, sparse ? nullptr : draft, sparse ? draft : nullptr, w, st));
  }
  {
    Workspace wz(sanitize_ws, zb);
    BPM_TRY(sanitize_run(env, draft, all_troughs_out, all_count, 0, items, sh, mult, sparse ? 1 : 0, troughs_out,
                         trough_count, wz, st));
  }
  {
    Workspace w = sub_ws(ws, 0);
    BPM_TRY(rolling_floor_run(env, troughs_out, trough_count, items, sh, window, floor_q, nullptr, nullptr, nullptr, nullptr,
                              nullptr, nullptr, nullptr, nullptr, floor_out, nullptr, w, st));
  }
  return BPM_OK;
}

// ------------------------------------------------------------------ K7 on its own
// Used when a long recording is processed as halo-overlapped time chunks (stream.py): the draft
// floor of a chunk exists only on the rank that owns it, so that rank sanitises its troughs
// (sanitize_run, floor.cu).

// ------------------------------------------------------------------ a3
size_t raw_peaks_workspace_bytes(int64_t total_m, int n) {
  Workspace ws(nullptr, 0);
  ws.take<double>(n);
  size_t a = quantile_workspace_bytes(n), b = find_peaks_workspace_bytes(total_m, n);
  return ws.used + (a > b ? a : b);
}

int raw_peaks_run(const double* env, const double* floor_, const BpmItem* items, const BatchShape& sh,
                  int distance, double prom_q, const double* q_ready /* optional: already computed */,
                  int64_t* peaks_out, int64_t* peak_count, Workspace& ws, cudaStream_t st) {
  if (!env || !floor_ || !items || !peaks_out || !peak_count || distance < 1) return BPM_ERR_ARG;
  double* q = ws.take<double>(sh.n_items);
  if (ws.overflow) return BPM_ERR_WORKSPACE;
  const double* thr = q_ready;
  if (!thr) {
    Workspace w = sub_ws(ws, 0);
    BPM_TRY(quantile_run(env, items, sh, prom_q, nullptr, q, w, st));                             // :225
    thr = q;
  }
  Workspace w = sub_ws(ws, 0);
  return find_peaks_run(env, +1, floor_, thr, distance, items, sh, peaks_out, peak_count, w, st);   // :227
}

}  // namespace bpm

// ------------------------------------------------------------------ per-kernel timing
// Diagnostic used by bench.py for the roofline line: while profiling is on, an event is
// recorded after every launch; on a serial stream consecutive events bracket one kernel.
#include <map>
#include <string>
#include <vector>
namespace bpm {
static std::vector<cudaEvent_t> g_prof_pool;
static std::vector<std::pair<const char*, int>> g_prof_marks;   // (kernel name, event index)
static size_t g_prof_used = 0;

static cudaEvent_t prof_event() {
  if (g_prof_used == g_prof_pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    g_prof_pool.push_back(e);
  }
  return g_prof_pool[g_prof_used++];
}

void profile_mark(const char* name, cudaStream_t st) {
  const int idx = static_cast<int>(g_prof_used);
  cudaEventRecord(prof_event(), st);
  g_prof_marks.emplace_back(name, idx);
}
}  // namespace bpm

// =========================================================================== C ABI
using namespace bpm;

extern "C" {

int bpm_abi_version(void) { return BPM_ABI_VERSION; }

const char* bpm_error_string(int code) {
  switch (code) {
    case BPM_OK: return "ok";
    case BPM_ERR_ARG: return "invalid argument";
    case BPM_ERR_WORKSPACE: return "workspace too small";
    case BPM_ERR_CUDA: return "CUDA error";
    case BPM_ERR_TOO_SHORT: return "recording not longer than the filter padding (15 samples)";
    default: return "unknown error";
  }
}

int64_t bpm_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }


int bpm_profile_begin(void* stream) {
  g_prof_used = 0;
  g_prof_marks.clear();
  g_profiling = true;
  profile_mark("<begin>", static_cast<cudaStream_t>(stream));
  return BPM_OK;
}

// Writes "name count total_ms\n" lines (aggregated per kernel) into `text`.
int bpm_profile_end(char* text, size_t cap) {
  g_profiling = false;
  if (!text || cap == 0) return BPM_ERR_ARG;
  if (g_prof_marks.empty()) { text[0] = 0; return BPM_OK; }
  if (cudaEventSynchronize(g_prof_pool[g_prof_marks.back().second]) != cudaSuccess) return BPM_ERR_CUDA;
  std::map<std::string, std::pair<long, double>> agg;
  for (size_t i = 1; i < g_prof_marks.size(); ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, g_prof_pool[g_prof_marks[i - 1].second], g_prof_pool[g_prof_marks[i].second]);
    auto& a = agg[g_prof_marks[i].first];
    a.first += 1;
    a.second += ms;
  }
  std::string out;
  for (auto& kv : agg) out += kv.first + " " + std::to_string(kv.second.first) + " " + std::to_string(kv.second.second) + "\n";
  if (out.size() + 1 > cap) return BPM_ERR_WORKSPACE;
  memcpy(text, out.c_str(), out.size() + 1);
  return BPM_OK;
}

size_t bpm_frontend_workspace_bytes(int64_t total_m, int n_items) { return frontend_workspace_bytes(total_m, n_items); }

int bpm_frontend(const void* pcm, int pcm_dtype, int channels, const BpmItem* items, const BpmItem* items_host,
                 int n_items, int64_t stride, const double* design, const double* design_host, int64_t design_words,
                 int env_window, double* filtered, double* envelope, double* absmax, void* workspace,
                 size_t workspace_bytes, void* stream) {
  if (!workspace || n_items <= 0 || !items_host) return BPM_ERR_ARG;
  // block length is word 0 of the design image; the caller states it through items' m
  // (validated in frontend_run), so read it from the host-visible relation m = ceil(n_dec / block)
  Workspace ws(workspace, workspace_bytes);
  // block is passed implicitly: design_words = header + 4 * (2 * block + 1)
  const int64_t rem = design_words - BPM_DESIGN_HEADER_WORDS - BPM_DESIGN_LANE_WORDS;
  if (rem < 28 || (rem - 12) % 16 != 0) return BPM_ERR_ARG;
  const int block = static_cast<int>((rem - 12) / 16);
  return frontend_run(pcm, pcm_dtype, channels, items, items_host, n_items, stride, design, design_host, design_words,
                      block, env_window, filtered, envelope, absmax, ws, static_cast<cudaStream_t>(stream));
}

int bpm_gather_frames(const void* pcm, int pcm_dtype, int channels, const BpmItem* items, const BpmItem* items_host,
                      int n_items, int64_t stride, double* frames_out, void* stream) {
  return gather_frames_run(pcm, pcm_dtype, channels, items, items_host, n_items, stride, frames_out,
                           static_cast<cudaStream_t>(stream));
}

int bpm_copy_frames(const void* pcm, int pcm_dtype, int channels, const BpmItem* items_host, int n_items,
                    int64_t stride, void* frames_out, void* stream) {
  static const size_t sample_bytes[5] = {2, 4, 1, 4, 8};               // BPM_PCM_I16, I32, U8, F32, F64
  if (!pcm || !items_host || !frames_out || n_items <= 0 || channels < 1 || stride < 1 || pcm_dtype < 0 ||
      pcm_dtype > BPM_PCM_F64)
    return BPM_ERR_ARG;
  const size_t fb = sample_bytes[pcm_dtype] * static_cast<size_t>(channels);
  for (int i = 0; i < n_items; ++i) {
    const BpmItem& it = items_host[i];
    if (it.n_in <= 0 || it.m != (it.n_in + stride - 1) / stride) return BPM_ERR_ARG;
  }
  for (int i = 0; i < n_items; ++i) {
    const BpmItem& it = items_host[i];
    const char* src = static_cast<const char*>(pcm) + static_cast<size_t>(it.in_off) * fb;
    char* dst = static_cast<char*>(frames_out) + static_cast<size_t>(it.m_off) * fb;
    if (cudaMemcpy2DAsync(dst, fb, src, static_cast<size_t>(stride) * fb, fb, static_cast<size_t>(it.m),
                          cudaMemcpyDefault, static_cast<cudaStream_t>(stream)) != cudaSuccess) {
      cudaGetLastError();
      return BPM_ERR_CUDA;
    }
  }
  return BPM_OK;
}

int bpm_debug_wav(const double* filtered, const double* absmax, const BpmItem* items, const BpmItem* items_host,
                  int n_items, int16_t* out, void* stream) {
  return debug_wav_run(filtered, absmax, items, items_host, n_items, out, static_cast<cudaStream_t>(stream));
}

size_t bpm_quantile_workspace_bytes(int n_items) { return quantile_workspace_bytes(n_items); }

int bpm_quantile(const double* x, const BpmItem* items, const BpmItem* items_host, int n_items, double q,
                 double* out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!workspace || !items_host || n_items <= 0) return BPM_ERR_ARG;
  Workspace ws(workspace, workspace_bytes);
  return quantile_run(x, items, batch_shape(items_host, n_items), q, nullptr, out, ws, static_cast<cudaStream_t>(stream));
}

size_t bpm_find_peaks_workspace_bytes(int64_t total_m, int n_items) { return find_peaks_workspace_bytes(total_m, n_items); }

int bpm_find_peaks(const double* x, int sign, const double* height, const double* prominence, int distance,
                   const BpmItem* items, const BpmItem* items_host, int n_items, int64_t* out_idx,
                   int64_t* out_count, void* workspace, size_t workspace_bytes, void* stream) {
  if (!workspace || !items_host || n_items <= 0 || (sign != 1 && sign != -1)) return BPM_ERR_ARG;
  Workspace ws(workspace, workspace_bytes);
  return find_peaks_run(x, sign, height, prominence, distance, items, batch_shape(items_host, n_items), out_idx,
                        out_count, ws, static_cast<cudaStream_t>(stream));
}

size_t bpm_rolling_floor_workspace_bytes(int64_t total_m, int n_items) { return rolling_floor_workspace_bytes(total_m, n_items); }

int bpm_rolling_floor(const double* envelope, const int64_t* knots, const int64_t* knot_count, const BpmItem* items,
                      const BpmItem* items_host, int n_items, int window, double q, double* floor_out,
                      void* workspace, size_t workspace_bytes, void* stream) {
  if (!workspace || !items_host || n_items <= 0) return BPM_ERR_ARG;
  Workspace ws(workspace, workspace_bytes);
  return rolling_floor_run(envelope, knots, knot_count, items, batch_shape(items_host, n_items), window, q, nullptr,
                           nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, floor_out, nullptr, ws,
                           static_cast<cudaStream_t>(stream));
}

size_t bpm_noise_floor_workspace_bytes(int64_t total_m, int n_items) { return noise_floor_workspace_bytes(total_m, n_items); }

int bpm_noise_floor(const double* envelope, const BpmItem* items, const BpmItem* items_host, int n_items,
                    int distance, double trough_prom_q, double floor_q, int window, double rejection_multiplier,
                    double* floor_out, int64_t* troughs_out, int64_t* trough_count, int64_t* trough_total,
                    int64_t* floor_mode, void* workspace, size_t workspace_bytes, void* stream) {
  if (!workspace || !items_host || n_items <= 0) return BPM_ERR_ARG;
  Workspace ws(workspace, workspace_bytes);
  return noise_floor_run(envelope, items, batch_shape(items_host, n_items), distance, trough_prom_q, floor_q, window,
                         rejection_multiplier, floor_out, troughs_out, trough_count, trough_total, floor_mode, nullptr,
                         ws, static_cast<cudaStream_t>(stream));
}

size_t bpm_sanitize_troughs_workspace_bytes(int64_t total_m, int n_items) { return sanitize_workspace_bytes(total_m, n_items); }

int bpm_sanitize_troughs(const double* envelope, const double* draft_floor, const int64_t* troughs,
                         const int64_t* trough_count, const BpmItem* items, const BpmItem* items_host, int n_items,
                         double rejection_multiplier, int64_t* kept_out, int64_t* kept_count, void* workspace,
                         size_t workspace_bytes, void* stream) {
  if (!workspace || !items_host || n_items <= 0) return BPM_ERR_ARG;
  Workspace ws(workspace, workspace_bytes);
  return sanitize_run(envelope, draft_floor, troughs, trough_count, 0, items, batch_shape(items_host, n_items),
                      rejection_multiplier, 0, kept_out, kept_count, ws, static_cast<cudaStream_t>(stream));
}

size_t bpm_raw_peaks_workspace_bytes(int64_t total_m, int n_items) { return raw_peaks_workspace_bytes(total_m, n_items); }

int bpm_raw_peaks(const double* envelope, const double* floor_, const BpmItem* items, const BpmItem* items_host,
                  int n_items, int distance, double prom_q, int64_t* peaks_out, int64_t* peak_count,
                  void* workspace, size_t workspace_bytes, void* stream) {
  if (!workspace || !items_host || n_items <= 0) return BPM_ERR_ARG;
  Workspace ws(workspace, workspace_bytes);
  return raw_peaks_run(envelope, floor_, items, batch_shape(items_host, n_items), distance, prom_q, nullptr, peaks_out,
                       peak_count, ws, static_cast<cudaStream_t>(stream));
}

int bpm_peak_metrics(const double* envelope, const double* floor_, const int64_t* peaks, const int64_t* peak_count,
                     const BpmItem* items, const BpmItem* items_host, int n_items, double smoothing_factor,
                     double* strength, double* deviation, double* smoothed, void* stream) {
  if (!items_host || n_items <= 0) return BPM_ERR_ARG;
  return peak_metrics_run(envelope, floor_, peaks, peak_count, items, batch_shape(items_host, n_items),
                          smoothing_factor, strength, deviation, smoothed, static_cast<cudaStream_t>(stream));
}

int bpm_peak_trough_noise(const double* envelope, const double* floor_, const int64_t* peaks,
                          const int64_t* peak_count, const int64_t* troughs, const int64_t* trough_count,
                          const BpmItem* items, const BpmItem* items_host, int n_items, double trough_noise_multiplier,
                          double trough_veto_multiplier, double* prev_amp, double* next_amp, double* ratio,
                          unsigned char* flags, void* stream) {
  if (!items_host || n_items <= 0) return BPM_ERR_ARG;
  return peak_trough_noise_run(envelope, floor_, peaks, peak_count, troughs, trough_count, items,
                               batch_shape(items_host, n_items), trough_noise_multiplier, trough_veto_multiplier,
                               prev_amp, next_amp, ratio, flags, static_cast<cudaStream_t>(stream));
}

int bpm_bpm_series(const int64_t* beats, const BpmItem* lists, const BpmItem* lists_host, int n_lists, int rate,
                   int64_t window_us, double* inst, double* smoothed, double* times_sec, int64_t* stamp_us,
                   int64_t* n_valid, void* stream) {
  if (!lists_host || n_lists <= 0 || window_us <= 0) return BPM_ERR_ARG;
  return bpm_series_run(beats, lists, batch_shape(lists_host, n_lists), rate, window_us, inst, smoothed, times_sec,
                        stamp_us, n_valid, static_cast<cudaStream_t>(stream));
}

size_t bpm_steepest_slope_workspace_bytes(int64_t total_beats, int n_lists) {
  (void)total_beats; (void)n_lists;
  return 256;
}

int bpm_steepest_slope(const double* smoothed, const int64_t* stamp_us, const int64_t* n_valid, const BpmItem* lists,
                       const BpmItem* lists_host, int n_lists, int sign, double window_sec, double* result,
                       void* workspace, size_t workspace_bytes, void* stream) {
  (void)workspace; (void)workspace_bytes;
  if (!lists_host || n_lists <= 0) return BPM_ERR_ARG;
  return steepest_run(smoothed, stamp_us, n_valid, lists, n_lists, batch_shape(lists_host, n_lists).max_m, sign,
                      window_sec, result, static_cast<cudaStream_t>(stream));
}

int bpm_windowed_hrv(const int64_t* beats, const BpmItem* lists, const BpmItem* lists_host, int n_lists, int rate,
                     int window_beats, int step_beats, double* out, int64_t* rows, void* stream) {
  if (!lists_host || n_lists <= 0) return BPM_ERR_ARG;
  return hrv_run(beats, lists, batch_shape(lists_host, n_lists), rate, window_beats, step_beats, out, rows,
                 static_cast<cudaStream_t>(stream));
}

int bpm_find_peaks_chunk(const double* x, int sign, const double* height, const double* prominence, int distance,
                         const BpmItem* items, const BpmItem* items_host, int64_t core_lo, int64_t core_hi,
                         int open_left, int open_right, int64_t* out_idx, int64_t* out_count, uint64_t* edge_hits,
                         int64_t* anchors, void* workspace, size_t workspace_bytes, void* stream) {
  if (!workspace || !items_host || !edge_hits || !anchors || (sign != 1 && sign != -1)) return BPM_ERR_ARG;
  Workspace ws(workspace, workspace_bytes);
  const ChunkInfo ci{core_lo, core_hi, open_left, open_right, reinterpret_cast<unsigned long long*>(edge_hits),
                     reinterpret_cast<long long*>(anchors)};
  return find_peaks_run(x, sign, height, prominence, distance, items, batch_shape(items_host, 1), out_idx, out_count,
                        ws, static_cast<cudaStream_t>(stream), nullptr, &ci);
}

size_t bpm_noise_floor_chunk_workspace_bytes(int64_t m) {
  return noise_floor_workspace_bytes(m, 1) + sizeof(double) * static_cast<size_t>(m) + 4096;
}

int bpm_noise_floor_chunk(const double* envelope, const BpmItem* items, const BpmItem* items_host, int distance,
                          const double* trough_prominence, double floor_q, int window, double rejection_multiplier,
                          int64_t core_lo, int64_t core_hi, int open_left, int open_right, double* floor_out,
                          int64_t* troughs_out, int64_t* trough_count, int64_t* all_troughs_out, int64_t* all_count,
                          uint64_t* edge_hits, int64_t* anchors, void* workspace, size_t workspace_bytes, void* stream) {
  if (!workspace || !items_host || !edge_hits || !anchors) return BPM_ERR_ARG;
  Workspace ws(workspace, workspace_bytes);
  const ChunkInfo ci{core_lo, core_hi, open_left, open_right, reinterpret_cast<unsigned long long*>(edge_hits),
                     reinterpret_cast<long long*>(anchors)};
  return noise_floor_chunk_run(envelope, items, batch_shape(items_host, 1), distance, trough_prominence, floor_q, window,
                               rejection_multiplier, ci, floor_out, troughs_out, trough_count, all_troughs_out, all_count,
                               ws, static_cast<cudaStream_t>(stream));
}

int bpm_deviation_series(const double* strength, const int64_t* peak_count, const BpmItem* items,
                         const BpmItem* items_host, int n_items, double smoothing_factor, double* deviation,
                         double* smoothed, void* stream) {
  if (!items_host || n_items <= 0) return BPM_ERR_ARG;
  return deviation_series_run(strength, peak_count, items, batch_shape(items_host, n_items), smoothing_factor,
                              deviation, smoothed, static_cast<cudaStream_t>(stream), true);
}

int bpm_peak_strength(const double* envelope, const double* floor_, const int64_t* peaks, const int64_t* peak_count,
                      const BpmItem* items, const BpmItem* items_host, int n_items, double* strength, void* stream) {
  if (!items_host || n_items <= 0) return BPM_ERR_ARG;
  return peak_strength_run(envelope, floor_, peaks, peak_count, items, batch_shape(items_host, n_items), strength,
                           static_cast<cudaStream_t>(stream));
}

int bpm_cast_f32(const double* src, float* dst, int64_t n, void* stream) {
  return cast_f32_run(src, dst, n, static_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------ a1..a4 in one call
size_t bpm_stage_a_workspace_bytes(int64_t total_m, int n_items) {
  size_t a = frontend_workspace_bytes(total_m, n_items);
  size_t b = noise_floor_workspace_bytes(total_m, n_items);
  size_t c = raw_peaks_workspace_bytes(total_m, n_items);
  size_t mx = a > b ? a : b;
  mx = mx > c ? mx : c;
  return mx + 256 * 4 + sizeof(double) * static_cast<size_t>(n_items);
}

int bpm_stage_a(const void* pcm, const BpmItem* items, const BpmItem* items_host, int n_items, const double* design,
                const double* design_host, int64_t design_words, const BpmStageAConfig* cfg,
                const BpmStageAOutputs* out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!workspace || !items_host || !cfg || !out || n_items <= 0) return BPM_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const BatchShape sh = batch_shape(items_host, n_items);
  Workspace top(workspace, workspace_bytes);
  double* q_tp = top.take<double>(n_items);
  if (top.overflow) return BPM_ERR_WORKSPACE;
  {
    Workspace ws(top.base + top.used, top.cap - top.used);
    BPM_TRY(frontend_run(pcm, cfg->pcm_dtype, cfg->channels, items, items_host, n_items, cfg->stride, design,
                         design_host, design_words, static_cast<int>(cfg->block), cfg->env_window, out->filtered,
                         out->envelope, out->absmax, ws, st));
  }
  if (cfg->want_debug_wav && out->debug_wav && out->filtered)
    BPM_TRY(debug_wav_run(out->filtered, out->absmax, items, items_host, n_items, out->debug_wav, st));
  {
    Workspace ws(top.base + top.used, top.cap - top.used);
    BPM_TRY(noise_floor_run(out->envelope, items, sh, cfg->distance, cfg->trough_prom_q, cfg->floor_q,
                            cfg->noise_window, cfg->rejection_multiplier, out->floor, out->troughs,
                            out->trough_count, out->trough_total, out->floor_mode, q_tp, ws, st));
  }
  {
    Workspace ws(top.base + top.used, top.cap - top.used);
    const double* ready = (cfg->peak_prom_q == cfg->trough_prom_q) ? q_tp : nullptr;   // same np.quantile call
    BPM_TRY(raw_peaks_run(out->envelope, out->floor, items, sh, cfg->distance, cfg->peak_prom_q, ready, out->peaks,
                          out->peak_count, ws, st));
  }
  BPM_TRY(peak_metrics_run(out->envelope, out->floor, out->peaks, out->peak_count, items, sh, cfg->smoothing_factor,
                           out->strength, out->deviation, out->smoothed_dev, st));
  return BPM_OK;
}

}  // extern "C"
