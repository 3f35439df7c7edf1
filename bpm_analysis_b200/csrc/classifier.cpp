// Sequential S1/S2 classifier of the reference as compiled host code (include/bpm_host.h).
//
// Follows PeakClassifier.classify_peaks and its helpers (bpm_analysis.py:113-329, :1120-1255):
// same decisions, same candidate list, same long-term BPM trace and byte-identical per-peak
// debug strings (the report writers parse them).  Every float expression is evaluated in the
// reference's operation order in IEEE double (build with -ffp-contract=off); Python's
// max()/min() argument-order behaviour with NaN, np.clip, np.interp (scalar path of numpy's
// arr_interp) and pandas' Series.asof are restated below.
//
// The reference tests membership of "S1 (Paired)", "Lone S1" and "Noise" in the debug string
// of a candidate beat / of the following raw peak (:140, :151, :161, :185).  Those substrings
// occur only in the type label that starts the string (the reason texts never contain them,
// and candidate beats are never "Noise"), so the tests are made on the stored type code.
#include "../../include/bpm_host.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

namespace {

const char SEP[] = "\xC2\xA7";   // the section sign the reference joins debug fields with

// ---- Python / numpy scalar semantics ---------------------------------------------------
inline double py_max(double a, double b) { return (b > a) ? b : a; }   // max(a, b): a unless b > a
inline double py_min(double a, double b) { return (b < a) ? b : a; }   // min(a, b): a unless b < a

inline double np_clip(double x, double lo, double hi) {                 // NaN propagates
  if (std::isnan(x)) return x;
  const double r = (x < lo) ? lo : x;
  return (r > hi) ? hi : r;
}

// np.interp(x, xp, fp) for one x and a short table: the branch of numpy's arr_interp that
// computes the slope on the fly (len(xp) > len(x), so no slopes are precomputed)
double np_interp(double x, const double* xp, const double* fp, int n) {
  if (std::isnan(x)) return x;
  if (x > xp[n - 1]) return fp[n - 1];
  if (x < xp[0]) return fp[0];
  int j = 0;
  while (j < n - 1 && x >= xp[j + 1]) ++j;                              // xp[j] <= x < xp[j+1]
  if (j == n - 1) return fp[j];
  if (xp[j] == x) return fp[j];
  const double slope = (fp[j + 1] - fp[j]) / (xp[j + 1] - xp[j]);
  double r = slope * (x - xp[j]) + fp[j];
  if (std::isnan(r)) {
    r = slope * (x - xp[j + 1]) + fp[j + 1];
    if (std::isnan(r) && fp[j] == fp[j + 1]) r = fp[j];
  }
  return r;
}

// Series.asof(t) on a float index: last row with index <= t, stepping back over NaN values
// (pandas core/generic.py, scalar branch); NaN before the first index
double series_asof(const double* idx, const double* val, int64_t n, double t) {
  if (n <= 0 || t < idx[0]) return std::nan("");
  int64_t lo = 0, hi = n;                                               // searchsorted(side="right")
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (idx[mid] <= t) lo = mid + 1; else hi = mid;
  }
  int64_t loc = lo > 0 ? lo - 1 : 0;
  while (loc > 0 && std::isnan(val[loc])) --loc;
  return val[loc];
}

// format(v, f".{prec}f") and format(v, ".0%").
// Fast path: x = |v| * 10^prec is formed in double (relative error <= 2^-53: below 1.3e-4
// absolute while x < 2^40) and rounded only when its fractional part is at least 1e-3 away
// from one half, where the correctly rounded decimal -- what printf and Python print -- is
// decided without ambiguity; the integer is then exact.  Anything else (ties and near-ties, huge
// values, inf, larger precisions) goes through snprintf.
void put_f(std::string& s, double v, int prec) {
  if (std::isnan(v)) { s += "nan"; return; }
  static const double p10[4] = {1.0, 10.0, 100.0, 1000.0};
  const double a = std::fabs(v);
  if (prec >= 0 && prec <= 3 && a < 1099511627776.0) {                  // 2^40
    const double x = a * p10[prec];
    const double fl = std::floor(x);
    const double frac = x - fl;
    if (x < 1099511627776.0 && std::fabs(frac - 0.5) > 1e-3) {
      unsigned long long r = static_cast<unsigned long long>(fl) + (frac > 0.5 ? 1ull : 0ull);
      char buf[32];
      int pos = 32;
      for (int d = 0; d < prec; ++d) { buf[--pos] = static_cast<char>('0' + r % 10); r /= 10; }
      if (prec > 0) buf[--pos] = '.';
      do { buf[--pos] = static_cast<char>('0' + r % 10); r /= 10; } while (r);
      if (std::signbit(v)) buf[--pos] = '-';
      s.append(buf + pos, static_cast<size_t>(32 - pos));
      return;
    }
  }
  char buf[400];
  const int k = std::snprintf(buf, sizeof buf, "%.*f", prec, v);
  if (k > 0) s.append(buf, static_cast<size_t>(k) < sizeof buf ? static_cast<size_t>(k) : sizeof buf - 1);
}
void put_pct0(std::string& s, double v) { put_f(s, v * 100.0, 0); s += '%'; }

// Debug text that is only assembled when the caller wants it (BPM_CLASSIFY_NO_TEXT skips every
// append and every number formatting; the decisions never read the text).
struct Text {
  std::string s;
  bool on = true;
  explicit Text(bool enabled) : on(enabled) {}
  Text& operator=(const char* c) { if (on) s = c; return *this; }
  Text& operator+=(const char* c) { if (on) s += c; return *this; }
  Text& operator+=(const std::string& c) { if (on) s += c; return *this; }
  Text& operator+=(const Text& c) { if (on) s += c.s; return *this; }
};
void put_f(Text& t, double v, int prec) { if (t.on) put_f(t.s, v, prec); }
void put_pct0(Text& t, double v) { if (t.on) put_pct0(t.s, v); }

struct Classifier {
  const double* env; const double* floor; const int64_t* peaks; int64_t n;
  const double* dev_t; const double* dev_v; int64_t n_dev;
  double sr; const BpmClassifierParams& p;

  std::vector<int64_t> cand;             // positions into peaks
  std::vector<int32_t> type;
  std::vector<std::string> text;
  std::vector<double> hist_t, hist_v;
  std::vector<BpmClassifierEvent> events;
  double lt_bpm; int64_t consecutive = 0; int64_t loop_idx = 0;
  bool want_text;

  Classifier(const double* e, const double* f, const int64_t* pk, int64_t n_, const double* dt, const double* dv,
             int64_t nd, double sr_, const BpmClassifierParams& pr)
      : env(e), floor(f), peaks(pk), n(n_), dev_t(dt), dev_v(dv), n_dev(nd), sr(sr_), p(pr),
        type(static_cast<size_t>(n_), BPM_PEAK_UNSET), text(static_cast<size_t>(n_)), lt_bpm(pr.start_bpm),
        want_text((pr.flags & BPM_CLASSIFY_NO_TEXT) == 0) {}

  double strength(int64_t pos) const {                                  // max(0, env[i] - floor.iloc[i])
    const int64_t i = peaks[pos];
    return py_max(0.0, env[i] - floor[i]);
  }

  // :135-141 and :180-186
  double pairing_ratio() const {
    const int64_t hw = p.stability_history_window;
    if (static_cast<int64_t>(cand.size()) < hw) return 0.5;
    // candidate_beats[-hw:] -- for hw <= 0 Python's slice is the whole list / the division fails;
    // the caller rejects hw < 1
    int64_t paired = 0;
    for (size_t k = cand.size() - static_cast<size_t>(hw); k < cand.size(); ++k)
      paired += (type[cand[k]] == BPM_PEAK_S1_PAIRED);
    return static_cast<double>(paired) / static_cast<double>(hw);
  }

  static bool is_lone(int32_t t) {
    return t == BPM_PEAK_LONE_S1 || t == BPM_PEAK_LONE_S1_CASCADE || t == BPM_PEAK_LONE_S1_LAST;
  }

  // :133-168 -- only observable effect: the log line (state['pairing_ratio_override'] is never read)
  void kickstart_check() {
    if (pairing_ratio() >= p.kickstart_check_threshold) return;
    const size_t history = 4, min_s1s = 3; const int min_matches = 3;
    if (cand.size() < history) return;
    int64_t lone[4]; size_t n_lone = 0;
    for (size_t k = cand.size() - history; k < cand.size(); ++k)
      if (is_lone(type[cand[k]])) lone[n_lone++] = cand[k];
    if (n_lone < min_s1s) return;
    int matches = 0;
    for (size_t k = 0; k < n_lone; ++k)
      if (lone[k] < n - 1 && type[lone[k] + 1] == BPM_PEAK_NOISE) ++matches;
    if (matches >= min_matches) events.push_back({BPM_EVENT_KICKSTART, matches, static_cast<int32_t>(n_lone), 0});
  }

  // :1120-1144
  double blended_confidence(double deviation, double bpm) const {
    static const double dev_pts[5] = {0.0, 0.25, 0.40, 0.80, 1.0};
    static const double c_low[5] = {0.9, 0.9, 0.7, 0.1, 0.1};
    static const double c_high[5] = {0.1, 0.5, 0.75, 0.65, 0.0};
    const double blend = np_clip((bpm - p.contractility_bpm_low) / (p.contractility_bpm_high - p.contractility_bpm_low),
                                 0.0, 1.0);
    double live[5];
    for (int i = 0; i < 5; ++i) live[i] = c_low[i] + (c_high[i] - c_low[i]) * blend;
    return np_interp(deviation, dev_pts, live, 5);
  }

  // :1147-1202
  double adjust_confidence(double confidence, int64_t s1, int64_t s2, double ratio, Text& reason) const {
    if (static_cast<int64_t>(cand.size()) >= 5) {
      const double xp[2] = {0.0, 1.0}, fp[2] = {p.stability_confidence_floor, p.stability_confidence_ceiling};
      const double factor = np_interp(ratio, xp, fp, 2);
      confidence *= factor;
      reason += "\n- Stability Pre-Adjust: x"; put_f(reason, factor, 2);
      reason += " (Pairing Ratio: "; put_pct0(reason, ratio); reason += ")";
    }
    const double s1_strength = strength(s1), s2_strength = strength(s2);
    const double cur_ratio = s2_strength / (s1_strength + 1e-9);
    const double t_s1 = static_cast<double>(peaks[s1]) / sr;
    const bool in_recovery = p.has_recovery_window && p.peak_bpm_time_sec < t_s1 && t_s1 < p.recovery_end_time_sec;
    const double eff_bpm = in_recovery ? py_max(lt_bpm, p.contractility_bpm_low) : lt_bpm;
    const double xp[2] = {p.contractility_bpm_low, p.contractility_bpm_high};
    const double fp[2] = {p.s2_s1_ratio_low_bpm, p.s2_s1_ratio_high_bpm};
    const double max_expected = np_interp(eff_bpm, xp, fp, 2);
    if (cur_ratio > max_expected) {
      const double severity = cur_ratio / max_expected;
      const double scale = np_clip((severity - 1.0) / 2.0, 0.0, 1.0);
      const double range = p.penalty_amount_max - p.penalty_amount_min;
      const double amount = p.penalty_amount_min + (scale * range);
      confidence -= amount;
      reason += "\n- PENALIZED by "; put_f(reason, amount, 2);
      reason += " (S2 Str. Ratio "; put_f(reason, cur_ratio, 1);
      reason += "x > Expected "; put_f(reason, max_expected, 1); reason += "x)";
    } else if (s1_strength > (s2_strength * p.s1_s2_boost_ratio)) {
      const double actual = s1_strength / (s2_strength + 1e-9);
      const double scale = np_clip((actual - p.s1_s2_boost_ratio) / (4.0 - p.s1_s2_boost_ratio), 0.0, 1.0);
      const double range = p.boost_amount_max - p.boost_amount_min;
      const double amount = p.boost_amount_min + (scale * range);
      confidence += amount;
      reason += "\n- BOOSTED by "; put_f(reason, amount, 2);
      reason += " (S1 Str. Ratio "; put_f(reason, actual, 1); reason += "x > S2)";
    }
    return py_max(0.0, py_min(1.0, confidence));
  }

  // :231-269
  bool attempt_pairing(int64_t s1, int64_t s2, double ratio, Text& reason) const {
    const double interval = static_cast<double>(peaks[s2] - peaks[s1]) / sr;
    const double deviation = series_asof(dev_t, dev_v, n_dev, static_cast<double>(peaks[s1]) / sr);
    double confidence = blended_confidence(deviation, lt_bpm);
    const double blend = np_clip((lt_bpm - p.contractility_bpm_low) / (p.contractility_bpm_high - p.contractility_bpm_low),
                                 0.0, 1.0);
    reason = "Base Conf (Blended Model "; put_pct0(reason, blend); reason += " High): "; put_f(reason, confidence, 2);
    confidence = adjust_confidence(confidence, s1, s2, ratio, reason);
    const double max_interval = py_min(p.s1_s2_interval_cap_sec, (60.0 / lt_bpm) * p.s1_s2_interval_rr_fraction);
    if (p.enable_interval_penalty && interval > max_interval) {
      const double zone_start = max_interval * p.interval_penalty_start_factor;
      const double zone_end = max_interval * p.interval_penalty_full_factor;
      if (interval > zone_start) {
        double scale = (interval - zone_start) / (zone_end - zone_start + 1e-9);
        scale = np_clip(scale, 0.0, 1.0);
        const double amount = scale * p.interval_max_penalty;
        confidence = py_max(0.0, confidence - amount);
        reason += "\n- Interval PENALTY by "; put_f(reason, amount, 2);
        reason += " (Interval "; put_f(reason, interval, 3);
        reason += "s > Max "; put_f(reason, max_interval, 3); reason += "s)";
      }
    }
    const bool paired = confidence >= p.pairing_confidence_threshold;
    reason += "\n- Final Score: "; put_f(reason, confidence, 2);
    reason += " vs Threshold "; put_f(reason, p.pairing_confidence_threshold, 2);
    reason += paired ? " -> Paired" : " -> Not Paired";
    return paired;
  }

  // :1206-1242
  double lone_s1_confidence(int64_t cur, int64_t last, Text& reason) const {
    const double expected_rr = 60.0 / lt_bpm;
    const double actual_rr = static_cast<double>(peaks[cur] - peaks[last]) / sr;
    const double dev_pct = std::fabs(actual_rr - expected_rr) / expected_rr;
    static const double rx[4] = {0.0, 0.15, 0.30, 0.50}, ry[4] = {1.0, 0.8, 0.4, 0.0};
    const double rhythm = np_interp(dev_pct, rx, ry, 4);
    const double last_strength = strength(last), cur_strength = strength(cur);
    const double amp_ratio = cur_strength / (last_strength + 1e-9);
    static const double ax[4] = {0.0, 0.4, 0.7, 1.0}, ay[4] = {0.0, 0.4, 0.8, 1.0};
    const double amplitude = np_interp(amp_ratio, ax, ay, 4);
    reason = "Rhythm Fit="; put_f(reason, rhythm, 2);
    reason += " (Interval "; put_f(reason, actual_rr, 3);
    reason += "s vs Expected "; put_f(reason, expected_rr, 3);
    reason += "s), Amplitude Fit="; put_f(reason, amplitude, 2);
    reason += " (Strength Ratio "; put_f(reason, amp_ratio, 2); reason += "x)";
    return (rhythm * p.lone_s1_rhythm_weight) + (amplitude * p.lone_s1_amplitude_weight);
  }

  // :304-329.  0 = valid, 1 = rejected on confidence (the detail then contains "Rhythm Fit", which is
  // what the caller's substring test at :286 looks for), 2 = rejected by the forward check
  int validate_lone_s1(int64_t cur, Text& detail) const {
    if (cand.empty()) { detail = "First beat"; return 0; }
    Text reason(want_text);
    const double confidence = lone_s1_confidence(cur, cand.back(), reason);
    const double thr = p.lone_s1_confidence_threshold;
    if (confidence < thr) {
      detail = "Rejected Lone S1: Confidence "; put_f(detail, confidence, 2);
      detail += " < Threshold "; put_f(detail, thr, 2); detail += ". ("; detail += reason; detail += ")";
      return 1;
    }
    if (cur < n - 1) {
      const int64_t i = peaks[cur], nx = peaks[cur + 1];
      const double forward = static_cast<double>(nx - i) / sr;
      const double expected_rr = 60.0 / lt_bpm;
      const double min_forward = expected_rr * p.lone_s1_forward_check_pct;
      if (forward < min_forward && !(env[i] > (env[nx] * 1.7))) {
        const double implied = forward > 0 ? 60.0 / forward : HUGE_VAL;
        detail = "Rejected Lone S1: Forward check failed (Implies "; put_f(detail, implied, 0); detail += " BPM)";
        return 2;
      }
    }
    detail = "Validated Lone S1: Confidence "; put_f(detail, confidence, 3);
    detail += " >= Threshold "; put_f(detail, thr, 2); detail += ". ("; detail += reason;
    detail += ", Weights: Rhythm="; put_f(detail, p.lone_s1_rhythm_weight, 2);
    detail += ", Amplitude="; put_f(detail, p.lone_s1_amplitude_weight, 2);
    detail += ", Final="; put_f(detail, confidence, 3); detail += ")";
    return 0;
  }

  // :271-302
  void classify_lone_peak(int64_t cur, const Text& fail_reason) {
    Text detail(want_text);
    const int verdict = validate_lone_s1(cur, detail);
    std::string info;
    if (want_text) {
      size_t skip = 0;                                                  // reason.lstrip(' |')
      const std::string& fr = fail_reason.s;
      while (skip < fr.size() && (fr[skip] == ' ' || fr[skip] == '|')) ++skip;
      info = std::string("PAIRING_FAIL_REASON") + SEP + fr.substr(skip);
    }
    std::string& out = text[cur];
    if (verdict == 0) {
      cand.push_back(cur);
      type[cur] = BPM_PEAK_LONE_S1;
      if (want_text) out = std::string("Lone S1") + SEP + info + SEP + "LONE_S1_VALIDATE_REASON" + SEP + detail.s;
      consecutive = 0;
      return;
    }
    if (verdict == 1) consecutive += 1; else consecutive = 0;           // "Rhythm Fit" in rejection_detail
    const std::string reject = want_text ? std::string("LONE_S1_REJECT_REASON") + SEP + detail.s : std::string();
    if (static_cast<double>(consecutive) >= p.cascade_reset_trigger_count) {
      events.push_back({BPM_EVENT_CASCADE_RESET, static_cast<int32_t>(cur), 0, 0});
      cand.push_back(cur);
      type[cur] = BPM_PEAK_LONE_S1_CASCADE;
      if (want_text) out = std::string("Lone S1 (Corrected by Cascade Reset)") + SEP + info + SEP + reject;
      consecutive = 0;
    } else {
      type[cur] = BPM_PEAK_NOISE;
      if (want_text) out = std::string("Noise") + SEP + info + SEP + reject;
    }
  }

  // :175-200
  void process_peak_pair(int64_t cur) {
    const int64_t next = cur + 1;
    const double ratio = pairing_ratio();
    Text reason(want_text);
    if (attempt_pairing(cur, next, ratio, reason)) {
      cand.push_back(cur);
      type[cur] = BPM_PEAK_S1_PAIRED;
      type[next] = BPM_PEAK_S2_PAIRED;
      if (want_text) {
        const std::string tag = std::string(SEP) + "PAIRING_SUCCESS_REASON" + SEP + reason.s;
        text[cur] = "S1 (Paired)" + tag;
        text[next] = "S2 (Paired)" + tag;
      }
      consecutive = 0;
      loop_idx += 2;
    } else {
      classify_lone_peak(cur, reason);
      loop_idx += 1;
    }
  }

  // :202-212 and update_long_term_bpm :1244-1258.  Runs after EVERY decision, so a rejected
  // peak re-applies the last accepted R-R interval, as in the reference.
  void update_long_term_bpm() {
    if (cand.size() > 1) {
      const double new_rr = static_cast<double>(peaks[cand[cand.size() - 1]] - peaks[cand[cand.size() - 2]]) / sr;
      if (new_rr > 0) {
        const double instant = 60.0 / new_rr;
        const double lr = 0.05, max_change_per_beat = 3.0;
        const double target = ((1 - lr) * lt_bpm) + (lr * instant);
        const double max_change = max_change_per_beat * new_rr;
        const double proposed = target - lt_bpm;
        const double limited = np_clip(proposed, -max_change, max_change);
        const double new_bpm = lt_bpm + limited;
        lt_bpm = py_max(p.min_bpm, py_min(new_bpm, p.max_bpm));
      }
    }
    if (!cand.empty()) {
      hist_t.push_back(static_cast<double>(peaks[cand.back()]) / sr);
      hist_v.push_back(lt_bpm);
    }
  }

  void run() {
    while (loop_idx < n) {
      kickstart_check();
      const int64_t cur = loop_idx;
      if (loop_idx >= n - 1) {                                          // :170-174
        cand.push_back(cur);
        type[cur] = BPM_PEAK_LONE_S1_LAST;
        if (want_text) text[cur] = "Lone S1 (Last Peak)";
        loop_idx += 1;
      } else {
        process_peak_pair(cur);
      }
      update_long_term_bpm();
    }
  }
};

struct Owned {
  BpmClassification pub;
  std::vector<int64_t> beats, offsets;
  std::vector<int32_t> types;
  std::string text;
  std::vector<double> hist_t, hist_v;
  std::vector<BpmClassifierEvent> events;
};

}  // namespace

extern "C" {

int bpm_host_abi_version(void) { return BPM_HOST_ABI_VERSION; }

int bpm_host_format_fixed(double v, int prec, char* out, size_t capacity) {
  if (!out || capacity == 0 || prec < 0 || prec > 17) return BPM_HOST_ERR_ARG;
  std::string s;
  put_f(s, v, prec);
  if (s.size() + 1 > capacity) return BPM_HOST_ERR_ARG;
  std::memcpy(out, s.c_str(), s.size() + 1);
  return static_cast<int>(s.size());
}

int bpm_classify_peaks(const double* envelope, const double* noise_floor, int64_t m, const int64_t* raw_peaks,
                       int64_t n_peaks, const double* dev_times, const double* dev_values, int64_t n_dev,
                       double sample_rate, const BpmClassifierParams* params, BpmClassification** out) {
  if (!envelope || !noise_floor || !raw_peaks || !params || !out || m <= 0 || n_peaks < 2 || n_dev < 0 ||
      (n_dev > 0 && (!dev_times || !dev_values)) || !(sample_rate > 0) || params->stability_history_window < 1 ||
      n_peaks > INT32_MAX)
    return BPM_HOST_ERR_ARG;
  for (int64_t i = 0; i < n_peaks; ++i)
    if (raw_peaks[i] < 0 || raw_peaks[i] >= m || (i > 0 && raw_peaks[i] <= raw_peaks[i - 1])) return BPM_HOST_ERR_ARG;
  *out = nullptr;
  try {
    Classifier c(envelope, noise_floor, raw_peaks, n_peaks, dev_times, dev_values, n_dev, sample_rate, *params);
    c.run();
    Owned* o = new Owned();
    o->beats = std::move(c.cand);
    o->types = std::move(c.type);
    o->hist_t = std::move(c.hist_t);
    o->hist_v = std::move(c.hist_v);
    o->events = std::move(c.events);
    o->offsets.resize(static_cast<size_t>(n_peaks) + 1);
    size_t total = 0;
    for (int64_t i = 0; i < n_peaks; ++i) total += c.text[i].size() + 1;
    o->text.reserve(total);
    for (int64_t i = 0; i < n_peaks; ++i) {                            // NUL-terminated entries, back to back
      o->offsets[i] = static_cast<int64_t>(o->text.size());
      o->text += c.text[i];
      o->text += '\0';
    }
    o->offsets[n_peaks] = static_cast<int64_t>(o->text.size());
    BpmClassification& r = o->pub;
    r.n_peaks = n_peaks;
    r.n_beats = static_cast<int64_t>(o->beats.size());
    r.n_history = static_cast<int64_t>(o->hist_t.size());
    r.n_events = static_cast<int64_t>(o->events.size());
    r.text_bytes = static_cast<int64_t>(o->text.size());
    r.beat_positions = o->beats.data();
    r.peak_types = o->types.data();
    r.text_offsets = o->offsets.data();
    r.text = o->text.data();
    r.history_times = o->hist_t.data();
    r.history_bpm = o->hist_v.data();
    r.events = o->events.data();
    r.final_long_term_bpm = c.lt_bpm;
    r.final_consecutive_rr_rejections = c.consecutive;
    *out = &o->pub;
    return BPM_HOST_OK;
  } catch (const std::bad_alloc&) {
    return BPM_HOST_ERR_NOMEM;
  } catch (...) {
    return BPM_HOST_ERR_ARG;
  }
}

int bpm_classify_peaks_batch(const BpmClassifyJob* jobs, int64_t n_jobs, int n_threads, BpmClassification** out,
                             int* status) {
  if (!jobs || !out || !status || n_jobs < 0) return BPM_HOST_ERR_ARG;
  if (n_jobs == 0) return BPM_HOST_OK;
  int64_t workers = n_threads > 0 ? n_threads : static_cast<int64_t>(std::thread::hardware_concurrency());
  if (workers < 1) workers = 1;
  if (workers > n_jobs) workers = n_jobs;
  std::atomic<int64_t> next{0};
  auto work = [&]() {
    for (;;) {
      const int64_t i = next.fetch_add(1);
      if (i >= n_jobs) return;
      const BpmClassifyJob& j = jobs[i];
      out[i] = nullptr;
      status[i] = bpm_classify_peaks(j.envelope, j.noise_floor, j.m, j.raw_peaks, j.n_peaks, j.dev_times, j.dev_values,
                                     j.n_dev, j.sample_rate, j.params, &out[i]);
    }
  };
  try {
    std::vector<std::thread> pool;
    for (int64_t t = 1; t < workers; ++t) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
  } catch (...) {
    return BPM_HOST_ERR_NOMEM;
  }
  return BPM_HOST_OK;
}

void bpm_classification_free(BpmClassification* c) {
  if (c) delete reinterpret_cast<Owned*>(c);     // pub is the first member of Owned
}

}  // extern "C"
