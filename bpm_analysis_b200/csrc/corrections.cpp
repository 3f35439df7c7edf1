// The per-beat loops of the reference's correction passes as compiled host code
// (include/bpm_host.h): correct_peaks_by_rhythm (bpm_analysis.py:1257-1306) and
// _fix_rhythmic_discontinuities (:1309-1412).  Same comparisons on the same IEEE doubles in the
// same order; decisions are returned as events for the Python mirror to apply and log.
#include "../../include/bpm_host.h"

#include <algorithm>
#include <new>
#include <vector>

namespace {

bool ascending_in_range(const int64_t* p, int64_t n, int64_t m) {
  for (int64_t i = 0; i < n; ++i)
    if (p[i] < 0 || p[i] >= m || (i > 0 && p[i] <= p[i - 1])) return false;
  return true;
}

struct EventSink {
  BpmCorrectionEvent* ev; int64_t cap; int64_t n = 0; bool overflow = false;
  void put(int32_t kind, int64_t a, int64_t b, double x) {
    if (n >= cap) { overflow = true; return; }
    ev[n++] = BpmCorrectionEvent{kind, 0, a, b, x};
  }
};

}  // namespace

extern "C" {

int bpm_correct_peaks_by_rhythm(const int64_t* peaks, int64_t n, const double* envelope, int64_t m, double sample_rate,
                                double threshold_sec, int64_t* out, int64_t* n_out, BpmCorrectionEvent* events,
                                int64_t* n_events) {
  if (!peaks || !envelope || !out || !n_out || !events || !n_events || n < 1 || m <= 0 || !(sample_rate > 0))
    return BPM_HOST_ERR_ARG;
  for (int64_t i = 0; i < n; ++i)
    if (peaks[i] < 0 || peaks[i] >= m) return BPM_HOST_ERR_ARG;
  EventSink sink{events, n};
  int64_t k = 0;
  out[k++] = peaks[0];                                                   // :1276
  for (int64_t i = 1; i < n; ++i) {
    const int64_t cur = peaks[i], last = out[k - 1];
    const double interval = static_cast<double>(cur - last) / sample_rate;
    if (interval < threshold_sec) {
      if (envelope[cur] > envelope[last]) {                              // the stronger peak wins (:1288)
        sink.put(BPM_CORR_REPLACED, cur, last, 0.0);
        out[k - 1] = cur;
      } else {
        sink.put(BPM_CORR_DISCARDED, cur, 0, 0.0);
      }
    } else {
      out[k++] = cur;
    }
  }
  *n_out = k;
  *n_events = sink.n;
  return BPM_HOST_OK;
}

int bpm_fix_rhythmic_discontinuities(const int64_t* s1, int64_t n_s1, const int64_t* raw, int64_t n_raw,
                                     const uint8_t* raw_is_noise, const double* envelope, const double* noise_floor,
                                     int64_t m, double sample_rate, double short_thr, double long_thr,
                                     double waiver_strength_ratio, double waiver_max_s2_s1_ratio, int64_t* out,
                                     int64_t* n_out, BpmCorrectionEvent* events, int64_t events_capacity,
                                     int64_t* n_events, int64_t* corrections_made) {
  if (!s1 || !raw || !raw_is_noise || !envelope || !noise_floor || !out || !n_out || !events || !n_events ||
      !corrections_made || n_s1 < 1 || n_raw < 1 || m <= 0 || !(sample_rate > 0) || events_capacity < 0)
    return BPM_HOST_ERR_ARG;
  if (!ascending_in_range(s1, n_s1, m) || !ascending_in_range(raw, n_raw, m)) return BPM_HOST_ERR_ARG;
  try {
    const int64_t margin = 3;
    EventSink sink{events, events_capacity};
    int64_t corrections = 0;
    std::vector<int64_t> to_add;                                         // sample indices, in order of discovery
    std::vector<uint8_t> added(static_cast<size_t>(n_raw), 0);           // by raw position

    // ---- pass 1: long intervals (missed beats), :1350-1382
    for (int64_t i = margin; i < n_s1 - 1 - margin; ++i) {
      const int64_t start = s1[i], end = s1[i + 1];
      if (!(static_cast<double>(end - start) / sample_rate > long_thr)) continue;
      sink.put(BPM_CORR_LONG_INTERVAL, start, 0, 0.0);
      // raw peaks strictly inside (start, end): the reference filters the whole raw list per gap
      const int64_t lo = std::upper_bound(raw, raw + n_raw, start) - raw;
      for (int64_t r = lo; r < n_raw && raw[r] < end; ++r) {
        if (!raw_is_noise[r] || added[r]) continue;
        if (r + 1 >= n_raw) continue;
        const int64_t c1 = raw[r], c2 = raw[r + 1];
        if (c2 >= end || !raw_is_noise[r + 1]) continue;
        const double diff = envelope[c1] - noise_floor[c1];
        const double s1_strength = (diff > 0) ? diff : 0.0;              // max(0, x)
        const bool strong = s1_strength > (waiver_strength_ratio * noise_floor[c1]);
        const bool plausible = (envelope[c2] / (envelope[c1] + 1e-9)) < waiver_max_s2_s1_ratio;
        if (strong && plausible) {
          sink.put(BPM_CORR_RELABELLED, r, r + 1, 0.0);
          corrections += 1;
          added[r] = 1;
          to_add.push_back(c1);
          break;
        }
      }
    }

    // ---- pass 2: short intervals (adjacent S1s), :1385-1407
    std::vector<int64_t> temp(s1, s1 + n_s1);
    temp.insert(temp.end(), to_add.begin(), to_add.end());
    std::sort(temp.begin(), temp.end());
    temp.erase(std::unique(temp.begin(), temp.end()), temp.end());
    const int64_t nt = static_cast<int64_t>(temp.size());
    std::vector<uint8_t> removed(static_cast<size_t>(nt), 0);
    for (int64_t i = margin; i < nt - 1 - margin; ++i) {
      if (removed[i] || removed[i + 1]) continue;
      const int64_t a = temp[i], b = temp[i + 1];
      const double interval = static_cast<double>(b - a) / sample_rate;
      if (interval < short_thr) {
        sink.put(BPM_CORR_SHORT_INTERVAL, a, b, interval);
        if (envelope[b] > envelope[a]) { removed[i] = 1; sink.put(BPM_CORR_REMOVED, a, 0, 0.0); }
        else { removed[i + 1] = 1; sink.put(BPM_CORR_REMOVED, b, 0, 0.0); }
        corrections += 1;
      }
    }
    int64_t k = 0;
    for (int64_t i = 0; i < nt; ++i)
      if (!removed[i]) out[k++] = temp[i];
    if (sink.overflow) return BPM_HOST_ERR_ARG;
    *n_out = k;
    *n_events = sink.n;
    *corrections_made = corrections;
    return BPM_HOST_OK;
  } catch (const std::bad_alloc&) {
    return BPM_HOST_ERR_NOMEM;
  }
}

}  // extern "C"
