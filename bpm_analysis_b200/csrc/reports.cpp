// libbpm_host.so -- text rows of the on-disk outputs.
//
// The reference writes its per-beat tables one Python f-string at a time: `_bpm_plot.csv`
// (bpm_analysis.py:458-473, csv.writer rows "t:.3f,bpm:.3f" terminated by "\r\n") and the
// "Heartbeat Data" table of `_Analysis_Summary.md` (:973-983, "| t:.2f | bpm:.1f |\n").  A 24-hour
// recording has ~1e5 beats; here a table is one call that formats every row with the same
// correctly rounded fixed-point formatter the classifier's debug strings use
// (bpm_host_format_fixed), skipping rows whose second value is NaN as the reference does.
#include "../../include/bpm_host.h"

#include <cmath>
#include <cstring>
#include <string>

extern "C" {

int64_t bpm_host_format_rows(const double* a, const double* b, int64_t n, int prec_a, int prec_b, const char* head,
                             const char* mid, const char* tail, int skip_nan_b, char* out, int64_t capacity) {
  if ((n > 0 && (!a || !b)) || n < 0 || !head || !mid || !tail || prec_a < 0 || prec_a > 17 || prec_b < 0 ||
      prec_b > 17 || capacity < 0 || (capacity > 0 && !out))
    return BPM_HOST_ERR_ARG;
  const size_t lh = std::strlen(head), lm = std::strlen(mid), lt = std::strlen(tail);
  int64_t need = 0;
  char num[400];
  for (int64_t i = 0; i < n; ++i) {
    if (skip_nan_b && std::isnan(b[i])) continue;
    const int ka = bpm_host_format_fixed(a[i], prec_a, num, sizeof num);
    if (ka < 0) return BPM_HOST_ERR_ARG;
    const bool fits_a = need + static_cast<int64_t>(lh) + ka <= capacity;
    if (fits_a) {
      std::memcpy(out + need, head, lh);
      std::memcpy(out + need + lh, num, static_cast<size_t>(ka));
    }
    need += static_cast<int64_t>(lh) + ka;
    const int kb = bpm_host_format_fixed(b[i], prec_b, num, sizeof num);
    if (kb < 0) return BPM_HOST_ERR_ARG;
    if (need + static_cast<int64_t>(lm) + kb + static_cast<int64_t>(lt) <= capacity) {
      std::memcpy(out + need, mid, lm);
      std::memcpy(out + need + lm, num, static_cast<size_t>(kb));
      std::memcpy(out + need + lm + kb, tail, lt);
    }
    need += static_cast<int64_t>(lm) + kb + static_cast<int64_t>(lt);
  }
  return need;   // > capacity: nothing usable was written, call again with this many bytes
}

}  // extern "C"
