// K8 (per-peak strength / deviation series), K9 (BPM series), K10 (steepest slopes),
// K12 (windowed HRV).  Small per-beat kernels; batches of recordings / beat lists run as
// blockIdx.y.  Reference lines: bpm_analysis.py:93-100, :1463-1484, :1552-1595, :1414-1461.
#include "common.cuh"

namespace bpm {

// ------------------------------------------------------------------ K8
// strength[k] = max(0, env[p_k] - floor[p_k])                                   (:93-95)
__global__ void k_peak_strength(const double* __restrict__ env, const double* __restrict__ floor_,
                                const int64_t* __restrict__ peaks, const int64_t* __restrict__ peak_count,
                                const BpmItem* __restrict__ items, double* __restrict__ strength) {
  const BpmItem it = items[blockIdx.y];
  const long long np = peak_count[blockIdx.y];
  for (long long k = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; k < np;
       k += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int64_t p = peaks[it.m_off + k];
    double s = __dsub_rn(env[it.m_off + p], floor_[it.m_off + p]);
    if (s < 0.0) s = 0.0;
    strength[it.m_off + k] = s;
  }
}

// deviation[k] = |s[k+1]-s[k]| / (max(s[k], s[k+1]) + 1e-9)                       (:96)
__global__ void k_peak_deviation(const double* __restrict__ strength, const int64_t* __restrict__ peak_count,
                                 const BpmItem* __restrict__ items, double* __restrict__ dev) {
  const BpmItem it = items[blockIdx.y];
  const long long np = peak_count[blockIdx.y];
  for (long long k = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; k + 1 < np;
       k += static_cast<long long>(gridDim.x) * blockDim.x) {
    const double a = strength[it.m_off + k], b = strength[it.m_off + k + 1];
    const double mx = (isnan(a) || isnan(b)) ? __longlong_as_double(0x7ff8000000000000ll) : fmax(a, b);
    dev[it.m_off + k] = __ddiv_rn(fabs(__dsub_rn(b, a)), __dadd_rn(mx, 1e-9));
  }
}

// centred rolling mean, window max(5, int((P-1)*factor)), min_periods=1          (:99-100)
// The window grows with the recording (5 % of the peak count: 640 values for a 60-min recording, 10^4
// for a 24-h stream).  A warp owns 32 consecutive outputs: its lanes add up the window of the first one
// together (coalesced), and output l is that sum plus the samples that entered minus those that left on
// the way from output 0 to l -- a warp scan of per-step differences (the way pandas' rolling mean
// itself proceeds: add one, remove one).  n / 32 windows are read instead of n.  (Round 1 re-added the
// window per output for short lists, 12 us at C2, and used a one-CTA prefix sum for long ones, 190 us
// at C4; this form takes 5 / 27 us.)
__global__ void __launch_bounds__(128) k_dev_smooth_slide(const double* __restrict__ dev,
                                                          const int64_t* __restrict__ peak_count,
                                                          const BpmItem* __restrict__ items, double factor,
                                                          double* __restrict__ out) {
  const BpmItem it = items[blockIdx.y];
  const long long n = peak_count[blockIdx.y] - 1;
  long long w = static_cast<long long>(__dmul_rn(static_cast<double>(n), factor));
  if (w < 5) w = 5;
  const long long off = (w - 1) / 2;
  const double* d = dev + it.m_off;
  const int lane = threadIdx.x & 31;
  const long long warps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  for (long long r = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); r * 32 < n; r += warps) {
    const long long i0 = r * 32;
    long long a0 = i0 + 1 + off - w, b0 = i0 + off;
    if (a0 < 0) a0 = 0;
    if (b0 > n - 1) b0 = n - 1;
    double s0 = 0.0, s1 = 0.0;
    long long k = a0 + lane;
    for (; k + 32 <= b0; k += 64) { s0 += d[k]; s1 += d[k + 32]; }
    if (k <= b0) s0 += d[k];
    double base = s0 + s1;
#pragma unroll
    for (int o = 16; o; o >>= 1) base += __shfl_xor_sync(0xffffffffu, base, o);
    // lane l: what output i0 + l gains / loses against output i0 + l - 1
    const long long i = i0 + lane;
    long long a = i + 1 + off - w, b = i + off;
    if (a < 0) a = 0;
    if (b > n - 1) b = n - 1;
    long long pa = i + off - w, pb = i - 1 + off;
    if (pa < 0) pa = 0;
    if (pb > n - 1) pb = n - 1;
    double delta = 0.0;
    if (lane > 0 && i < n) {
      if (b > pb) delta += d[b];
      if (a > pa) delta -= d[pa];
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, delta, o);
      if (lane >= o) delta += t;
    }
    if (i < n) out[it.m_off + i] = __ddiv_rn(base + delta, static_cast<double>(b - a + 1));
  }
}

int deviation_series_run(const double* strength, const int64_t* peak_count, const BpmItem* items, const BatchShape& sh,
                         double factor, double* deviation, double* smoothed, cudaStream_t st, bool list_sized = false);

int peak_strength_run(const double* env, const double* floor_, const int64_t* peaks, const int64_t* peak_count,
                      const BpmItem* items, const BatchShape& sh, double* strength, cudaStream_t st) {
  if (!env || !floor_ || !peaks || !peak_count || !items || !strength) return BPM_ERR_ARG;
  int64_t gx = (sh.max_m / 2 + 2 + 255) / 256;
  const int64_t cap = (148 * 4 + sh.n_items - 1) / sh.n_items;
  if (gx > cap) gx = cap;
  BPM_KERNEL(k_peak_strength);
  k_peak_strength<<<dim3(static_cast<unsigned>(gx < 1 ? 1 : gx), sh.n_items), 256, 0, st>>>(env, floor_, peaks, peak_count,
                                                                                          items, strength);
  BPM_LAUNCH_OK();
  return BPM_OK;
}

int peak_metrics_run(const double* env, const double* floor_, const int64_t* peaks, const int64_t* peak_count,
                     const BpmItem* items, const BatchShape& sh, double factor, double* strength,
                     double* deviation, double* smoothed, cudaStream_t st) {
  if (!env || !floor_ || !peaks || !peak_count || !items || !strength || !deviation || !smoothed) return BPM_ERR_ARG;
  // peak counts live on the device: bounded grids, grid-stride loops
  int64_t gx = (sh.max_m / 2 + 2 + 255) / 256;
  const int64_t cap = (148 * 4 + sh.n_items - 1) / sh.n_items;
  if (gx > cap) gx = cap;
  const dim3 grid(static_cast<unsigned>(gx < 1 ? 1 : gx), sh.n_items);
  BPM_KERNEL(k_peak_strength);
  k_peak_strength<<<grid, 256, 0, st>>>(env, floor_, peaks, peak_count, items, strength);
  BPM_LAUNCH_OK();
  return deviation_series_run(strength, peak_count, items, sh, factor, deviation, smoothed, st);
}

// deviation[k] and its centred rolling mean from a given strength list (:96-100).
// list_sized: the descriptors give the LIST lengths (not envelope lengths), i.e. max_m = P.
int deviation_series_run(const double* strength, const int64_t* peak_count, const BpmItem* items, const BatchShape& sh,
                         double factor, double* deviation, double* smoothed, cudaStream_t st, bool list_sized) {
  if (!strength || !peak_count || !items || !deviation || !smoothed) return BPM_ERR_ARG;
  int64_t gx = (sh.max_m / 2 + 2 + 255) / 256;
  const int64_t cap = (148 * 4 + sh.n_items - 1) / sh.n_items;
  if (gx > cap) gx = cap;
  const dim3 grid(static_cast<unsigned>(gx < 1 ? 1 : gx), sh.n_items);
  BPM_KERNEL(k_peak_deviation);
  k_peak_deviation<<<grid, 256, 0, st>>>(strength, peak_count, items, deviation);
  BPM_LAUNCH_OK();
  {
    // peak counts live on the device: the grid is sized for the most peaks the list can hold
    int64_t runs = ((list_sized ? sh.max_m : sh.max_m / 2 + 1) + 31) / 32;      // one warp each, 4 per CTA
    int64_t gs = (runs + 3) / 4;
    const int64_t cap2 = (148 * 8 + sh.n_items - 1) / sh.n_items;
    if (gs > cap2) gs = cap2;
    BPM_KERNEL(k_dev_smooth_slide);
    k_dev_smooth_slide<<<dim3(static_cast<unsigned>(gs < 1 ? 1 : gs), sh.n_items), 128, 0, st>>>(
        deviation, peak_count, items, factor, smoothed);
    BPM_LAUNCH_OK();
  }
  return BPM_OK;
}

// ------------------------------------------------------------------ K8b: surrounding-trough noise
// The north star's "trough-noise" per-peak metric.  No function of bpm_analysis.py computes it any
// more (SURVEY.md §8a note); the definitions below are the ones its documentation and surviving
// config keys give (PARITY UNPINNED, oracle = oracle/ref_port.py::peak_trough_noise):
//   prev / next = amplitude of the sanitised trough just before / after the peak (NaN if none);
//   deeper      = the lower of the two ("analyzes the deeper of the two troughs surrounding a peak",
//                 Documentation/Changelog.md:454);
//   ratio       = deeper / floor[peak];   bit 0: ratio > trough_noise_multiplier (config.py:31,
//                 "BPM Detection logic explained.md":262);
//   bit 1       : look-ahead veto  m * (env[p] - next) < (env[p_next] - next)  with m =
//                 trough_veto_multiplier (config.py:30, "...logic explained.md":276-278), only when
//                 the next trough lies before the next peak.
__global__ void k_peak_trough_noise(const double* __restrict__ env, const double* __restrict__ floor_,
                                    const int64_t* __restrict__ peaks, const int64_t* __restrict__ peak_count,
                                    const int64_t* __restrict__ troughs, const int64_t* __restrict__ trough_count,
                                    const BpmItem* __restrict__ items, double noise_mult, double veto_mult,
                                    double* __restrict__ prev_amp, double* __restrict__ next_amp,
                                    double* __restrict__ ratio, unsigned char* __restrict__ flags) {
  const BpmItem it = items[blockIdx.y];
  const long long np = peak_count[blockIdx.y], nt = trough_count[blockIdx.y];
  const int64_t* tr = troughs + it.m_off;
  const double* e = env + it.m_off;
  const double qnan = __longlong_as_double(0x7ff8000000000000ll);
  for (long long k = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; k < np;
       k += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int64_t p = peaks[it.m_off + k];
    long long lo = 0, hi = nt;                       // first trough index with tr[i] > p
    while (lo < hi) {
      const long long mid = (lo + hi) >> 1;
      if (tr[mid] > p) hi = mid; else lo = mid + 1;
    }
    long long ip = lo - 1;                           // last trough <= p; a trough is never a peak, but be strict
    while (ip >= 0 && tr[ip] >= p) --ip;
    const double pa = (ip >= 0) ? e[tr[ip]] : qnan;
    const double na = (lo < nt) ? e[tr[lo]] : qnan;
    double deeper = qnan;
    if (ip >= 0 && lo < nt) deeper = fmin(pa, na);
    else if (ip >= 0) deeper = pa;
    else if (lo < nt) deeper = na;
    const double r = __ddiv_rn(deeper, floor_[it.m_off + p]);
    unsigned char f = 0;
    if (r > noise_mult) f |= 1;
    if (k + 1 < np && lo < nt) {
      const int64_t pn = peaks[it.m_off + k + 1];
      if (tr[lo] < pn && __dmul_rn(veto_mult, __dsub_rn(e[p], na)) < __dsub_rn(e[pn], na)) f |= 2;
    }
    prev_amp[it.m_off + k] = pa;
    next_amp[it.m_off + k] = na;
    ratio[it.m_off + k] = r;
    flags[it.m_off + k] = f;
  }
}

int peak_trough_noise_run(const double* env, const double* floor_, const int64_t* peaks, const int64_t* peak_count,
                          const int64_t* troughs, const int64_t* trough_count, const BpmItem* items,
                          const BatchShape& sh, double noise_mult, double veto_mult, double* prev_amp,
                          double* next_amp, double* ratio, unsigned char* flags, cudaStream_t st) {
  if (!env || !floor_ || !peaks || !peak_count || !troughs || !trough_count || !items || !prev_amp || !next_amp ||
      !ratio || !flags)
    return BPM_ERR_ARG;
  int64_t gx = (sh.max_m / 2 + 2 + 255) / 256;
  const int64_t cap = (148 * 4 + sh.n_items - 1) / sh.n_items;
  if (gx > cap) gx = cap;
  BPM_KERNEL(k_peak_trough_noise);
  k_peak_trough_noise<<<dim3(static_cast<unsigned>(gx < 1 ? 1 : gx), sh.n_items), 256, 0, st>>>(
      env, floor_, peaks, peak_count, troughs, trough_count, items, noise_mult, veto_mult, prev_amp, next_amp, ratio,
      flags);
  BPM_LAUNCH_OK();
  return BPM_OK;
}

// ------------------------------------------------------------------ K9
// datetime.timedelta(seconds=t) keeps whole seconds exactly and rounds the fraction to the
// nearest microsecond, ties to even (CPython Modules/_datetimemodule.c accum()/delta_new).
__device__ __forceinline__ long long seconds_to_us(double t) {
  const double ip = trunc(t);
  const double fr = __dsub_rn(t, ip);
  return static_cast<long long>(ip) * 1000000ll + static_cast<long long>(rint(__dmul_rn(fr, 1.0e6)));
}

// one CTA per beat list: instantaneous BPM of the valid intervals, ordered (:1466-1475)
__global__ void __launch_bounds__(1024) k_bpm_instant(const int64_t* __restrict__ beats,
                                                      const BpmItem* __restrict__ lists, int rate,
                                                      double* __restrict__ inst, double* __restrict__ times_sec,
                                                      int64_t* __restrict__ stamp_us, int64_t* __restrict__ n_valid) {
  __shared__ int s_scan[34];
  __shared__ long long s_base;
  const BpmItem it = lists[blockIdx.x];
  const int64_t* p = beats + it.m_off;
  const long long B = it.m;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  const double r = static_cast<double>(rate);
  for (long long base = 0; base + 1 < B; base += blockDim.x) {
    const long long i = base + threadIdx.x;
    bool ok = false;
    double dt = 0.0, t1 = 0.0;
    if (i + 1 < B) {
      const double t0 = __ddiv_rn(static_cast<double>(p[i]), r);
      t1 = __ddiv_rn(static_cast<double>(p[i + 1]), r);
      dt = __dsub_rn(t1, t0);
      ok = dt > 1e-6;
    }
    int total;
    const int ex = block_exclusive_scan(ok ? 1 : 0, &total, s_scan);
    const long long b0 = s_base;
    if (ok) {
      const long long o = it.m_off + b0 + ex;
      inst[o] = __ddiv_rn(60.0, dt);
      times_sec[o] = t1;
      stamp_us[o] = seconds_to_us(t1);
    }
    __syncthreads();
    if (threadIdx.x == 0) s_base = b0 + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) n_valid[blockIdx.x] = s_base;
}

// centred time window (t - w/2, t + w/2], min_periods=1 (:1477-1479)
__global__ void k_bpm_smooth(const double* __restrict__ inst, const int64_t* __restrict__ stamp_us,
                             const int64_t* __restrict__ n_valid, const BpmItem* __restrict__ lists,
                             int64_t window_us, double* __restrict__ smoothed) {
  const BpmItem it = lists[blockIdx.y];
  const long long n = n_valid[blockIdx.y];
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t* us = stamp_us + it.m_off;
  const double* v = inst + it.m_off;
  const int64_t half = window_us / 2;
  const int64_t lo = us[i] - half, hi = us[i] + half;
  long long a = i, b = i;
  while (a - 1 >= 0 && us[a - 1] > lo) --a;
  while (b + 1 < n && us[b + 1] <= hi) ++b;
  double s = 0.0;
  for (long long k = a; k <= b; ++k) s = __dadd_rn(s, v[k]);
  smoothed[it.m_off + i] = __ddiv_rn(s, static_cast<double>(b - a + 1));
}

int bpm_series_run(const int64_t* beats, const BpmItem* lists, const BatchShape& sh, int rate,
                   int64_t window_us, double* inst, double* smoothed, double* times_sec, int64_t* stamp_us,
                   int64_t* n_valid, cudaStream_t st) {
  if (!beats || !lists || !inst || !smoothed || !times_sec || !stamp_us || !n_valid || rate <= 0) return BPM_ERR_ARG;
  BPM_KERNEL(k_bpm_instant);
  k_bpm_instant<<<sh.n_items, 1024, 0, st>>>(beats, lists, rate, inst, times_sec, stamp_us, n_valid);
  BPM_LAUNCH_OK();
  BPM_KERNEL(k_bpm_smooth);
  k_bpm_smooth<<<dim3(cdiv(sh.max_m, 256), sh.n_items), 256, 0, st>>>(inst, stamp_us, n_valid, lists, window_us, smoothed);
  BPM_LAUNCH_OK();
  return BPM_OK;
}

// ------------------------------------------------------------------ K10
// one CTA per series.  sign=+1: for every i the first j with t_j >= t_i + window, slope =
// (v_j - v_i)/(t_j - t_i); the strictly largest positive slope, first i on ties (:1582-1594).
// sign=-1: the same on the sub-series starting at the first maximum, strictly most negative
// slope (:1555-1573).  result = {found, i, j, slope}.
__global__ void __launch_bounds__(1024) k_steepest(const double* __restrict__ smoothed,
                                                   const int64_t* __restrict__ stamp_us,
                                                   const int64_t* __restrict__ n_valid,
                                                   const BpmItem* __restrict__ lists, int sign, double window_sec,
                                                   double* __restrict__ result, int stage_cap) {
  extern __shared__ __align__(16) unsigned char st_raw[];   // optional staging: stamps then values
  __shared__ double s_val[32];
  __shared__ long long s_idx[32];
  __shared__ long long s_j[32];
  __shared__ long long s_start;
  // sign == 0: blockIdx.y selects the direction (0: +1 exertion, 1: -1 recovery), results
  // interleaved as [list][2][4]
  const bool both = (sign == 0);
  if (both) sign = blockIdx.y == 0 ? +1 : -1;
  const BpmItem it = lists[blockIdx.x];
  const long long n = n_valid[blockIdx.x];
  const double* v = smoothed + it.m_off;
  const int64_t* us = stamp_us + it.m_off;
  bool staged = false;
  if (stage_cap > 0 && n <= stage_cap) {
    int64_t* s_us = reinterpret_cast<int64_t*>(st_raw);
    double* s_v = reinterpret_cast<double*>(st_raw) + stage_cap;
    for (long long t = threadIdx.x; t < n; t += blockDim.x) { s_us[t] = us[t]; s_v[t] = v[t]; }
    __syncthreads();
    us = s_us;
    v = s_v;
    staged = true;
  }
  double* res = both ? result + 8 * blockIdx.x + 4 * blockIdx.y : result + 4 * blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (n < 2) {
    if (threadIdx.x == 0) { res[0] = 0; res[1] = 0; res[2] = 0; res[3] = 0; }
    return;
  }
  long long start = 0;
  if (sign < 0) {
    // first index of the maximum (Series.idxmax)
    double bv = -INFINITY; long long bi = 0x7fffffffffffffffll;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
      const double x = v[i];
      if (x > bv) { bv = x; bi = i; }
    }
    for (int o = 16; o; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { s_val[warp] = bv; s_idx[warp] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < nw; ++w)
        if (s_val[w] > bv || (s_val[w] == bv && s_idx[w] < bi)) { bv = s_val[w]; bi = s_idx[w]; }
      s_start = bi;
    }
    __syncthreads();
    start = s_start;
    __syncthreads();
  }
  const long long len = n - start;
  const int64_t us0 = us[start];
  const double t_last = __ddiv_rn(static_cast<double>(us[n - 1] - us0), 1.0e6);
  if (len < 1 || t_last < window_sec) {
    if (threadIdx.x == 0) { res[0] = 0; res[1] = 0; res[2] = 0; res[3] = 0; }
    return;
  }
  // seconds since the sub-series start, total_seconds() = us / 1e6, tabulated once when staged
  double* s_t = reinterpret_cast<double*>(st_raw) + 2 * static_cast<size_t>(stage_cap);
  if (staged) {
    __syncthreads();
    for (long long t = start + threadIdx.x; t < n; t += blockDim.x)
      s_t[t] = __ddiv_rn(static_cast<double>(us[t] - us0), 1.0e6);
    __syncthreads();
  }
  double best = 0.0; long long best_i = 0x7fffffffffffffffll, best_j = 0;
  for (long long ii = threadIdx.x; ii + 1 < len; ii += blockDim.x) {
    const long long i = start + ii;
    const double ti = staged ? s_t[i] : __ddiv_rn(static_cast<double>(us[i] - us0), 1.0e6);
    const double target = __dadd_rn(ti, window_sec);
    // first j in [start, n) with t_j >= target (times are non-decreasing)
    long long lo = start, hi = n;
    while (lo < hi) {
      const long long mid = (lo + hi) >> 1;
      const double tm = staged ? s_t[mid] : __ddiv_rn(static_cast<double>(us[mid] - us0), 1.0e6);
      if (tm >= target) hi = mid; else lo = mid + 1;
    }
    if (lo >= n) continue;           // the reference breaks here; later i cannot succeed either
    const double tj = staged ? s_t[lo] : __ddiv_rn(static_cast<double>(us[lo] - us0), 1.0e6);
    const double dur = __dsub_rn(tj, ti);
    if (dur > 0.0) {
      const double slope = __ddiv_rn(__dsub_rn(v[lo], v[i]), dur);
      const bool better = sign > 0 ? (slope > best) : (slope < best);
      if (better) { best = slope; best_i = i; best_j = lo; }   // ii ascending per thread: first wins
    }
  }
  for (int o = 16; o; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, best, o);
    const long long oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    const long long oj = __shfl_xor_sync(0xffffffffu, best_j, o);
    const bool better = sign > 0 ? (ov > best) : (ov < best);
    if (better || (ov == best && oi < best_i)) { best = ov; best_i = oi; best_j = oj; }
  }
  if (lane == 0) { s_val[warp] = best; s_idx[warp] = best_i; s_j[warp] = best_j; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < nw; ++w) {
      const bool better = sign > 0 ? (s_val[w] > best) : (s_val[w] < best);
      if (better || (s_val[w] == best && s_idx[w] < best_i)) { best = s_val[w]; best_i = s_idx[w]; best_j = s_j[w]; }
    }
    const bool found = best_i != 0x7fffffffffffffffll && best != 0.0;
    res[0] = found ? 1.0 : 0.0;
    res[1] = found ? static_cast<double>(best_i) : 0.0;
    res[2] = found ? static_cast<double>(best_j) : 0.0;
    res[3] = found ? best : 0.0;
  }
}

int steepest_run(const double* smoothed, const int64_t* stamp_us, const int64_t* n_valid, const BpmItem* lists,
                 int n_lists, int64_t max_len, int sign, double window_sec, double* result, cudaStream_t st) {
  if (!smoothed || !stamp_us || !n_valid || !lists || !result || n_lists <= 0 || sign < -1 || sign > 1)
    return BPM_ERR_ARG;
  // series up to 6144 points are staged in shared memory (the binary searches are latency-bound)
  const int stage_cap = max_len <= 6144 ? 6144 : 0;
  const size_t smem = static_cast<size_t>(stage_cap) * 24;
  if (smem) cudaFuncSetAttribute(k_steepest, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  BPM_KERNEL(k_steepest);
  k_steepest<<<dim3(n_lists, sign == 0 ? 2 : 1), 1024, smem, st>>>(smoothed, stamp_us, n_valid, lists, sign,
                                                                  window_sec, result, stage_cap);
  BPM_LAUNCH_OK();
  return BPM_OK;
}

// ------------------------------------------------------------------ K12
// one thread per window of `win` RR intervals (:1424-1454)
__global__ void k_hrv(const int64_t* __restrict__ beats, const BpmItem* __restrict__ lists, int rate, int win,
                      int step, double* __restrict__ out, int64_t* __restrict__ rows) {
  const BpmItem it = lists[blockIdx.y];
  const long long B = it.m;
  long long nrows = 0;
  if (B >= win) {
    const long long span = (B - 1) - win + 1;          // range(0, span, step)
    nrows = span > 0 ? (span + step - 1) / step : 0;
  }
  const long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r == 0) rows[blockIdx.y] = nrows;
  if (r >= nrows) return;
  const int64_t* p = beats + it.m_off;
  const long long i = r * step;
  const double rt = static_cast<double>(rate);
  double sum = 0.0;
  for (int k = 0; k < win; ++k) {
    const double ms = __dmul_rn(__ddiv_rn(static_cast<double>(p[i + k + 1] - p[i + k]), rt), 1000.0);
    sum = __dadd_rn(sum, ms);
  }
  const double mean_ms = __ddiv_rn(sum, static_cast<double>(win));
  double ssd = 0.0, sdiff = 0.0, prev = 0.0;
  for (int k = 0; k < win; ++k) {
    const double ms = __dmul_rn(__ddiv_rn(static_cast<double>(p[i + k + 1] - p[i + k]), rt), 1000.0);
    const double d = __dsub_rn(ms, mean_ms);
    ssd = __dadd_rn(ssd, __dmul_rn(d, d));
    if (k > 0) { const double e = __dsub_rn(ms, prev); sdiff = __dadd_rn(sdiff, __dmul_rn(e, e)); }
    prev = ms;
  }
  const double sdnn = sqrt(__ddiv_rn(ssd, static_cast<double>(win)));
  const double rmssd = sqrt(__ddiv_rn(sdiff, static_cast<double>(win - 1)));
  const double mean_s = __ddiv_rn(mean_ms, 1000.0);
  const double t0 = __ddiv_rn(static_cast<double>(p[i]), rt), t1 = __ddiv_rn(static_cast<double>(p[i + win]), rt);
  double* o = out + 4 * (it.m_off + r);
  o[0] = __ddiv_rn(__dadd_rn(t0, t1), 2.0);
  o[1] = mean_s > 0.0 ? __ddiv_rn(rmssd, mean_s) : 0.0;
  o[2] = sdnn;
  o[3] = mean_s > 0.0 ? __ddiv_rn(60.0, mean_s) : 0.0;
}

// ------------------------------------------------------------------ float32 outputs
// "float32 mode" of the north star: every stage computes in float64 (the filter states and the scan
// carries need it: a float32 recurrence at these pole radii is off by 1e-3); only the signals handed
// back to the host are rounded to float32 -- half the read-back bytes, 6e-8 relative error.
__global__ void k_cast_f32(const double* __restrict__ src, float* __restrict__ dst, int64_t n) {
  const int64_t T = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += T)
    dst[i] = static_cast<float>(src[i]);
}

int cast_f32_run(const double* src, float* dst, int64_t n, cudaStream_t st) {
  if (!src || !dst || n < 0) return BPM_ERR_ARG;
  if (n == 0) return BPM_OK;
  int64_t gx = (n + 1023) / 1024;
  if (gx > 148 * 8) gx = 148 * 8;
  BPM_KERNEL(k_cast_f32);
  k_cast_f32<<<static_cast<unsigned>(gx), 256, 0, st>>>(src, dst, n);
  BPM_LAUNCH_OK();
  return BPM_OK;
}

int hrv_run(const int64_t* beats, const BpmItem* lists, const BatchShape& sh, int rate, int win, int step,
            double* out, int64_t* rows, cudaStream_t st) {
  if (!beats || !lists || !out || !rows || rate <= 0 || win < 2 || step < 1) return BPM_ERR_ARG;
  BPM_KERNEL(k_hrv);
  k_hrv<<<dim3(cdiv(sh.max_m / step + 1, 128), sh.n_items), 128, 0, st>>>(beats, lists, rate, win, step, out, rows);
  BPM_LAUNCH_OK();
  return BPM_OK;
}

}  // namespace bpm
