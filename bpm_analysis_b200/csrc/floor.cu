// K5 + K6 + K7: the dynamic noise floor (bpm_analysis.py:1081-1086, :1090-1097, :1103-1106).
//
//   pd.Series(env[knots], index=knots).reindex(arange(m)).interpolate()
//     .rolling(window=W, min_periods=3, center=True).quantile(q).bfill().ffill()
//
// The interpolated series is piecewise linear between the knots (troughs): NaN before the
// first knot, np.interp's  slope*(i - t_k) + v_k  inside segment k, the last knot's value
// held to the end.  It is never materialised.  A window holds a few dozen segments, each an
// arithmetic progression, so for a value v the number of window samples < v / <= v and the
// nearest sample above / below v are closed-form per segment (estimate by the inverse slope,
// then corrected against the actual float64 sample values, so ranks are exact).  Each thread
// owns a run of consecutive outputs: the first is located by bisection on v, every later one
// starts from its predecessor's answer and moves by successor / predecessor steps -- the
// window only gained and lost one sample.  pandas' order-statistic interpolation
// (vlow + (vhigh - vlow) * frac, roll_quantile 'linear') is evaluated unfused.
//
// ALU-bound, not HBM-bound: algorithmic traffic is 8 B per output plus the knot table.
#include "common.cuh"

namespace bpm {

struct FloorMeta {
  long long n_knots;
  long long iv0, iv1;     // outputs with >= min_periods observations: [iv0, iv1]; bfill/ffill clamp to it
  long long valid;        // 0: no output is valid (all NaN)
};

struct KnotTable {
  const int64_t* t;       // positions
  const double* v;        // env at the knot
  const double* slope;    // np.interp slope of the segment starting here (0 for the last)
  const double* inv;      // 1 / slope (0 when slope == 0)
};

constexpr int MIN_PERIODS = 3;   // bpm_analysis.py:1085

__device__ __forceinline__ long long n_obs_at(long long i, long long m, long long t0, int left, int off) {
  long long hi = i + off; if (hi > m - 1) hi = m - 1;
  long long lo = i - left; if (lo < 0) lo = 0; if (lo < t0) lo = t0;
  return hi - lo + 1;
}

__global__ void k_knot_table(const double* __restrict__ env, const int64_t* __restrict__ knots,
                             const int64_t* __restrict__ knot_count, const BpmItem* __restrict__ items,
                             int window, double* __restrict__ kv, double* __restrict__ ks,
                             double* __restrict__ kinv, FloorMeta* __restrict__ meta) {
  const int item = blockIdx.y;
  const BpmItem it = items[item];
  const long long T = knot_count[item];
  const int64_t* kt = knots + it.m_off;
  const double* e = env + it.m_off;
  const long long k = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k < T) {
    const double v0 = e[kt[k]];
    double s = 0.0, inv = 0.0;
    if (k + 1 < T) {
      const double v1 = e[kt[k + 1]];
      s = __ddiv_rn(__dsub_rn(v1, v0), static_cast<double>(kt[k + 1] - kt[k]));
      inv = (s != 0.0) ? 1.0 / s : 0.0;
    }
    kv[it.m_off + k] = v0;
    ks[it.m_off + k] = s;
    kinv[it.m_off + k] = inv;
  }
  if (k == 0) {
    FloorMeta mt;
    mt.n_knots = T;
    mt.valid = 0; mt.iv0 = 0; mt.iv1 = -1;
    if (T >= 1 && window >= MIN_PERIODS) {
      const int off = (window - 1) / 2, left = window - 1 - off;
      const long long t0 = kt[0], m = it.m;
      long long a = t0 + 2 - off; if (a < 0) a = 0;
      if (a <= m - 1 && n_obs_at(a, m, t0, left, off) >= MIN_PERIODS) {
        long long b = m - 1;
        while (b > a && n_obs_at(b, m, t0, left, off) < MIN_PERIODS) --b;
        mt.valid = 1; mt.iv0 = a; mt.iv1 = b;
      }
    }
    meta[item] = mt;
  }
}

struct WinEval {
  long long lt, le;       // window samples < v, <= v
  double succ, pred;      // nearest sample above / below v (+inf / -inf when none)
};

struct WinCtx {
  const int64_t* t; const double* v; const double* s; const double* inv;
  long long T;
};

// value of the interpolated series at i inside segment k (np.interp formula, unfused)
__device__ __forceinline__ double seg_val(const WinCtx& c, long long k, long long i) {
  return __dadd_rn(__dmul_rn(c.s[k], static_cast<double>(i - c.t[k])), c.v[k]);
}

__device__ __forceinline__ long long clamp_est(double e, long long lo, long long hi) {
  if (!(e > static_cast<double>(lo))) return lo;      // also catches NaN
  if (e >= static_cast<double>(hi)) return hi;
  return static_cast<long long>(e);
}

// counts / neighbours of v among the samples at indices [a, b], segments ka..kb
__device__ WinEval win_eval(const WinCtx& c, long long a, long long b, long long ka, long long kb, double v) {
  WinEval r;
  r.lt = 0; r.le = 0; r.succ = INFINITY; r.pred = -INFINITY;
  for (long long k = ka; k <= kb; ++k) {
    const long long tk = c.t[k];
    const long long i0 = tk > a ? tk : a;
    long long i1 = b;
    if (k + 1 < c.T) { const long long e = c.t[k + 1] - 1; if (e < i1) i1 = e; }
    if (i1 < i0) continue;
    const double sl = c.s[k];
    const double vk = c.v[k];
    if (sl == 0.0 || k + 1 >= c.T) {
      const long long cnt = i1 - i0 + 1;
      if (vk < v) { r.lt += cnt; r.le += cnt; if (vk > r.pred) r.pred = vk; }
      else if (vk == v) { r.le += cnt; }
      else { if (vk < r.succ) r.succ = vk; }
      continue;
    }
    const double est = (v - vk) * c.inv[k];
    if (sl > 0.0) {
      // largest i in [i0, i1] with f(i) <= v  (i0 - 1 when none)
      long long ie = clamp_est(floor(est) + static_cast<double>(tk), i0 - 1, i1);
      while (ie < i1 && seg_val(c, k, ie + 1) <= v) ++ie;
      while (ie >= i0 && seg_val(c, k, ie) > v) --ie;
      long long il = ie;                                  // largest with f(i) < v
      while (il >= i0 && !(seg_val(c, k, il) < v)) --il;
      r.le += ie - i0 + 1;
      r.lt += il - i0 + 1;
      if (ie < i1) { const double s1 = seg_val(c, k, ie + 1); if (s1 < r.succ) r.succ = s1; }
      if (il >= i0) { const double p1 = seg_val(c, k, il); if (p1 > r.pred) r.pred = p1; }
    } else {
      // smallest i in [i0, i1] with f(i) <= v  (i1 + 1 when none)
      long long ie = clamp_est(ceil(est) + static_cast<double>(tk), i0, i1 + 1);
      while (ie > i0 && seg_val(c, k, ie - 1) <= v) --ie;
      while (ie <= i1 && seg_val(c, k, ie) > v) ++ie;
      long long il = ie;                                  // smallest with f(i) < v
      while (il <= i1 && !(seg_val(c, k, il) < v)) ++il;
      r.le += i1 - ie + 1;
      r.lt += i1 - il + 1;
      if (ie > i0) { const double s1 = seg_val(c, k, ie - 1); if (s1 < r.succ) r.succ = s1; }
      if (il <= i1) { const double p1 = seg_val(c, k, il); if (p1 > r.pred) r.pred = p1; }
    }
  }
  return r;
}

// last knot index with t[k] <= i (requires t[0] <= i)
__device__ __forceinline__ long long knot_at_or_before(const WinCtx& c, long long i) {
  long long lo = 0, hi = c.T - 1;
  while (lo < hi) {
    const long long mid = (lo + hi + 1) >> 1;
    if (c.t[mid] <= i) lo = mid; else hi = mid - 1;
  }
  return lo;
}

constexpr int RF_THREADS = 128;

// mode[item]: 0 = rolling quantile over the knots; 1 = copy alt[]; 2 = constant cval[item].
// nan_fill (optional): value written instead of NaN when no output is valid.
__global__ void __launch_bounds__(RF_THREADS) k_rolling_floor(
    const BpmItem* __restrict__ items, KnotTable kt, const FloorMeta* __restrict__ meta, int window, double q,
    int run, const int* __restrict__ mode, const double* __restrict__ alt, const double* __restrict__ cval,
    const double* __restrict__ nan_fill, double* __restrict__ out) {
  const int item = blockIdx.y;
  const BpmItem it = items[item];
  const long long m = it.m;
  const long long first = (static_cast<long long>(blockIdx.x) * RF_THREADS + threadIdx.x) * run;
  if (first >= m) return;
  const long long last = min(m, first + run);
  double* o = out + it.m_off;
  const int md = mode ? mode[item] : 0;
  const double nanv = nan_fill ? nan_fill[item] : __longlong_as_double(0x7ff8000000000000ll);
  if (md == 2) {
    const double c = cval[item];
    for (long long i = first; i < last; ++i) o[i] = c;
    return;
  }
  if (md == 1) {
    const double* a = alt + it.m_off;
    for (long long i = first; i < last; ++i) { const double v = a[i]; o[i] = isnan(v) ? nanv : v; }
    return;
  }
  const FloorMeta mt = meta[item];
  if (!mt.valid) {
    for (long long i = first; i < last; ++i) o[i] = nanv;
    return;
  }
  WinCtx c;
  c.t = kt.t + it.m_off; c.v = kt.v + it.m_off; c.s = kt.slope + it.m_off; c.inv = kt.inv + it.m_off;
  c.T = mt.n_knots;
  const long long t0 = c.t[0];
  const int off = (window - 1) / 2, left = window - 1 - off;

  bool have = false;
  double v = 0.0, prev_out = 0.0;
  long long prev_i = -1, ka = 0, kb = 0;
  for (long long io = first; io < last; ++io) {
    long long i = io;
    if (i < mt.iv0) i = mt.iv0;
    if (i > mt.iv1) i = mt.iv1;
    if (i == prev_i) { o[io] = prev_out; continue; }
    long long b = i + off; if (b > m - 1) b = m - 1;
    long long a = i - left; if (a < 0) a = 0; if (a < t0) a = t0;
    const long long n = b - a + 1;
    // pandas roll_quantile: idx_with_fraction = q * (nobs - 1)
    const double fq = __dmul_rn(q, static_cast<double>(n - 1));
    const long long idx = static_cast<long long>(fq);
    const double frac = __dsub_rn(fq, static_cast<double>(idx));
    if (!have || i != prev_i + 1) {
      ka = knot_at_or_before(c, a);
      kb = knot_at_or_before(c, b);
    } else {
      while (ka + 1 < c.T && c.t[ka + 1] <= a) ++ka;
      while (kb + 1 < c.T && c.t[kb + 1] <= b) ++kb;
    }
    if (!have) {
      // bracket by the extreme segment end values, then bisect on v
      double vlo = INFINITY, vhi = -INFINITY;
      for (long long k = ka; k <= kb; ++k) {
        const long long i0 = c.t[k] > a ? c.t[k] : a;
        long long i1 = b;
        if (k + 1 < c.T) { const long long e = c.t[k + 1] - 1; if (e < i1) i1 = e; }
        if (i1 < i0) continue;
        const bool flat = (k + 1 >= c.T);
        const double f0 = flat ? c.v[k] : seg_val(c, k, i0), f1 = flat ? c.v[k] : seg_val(c, k, i1);
        vlo = fmin(vlo, fmin(f0, f1));
        vhi = fmax(vhi, fmax(f0, f1));
      }
      v = vhi;
      for (int itn = 0; itn < 22 && vlo < vhi; ++itn) {
        const double mid = vlo + 0.5 * (vhi - vlo);
        if (!(mid > vlo && mid < vhi)) break;
        const WinEval e = win_eval(c, a, b, ka, kb, mid);
        if (idx < e.lt) { vhi = mid; v = mid; }
        else if (idx >= e.le) { vlo = mid; }
        else { v = mid; break; }
      }
      have = true;
    }
    WinEval e = win_eval(c, a, b, ka, kb, v);
    for (int guard = 0; guard < (1 << 20); ++guard) {
      if (idx < e.lt) v = e.pred;
      else if (idx >= e.le) v = e.succ;
      else break;
      e = win_eval(c, a, b, ka, kb, v);
    }
    double res = v;
    if (fq != static_cast<double>(idx)) {
      const double vhigh = (idx + 1 < e.le) ? v : e.succ;
      res = __dadd_rn(v, __dmul_rn(__dsub_rn(vhigh, v), frac));
    }
    o[io] = res;
    prev_out = res;
    prev_i = i;
  }
}

// K7: keep trough t iff the draft floor there is not NaN and env[t] <= mult * floor[t]
// (bpm_analysis.py:1090-1097).  keep_all[item] != 0 keeps every trough (the <5 troughs path
// returns the unsanitised list, :1077).
__global__ void k_sanitize_flags(const double* __restrict__ env, const double* __restrict__ draft,
                                 const int64_t* __restrict__ troughs, const int64_t* __restrict__ trough_count,
                                 const int* __restrict__ keep_all, const BpmItem* __restrict__ items, double mult,
                                 unsigned char* __restrict__ flags) {
  const int item = blockIdx.y;
  const BpmItem it = items[item];
  const long long k = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= trough_count[item]) return;
  bool keep = true;
  if (!(keep_all && keep_all[item])) {
    const int64_t t = troughs[it.m_off + k];
    const double f = draft[it.m_off + t];
    keep = !isnan(f) && env[it.m_off + t] <= __dmul_rn(mult, f);
  }
  flags[it.m_off + k] = keep ? 1 : 0;
}

// per-item control words of _calculate_dynamic_noise_floor
//   stage 0 (after the trough search):  few[i] = n_all < 5 ; draft_mode[i] = few ? skip(2 -> constant) : 0
//   stage 1 (after sanitisation):       final_mode[i] = few ? 2 : (n_kept > 2 ? 0 : 1)
__global__ void k_floor_modes(const int64_t* __restrict__ n_all, const int64_t* __restrict__ n_kept, int n_items,
                              int stage, int* __restrict__ few, int* __restrict__ mode) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_items) return;
  if (stage == 0) {
    const int f = n_all[i] < 5 ? 1 : 0;
    few[i] = f;
    mode[i] = f ? 2 : 0;
  } else {
    mode[i] = few[i] ? 2 : (n_kept[i] > 2 ? 0 : 1);
  }
}

// ------------------------------------------------------------------ host side
struct FloorBuffers {
  double *kv, *ks, *kinv;
  FloorMeta* meta;
};

static int carve_floor(Workspace& ws, int64_t total_m, int n_items, FloorBuffers* b) {
  b->kv = ws.take<double>(total_m);
  b->ks = ws.take<double>(total_m);
  b->kinv = ws.take<double>(total_m);
  b->meta = ws.take<FloorMeta>(n_items);
  return ws.overflow ? BPM_ERR_WORKSPACE : BPM_OK;
}

size_t rolling_floor_workspace_bytes(int64_t total_m, int n_items) {
  Workspace ws(nullptr, 0);
  FloorBuffers b;
  carve_floor(ws, total_m, n_items, &b);
  return ws.used;
}

int rolling_floor_run(const double* env, const int64_t* knots, const int64_t* knot_count, const BpmItem* items,
                      const BatchShape& sh, int window, double q, const int* mode, const double* alt,
                      const double* cval, const double* nan_fill, double* out, Workspace& ws, cudaStream_t st) {
  if (!env || !knots || !knot_count || !items || !out || sh.n_items <= 0 || window < 1) return BPM_ERR_ARG;
  FloorBuffers b;
  BPM_TRY(carve_floor(ws, sh.total_m, sh.n_items, &b));
  const int64_t max_k = sh.max_m / 2 + 2;
  BPM_KERNEL(k_knot_table);
  k_knot_table<<<dim3(cdiv(max_k, 256), sh.n_items), 256, 0, st>>>(env, knots, knot_count, items, window,
                                                                    b.kv, b.ks, b.kinv, b.meta);
  BPM_LAUNCH_OK();
  // outputs per thread: long enough to amortise the bisection, short enough to fill the GPU
  int64_t run = sh.total_m / (148 * 2048);
  if (run < 16) run = 16;
  if (run > 128) run = 128;
  KnotTable kt{knots, b.kv, b.ks, b.kinv};
  BPM_KERNEL(k_rolling_floor);
  k_rolling_floor<<<dim3(cdiv(sh.max_m, RF_THREADS * run), sh.n_items), RF_THREADS, 0, st>>>(
      items, kt, b.meta, window, q, static_cast<int>(run), mode, alt, cval, nan_fill, out);
  BPM_LAUNCH_OK();
  return BPM_OK;
}

}  // namespace bpm
