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
  const int* t;           // positions (int32: a recording has < 2^31 envelope samples)
  const double* v;        // env at the knot
  const double* slope;    // np.interp slope of the segment starting here (0 for the last)
  const double* inv;      // 1 / slope (0 when slope == 0)
  const double* end;      // value of the last sample of the segment (index t[k+1]-1)
};

constexpr int MIN_PERIODS = 3;   // bpm_analysis.py:1085

__device__ __forceinline__ long long n_obs_at(long long i, long long m, long long t0, int left, int off) {
  long long hi = i + off; if (hi > m - 1) hi = m - 1;
  long long lo = i - left; if (lo < 0) lo = 0; if (lo < t0) lo = t0;
  return hi - lo + 1;
}

__global__ void k_knot_table(const double* __restrict__ env, const int64_t* __restrict__ knots,
                             const int64_t* __restrict__ knot_count, const BpmItem* __restrict__ items,
                             int window, int* __restrict__ kt32, double* __restrict__ kv, double* __restrict__ ks,
                             double* __restrict__ kinv, double* __restrict__ kend, FloorMeta* __restrict__ meta) {
  const int item = blockIdx.y;
  const BpmItem it = items[item];
  const long long T = knot_count[item];
  const int64_t* kt = knots + it.m_off;
  const double* e = env + it.m_off;
  const long long k = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k < T) {
    const double v0 = e[kt[k]];
    double s = 0.0, inv = 0.0, ve = v0;
    if (k + 1 < T) {
      const double v1 = e[kt[k + 1]];
      const long long len = kt[k + 1] - kt[k];
      s = __ddiv_rn(__dsub_rn(v1, v0), static_cast<double>(len));
      inv = (s != 0.0) ? 1.0 / s : 0.0;
      ve = __dadd_rn(__dmul_rn(s, static_cast<double>(len - 1)), v0);
    }
    kt32[it.m_off + k] = static_cast<int>(kt[k]);
    kv[it.m_off + k] = v0;
    ks[it.m_off + k] = s;
    kinv[it.m_off + k] = inv;
    kend[it.m_off + k] = ve;
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
  int lt, le;             // window samples < v, <= v
  double succ, pred;      // nearest sample above / below v (+inf / -inf when none)
};

struct WinCtx {
  const int* t; const double* v; const double* s; const double* inv; const double* e;
  int T;
};

// value of the interpolated series at i inside segment k (np.interp formula, unfused)
__device__ __forceinline__ double seg_val(const WinCtx& c, int k, int i) {
  return __dadd_rn(__dmul_rn(c.s[k], static_cast<double>(i - c.t[k])), c.v[k]);
}
// ... at any index of segment k, including the flat tail after the last knot
__device__ __forceinline__ double val_at(const WinCtx& c, int k, int i) {
  return (k + 1 >= c.T) ? c.v[k] : seg_val(c, k, i);
}

__device__ __forceinline__ int clamp_est(double e, int lo, int hi) {
  if (!(e > static_cast<double>(lo))) return lo;      // also catches NaN
  if (e >= static_cast<double>(hi)) return hi;
  return static_cast<int>(e);
}

// counts / neighbours of v among the samples at indices [a, b], segments ka..kb
__device__ WinEval win_eval(const WinCtx& c, int a, int b, int ka, int kb, double v) {
  WinEval r;
  r.lt = 0; r.le = 0; r.succ = INFINITY; r.pred = -INFINITY;
  for (int k = ka; k <= kb; ++k) {
    const int tk = c.t[k];
    const bool last = (k + 1 >= c.T);
    const int tn = last ? 0x7fffffff : c.t[k + 1];
    const int i0 = tk > a ? tk : a;
    const int i1 = (tn - 1 < b) ? tn - 1 : b;
    if (i1 < i0) continue;
    const int cnt = i1 - i0 + 1;
    const double sl = c.s[k];
    const double vk = c.v[k];
    if (last || sl == 0.0) {
      if (vk < v) { r.lt += cnt; r.le += cnt; if (vk > r.pred) r.pred = vk; }
      else if (vk == v) { r.le += cnt; }
      else { if (vk < r.succ) r.succ = vk; }
      continue;
    }
    // quick accept / reject on the end values of the (monotone) piece inside the window
    const double fa = (i0 == tk) ? vk : seg_val(c, k, i0);
    const double fb = (i1 == tn - 1) ? c.e[k] : seg_val(c, k, i1);
    const double lo = fmin(fa, fb), hi = fmax(fa, fb);
    if (hi < v) { r.lt += cnt; r.le += cnt; if (hi > r.pred) r.pred = hi; continue; }
    if (lo > v) { if (lo < r.succ) r.succ = lo; continue; }
    const double est = (v - vk) * c.inv[k];
    if (sl > 0.0) {
      // largest i in [i0, i1] with f(i) <= v  (i0 - 1 when none)
      int ie = clamp_est(floor(est) + static_cast<double>(tk), i0 - 1, i1);
      while (ie < i1 && seg_val(c, k, ie + 1) <= v) ++ie;
      while (ie >= i0 && seg_val(c, k, ie) > v) --ie;
      int il = ie;                                        // largest with f(i) < v
      while (il >= i0 && !(seg_val(c, k, il) < v)) --il;
      r.le += ie - i0 + 1;
      r.lt += il - i0 + 1;
      if (ie < i1) { const double s1 = seg_val(c, k, ie + 1); if (s1 < r.succ) r.succ = s1; }
      if (il >= i0) { const double p1 = seg_val(c, k, il); if (p1 > r.pred) r.pred = p1; }
    } else {
      // smallest i in [i0, i1] with f(i) <= v  (i1 + 1 when none)
      int ie = clamp_est(ceil(est) + static_cast<double>(tk), i0, i1 + 1);
      while (ie > i0 && seg_val(c, k, ie - 1) <= v) --ie;
      while (ie <= i1 && seg_val(c, k, ie) > v) ++ie;
      int il = ie;                                        // smallest with f(i) < v
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
__device__ __forceinline__ int knot_at_or_before(const WinCtx& c, int i) {
  int lo = 0, hi = c.T - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (c.t[mid] <= i) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// first output of a run: the order statistic `idx` of the window, located by a bracket on
// the value axis narrowed with counts (secant and bisection steps alternate)
__device__ double locate_rank(const WinCtx& c, int a, int b, int ka, int kb, int n, int idx) {
  double vmin = INFINITY, vmax = -INFINITY;
  for (int k = ka; k <= kb; ++k) {
    const int tk = c.t[k];
    const bool last = (k + 1 >= c.T);
    const int tn = last ? 0x7fffffff : c.t[k + 1];
    const int i0 = tk > a ? tk : a;
    const int i1 = (tn - 1 < b) ? tn - 1 : b;
    if (i1 < i0) continue;
    const double fa = val_at(c, k, i0), fb = val_at(c, k, i1);
    vmin = fmin(vmin, fmin(fa, fb));
    vmax = fmax(vmax, fmax(fa, fb));
  }
  double lo = vmin, hi = vmax;
  WinEval e = win_eval(c, a, b, ka, kb, lo);
  if (idx < e.le) return lo;                              // the minimum already covers the rank
  int cle_lo = e.le, cle_hi = n;                          // count(<= lo) <= idx < count(<= hi)
  for (int itn = 0; itn < 64 && cle_hi - cle_lo > 3 && lo < hi; ++itn) {
    double g = (itn & 1) ? lo + 0.5 * (hi - lo)
                         : lo + (hi - lo) * ((static_cast<double>(idx - cle_lo) + 0.5) / static_cast<double>(cle_hi - cle_lo));
    if (!(g > lo && g < hi)) break;
    e = win_eval(c, a, b, ka, kb, g);
    if (idx < e.lt) { hi = g; cle_hi = e.lt; }
    else if (idx >= e.le) { lo = g; cle_lo = e.le; }
    else return g;
  }
  return hi;                                              // at or just above the target: callers walk down
}

constexpr int RF_THREADS = 64;
#ifdef BPM_DEBUG_COUNTERS
__device__ unsigned long long g_dbg[16];
#define DBG(i) atomicAdd(&g_dbg[i], 1ull)
#define DBGN(i, n) atomicAdd(&g_dbg[i], (unsigned long long)(n))
#else
#define DBG(i)
#define DBGN(i, n)
#endif
constexpr int RF_SMAX = 128;       // segments a window may span on the cached path (power of two)

// ---- cached path --------------------------------------------------------------------------
// Per thread and per segment of its window: cnt_le[k] = number of in-window samples of segment
// k that are <= v.  Inside a window a segment is monotone, so those samples are a prefix
// (rising) or a suffix (falling) of its in-window range and the nearest sample above / below v
// in that segment is one seg_val() away.  Sliding by one output touches only the two edge
// segments (O(1)); when a sample crosses the quantile level the order statistic moves to its
// neighbour, found by ONE pass over the window's segments (no division, no search).
// The run loop works in rounds: every lane first advances through the O(1) steps until it
// needs such a pass, then the lanes of the warp do their passes together.
struct SegRange {
  int i0, i1, cnt;
  double sl, vk;
  bool flat;
};

__device__ __forceinline__ SegRange seg_range(const WinCtx& c, int k, int a, int b) {
  SegRange r;
  const int tk = c.t[k];
  const bool last = (k + 1 >= c.T);
  const int tn = last ? 0x7fffffff : c.t[k + 1];
  r.i0 = tk > a ? tk : a;
  r.i1 = (tn - 1 < b) ? tn - 1 : b;
  r.cnt = r.i1 - r.i0 + 1;
  r.sl = c.s[k];
  r.vk = c.v[k];
  r.flat = last || r.sl == 0.0;
  return r;
}

// index of the j-th smallest in-window sample of a (monotone, non-flat) segment, j = 1..cnt
__device__ __forceinline__ int group_index(const SegRange& r, int j) {
  return r.sl > 0.0 ? r.i0 + j - 1 : r.i1 - j + 1;
}

#define RF_C(k) cs[((k) & (RF_SMAX - 1)) * RF_THREADS]

struct RollState {
  double v, succ;
  int lt, le;
  bool succ_ok;
};

// One pass over the segments: nearest sample above v (and the runner-up), nearest sample below v,
// which segments own them, and whether ties make the simple single-owner update insufficient.
struct Neighbours {
  double up1, up2, dn1;
  int own_up, own_dn, tie_seg;      // tie_seg: a segment holding samples equal to v (-1 none)
  bool multi_up, multi_dn, multi_tie;
};

__device__ Neighbours scan_neighbours(const WinCtx& c, const short* cs, int a, int b, int ka, int kb, double v) {
  Neighbours nb;
  nb.up1 = INFINITY; nb.up2 = INFINITY; nb.dn1 = -INFINITY;
  nb.own_up = -1; nb.own_dn = -1; nb.tie_seg = -1;
  nb.multi_up = false; nb.multi_dn = false; nb.multi_tie = false;
  for (int k = ka; k <= kb; ++k) {
    const SegRange r = seg_range(c, k, a, b);
    if (r.cnt <= 0) continue;
    const int ck = RF_C(k);
    if (ck < r.cnt) {
      const double cand = r.flat ? r.vk : seg_val(c, k, group_index(r, ck + 1));
      if (cand < nb.up1) { nb.up2 = nb.up1; nb.up1 = cand; nb.own_up = k; nb.multi_up = false; }
      else if (cand == nb.up1) { nb.multi_up = true; }
      else if (cand < nb.up2) { nb.up2 = cand; }
    }
    if (ck > 0) {
      double top = r.flat ? r.vk : seg_val(c, k, group_index(r, ck));
      int j = ck;
      if (top == v) {
        // this segment holds the sample(s) equal to v: its candidate below lies under them
        if (nb.tie_seg >= 0) nb.multi_tie = true;
        nb.tie_seg = k;
        if (r.flat) { j = 0; }
        else {
          do { --j; } while (j > 0 && seg_val(c, k, group_index(r, j)) == v);
          if (j > 0) top = seg_val(c, k, group_index(r, j));
        }
      }
      if (j > 0) {
        if (top > nb.dn1) { nb.dn1 = top; nb.own_dn = k; nb.multi_dn = false; }
        else if (top == nb.dn1) { nb.multi_dn = true; }
      }
    }
  }
  return nb;
}

// exact state at value v: per-segment counts, totals, successor
__device__ void cached_init(const WinCtx& c, short* cs, int a, int b, int ka, int kb, double v, RollState& st) {
  st.v = v; st.lt = 0; st.le = 0; st.succ = INFINITY; st.succ_ok = true;
  for (int k = ka; k <= kb; ++k) {
    const WinEval e = win_eval(c, a, b, k, k, v);
    RF_C(k) = static_cast<short>(e.le);
    st.le += e.le;
    st.lt += e.lt;
    if (e.succ < st.succ) st.succ = e.succ;
  }
}

// move to the next larger / smaller order statistic until lt <= idx < le, refreshing succ
__device__ void cached_settle(const WinCtx& c, short* cs, int a, int b, int ka, int kb, int idx, RollState& st) {
  for (int guard = 0; guard < (1 << 22); ++guard) {
    const bool up = idx >= st.le, down = idx < st.lt;
    if (!up && !down && st.succ_ok) return;
    const Neighbours nb = scan_neighbours(c, cs, a, b, ka, kb, st.v);
    DBG(0); DBGN(1, kb - ka + 1);
    if (up) DBG(2);
    if (down) DBG(3);
    if (!up && !down) { st.succ = nb.up1; st.succ_ok = true; return; }
    if (up) {
      st.lt = st.le;
      st.v = nb.up1;
      if (nb.own_up < 0) { st.succ = INFINITY; st.succ_ok = true; return; }
      if (!nb.multi_up) {
        const int k = nb.own_up;
        const SegRange r = seg_range(c, k, a, b);
        int ck = RF_C(k);
        const int c0 = ck;
        if (r.flat) ck = r.cnt;
        else { ++ck; while (ck < r.cnt && seg_val(c, k, group_index(r, ck + 1)) == st.v) ++ck; }
        RF_C(k) = static_cast<short>(ck);
        st.le += ck - c0;
        double nxt = INFINITY;
        if (ck < r.cnt) nxt = seg_val(c, k, group_index(r, ck + 1));
        st.succ = nxt < nb.up2 ? nxt : nb.up2;
        st.succ_ok = true;
      } else {
        for (int k = ka; k <= kb; ++k) {
          const SegRange r = seg_range(c, k, a, b);
          int ck = RF_C(k);
          if (r.cnt <= 0 || ck >= r.cnt) continue;
          const int c0 = ck;
          if (r.flat) { if (r.vk == st.v) ck = r.cnt; }
          else { while (ck < r.cnt && seg_val(c, k, group_index(r, ck + 1)) == st.v) ++ck; }
          RF_C(k) = static_cast<short>(ck);
          st.le += ck - c0;
        }
        st.succ_ok = false;                                // recomputed by the next scan if needed
      }
    } else {
      // down: samples equal to the old v leave the "<= v" groups; the new v is the best candidate
      const double old_v = st.v;
      const bool had_old = (st.le - st.lt) > 0;
      if (nb.multi_tie || nb.multi_dn) {
        int mult = 0;
        for (int k = ka; k <= kb; ++k) {
          const SegRange r = seg_range(c, k, a, b);
          int ck = RF_C(k);
          if (r.cnt <= 0 || ck <= 0) continue;
          if (r.flat) {
            if (r.vk == old_v) ck = 0;
            else if (r.vk == nb.dn1) mult += ck;
          } else {
            while (ck > 0 && seg_val(c, k, group_index(r, ck)) == old_v) --ck;
            int j = ck;
            while (j > 0 && seg_val(c, k, group_index(r, j)) == nb.dn1) { ++mult; --j; }
          }
          RF_C(k) = static_cast<short>(ck);
        }
        st.le = st.lt;
        st.lt = st.le - mult;
      } else {
        if (nb.tie_seg >= 0) {
          const int k = nb.tie_seg;
          const SegRange r = seg_range(c, k, a, b);
          int ck = RF_C(k);
          if (r.flat) ck = 0;
          else { while (ck > 0 && seg_val(c, k, group_index(r, ck)) == old_v) --ck; }
          RF_C(k) = static_cast<short>(ck);
        }
        int mult = 1;
        if (nb.own_dn >= 0) {
          const int k = nb.own_dn;
          const SegRange r = seg_range(c, k, a, b);
          const int ck = RF_C(k);
          if (r.flat) mult = ck;
          else { int j = ck - 1; while (j > 0 && seg_val(c, k, group_index(r, j)) == nb.dn1) { ++mult; --j; } }
        }
        st.le = st.lt;
        st.lt = st.le - mult;
      }
      st.v = nb.dn1;
      if (had_old) { st.succ = old_v; st.succ_ok = true; }
      else { st.succ = nb.up1; st.succ_ok = true; }
    }
  }
}

// continuous approximation of the window's value distribution (each segment uniform between
// its end values): a cheap starting guess for the order statistic `idx`
__device__ double approx_rank_value(const WinCtx& c, int a, int b, int ka, int kb, int idx) {
  double vmin = INFINITY, vmax = -INFINITY;
  for (int k = ka; k <= kb; ++k) {
    const SegRange r = seg_range(c, k, a, b);
    if (r.cnt <= 0) continue;
    const double fa = val_at(c, k, r.i0), fb = val_at(c, k, r.i1);
    vmin = fmin(vmin, fmin(fa, fb));
    vmax = fmax(vmax, fmax(fa, fb));
  }
  double lo = vmin, hi = vmax;
  const double target = static_cast<double>(idx) + 0.5;
  for (int itn = 0; itn < 10 && lo < hi; ++itn) {
    const double g = lo + 0.5 * (hi - lo);
    double cntf = 0.0;
    for (int k = ka; k <= kb; ++k) {
      const SegRange r = seg_range(c, k, a, b);
      if (r.cnt <= 0) continue;
      const double fa = val_at(c, k, r.i0), fb = val_at(c, k, r.i1);
      const double l = fmin(fa, fb), h = fmax(fa, fb);
      if (g >= h) cntf += r.cnt;
      else if (g > l) cntf += static_cast<double>(r.cnt) * ((g - l) / (h - l));
    }
    if (cntf < target) lo = g; else hi = g;
  }
  return lo + 0.5 * (hi - lo);
}

// start of a run: approximate guess, a few exact secant corrections, then neighbour moves
__device__ void cached_start(const WinCtx& c, short* cs, int a, int b, int ka, int kb, int n, int idx, RollState& st) {
  double g = approx_rank_value(c, a, b, ka, kb, idx);
  DBG(4);
  for (int attempt = 0; attempt < 6; ++attempt) {
    WinEval e = win_eval(c, a, b, ka, kb, g);
    DBG(5);
    if (attempt == 0) DBGN(6, abs(idx < e.lt ? idx - e.lt : (idx >= e.le ? idx - e.le + 1 : 0)));
    int dist = 0;
    if (idx < e.lt) dist = idx - e.lt;                      // negative: target lies below g
    else if (idx >= e.le) dist = idx - e.le + 1;            // positive: target lies above g
    if (dist > -6 && dist < 6) break;
    // local spacing from the neighbours of g
    double sp = 0.0;
    if (isfinite(e.succ) && isfinite(e.pred)) sp = 0.5 * (e.succ - e.pred);
    else if (isfinite(e.succ)) sp = e.succ - g;
    else if (isfinite(e.pred)) sp = g - e.pred;
    if (!(sp > 0.0)) break;
    g += static_cast<double>(dist) * sp;
  }
  if (!isfinite(g)) g = locate_rank(c, a, b, ka, kb, n, idx);
  cached_init(c, cs, a, b, ka, kb, g, st);
  DBGN(7, abs(idx < st.lt ? idx - st.lt : (idx >= st.le ? idx - st.le + 1 : 0)));
  if (idx - st.le > 64 || st.lt - idx > 64) {
    DBG(8);               // secant went astray: bracketed search
    g = locate_rank(c, a, b, ka, kb, n, idx);
    cached_init(c, cs, a, b, ka, kb, g, st);
  }
  cached_settle(c, cs, a, b, ka, kb, idx, st);
}

// mode[item]: 0 = rolling quantile over the knots; 1 = copy alt[]; 2 = constant cval[item].
// nan_fill (optional): value written instead of NaN when no output is valid.
__global__ void __launch_bounds__(RF_THREADS) k_rolling_floor(
    const BpmItem* __restrict__ items, KnotTable kt, const FloorMeta* __restrict__ meta, int window, double q,
    int run, const int* __restrict__ mode, const double* __restrict__ alt, const double* __restrict__ cval,
    const double* __restrict__ nan_fill, double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char rf_smem[];
  short* cs = reinterpret_cast<short*>(rf_smem) + threadIdx.x;       // [RF_SMAX][RF_THREADS]

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
  c.e = kt.end + it.m_off;
  c.T = static_cast<int>(mt.n_knots);
  const int t0 = c.t[0];
  const int off = (window - 1) / 2, left = window - 1 - off;
  const int mi = static_cast<int>(m);

  // can the whole run stay on the cached path?  (segments spanned by the union of its windows)
  bool cached = window < 32000;
  {
    int i_lo = static_cast<int>(first), i_hi = static_cast<int>(last - 1);
    if (i_lo < mt.iv0) i_lo = static_cast<int>(mt.iv0);
    if (i_lo > mt.iv1) i_lo = static_cast<int>(mt.iv1);
    if (i_hi < mt.iv0) i_hi = static_cast<int>(mt.iv0);
    if (i_hi > mt.iv1) i_hi = static_cast<int>(mt.iv1);
    int a_lo = i_lo - left; if (a_lo < t0) a_lo = t0; if (a_lo < 0) a_lo = 0;
    int b_hi = i_hi + off; if (b_hi > mi - 1) b_hi = mi - 1;
    if (knot_at_or_before(c, b_hi) - knot_at_or_before(c, a_lo) + 1 > RF_SMAX - 2) cached = false;
  }

  RollState st;
  st.v = 0.0; st.succ = INFINITY; st.lt = 0; st.le = 0; st.succ_ok = false;
  bool have = false;
  double prev_out = 0.0;
  int prev_i = -2, ka = 0, kb = 0, a_prev = 0, b_prev = 0;
  long long io = first;

  if (cached) {
    // rounds: O(1) steps until a neighbour scan is needed, then the scan, together
    while (io < last) {
      int a = 0, b = 0, n = 0, idx = 0, i = 0;
      double fq = 0.0, frac = 0.0;
      bool heavy = false, fresh = false;
      while (io < last) {
        i = static_cast<int>(io);
        if (i < mt.iv0) i = static_cast<int>(mt.iv0);
        if (i > mt.iv1) i = static_cast<int>(mt.iv1);
        if (i == prev_i) { o[io] = prev_out; ++io; continue; }
        b = i + off; if (b > mi - 1) b = mi - 1;
        a = i - left; if (a < 0) a = 0; if (a < t0) a = t0;
        n = b - a + 1;
        fq = __dmul_rn(q, static_cast<double>(n - 1));          // pandas: q * (nobs - 1)
        idx = static_cast<int>(fq);
        frac = __dsub_rn(fq, static_cast<double>(idx));
        if (!have || i != prev_i + 1) { fresh = true; heavy = true; break; }
        // one sample may enter at b, one may leave at a_prev
        if (b > b_prev) {
          if (kb + 1 < c.T && c.t[kb + 1] <= b) { ++kb; RF_C(kb) = 0; }
          const double en = val_at(c, kb, b);
          if (en <= st.v) { ++st.le; if (en < st.v) ++st.lt; RF_C(kb) += 1; }
          else if (en < st.succ) st.succ = en;
        }
        if (a > a_prev) {
          const double lv = val_at(c, ka, a_prev);
          if (lv <= st.v) { --st.le; if (lv < st.v) --st.lt; RF_C(ka) -= 1; }
          else if (lv == st.succ) st.succ_ok = false;
          if (ka + 1 < c.T && c.t[ka + 1] <= a) ++ka;
        }
        const bool need_succ = (fq != static_cast<double>(idx)) && !(idx + 1 < st.le);
        DBG(9);
        if (idx < st.lt || idx >= st.le || (need_succ && !st.succ_ok)) { DBG(10); heavy = true; break; }
        double res = st.v;
        if (fq != static_cast<double>(idx)) {
          const double vhigh = (idx + 1 < st.le) ? st.v : st.succ;
          res = __dadd_rn(st.v, __dmul_rn(__dsub_rn(vhigh, st.v), frac));
        }
        o[io] = res; prev_out = res; prev_i = i; a_prev = a; b_prev = b;
        ++io;
      }
      if (!heavy) break;
      if (fresh) {
        ka = knot_at_or_before(c, a);
        kb = knot_at_or_before(c, b);
        for (int k = 0; k < RF_SMAX; ++k) cs[k * RF_THREADS] = 0;
        cached_start(c, cs, a, b, ka, kb, n, idx, st);
        have = true;
      } else {
        cached_settle(c, cs, a, b, ka, kb, idx, st);
      }
      if (fq != static_cast<double>(idx) && !(idx + 1 < st.le) && !st.succ_ok)
        cached_settle(c, cs, a, b, ka, kb, idx, st);
      double res = st.v;
      if (fq != static_cast<double>(idx)) {
        const double vhigh = (idx + 1 < st.le) ? st.v : st.succ;
        res = __dadd_rn(st.v, __dmul_rn(__dsub_rn(vhigh, st.v), frac));
      }
      o[io] = res; prev_out = res; prev_i = i; a_prev = a; b_prev = b;
      ++io;
    }
    return;
  }

  // windows spanning more segments than the cache holds: exact evaluation per move
  double v = 0.0, succ = INFINITY;
  int lt = 0, le = 0;
  bool dirty = false;
  for (; io < last; ++io) {
    int i = static_cast<int>(io);
    if (i < mt.iv0) i = static_cast<int>(mt.iv0);
    if (i > mt.iv1) i = static_cast<int>(mt.iv1);
    if (i == prev_i) { o[io] = prev_out; continue; }
    int b = i + off; if (b > mi - 1) b = mi - 1;
    int a = i - left; if (a < 0) a = 0; if (a < t0) a = t0;
    const int n = b - a + 1;
    const double fq = __dmul_rn(q, static_cast<double>(n - 1));
    const int idx = static_cast<int>(fq);
    const double frac = __dsub_rn(fq, static_cast<double>(idx));
    if (!have || i != prev_i + 1) {
      ka = knot_at_or_before(c, a);
      kb = knot_at_or_before(c, b);
      v = locate_rank(c, a, b, ka, kb, n, idx);
      const WinEval e = win_eval(c, a, b, ka, kb, v);
      lt = e.lt; le = e.le; succ = e.succ; dirty = false;
      have = true;
    } else {
      if (b > b_prev) {
        while (kb + 1 < c.T && c.t[kb + 1] <= b) ++kb;
        const double en = val_at(c, kb, b);
        if (en < v) { ++lt; ++le; }
        else if (en == v) { ++le; }
        else if (en < succ) succ = en;
      }
      if (a > a_prev) {
        const double lv = val_at(c, ka, a_prev);
        if (lv < v) { --lt; --le; }
        else if (lv == v) { --le; }
        else if (lv == succ) dirty = true;
        while (ka + 1 < c.T && c.t[ka + 1] <= a) ++ka;
      }
    }
    for (int guard = 0; guard < (1 << 22); ++guard) {
      if (idx < lt) {
        WinEval e = win_eval(c, a, b, ka, kb, v);
        v = e.pred;
        e = win_eval(c, a, b, ka, kb, v);
        lt = e.lt; le = e.le; succ = e.succ; dirty = false;
      } else if (idx >= le) {
        if (dirty) {
          const WinEval e = win_eval(c, a, b, ka, kb, v);
          lt = e.lt; le = e.le; succ = e.succ; dirty = false;
          continue;
        }
        v = succ;
        const WinEval e = win_eval(c, a, b, ka, kb, v);
        lt = e.lt; le = e.le; succ = e.succ;
      } else {
        break;
      }
    }
    double res = v;
    if (fq != static_cast<double>(idx)) {
      double vhigh = v;
      if (!(idx + 1 < le)) {
        if (dirty) {
          const WinEval e = win_eval(c, a, b, ka, kb, v);
          lt = e.lt; le = e.le; succ = e.succ; dirty = false;
        }
        vhigh = succ;
      }
      res = __dadd_rn(v, __dmul_rn(__dsub_rn(vhigh, v), frac));
    }
    o[io] = res;
    prev_out = res;
    prev_i = i;
    a_prev = a;
    b_prev = b;
  }
}
#undef RF_C

// ---- block-cooperative path -----------------------------------------------------------------
// A CTA owns RB_THREADS * run consecutive outputs.  It materialises the interpolated samples its
// windows touch (n = outputs + window - 1 <= RB_NCAP) in shared memory and SORTS them by value
// (sample sort: RB_G - 1 splitters from a sorted strided sample, counting sort into buckets,
// then each sample counts the members of its small bucket that precede it), keeping perm[] (sorted order -> sample index) and
// rank[] (sample index -> sorted position).  A coarse 2-D prefix table over (index chunk, rank
// band) lets every thread place its first window's order statistic in O(log) steps.  After that
// a thread slides over its run with a pointer p into the sorted order: the entering / leaving
// sample changes the number of in-window entries before p by at most one each, and p walks a few
// entries to the new order statistic.  Uniform O(1) work per output, all in shared memory, and
// exact: the samples are the float64 np.interp values.
constexpr int RB_THREADS = 256;
constexpr int RB_NCAP = 6144;          // samples a CTA can stage
constexpr int RB_G = 1024;             // sort buckets (RB_G - 1 splitters)
constexpr int RB_NSUP = 32;            // rank bands of the coarse table
constexpr int RB_MAXCH = 224;          // index chunks of the coarse table

struct RbShared {
  double d[RB_NCAP];
  unsigned short rank[RB_NCAP];        // bucket id while sorting, then position in sorted order
  unsigned short perm[RB_NCAP];
  unsigned short start[RB_G + 1];
  unsigned int hist[RB_G];
  union {
    double piv[RB_G];                                     // sorted splitters, piv[RB_G-1] = +inf
    unsigned short pc[RB_MAXCH + 1][RB_NSUP];             // samples with index < c*ch and band <= s
  } u;
  unsigned short ragged[RB_NSUP][RB_THREADS];   // per-thread band counts of a window's ragged ends
  double red[2 * (RB_THREADS / 32)];
  int scan_tmp[40];
};

__global__ void __launch_bounds__(RB_THREADS) k_rolling_floor_blk(
    const BpmItem* __restrict__ items, KnotTable kt, const FloorMeta* __restrict__ meta, int window, double q,
    int run, const int* __restrict__ mode, const double* __restrict__ alt, const double* __restrict__ cval,
    const double* __restrict__ nan_fill, double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char rb_raw[];
  RbShared& sh = *reinterpret_cast<RbShared*>(rb_raw);
  const int item = blockIdx.y;
  const BpmItem it = items[item];
  const long long m = it.m;
  const long long blk_first = static_cast<long long>(blockIdx.x) * RB_THREADS * run;
  if (blk_first >= m) return;
  const long long blk_last = min(m, blk_first + static_cast<long long>(RB_THREADS) * run);   // exclusive
  double* o = out + it.m_off;
  const int tid = threadIdx.x;
  const int md = mode ? mode[item] : 0;
  const double nanv = nan_fill ? nan_fill[item] : __longlong_as_double(0x7ff8000000000000ll);
  if (md == 2) {
    const double c = cval[item];
    for (long long i = blk_first + tid; i < blk_last; i += RB_THREADS) o[i] = c;
    return;
  }
  if (md == 1) {
    const double* al = alt + it.m_off;
    for (long long i = blk_first + tid; i < blk_last; i += RB_THREADS) { const double v = al[i]; o[i] = isnan(v) ? nanv : v; }
    return;
  }
  const FloorMeta mt = meta[item];
  if (!mt.valid) {
    for (long long i = blk_first + tid; i < blk_last; i += RB_THREADS) o[i] = nanv;
    return;
  }
  WinCtx c;
  c.t = kt.t + it.m_off; c.v = kt.v + it.m_off; c.s = kt.slope + it.m_off; c.inv = kt.inv + it.m_off;
  c.e = kt.end + it.m_off;
  c.T = static_cast<int>(mt.n_knots);
  const int t0 = c.t[0];
  const int off = (window - 1) / 2, left = window - 1 - off;
  const int mi = static_cast<int>(m);

  // effective (bfill / ffill clamped) output range of the CTA and the samples its windows touch
  int ie0 = static_cast<int>(blk_first), ie1 = static_cast<int>(blk_last - 1);
  if (ie0 < mt.iv0) ie0 = static_cast<int>(mt.iv0);
  if (ie0 > mt.iv1) ie0 = static_cast<int>(mt.iv1);
  if (ie1 < mt.iv0) ie1 = static_cast<int>(mt.iv0);
  if (ie1 > mt.iv1) ie1 = static_cast<int>(mt.iv1);
  int x0 = ie0 - left; if (x0 < 0) x0 = 0; if (x0 < t0) x0 = t0;
  int x1 = ie1 + off; if (x1 > mi - 1) x1 = mi - 1;
  const int n = x1 - x0 + 1;                                  // <= RB_NCAP (host guarantees)

  // ---- S1: interpolated samples -> shared memory
  {
    const int per = (n + RB_THREADS - 1) / RB_THREADS;
    const int j0 = tid * per, j1 = min(n, j0 + per);
    if (j0 < j1) {
      int k = knot_at_or_before(c, x0 + j0);
      for (int j = j0; j < j1; ++j) {
        const int x = x0 + j;
        while (k + 1 < c.T && c.t[k + 1] <= x) ++k;
        sh.d[j] = val_at(c, k, x);
      }
    }
  }
  for (int t = tid; t < RB_G; t += RB_THREADS) sh.hist[t] = 0;
  __syncthreads();

  // ---- S2: splitters = sorted strided sample of the staged values; bucket ids; histogram
  for (int t = tid; t < RB_G; t += RB_THREADS)
    sh.u.piv[t] = (t < RB_G - 1) ? sh.d[static_cast<int>((static_cast<long long>(t) * (n - 1)) / (RB_G - 2))] : INFINITY;
  __syncthreads();
  for (int k2 = 2; k2 <= RB_G; k2 <<= 1) {                    // bitonic sort, RB_G / 2 compare-exchanges per step
    for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
      for (int t = tid; t < RB_G / 2; t += RB_THREADS) {
        const int lo_i = ((t / j2) * 2 * j2) + (t % j2);
        const int hi_i = lo_i + j2;
        const bool up = ((lo_i & k2) == 0);
        const double x = sh.u.piv[lo_i], y = sh.u.piv[hi_i];
        if ((x > y) == up) { sh.u.piv[lo_i] = y; sh.u.piv[hi_i] = x; }
      }
      __syncthreads();
    }
  }
  for (int j = tid; j < n; j += RB_THREADS) {
    // bucket = number of splitters strictly below the value (equal values share a bucket)
    const double x = sh.d[j];
    int lo_i = 0, hi_i = RB_G - 1;
    while (lo_i < hi_i) {
      const int mid = (lo_i + hi_i) >> 1;
      if (sh.u.piv[mid] < x) lo_i = mid + 1; else hi_i = mid;
    }
    sh.rank[j] = static_cast<unsigned short>(lo_i);
    atomicAdd(&sh.hist[lo_i], 1u);
  }
  __syncthreads();
  // ---- S3: exclusive scan of the histogram -> start[], scatter (hist becomes the cursor),
  //          order inside every bucket, inverse permutation
  {
    constexpr int PER = RB_G / RB_THREADS;
    unsigned int loc[PER];
    int sum = 0;
#pragma unroll
    for (int u2 = 0; u2 < PER; ++u2) { loc[u2] = sh.hist[tid * PER + u2]; sum += loc[u2]; }
    int total;
    int ex = block_exclusive_scan(sum, &total, sh.scan_tmp);
#pragma unroll
    for (int u2 = 0; u2 < PER; ++u2) {
      sh.start[tid * PER + u2] = static_cast<unsigned short>(ex);
      sh.hist[tid * PER + u2] = ex;
      ex += loc[u2];
    }
    if (tid == RB_THREADS - 1) sh.start[RB_G] = static_cast<unsigned short>(ex);
  }
  __syncthreads();
  for (int j = tid; j < n; j += RB_THREADS) {
    const unsigned int pos = atomicAdd(&sh.hist[sh.rank[j]], 1u);
    sh.perm[pos] = static_cast<unsigned short>(j);
  }
  __syncthreads();
  // order inside the buckets by counting: final position = bucket start + #{bucket members that
  // sort before this sample}.  Work is dealt out in bucket order (thread <-> scattered position),
  // so the lanes of a warp walk the same bucket: equal trip counts, broadcast shared-memory reads.
  {
    constexpr int MAXPER = (RB_NCAP + RB_THREADS - 1) / RB_THREADS;
    unsigned short fin[MAXPER];
#pragma unroll 1
    for (int u2 = 0, e = tid; e < n; e += RB_THREADS, ++u2) {
      const int j = sh.perm[e];
      const int g = sh.rank[j];
      const int e0 = sh.start[g], e1 = sh.start[g + 1];
      const double x = sh.d[j];
      int before = 0;
      for (int f = e0; f < e1; ++f) {
        const int y = sh.perm[f];
        const double dy = sh.d[y];
        before += (dy < x || (dy == x && y < j)) ? 1 : 0;
      }
      fin[u2] = static_cast<unsigned short>(e0 + before);
    }
    __syncthreads();
    // perm is rewritten in place: every thread first reads the sample ids it owns
    unsigned short own[MAXPER];
#pragma unroll 1
    for (int u2 = 0, e = tid; e < n; e += RB_THREADS, ++u2) own[u2] = sh.perm[e];
    __syncthreads();
#pragma unroll 1
    for (int u2 = 0, e = tid; e < n; e += RB_THREADS, ++u2) {
      sh.rank[own[u2]] = fin[u2];
      sh.perm[fin[u2]] = own[u2];
    }
  }
  // ---- S4: coarse table over (index chunk, rank band)
  int ch = run > 32 ? run : 32;
  while ((n + ch - 1) / ch > RB_MAXCH) ch *= 2;
  const int nch = (n + ch - 1) / ch;
  int bshift = 0;
  while (((n - 1) >> bshift) >= RB_NSUP) ++bshift;            // band = rank >> bshift  in [0, RB_NSUP)
  for (int t = tid; t < (nch + 1) * RB_NSUP; t += RB_THREADS) (&sh.u.pc[0][0])[t] = 0;
  __syncthreads();
  for (int j = tid; j < n; j += RB_THREADS) {
    // counts land one row down so that an inclusive column scan yields the exclusive prefix
    unsigned short* cell = &sh.u.pc[j / ch + 1][sh.rank[j] >> bshift];
    // 16-bit shared atomics do not exist: add into the containing 32-bit word
    unsigned int* word = reinterpret_cast<unsigned int*>(reinterpret_cast<uintptr_t>(cell) & ~uintptr_t(3));
    const unsigned int add = (reinterpret_cast<uintptr_t>(cell) & 2) ? (1u << 16) : 1u;
    atomicAdd(word, add);
  }
  __syncthreads();
  if (tid < RB_NSUP) {
    unsigned int acc = 0;
    for (int r2 = 0; r2 <= nch; ++r2) { acc += sh.u.pc[r2][tid]; sh.u.pc[r2][tid] = static_cast<unsigned short>(acc); }
  }
  __syncthreads();
  for (int r2 = tid; r2 <= nch; r2 += RB_THREADS) {
    unsigned int acc = 0;
    for (int s2 = 0; s2 < RB_NSUP; ++s2) { acc += sh.u.pc[r2][s2]; sh.u.pc[r2][s2] = static_cast<unsigned short>(acc); }
  }
  __syncthreads();

  // ---- S5 / S6: every thread slides over its run
  const long long first = blk_first + static_cast<long long>(tid) * run;
  if (first >= blk_last) return;
  const long long last = min(blk_last, first + run);
  bool have = false;
  int p = 0, cb = 0;            // position in sorted order; in-window entries at positions < p
  double prev_out = 0.0;
  int prev_i = -2, a_prev = 0, b_prev = 0;
  for (long long io = first; io < last; ++io) {
    int i = static_cast<int>(io);
    if (i < mt.iv0) i = static_cast<int>(mt.iv0);
    if (i > mt.iv1) i = static_cast<int>(mt.iv1);
    if (i == prev_i) { o[io] = prev_out; continue; }
    int bb = i + off; if (bb > mi - 1) bb = mi - 1;
    int aa = i - left; if (aa < 0) aa = 0; if (aa < t0) aa = t0;
    const int nw = bb - aa + 1;
    const double fq = __dmul_rn(q, static_cast<double>(nw - 1));      // pandas: q * (nobs - 1)
    const int idx = static_cast<int>(fq);
    const double frac = __dsub_rn(fq, static_cast<double>(idx));
    const int a = aa - x0, b = bb - x0;                        // local indices, inclusive
    if (!have || i != prev_i + 1) {
      // #{window samples in rank bands <= s}: whole chunks from the table + the ragged ends.
      // The ragged ends are counted once into a per-thread band histogram (cumulated in place).
      const int ca = (a + ch - 1) / ch;                        // first whole chunk
      const int cbk = (b + 1) / ch;                            // one past the last whole chunk
      unsigned short* rg = &sh.ragged[0][0] + tid;                   // [RB_NSUP][RB_THREADS]
      for (int s2 = 0; s2 < RB_NSUP; ++s2) rg[s2 * RB_THREADS] = 0;
      if (cbk > ca) {
        for (int j = a; j < ca * ch; ++j) rg[(sh.rank[j] >> bshift) * RB_THREADS] += 1;
        for (int j = cbk * ch; j <= b; ++j) rg[(sh.rank[j] >> bshift) * RB_THREADS] += 1;
      } else {
        for (int j = a; j <= b; ++j) rg[(sh.rank[j] >> bshift) * RB_THREADS] += 1;
      }
      {
        unsigned int acc = 0;
        for (int s2 = 0; s2 < RB_NSUP; ++s2) { acc += rg[s2 * RB_THREADS]; rg[s2 * RB_THREADS] = static_cast<unsigned short>(acc); }
      }
      auto cum = [&](int s_) -> int {
        if (s_ < 0) return 0;
        int cs_ = rg[s_ * RB_THREADS];
        if (cbk > ca) cs_ += static_cast<int>(sh.u.pc[cbk][s_]) - static_cast<int>(sh.u.pc[ca][s_]);
        return cs_;
      };
      int slo = 0, shi = RB_NSUP - 1;                          // smallest band with cum > idx
      while (slo < shi) {
        const int mid = (slo + shi) >> 1;
        if (cum(mid) > idx) shi = mid; else slo = mid + 1;
      }
      p = slo << bshift;
      cb = cum(slo - 1);
      have = true;
    } else {
      if (b > b_prev) cb += (sh.rank[b] < p) ? 1 : 0;
      if (a > a_prev) cb -= (sh.rank[a_prev] < p) ? 1 : 0;
    }
    // walk p to the in-window entry that has exactly idx in-window entries before it
    while (cb > idx) {
      --p;
      const int j = sh.perm[p];
      if (j >= a && j <= b) --cb;
    }
    while (true) {
      const int j = sh.perm[p];
      const bool in = (j >= a && j <= b);
      if (in && cb == idx) break;
      if (in) ++cb;
      ++p;
    }
    const double vlow = sh.d[sh.perm[p]];
    double res = vlow;
    if (fq != static_cast<double>(idx)) {
      int p2 = p + 1;
      while (true) {                                           // next in-window entry (idx + 1 < nw here)
        const int j = sh.perm[p2];
        if (j >= a && j <= b) break;
        ++p2;
      }
      const double vhigh = sh.d[sh.perm[p2]];
      res = __dadd_rn(vlow, __dmul_rn(__dsub_rn(vhigh, vlow), frac));
    }
    o[io] = res;
    prev_out = res; prev_i = i; a_prev = a; b_prev = b;
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
  int* kt32;
  double *kv, *ks, *kinv, *kend;
  FloorMeta* meta;
};

static int carve_floor(Workspace& ws, int64_t total_m, int n_items, FloorBuffers* b) {
  b->kt32 = ws.take<int>(total_m);
  b->kv = ws.take<double>(total_m);
  b->ks = ws.take<double>(total_m);
  b->kinv = ws.take<double>(total_m);
  b->kend = ws.take<double>(total_m);
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
                                                                    b.kt32, b.kv, b.ks, b.kinv, b.kend, b.meta);
  BPM_LAUNCH_OK();
  // outputs per thread: long enough to amortise the bisection, short enough to fill the GPU
  int64_t run = sh.total_m / (148 * 1024);
  if (run < 48) run = 48;
  if (run > 512) run = 512;
  if (sh.max_m >= (1ll << 31)) return BPM_ERR_ARG;
  KnotTable kt{b.kt32, b.kv, b.ks, b.kinv, b.kend};
  // block-cooperative kernel when a CTA can stage its windows' samples; else the per-thread kernel
  if (window + RB_THREADS <= RB_NCAP) {
    int64_t rb = (RB_NCAP - window + 1) / RB_THREADS;
    if (rb > 64) rb = 64;
    // keep the grid large enough to fill the GPU on a single recording
    while (rb > 8 && sh.total_m / (RB_THREADS * rb) < 2 * 148) rb /= 2;
    cudaFuncSetAttribute(k_rolling_floor_blk, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(RbShared)));
    BPM_KERNEL(k_rolling_floor_blk);
    k_rolling_floor_blk<<<dim3(cdiv(sh.max_m, RB_THREADS * rb), sh.n_items), RB_THREADS, sizeof(RbShared), st>>>(
        items, kt, b.meta, window, q, static_cast<int>(rb), mode, alt, cval, nan_fill, out);
    BPM_LAUNCH_OK();
    return BPM_OK;
  }
  const size_t smem = sizeof(short) * RF_SMAX * RF_THREADS;
  cudaFuncSetAttribute(k_rolling_floor, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  BPM_KERNEL(k_rolling_floor);
  k_rolling_floor<<<dim3(cdiv(sh.max_m, RF_THREADS * run), sh.n_items), RF_THREADS, smem, st>>>(
      items, kt, b.meta, window, q, static_cast<int>(run), mode, alt, cval, nan_fill, out);
  BPM_LAUNCH_OK();
  return BPM_OK;
}

}  // namespace bpm
