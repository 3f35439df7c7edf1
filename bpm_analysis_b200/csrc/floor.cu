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

// Which branch of _calculate_dynamic_noise_floor a recording takes, derived on the device from its
// trough counts (no separate launch): 2 = fewer than 5 troughs: constant floor (:1073-1077);
// 1 = at most 2 troughs survive sanitisation: the floor over ALL troughs (:1107-1110); 0 = the
// regular case.  n_all == nullptr: mode 0 (stand-alone rolling floor).  n_kept == nullptr: the draft
// stage (only the "< 5" test applies).
struct FloorModeSrc {
  const int64_t* n_all;
  const int64_t* n_kept;
};
__device__ __forceinline__ int floor_mode_of(const FloorModeSrc& ms, int item) {
  if (ms.n_all == nullptr) return 0;
  if (ms.n_all[item] < 5) return 2;
  if (ms.n_kept == nullptr) return 0;
  return ms.n_kept[item] > 2 ? 0 : 1;
}

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

__device__ __forceinline__ long long n_obs_at(long long i, long long m, long long t0, int left, int off) {
  long long hi = i + off; if (hi > m - 1) hi = m - 1;
  long long lo = i - left; if (lo < 0) lo = 0; if (lo < t0) lo = t0;
  return hi - lo + 1;
}

// mode[item] == 1 selects the alternative knot list (the reference falls back to the draft floor,
// i.e. the floor over ALL troughs, when <= 2 troughs survive sanitisation, :1107-1110).
__global__ void k_knot_table(const double* __restrict__ env, const int64_t* __restrict__ knots,
                             const int64_t* __restrict__ knot_count, const int64_t* __restrict__ alt_knots,
                             const int64_t* __restrict__ alt_count, FloorModeSrc ms,
                             const BpmItem* __restrict__ items,
                             int window, int* __restrict__ kt32, double* __restrict__ kv, double* __restrict__ ks,
                             double* __restrict__ kinv, double* __restrict__ kend, FloorMeta* __restrict__ meta,
                             int64_t* __restrict__ total_out, int64_t* __restrict__ mode_out) {
  const int item = blockIdx.y;
  const BpmItem it = items[item];
  const int md_item = floor_mode_of(ms, item);
  const bool use_alt = (alt_knots != nullptr && md_item == 1);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    // what the caller's log lines report (bpm_analysis.py:1074, :1099, :1109)
    if (total_out && ms.n_all) total_out[item] = ms.n_all[item];
    if (mode_out) mode_out[item] = md_item;
  }
  const long long T = use_alt ? alt_count[item] : knot_count[item];
  const int64_t* kt = (use_alt ? alt_knots : knots) + it.m_off;
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

// first index r in [0, T] with t[r] >= bound (T: none), searched by a whole warp 32 ways per step: a
// 13 k-knot table is three dependent loads deep instead of fourteen (the CTA waits for this at its start)
__device__ __forceinline__ int warp_first_knot_ge(const int* __restrict__ t, int T, long long bound) {
  const int lane = threadIdx.x & 31;
  int lo = 0, hi = T;                     // t[k] < bound for k < lo;  hi == T or t[hi] >= bound
  while (hi - lo > 32) {
    const int step = (hi - lo + 31) / 32;
    const int idx = lo + lane * step;
    const bool ge = (idx < hi) ? (static_cast<long long>(t[idx]) >= bound) : true;
    const unsigned m = __ballot_sync(0xffffffffu, ge);
    const int f = m ? __ffs(m) - 1 : 32;
    if (f == 0) return lo;
    const int nlo = lo + (f - 1) * step + 1, nhi = lo + f * step;
    lo = nlo;
    if (nhi < hi) hi = nhi;
  }
  const int idx = lo + lane;
  const bool ge = (idx < hi) ? (static_cast<long long>(t[idx]) >= bound) : true;
  const unsigned m = __ballot_sync(0xffffffffu, ge);
  return m ? lo + __ffs(m) - 1 : hi;
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

// mode[item]: 0 / 1 = rolling quantile over the knot table (k_knot_table picked the list);
// 2 = constant cval[item].
// nan_fill (optional): value written instead of NaN when no output is valid.
__global__ void __launch_bounds__(RF_THREADS) k_rolling_floor(
    const BpmItem* __restrict__ items, KnotTable kt, const FloorMeta* __restrict__ meta, int window, double q,
    int run, FloorModeSrc ms, const double* __restrict__ cval,
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
  const int md = floor_mode_of(ms, item);
  const double nanv = nan_fill ? nan_fill[item] : __longlong_as_double(0x7ff8000000000000ll);
  if (md == 2) {
    const double c = cval[item];
    for (long long i = first; i < last; ++i) o[i] = c;
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
// One CTA per SM (RB_THREADS threads, ~215 KB of shared memory) owns `outs` consecutive outputs;
// the host sizes `outs` so that the grid is a whole number of waves of 148 CTAs.  The CTA
//   S1  materialises the n = outs + window - 1 interpolated samples its windows touch in shared
//       memory (float64 np.interp values, never written to HBM);
//   S1b estimates a PIVOT value from 16 probe windows (128 strided samples each, the element of
//       rank (q + margin) * 128): order statistic idx (and idx + 1) of every window of the tile
//       is expected at or below it, so only the samples <= pivot ("kept", ~1/3 for q = 0.2) have
//       to be sorted -- the others can never be the answer and only ever count as "above";
//   S2/S3 sample-sorts the kept samples: splitters = the part of a sorted 2048-sample below the
//       pivot, counting sort into buckets, exact order inside each (small) bucket by counting;
//       keeps perm[] (sorted position -> sample) and rank[] (sample -> sorted position, 0xFFFF
//       for samples above the pivot);
//   S4  builds a coarse (index chunk x rank band) prefix table so that a thread can place the
//       order statistic of its first window in O(log) steps;
//   S5  every thread slides over its run of outputs with a pointer p into the sorted order: the
//       entering / leaving sample changes the number of in-window entries before p by at most
//       one each, and p walks a few entries to the new order statistic.
// The pivot is validated lazily: if any walk runs off the end of the kept set (a window with
// fewer than idx + 2 kept samples) the CTA repeats S2..S5 with pivot = +inf, i.e. sorts
// everything.  Results are exact either way: the samples are the float64 np.interp values and
// pandas' interpolation between the two order statistics is evaluated unfused.
#ifndef BPM_RB_THREADS
#define BPM_RB_THREADS 1024
#endif
constexpr int RB_THREADS = BPM_RB_THREADS;
constexpr int RB_NCAP = RB_THREADS == 1024 ? 13312 : 14336;   // samples a CTA can stage (shared-memory budget)
constexpr int RB_S = 2048;             // sorted sample the splitters come from (power of two)
constexpr int RB_NSUP = 32;            // rank bands of the coarse table
constexpr int RB_CH = 32;              // samples per index chunk of the coarse table
constexpr int RB_MAXCH = RB_NCAP / RB_CH;
constexpr int RB_NPROBE = RB_THREADS / 32;
constexpr int RB_CW = RB_S / 2 / (RB_THREADS / 32);   // bitonic compare-exchanges per warp and step
constexpr int RB_PROBE_N = 64;         // samples per probe window (2 per lane)
constexpr unsigned short RB_ABOVE = 0xFFFFu;
constexpr int RB_NFINE = 4096;         // equal-width bins of the equalised bucketing (16 KB as uint32)
constexpr int RB_EQ_MAX = 384;         // more samples than this in one fine bin: use the splitter path

struct RbSortPhase {
  double piv[RB_S];                    // sorted splitter candidates, piv[RB_S-1] = +inf
  unsigned int hist[RB_S];
  unsigned short start[RB_S + 1];
};
struct RbSlidePhase {
  unsigned short pc[RB_MAXCH + 1][RB_NSUP];       // kept samples with index < c*ch and band <= s
  unsigned char ragged[RB_NSUP][RB_THREADS];      // per-thread band counts of a window's ragged ends (< 2 RB_CH)
};
constexpr int RB_KNOT_STAGE = 1280;    // knots of a tile staged in shared memory (min distance 15 => <= 888 for a full tile)
struct RbKnotStage {
  double v[RB_KNOT_STAGE], s[RB_KNOT_STAGE];
  int t[RB_KNOT_STAGE];
};
struct RbShared {
  double d[RB_NCAP];
  unsigned short rank[RB_NCAP];        // bucket id while sorting, then position in sorted order
  unsigned short perm[RB_NCAP + 8];    // 8-byte aligned; [nk, nk+4) padded with RB_ABOVE for the 4-entry loads
  union {
    RbSortPhase sort;
    RbSlidePhase slide;
    RbKnotStage knots;
  } u;
  unsigned long long pivot_key;        // max over the probes (order-preserving key of a double)
  int scan_tmp[40];
  int fail;
  int probes_ok;
  int eq_max;                          // fullest fine bin of the current attempt
  unsigned long long vmin_key, vmax_key;   // range of the staged samples
  int k_lo, k_hi;                      // knots around the tile's first / last sample
  int kf, kl;                          // knots positioned inside the tile's output range
};

// in-window test on a perm entry: j in [a, a + span]  (RB_ABOVE never is: 65535 - a > span)
__device__ __forceinline__ bool rb_in(unsigned int j, int a, unsigned int span) {
  return (j - static_cast<unsigned int>(a)) <= span;
}
#ifdef BPM_DEBUG_COUNTERS
#define RB_TICK(i) do { __syncthreads(); if (threadIdx.x == 0) { long long t_ = clock64(); atomicAdd(&g_dbg[i], (unsigned long long)(t_ - t_phase)); t_phase = t_; } } while (0)
#else
#define RB_TICK(i)
#endif

__global__ void __launch_bounds__(RB_THREADS, 1) k_rolling_floor_blk(
    const BpmItem* __restrict__ items, KnotTable kt, const FloorMeta* __restrict__ meta, int window, double q,
    int outs, FloorModeSrc ms, const double* __restrict__ cval,
    const double* __restrict__ nan_fill, double* __restrict__ out, double* __restrict__ sparse_out) {
  extern __shared__ __align__(16) unsigned char rb_raw[];
  RbShared& sh = *reinterpret_cast<RbShared*>(rb_raw);
  const int item = blockIdx.y;
  const BpmItem it = items[item];
  const long long m = it.m;
  const long long blk_first = static_cast<long long>(blockIdx.x) * outs;
  if (blk_first >= m) return;
  const long long blk_last = min(m, blk_first + static_cast<long long>(outs));   // exclusive
  double* o = out + it.m_off;
  const int tid = threadIdx.x;
  const int md = floor_mode_of(ms, item);
  const double nanv = nan_fill ? nan_fill[item] : __longlong_as_double(0x7ff8000000000000ll);
  // SPARSE mode (sparse_out != nullptr): outputs are wanted only AT the knots (the draft floor is
  // read nowhere else, :1093) and are written per knot number; tiles without knots do nothing.
  const bool sparse = (sparse_out != nullptr);
  if (md == 2) {
    if (sparse) return;                                       // the "<5 troughs" path never reads the draft
    const double c = cval[item];
    for (long long i = blk_first + tid; i < blk_last; i += RB_THREADS) o[i] = c;
    return;
  }
  const FloorMeta mt = meta[item];
  if (!mt.valid) {
    if (sparse) {
      if (blockIdx.x == 0) for (long long k = tid; k < mt.n_knots; k += RB_THREADS) sparse_out[it.m_off + k] = nanv;
      return;
    }
    for (long long i = blk_first + tid; i < blk_last; i += RB_THREADS) o[i] = nanv;
    return;
  }
#ifdef BPM_DEBUG_COUNTERS
  long long t_phase = clock64();
#endif
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
  if (tid < 128) {
    // four warps, one boundary each: the knots around the tile's first / last staged sample (last knot
    // <= x: one before the first knot >= x + 1; t[0] <= x0 holds) and the knots inside its output range
    const int w = tid >> 5;
    const long long bound = (w == 0) ? static_cast<long long>(x0) + 1 : (w == 1) ? static_cast<long long>(x1) + 1
                                                                      : (w == 2) ? blk_first : blk_last;
    const int r = warp_first_knot_ge(c.t, c.T, bound);
    if ((tid & 31) == 0) {
      if (w == 0) sh.k_lo = r - 1;
      else if (w == 1) sh.k_hi = r - 1;
      else if (w == 2) sh.kf = r;
      else sh.kl = r;
    }
  }
  __syncthreads();
  const int kf = sh.kf, kcnt = sh.kl - sh.kf;
  if (sparse && kcnt <= 0) return;
  {
    // the tile's knots [k_lo, k_hi + 1] go to shared memory first (the sort phase's area is idle): the
    // per-thread search and the interpolation then never wait for global memory
    RbKnotStage& ks = sh.u.knots;
    const int k_lo = sh.k_lo;
    const int nkn = min(sh.k_hi + 1, c.T - 1) - k_lo + 1;
    const bool staged = nkn <= RB_KNOT_STAGE;
    if (staged) {
      for (int k = tid; k < nkn; k += RB_THREADS) {
        ks.t[k] = c.t[k_lo + k];
        ks.v[k] = c.v[k_lo + k];
        ks.s[k] = c.s[k_lo + k];
      }
      __syncthreads();
    }
    WinCtx cs = c;
    if (staged) { cs.t = ks.t - k_lo; cs.v = ks.v - k_lo; cs.s = ks.s - k_lo; }
    const int per = (n + RB_THREADS - 1) / RB_THREADS;
    const int j0 = tid * per, j1 = min(n, j0 + per);
    if (j0 < j1) {
      int k = k_lo;
      {
        int hi = sh.k_hi;                                     // last knot <= x0 + j0, searched inside the tile's knots
        const int xi = x0 + j0;
        while (k < hi) {
          const int mid = (k + hi + 1) >> 1;
          if (cs.t[mid] <= xi) k = mid; else hi = mid - 1;
        }
      }
      for (int j = j0; j < j1; ++j) {
        const int x = x0 + j;
        while (k + 1 < c.T && cs.t[k + 1] <= x) ++k;
        sh.d[j] = val_at(cs, k, x);
      }
    }
  }
  if (tid == 0) { sh.pivot_key = 0ull; sh.fail = 0; sh.probes_ok = 1; sh.vmin_key = ~0ull; sh.vmax_key = 0ull; }
  __syncthreads();
  {
    unsigned long long kmin = ~0ull, kmax = 0ull;
    for (int j = tid; j < n; j += RB_THREADS) {
      const unsigned long long k = f64_key(sh.d[j]);
      kmin = k < kmin ? k : kmin;
      kmax = k > kmax ? k : kmax;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const unsigned long long a = __shfl_xor_sync(0xffffffffu, kmin, o), b = __shfl_xor_sync(0xffffffffu, kmax, o);
      kmin = a < kmin ? a : kmin;
      kmax = b > kmax ? b : kmax;
    }
    if ((tid & 31) == 0) { atomicMin(&sh.vmin_key, kmin); atomicMax(&sh.vmax_key, kmax); }
  }
  __syncthreads();
  const double vmin_all = key_f64(sh.vmin_key), vmax_all = key_f64(sh.vmax_key);
  RB_TICK(0);

  // ---- S1b: pivot estimate.  Warp w probes the window of output ie0 + (2w+1)(ie1-ie0)/(2 NPROBE).
  const double keep_frac = q + 0.15;
  bool use_pivot = keep_frac < 0.8;
  if (use_pivot) {
    const int w = tid >> 5, lane = tid & 31;
    const int ip = ie0 + static_cast<int>((static_cast<long long>(2 * w + 1) * (ie1 - ie0)) / (2 * RB_NPROBE));
    int pb = ip + off; if (pb > mi - 1) pb = mi - 1;
    int pa = ip - left; if (pa < 0) pa = 0; if (pa < t0) pa = t0;
    const int nwp = pb - pa + 1;
    if (nwp < 4 * RB_PROBE_N) {
      if (lane == 0) sh.probes_ok = 0;                          // tiny windows: not worth a pivot
    } else {
      // ranks are taken on the high 32 bits of the order-preserving key (integer compares; an
      // estimate does not need the low mantissa bits), ties broken by sample number
      double mine[2];
      unsigned int hk[2];
      int below[2];
#pragma unroll
      for (int u2 = 0; u2 < 2; ++u2) {
        const int k2 = lane * 2 + u2;
        mine[u2] = sh.d[(pa - x0) + static_cast<int>((static_cast<long long>(k2) * (nwp - 1)) / (RB_PROBE_N - 1))];
        hk[u2] = static_cast<unsigned int>(f64_key(mine[u2]) >> 32);
        below[u2] = 0;
      }
      for (int src = 0; src < 32; ++src) {
#pragma unroll
        for (int v2 = 0; v2 < 2; ++v2) {
          const unsigned int y = __shfl_sync(0xffffffffu, hk[v2], src);
          const int ky = src * 2 + v2;
#pragma unroll
          for (int u2 = 0; u2 < 2; ++u2) {
            const int k2 = lane * 2 + u2;
            below[u2] += (y < hk[u2] || (y == hk[u2] && ky < k2)) ? 1 : 0;
          }
        }
      }
      int target = static_cast<int>(ceil(keep_frac * RB_PROBE_N));
      if (target > RB_PROBE_N - 1) target = RB_PROBE_N - 1;
#pragma unroll
      for (int u2 = 0; u2 < 2; ++u2)
        if (below[u2] == target) atomicMax(&sh.pivot_key, f64_key(mine[u2]));
    }
  }
  __syncthreads();
  double pivot = INFINITY;
  if (use_pivot && sh.probes_ok) pivot = key_f64(sh.pivot_key);
  RB_TICK(1);

  const int outs_here = static_cast<int>(blk_last - blk_first);
  const int run = (outs_here + RB_THREADS - 1) / RB_THREADS;
  const long long first = blk_first + static_cast<long long>(tid) * run;
  const long long last = min(blk_last, first + run);

  for (int attempt = 0; attempt < 2; ++attempt) {
    RbSortPhase& so = sh.u.sort;
    // ---- S2: bucket ids of the kept samples (monotone in the value, equal values share a bucket) + histogram.
    // Fast path: a 4096-bin equal-width histogram over [vmin, min(pivot, vmax)] is EQUALISED through its
    // prefix sum -- bucket = cdf[bin] >> shift -- which gives buckets of a few samples without sorting
    // anything.  It cannot split one fine bin, so when a single bin holds more than RB_EQ_MAX samples
    // (a heavily skewed tile) the tile takes the splitter path instead: splitters = the part of a
    // bitonic-sorted 2048-sample below the pivot, bucket by binary search.
    unsigned int* fine = reinterpret_cast<unsigned int*>(so.piv);          // [RB_NFINE], aliases the splitters
    for (int t = tid; t < RB_NFINE; t += RB_THREADS) fine[t] = 0;
    for (int t = tid; t < RB_S; t += RB_THREADS) so.hist[t] = 0;
    if (tid == 0) sh.eq_max = 0;
    __syncthreads();
    const double vhi = pivot < vmax_all ? pivot : vmax_all;
    const double eq_scale = (vhi > vmin_all) ? (static_cast<double>(RB_NFINE) / (vhi - vmin_all)) : 0.0;
    for (int j = tid; j < n; j += RB_THREADS) {
      const double x = sh.d[j];
      unsigned short fb = RB_ABOVE;
      if (x <= pivot) {
        int bi = static_cast<int>((x - vmin_all) * eq_scale);
        if (bi > RB_NFINE - 1) bi = RB_NFINE - 1;
        if (bi < 0) bi = 0;
        fb = static_cast<unsigned short>(bi);
        atomicAdd(&fine[bi], 1u);
      }
      sh.rank[j] = fb;
    }
    __syncthreads();
    int nbuckets;
    {
      constexpr int PERF = RB_NFINE / RB_THREADS;
      unsigned int loc[PERF];
      int sum = 0, mx = 0;
#pragma unroll
      for (int u2 = 0; u2 < PERF; ++u2) {
        loc[u2] = fine[tid * PERF + u2];
        sum += loc[u2];
        mx = mx > static_cast<int>(loc[u2]) ? mx : static_cast<int>(loc[u2]);
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) { const int t = __shfl_xor_sync(0xffffffffu, mx, o); mx = t > mx ? t : mx; }
      if ((tid & 31) == 0 && mx > 0) atomicMax(&sh.eq_max, mx);
      int nk_eq;
      int ex = block_exclusive_scan(sum, &nk_eq, sh.scan_tmp);               // (contains the barriers)
#pragma unroll
      for (int u2 = 0; u2 < PERF; ++u2) { fine[tid * PERF + u2] = ex; ex += loc[u2]; }
      __syncthreads();
      if (sh.eq_max <= RB_EQ_MAX) {
        int shift = 0;
        while (((nk_eq > 0 ? nk_eq - 1 : 0) >> shift) >= RB_S) ++shift;
        nbuckets = ((nk_eq > 0 ? nk_eq - 1 : 0) >> shift) + 1;
        for (int j = tid; j < n; j += RB_THREADS) {
          const unsigned short fb = sh.rank[j];
          if (fb != RB_ABOVE) {
            const unsigned int bk = fine[fb] >> shift;
            sh.rank[j] = static_cast<unsigned short>(bk);
            atomicAdd(&so.hist[bk], 1u);
          }
        }
        RB_TICK(2);
      } else {
        DBG(14);
        __syncthreads();                                         // everyone has read eq_max / the cdf
        for (int t = tid; t < RB_S; t += RB_THREADS)
          so.piv[t] = (t < RB_S - 1) ? sh.d[static_cast<int>((static_cast<long long>(t) * (n - 1)) / (RB_S - 2))] : INFINITY;
        __syncthreads();
        // bitonic sort; compare-exchange c of a step touches elements inside the (2 RB_CW)-element block
        // of c / RB_CW whenever j2 <= RB_CW, so warp w owns CEs [RB_CW w, RB_CW (w+1)) and those steps
        // only need __syncwarp
        for (int k2 = 2; k2 <= RB_S; k2 <<= 1) {
          for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
#pragma unroll
            for (int h = 0; h < RB_S / 2 / RB_THREADS; ++h) {
              const int t = (tid >> 5) * RB_CW + h * 32 + (tid & 31);
              const int lo_i = ((t & ~(j2 - 1)) << 1) | (t & (j2 - 1));      // j2 is a power of two
              const int hi_i = lo_i + j2;
              const bool up = ((lo_i & k2) == 0);
              const double x = so.piv[lo_i], y = so.piv[hi_i];
              if ((x > y) == up) { so.piv[lo_i] = y; so.piv[hi_i] = x; }
            }
            if (j2 > RB_CW || (j2 == 1 && k2 > RB_CW)) __syncthreads(); else __syncwarp();
          }
        }
        __syncthreads();
        // usable splitters: piv[0 .. nsplit) are <= pivot (the +inf sentinel is never usable)
        int nsplit;
        {
          int lo_i = 0, hi_i = RB_S - 1;
          while (lo_i < hi_i) {
            const int mid = (lo_i + hi_i) >> 1;
            if (so.piv[mid] <= pivot) lo_i = mid + 1; else hi_i = mid;
          }
          nsplit = lo_i;                                           // in [0, RB_S - 1]
        }
        nbuckets = nsplit + 1;
        RB_TICK(2);
        for (int j = tid; j < n; j += RB_THREADS) {
          const double x = sh.d[j];
          unsigned short bk = RB_ABOVE;
          if (x <= pivot) {
            // bucket = number of usable splitters strictly below the value (equal values share a bucket)
            int lo_i = 0, hi_i = nsplit;
            while (lo_i < hi_i) {
              const int mid = (lo_i + hi_i) >> 1;
              if (so.piv[mid] < x) lo_i = mid + 1; else hi_i = mid;
            }
            bk = static_cast<unsigned short>(lo_i);
            atomicAdd(&so.hist[lo_i], 1u);
          }
          sh.rank[j] = bk;
        }
      }
    }
    __syncthreads();
    RB_TICK(3);
    // ---- S3: exclusive scan of the histogram -> start[], scatter (hist becomes the cursor),
    //          exact order inside every bucket, inverse permutation
    int nk;
    {
      constexpr int PER = RB_S / RB_THREADS;
      unsigned int loc[PER];
      int sum = 0;
#pragma unroll
      for (int u2 = 0; u2 < PER; ++u2) {
        const int b2 = tid * PER + u2;
        loc[u2] = (b2 < nbuckets) ? so.hist[b2] : 0u;
        sum += loc[u2];
      }
      int ex = block_exclusive_scan(sum, &nk, sh.scan_tmp);
#pragma unroll
      for (int u2 = 0; u2 < PER; ++u2) {
        const int b2 = tid * PER + u2;
        so.start[b2] = static_cast<unsigned short>(ex);
        so.hist[b2] = ex;
        ex += loc[u2];
      }
      if (tid == RB_THREADS - 1) so.start[RB_S] = static_cast<unsigned short>(ex);
    }
    __syncthreads();
    for (int j = tid; j < n; j += RB_THREADS) {
      const unsigned short bk = sh.rank[j];
      if (bk != RB_ABOVE) {
        const unsigned int pos = atomicAdd(&so.hist[bk], 1u);
        sh.perm[pos] = static_cast<unsigned short>(j);
      }
    }
    __syncthreads();
    RB_TICK(4);
    // order inside the buckets by counting: final position = bucket start + #{bucket members that
    // sort before this sample}.  Work is dealt out in bucket order (thread <-> scattered position),
    // so the lanes of a warp walk the same bucket: equal trip counts, broadcast shared-memory reads.
    // Only the owner of sample j reads rank[j] (its bucket), so it can be overwritten in place.
    for (int e = tid; e < nk; e += RB_THREADS) {
      const int j = sh.perm[e];
      const int g = sh.rank[j];
      const int e0 = so.start[g], e1 = so.start[g + 1];
      const double x = sh.d[j];
      int before = 0;
      for (int f = e0; f < e1; ++f) {
        const int y = sh.perm[f];
        const double dy = sh.d[y];
        before += (dy < x || (dy == x && y < j)) ? 1 : 0;
      }
      sh.rank[j] = static_cast<unsigned short>(e0 + before);
    }
    __syncthreads();
    for (int j = tid; j < n; j += RB_THREADS) {
      const unsigned short r2 = sh.rank[j];
      if (r2 != RB_ABOVE) sh.perm[r2] = static_cast<unsigned short>(j);
    }
    if (tid < 8) sh.perm[nk + tid] = RB_ABOVE;
    __syncthreads();                                           // sort phase of the union is dead from here
    RB_TICK(5);

    // ---- S4: coarse table over (index chunk, rank band) of the kept samples
    RbSlidePhase& sl = sh.u.slide;
    constexpr int ch = RB_CH;
    const int nch = (n + ch - 1) / ch;
    int bshift = 0;
    while (((nk > 0 ? nk - 1 : 0) >> bshift) >= RB_NSUP) ++bshift;   // band = rank >> bshift  in [0, RB_NSUP)
    for (int t = tid; t < (nch + 1) * RB_NSUP; t += RB_THREADS) (&sl.pc[0][0])[t] = 0;
    __syncthreads();
    for (int j = tid; j < n; j += RB_THREADS) {
      const unsigned short r2 = sh.rank[j];
      if (r2 == RB_ABOVE) continue;
      // counts land one row down so that an inclusive column scan yields the exclusive prefix
      unsigned short* cell = &sl.pc[j / ch + 1][r2 >> bshift];
      // 16-bit shared atomics do not exist: add into the containing 32-bit word
      unsigned int* word = reinterpret_cast<unsigned int*>(reinterpret_cast<uintptr_t>(cell) & ~uintptr_t(3));
      const unsigned int add = (reinterpret_cast<uintptr_t>(cell) & 2) ? (1u << 16) : 1u;
      atomicAdd(word, add);
    }
    __syncthreads();
    if (tid < RB_NSUP) {
      unsigned int acc = 0;
      for (int r2 = 0; r2 <= nch; ++r2) { acc += sl.pc[r2][tid]; sl.pc[r2][tid] = static_cast<unsigned short>(acc); }
    }
    __syncthreads();
    for (int r2 = tid; r2 <= nch; r2 += RB_THREADS) {
      unsigned int acc = 0;
      for (int s2 = 0; s2 < RB_NSUP; ++s2) { acc += sl.pc[r2][s2]; sl.pc[r2][s2] = static_cast<unsigned short>(acc); }
    }
    __syncthreads();
    RB_TICK(6);

    // ---- S5: every thread slides over its run
    bool failed = false;
    const long long nsteps = sparse ? ((kcnt > tid) ? (kcnt - tid + RB_THREADS - 1) / RB_THREADS : 0)
                                    : ((first < blk_last) ? (last - first) : 0);
    if (nsteps > 0) {
      bool have = false;
      int p = 0, cb = 0;            // position in sorted order; in-window entries at positions < p
      double prev_out = 0.0;
      int prev_i = -2, a_prev = 0, b_prev = 0;
      for (long long step = 0; step < nsteps && !failed; ++step) {
        const long long ks = kf + tid + step * RB_THREADS;      // sparse: the knot this output belongs to
        const long long io = sparse ? static_cast<long long>(c.t[ks]) : first + step;
        double* dst = sparse ? (sparse_out + it.m_off + ks) : (o + io);
        int i = static_cast<int>(io);
        if (i < mt.iv0) i = static_cast<int>(mt.iv0);
        if (i > mt.iv1) i = static_cast<int>(mt.iv1);
        if (i == prev_i) { *dst = prev_out; continue; }
        int bb = i + off; if (bb > mi - 1) bb = mi - 1;
        int aa = i - left; if (aa < 0) aa = 0; if (aa < t0) aa = t0;
        const int nw = bb - aa + 1;
        const double fq = __dmul_rn(q, static_cast<double>(nw - 1));      // pandas: q * (nobs - 1)
        const int idx = static_cast<int>(fq);
        const double frac = __dsub_rn(fq, static_cast<double>(idx));
        const int a = aa - x0, b = bb - x0;                        // local indices, inclusive
        if (!have || i != prev_i + 1) {
          // #{kept window samples in rank bands <= s}: whole chunks from the table + the ragged ends.
          // The ragged ends are counted once into a per-thread band histogram (cumulated in place).
          const int ca = (a + ch - 1) / ch;                        // first whole chunk
          const int cbk = (b + 1) / ch;                            // one past the last whole chunk
          unsigned char* rg = &sl.ragged[0][0] + tid;              // [RB_NSUP][RB_THREADS]
          for (int s2 = 0; s2 < RB_NSUP; ++s2) rg[s2 * RB_THREADS] = 0;
          auto tally = [&](int j) {
            const unsigned short r2 = sh.rank[j];
            if (r2 != RB_ABOVE) rg[(r2 >> bshift) * RB_THREADS] += 1;
          };
          if (cbk > ca) {
            for (int j = a; j < ca * ch; ++j) tally(j);
            for (int j = cbk * ch; j <= b; ++j) tally(j);
          } else {
            for (int j = a; j <= b; ++j) tally(j);
          }
          {
            unsigned int acc = 0;
            for (int s2 = 0; s2 < RB_NSUP; ++s2) { acc += rg[s2 * RB_THREADS]; rg[s2 * RB_THREADS] = static_cast<unsigned char>(acc); }
          }
          auto cum = [&](int s_) -> int {
            if (s_ < 0) return 0;
            int cs_ = rg[s_ * RB_THREADS];
            if (cbk > ca) cs_ += static_cast<int>(sl.pc[cbk][s_]) - static_cast<int>(sl.pc[ca][s_]);
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
          if (b > b_prev) cb += (sh.rank[b] < p) ? 1 : 0;          // RB_ABOVE is never < p
          if (a > a_prev) cb -= (sh.rank[a_prev] < p) ? 1 : 0;
        }
        // walk p to the in-window entry that has exactly idx in-window entries before it
        const unsigned int span = static_cast<unsigned int>(b - a);
        while (cb > idx) {
          --p;
          if (rb_in(sh.perm[p], a, span)) --cb;
        }
        while (true) {
          const bool in = rb_in(sh.perm[p], a, span);              // perm[nk] is RB_ABOVE: never in
          if (in && cb == idx) break;
          if (p >= nk) { failed = true; break; }                   // fewer than idx+1 kept samples in this window
          if (in) ++cb;
          ++p;
        }
        if (failed) break;
        const double vlow = sh.d[sh.perm[p]];
        double res = vlow;
        if (fq != static_cast<double>(idx)) {
          int p2 = p + 1;
          while (!rb_in(sh.perm[p2], a, span)) {                   // next in-window entry (idx + 1 < nw here)
            if (p2 >= nk) { failed = true; break; }
            ++p2;
          }
          if (failed) break;
          const double vhigh = sh.d[sh.perm[p2]];
          res = __dadd_rn(vlow, __dmul_rn(__dsub_rn(vhigh, vlow), frac));
        }
        *dst = res;
        prev_out = res; prev_i = i; a_prev = a; b_prev = b;
      }
    }
    if (failed) sh.fail = 1;
    __syncthreads();
    RB_TICK(7);
    if (!sh.fail) break;
    // the pivot was too low for some window: sort everything (attempt 1 cannot fail)
    DBG(15);
    __syncthreads();
    if (tid == 0) sh.fail = 0;
    pivot = INFINITY;
    __syncthreads();
  }
}

// K7: keep trough t iff the draft floor there is not NaN and env[t] <= mult * floor[t]
// (bpm_analysis.py:1090-1097), written in order by a single-pass compaction (decoupled look-back
// over tile counts, common.cuh).  keep_few: a recording with fewer than 5 troughs keeps them all
// (the "<5 troughs" path returns the unsanitised list, :1077).
constexpr int SZ_THREADS = 256;
constexpr int SZ_PER = 4;
constexpr int SZ_TILE = SZ_THREADS * SZ_PER;

__global__ void __launch_bounds__(SZ_THREADS) k_sanitize_compact(
    const double* __restrict__ env, const double* __restrict__ draft, const int64_t* __restrict__ troughs,
    const int64_t* __restrict__ trough_count, int keep_few, const BpmItem* __restrict__ items,
    double mult, int draft_by_knot, unsigned long long* __restrict__ status, int64_t status_stride,
    int64_t* __restrict__ kept_out, int64_t* __restrict__ kept_count) {
  __shared__ int s_scan[34];
  __shared__ long long s_off;
  const int item = blockIdx.y;
  const BpmItem it = items[item];
  const long long nt = trough_count[item];
  const long long k0 = static_cast<long long>(blockIdx.x) * SZ_TILE;
  if (k0 >= nt) {
    if (nt == 0 && blockIdx.x == 0 && threadIdx.x == 0) kept_count[item] = 0;
    return;
  }
  const bool all = keep_few && nt < 5;                       // the "<5 troughs" path returns them all (:1077)
  int64_t t[SZ_PER];
  unsigned mask = 0;
#pragma unroll
  for (int u = 0; u < SZ_PER; ++u) {
    const long long k = k0 + threadIdx.x * SZ_PER + u;
    t[u] = 0;
    if (k >= nt) continue;
    t[u] = troughs[it.m_off + k];
    bool keep = true;
    if (!all) {
      const double f = draft_by_knot ? draft[it.m_off + k] : draft[it.m_off + t[u]];   // one value per trough, or dense
      keep = !isnan(f) && env[it.m_off + t[u]] <= __dmul_rn(mult, f);
    }
    if (keep) mask |= 1u << u;
  }
  int total;
  const int ex = block_exclusive_scan(__popc(mask), &total, s_scan);
  if (threadIdx.x < 32) {
    const long long off = lookback_exclusive(status + static_cast<int64_t>(item) * status_stride, blockIdx.x, total);
    if (threadIdx.x == 0) s_off = off;
  }
  __syncthreads();
  int64_t o = it.m_off + s_off + ex;
#pragma unroll
  for (int u = 0; u < SZ_PER; ++u)
    if ((mask >> u) & 1) kept_out[o++] = t[u];
  if (threadIdx.x == 0 && k0 + SZ_TILE >= nt) kept_count[item] = s_off + total;
}

// host launchers of the small kernels above (called from pipeline.cu: every translation unit
// launches only kernels it defines, so no relocatable device code is needed)
size_t sanitize_workspace_bytes(int64_t total_m, int n) {
  Workspace ws(nullptr, 0);
  ws.take<unsigned long long>(static_cast<size_t>(n) * ((total_m / 2 + 2) / SZ_TILE + 2));
  return ws.used;
}

int sanitize_run(const double* env, const double* draft, const int64_t* troughs, const int64_t* trough_count,
                 int keep_few, const BpmItem* items, const BatchShape& sh, double mult, int draft_by_knot,
                 int64_t* kept_out, int64_t* kept_count, Workspace& ws, cudaStream_t st) {
  if (!env || !draft || !troughs || !trough_count || !items || !kept_out || !kept_count) return BPM_ERR_ARG;
  const int64_t max_t = sh.max_m / 2 + 2;
  const int64_t stride = max_t / SZ_TILE + 2;
  unsigned long long* status = ws.take<unsigned long long>(static_cast<size_t>(sh.n_items) * stride);
  if (ws.overflow) return BPM_ERR_WORKSPACE;
  if (cudaMemsetAsync(status, 0, sizeof(unsigned long long) * sh.n_items * stride, st) != cudaSuccess) return BPM_ERR_CUDA;
  BPM_KERNEL(k_sanitize_compact);
  k_sanitize_compact<<<dim3(cdiv(max_t, SZ_TILE), sh.n_items), SZ_THREADS, 0, st>>>(
      env, draft, troughs, trough_count, keep_few, items, mult, draft_by_knot, status, stride, kept_out, kept_count);
  BPM_LAUNCH_OK();
  return BPM_OK;
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

// can the draft floor be computed at the knots only (block-cooperative kernel available)?
bool rolling_floor_sparse_ok(int window) { return static_cast<int64_t>(RB_NCAP) - window + 1 >= RB_THREADS; }

// sparse_out != nullptr (only with rolling_floor_sparse_ok): values at the knots, per knot number;
// `out` is then not written.  alt_knots / alt_count: the list used for items whose mode is 1;
// mode_n_all / mode_n_kept: the trough counts the mode is derived from (floor_mode_of).
int rolling_floor_run(const double* env, const int64_t* knots, const int64_t* knot_count, const BpmItem* items,
                      const BatchShape& sh, int window, double q, const int64_t* mode_n_all,
                      const int64_t* mode_n_kept, const int64_t* alt_knots, const int64_t* alt_count, const double* cval,
                      const double* nan_fill, int64_t* total_out, int64_t* mode_out, double* out, double* sparse_out,
                      Workspace& ws, cudaStream_t st) {
  if (!env || !knots || !knot_count || !items || (!out && !sparse_out) || sh.n_items <= 0 || window < 1) return BPM_ERR_ARG;
  if (sparse_out && !rolling_floor_sparse_ok(window)) return BPM_ERR_ARG;
  FloorBuffers b;
  BPM_TRY(carve_floor(ws, sh.total_m, sh.n_items, &b));
  const int64_t max_k = sh.max_m / 2 + 2;
  BPM_KERNEL(k_knot_table);
  const FloorModeSrc ms{mode_n_all, mode_n_kept};
  k_knot_table<<<dim3(cdiv(max_k, 256), sh.n_items), 256, 0, st>>>(env, knots, knot_count, alt_knots, alt_count, ms,
                                                                    items, window, b.kt32, b.kv, b.ks, b.kinv, b.kend,
                                                                    b.meta, total_out, mode_out);
  BPM_LAUNCH_OK();
  // outputs per thread of the per-thread kernel: long enough to amortise the bisection
  int64_t run = sh.total_m / (148 * 1024);
  if (run < 48) run = 48;
  if (run > 512) run = 512;
  if (sh.max_m >= (1ll << 31)) return BPM_ERR_ARG;
  KnotTable kt{b.kt32, b.kv, b.ks, b.kinv, b.kend};
  // block-cooperative kernel when a CTA can stage its windows' samples; else the per-thread kernel
  const int64_t max_outs = static_cast<int64_t>(RB_NCAP) - window + 1;
  if (max_outs >= RB_THREADS) {
    // one CTA per SM: size the tiles so that the grid is a whole number of waves
    const int64_t sms = 148;
    const int64_t waves = (sh.total_m + sms * max_outs - 1) / (sms * max_outs);
    int64_t outs = (sh.total_m + sms * waves - 1) / (sms * waves);
    if (outs < RB_THREADS) outs = RB_THREADS;
    if (outs > max_outs) outs = max_outs;
    cudaFuncSetAttribute(k_rolling_floor_blk, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(RbShared)));
    BPM_KERNEL(k_rolling_floor_blk);
    k_rolling_floor_blk<<<dim3(cdiv(sh.max_m, outs), sh.n_items), RB_THREADS, sizeof(RbShared), st>>>(
        items, kt, b.meta, window, q, static_cast<int>(outs), ms, cval, nan_fill, out, sparse_out);
    BPM_LAUNCH_OK();
    return BPM_OK;
  }
  const size_t smem = sizeof(short) * RF_SMAX * RF_THREADS;
  cudaFuncSetAttribute(k_rolling_floor, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  BPM_KERNEL(k_rolling_floor);
  k_rolling_floor<<<dim3(cdiv(sh.max_m, RF_THREADS * run), sh.n_items), RF_THREADS, smem, st>>>(
      items, kt, b.meta, window, q, static_cast<int>(run), ms, cval, nan_fill, out);
  BPM_LAUNCH_OK();
  return BPM_OK;
}

}  // namespace bpm

#ifdef BPM_DEBUG_COUNTERS
extern "C" int bpm_debug_counters(unsigned long long* out_host, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out_host, bpm::g_dbg, sizeof(unsigned long long) * 16);
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(bpm::g_dbg, z, sizeof(z)); }
  return 0;
}
#endif
