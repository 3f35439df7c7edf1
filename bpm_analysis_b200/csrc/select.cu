// K3: np.quantile(x, q), method 'linear' (bpm_analysis.py:225, :1067, :1075, :1114;
// numpy/lib/_function_base_impl.py:126-129, 4657-4678).
//
// Exact order statistics by most-significant-digit radix select on the order-preserving
// 64-bit image of each float64: two 11-bit digit passes, then the (tiny) bucket holding the
// target is collected and sorted by one CTA (a bucket larger than SEL_CAP -- massive duplicates
// -- falls back to the remaining digit passes 11,11,11,9 inside that CTA), then numpy's
// two-branch lerp evaluated with unfused float64 operations so the result is bit-identical to
// numpy's.  Up to SEL_MAXQ quantile levels of the same data are
// resolved in the same passes (the noise-floor stage needs q(trough_prominence) always and
// q(noise_floor_quantile) only for recordings with fewer than 5 troughs: `cond` switches a
// level off per recording).  Histogram updates are warp-aggregated (envelope values share
// their leading digits, so plain shared-memory atomics would serialise).
#include "common.cuh"

namespace bpm {

constexpr int SEL_PASSES = 6;
constexpr int SEL_BINS = 2048;
constexpr int SEL_THREADS = 256;
constexpr int SEL_MAXQ = 3;
constexpr int SEL_BATCH = 8;                // loads in flight per thread in the sweeps

__host__ __device__ __forceinline__ int sel_shift(int p) { return p < 5 ? 53 - 11 * p : 0; }
__host__ __device__ __forceinline__ int sel_bits(int p) { return p < 5 ? 11 : 9; }

struct SelState {
  unsigned long long prefix;   // digits resolved so far (right aligned)
  long long rank;              // rank still to resolve inside the prefix bucket
  long long below;             // elements strictly below the bucket
  long long count;             // elements in the bucket
  unsigned long long next_key; // smallest key above the selected one (k_select_next)
  long long k;                 // target rank (0-based)
  double gamma;                // fractional part of the virtual index
  long long active;            // 0: skipped for this item
};

struct SelLevels {
  int nq;
  double q[SEL_MAXQ];
  const int* cond[SEL_MAXQ];   // per recording on/off (nullptr: always on)
  double* out[SEL_MAXQ];       // per recording result
};

// index helpers: states[pass][level][item], hist[item][level][pass][bin]
__device__ __forceinline__ size_t st_idx(int pass, int lvl, int item, int n_items) {
  return (static_cast<size_t>(pass) * SEL_MAXQ + lvl) * n_items + item;
}
__device__ __forceinline__ size_t hist_idx(int item, int lvl, int pass) {
  return ((static_cast<size_t>(item) * SEL_MAXQ + lvl) * SEL_PASSES + pass) * SEL_BINS;
}

// state after resolving one pass, from the state before it and that pass's histogram
// (block-wide; every block computes it redundantly)
__device__ void sel_advance(const SelState& prev, const unsigned int* __restrict__ hist, int bits, SelState* out_shared,
                            long long* s_cum /* SEL_THREADS + 1 */) {
  const int bins = 1 << bits;
  const int per = (bins + SEL_THREADS - 1) / SEL_THREADS;
  const int b0 = threadIdx.x * per;
  long long local = 0;
  for (int b = b0; b < min(b0 + per, bins); ++b) local += hist[b];
  s_cum[threadIdx.x + 1] = local;
  if (threadIdx.x == 0) s_cum[0] = 0;
  __syncthreads();
  if (threadIdx.x < 32) {
    // inclusive scan of SEL_THREADS partial sums by one warp, 8 per lane
    constexpr int PW = SEL_THREADS / 32;
    long long v[PW];
    long long s = 0;
#pragma unroll
    for (int u = 0; u < PW; ++u) { s += s_cum[1 + threadIdx.x * PW + u]; v[u] = s; }
    long long inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, inc, o);
      if (threadIdx.x >= o) inc += t;
    }
    const long long base = inc - s;
#pragma unroll
    for (int u = 0; u < PW; ++u) s_cum[1 + threadIdx.x * PW + u] = base + v[u];
  }
  __syncthreads();
  const long long rank = prev.rank;
  const long long lo = s_cum[threadIdx.x], hi = s_cum[threadIdx.x + 1];
  if (rank >= lo && rank < hi) {
    long long c = lo;
    for (int b = b0; b < min(b0 + per, bins); ++b) {
      const long long h = hist[b];
      if (rank < c + h) {
        out_shared->prefix = (prev.prefix << bits) | static_cast<unsigned long long>(b);
        out_shared->rank = rank - c;
        out_shared->below = prev.below + c;
        out_shared->count = h;
        break;
      }
      c += h;
    }
    out_shared->k = prev.k;
    out_shared->gamma = prev.gamma;
    out_shared->active = prev.active;
    out_shared->next_key = ~0ull;
  }
  __syncthreads();
}

// state before the first digit pass: target rank k = floor((n - 1) q), fraction gamma (numpy's virtual index)
__device__ __forceinline__ SelState sel_initial(long long n, double q, bool on) {
  SelState s;
  const double v = __dmul_rn(static_cast<double>(n - 1), q);     // (n - 1) * q
  const double fl = floor(v);
  s.prefix = 0;
  s.k = static_cast<long long>(fl);
  if (s.k > n - 1) s.k = n - 1;
  if (s.k < 0) s.k = 0;
  s.rank = s.k;
  s.below = 0;
  s.count = n;
  s.gamma = __dsub_rn(v, fl);
  s.next_key = ~0ull;
  s.active = (n > 0 && on) ? 1 : 0;
  return s;
}

// add 1 to s_hist[bin] for every lane with `on`.  Envelope values share their leading digits (a warp's
// 32 consecutive samples usually fall into ONE bin of the first digit), later digits are close to
// random: up to ROUNDS distinct bins per warp (two in the first pass, one later -- enough for a warp of
// equal samples) are added by one lane each with the group's size, whatever is left goes through plain
// shared-memory atomics (rarely conflicting).  Called by whole warps.
template <int ROUNDS>
__device__ __forceinline__ void warp_hist_add(unsigned int* s_hist, bool on, unsigned int bin, int lane) {
  unsigned rem = __ballot_sync(0xffffffffu, on);
  if (rem == 0u) return;
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const int leader = __ffs(rem) - 1;
    const unsigned int b = __shfl_sync(0xffffffffu, bin, leader);
    const unsigned grp = __ballot_sync(0xffffffffu, on && bin == b) & rem;
    if (lane == leader) atomicAdd(&s_hist[b], static_cast<unsigned int>(__popc(grp)));
    rem &= ~grp;
    if (rem == 0u) return;
  }
  if ((rem >> lane) & 1u) atomicAdd(&s_hist[bin], 1u);
}

// Levels whose resolved prefixes agree look at the same bucket (the default parameters ask for the 0.1
// quantile twice): they are served by ONE histogram / ONE comparison per sample.  rep[l] = the first
// active level with level l's prefix (-1: level off); the distinct representatives are the "groups".
struct SelGroups {
  int n;
  int lvl[SEL_MAXQ];     // representative level of group g
  int of[SEL_MAXQ];      // group of level l (-1: off)
};
__device__ __forceinline__ SelGroups sel_groups(const SelState* s_cur, int nq, bool first) {
  SelGroups g;
  g.n = 0;
#pragma unroll
  for (int l = 0; l < SEL_MAXQ; ++l) {
    g.of[l] = -1;
    if (l >= nq || s_cur[l].active == 0) continue;
#pragma unroll
    for (int h = 0; h < SEL_MAXQ; ++h)
      if (h < g.n && g.of[l] < 0 && (first || s_cur[g.lvl[h]].prefix == s_cur[l].prefix)) g.of[l] = h;
    if (g.of[l] < 0) {
#pragma unroll
      for (int h = 0; h < SEL_MAXQ; ++h) if (h == g.n) g.lvl[h] = l;
      g.of[l] = g.n;
      ++g.n;
    }
  }
  return g;
}

// pass p: (p > 0) resolve pass p-1, then histogram digit p of the elements under each level's prefix.
// A CTA covers per * SEL_THREADS samples (the host picks `per` so that small recordings still fill the
// GPU and long ones amortise the 2048-bin flush).  HI32: the digit and everything above it lie in the
// upper half of the key (passes 0 and 1), so only that half of each sample is loaded and converted.
struct SelShared {
  unsigned int hist[SEL_MAXQ][SEL_BINS];
  SelState cur[SEL_MAXQ];
  long long cum[SEL_THREADS + 1];
  unsigned long long red_min[SEL_MAXQ][SEL_THREADS / 32], red_lo[SEL_MAXQ][SEL_THREADS / 32],
      red_hi[SEL_MAXQ][SEL_THREADS / 32];
};

template <bool HI32>
__device__ __forceinline__ void sel_pass_body(SelShared& sm, const double* __restrict__ x, const BpmItem* __restrict__ items,
                                              int p, int nq, const SelLevels& lv, SelState* __restrict__ states,
                                              unsigned int* __restrict__ hist, int n_items, int per) {
  unsigned int (*s_hist)[SEL_BINS] = sm.hist;
  SelState* s_cur = sm.cur;
  long long* s_cum = sm.cum;
  const int item = blockIdx.y;
  const BpmItem it = items[item];
  for (int l = 0; l < nq; ++l) {
    if (p > 0) {
      const SelState before = states[st_idx(p - 1, l, item, n_items)];
      if (before.active == 0) {
        if (threadIdx.x == 0) {
          s_cur[l].active = 0;
          if (blockIdx.x == 0) states[st_idx(p, l, item, n_items)].active = 0;   // keep the chain defined
        }
        __syncthreads();
        continue;
      }
      sel_advance(before, hist + hist_idx(item, (p - 1 == 0) ? 0 : l, p - 1), sel_bits(p - 1), &s_cur[l], s_cum);
      if (blockIdx.x == 0 && threadIdx.x == 0) states[st_idx(p, l, item, n_items)] = s_cur[l];
    } else {
      // first pass: every CTA derives the initial state itself (no separate init launch)
      if (threadIdx.x == 0) {
        s_cur[l] = sel_initial(it.m, lv.q[l], lv.cond[l] == nullptr || lv.cond[l][item] != 0);
        if (blockIdx.x == 0) states[st_idx(0, l, item, n_items)] = s_cur[l];
      }
      __syncthreads();
    }
  }
  // in the first pass no digit is resolved yet: ONE histogram of the whole recording serves every level
  // (sel_advance of the next pass reads it through hist_idx(item, 0, 0))
  const SelGroups gr = sel_groups(s_cur, nq, p == 0);
  const int64_t i0 = static_cast<int64_t>(blockIdx.x) * per * SEL_THREADS;
  if (gr.n == 0 || i0 >= it.m) return;
  const int sh = sel_shift(p), bits = sel_bits(p);
  const int up = sh + bits;                       // bits above this digit
  const unsigned int mask = (1u << bits) - 1u;
  unsigned long long pref[SEL_MAXQ];
#pragma unroll
  for (int g = 0; g < SEL_MAXQ; ++g) pref[g] = (g < gr.n) ? s_cur[gr.lvl[g]].prefix : 0ull;
  for (int t = threadIdx.x; t < gr.n * SEL_BINS; t += SEL_THREADS) (&s_hist[0][0])[t] = 0;
  __syncthreads();
  const double* __restrict__ xi = x + it.m_off;
  const int lane = threadIdx.x & 31;
  // SEL_BATCH independent loads are issued before the first is consumed (`per` is a run-time value: without
  // the explicit batch every load would wait for the previous sample's histogram update)
  for (int k0 = 0; k0 < per; k0 += SEL_BATCH) {
    unsigned int hi32[SEL_BATCH];
    unsigned long long key64[SEL_BATCH];
    bool inb[SEL_BATCH];
#pragma unroll
    for (int u = 0; u < SEL_BATCH; ++u) {
      const int64_t i = i0 + static_cast<int64_t>(k0 + u) * SEL_THREADS + threadIdx.x;
      inb[u] = (k0 + u < per) && i < it.m;
      if (HI32) hi32[u] = inb[u] ? __ldg(reinterpret_cast<const unsigned int*>(xi) + 2 * i + 1) : 0u;
      else key64[u] = inb[u] ? f64_key(xi[i]) : 0ull;
    }
#pragma unroll
    for (int u = 0; u < SEL_BATCH; ++u) {
      const bool in = inb[u];
      unsigned int bin;
      unsigned long long top;
      if (HI32) {
        const unsigned int hi = hi32[u];
        const unsigned int khi = (hi & 0x80000000u) ? ~hi : (hi | 0x80000000u);      // upper half of f64_key
        bin = (khi >> (sh - 32)) & mask;
        top = (up >= 64) ? 0ull : static_cast<unsigned long long>(khi >> (up - 32));
      } else {
        bin = static_cast<unsigned int>(key64[u] >> sh) & mask;
        top = (up >= 64) ? 0ull : (key64[u] >> up);
      }
#pragma unroll
      for (int g = 0; g < SEL_MAXQ; ++g) {
        if (g >= gr.n) break;
        if (p == 0) warp_hist_add<2>(s_hist[g], in, bin, lane);
        else warp_hist_add<1>(s_hist[g], in && top == pref[g], bin, lane);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int l = 0; l < SEL_MAXQ; ++l) {
    if (l >= ((p == 0) ? 1 : nq)) break;
    const int g = (p == 0) ? 0 : gr.of[l];
    if (g < 0) continue;
    unsigned int* gh = hist + hist_idx(item, l, p);
    for (int t = threadIdx.x; t < (1 << bits); t += SEL_THREADS) {
      const unsigned int c = s_hist[g][t];
      if (c) atomicAdd(gh + t, c);
    }
  }
}

template <bool HI32>
__global__ void __launch_bounds__(SEL_THREADS) k_select_pass(const double* __restrict__ x,
                                                             const BpmItem* __restrict__ items, int p, int nq,
                                                             SelLevels lv, SelState* __restrict__ states,
                                                             unsigned int* __restrict__ hist, int n_items, int per) {
  __shared__ SelShared sm;
  sel_pass_body<HI32>(sm, x, items, p, nq, lv, states, hist, n_items, per);
}

// ---- after two digit passes (22 bits) the bucket holding the target is almost always tiny:
// COLLECT its keys in one more sweep (plus the smallest key above the bucket), then one CTA per
// recording sorts them and evaluates numpy's lerp.  A bucket larger than SEL_CAP (massive
// duplicates, e.g. digital silence) is resolved by the same finishing CTA with the remaining four
// digit passes over the whole recording -- slow, but rare and with no extra launches.
constexpr int SEL_CAP = 4096;
constexpr int SEL_FIN_THREADS = 1024;
constexpr int SEL_RANK_MAX = 384;                 // buckets up to this size are ranked by counting (measured: 2048 doubled the time at ~1.5 k keys), larger ones sorted

struct SelCollect {
  unsigned long long* buf;          // [item][level][SEL_CAP] keys of the bucket
  unsigned int* count;              // [item][level]
  unsigned long long* next_above;   // [item][level] smallest key whose resolved prefix is above the bucket's
  unsigned long long* bmin;         // [item][level] smallest / largest key inside a bucket too large to collect:
  unsigned long long* bmax;         //   equal => the bucket is one value repeated (digital silence, clipping)
};

__device__ __forceinline__ void sel_collect_body(SelShared& sm, const double* __restrict__ x,
                                                 const BpmItem* __restrict__ items, int nq, SelState* __restrict__ states,
                                                 const unsigned int* __restrict__ hist, const SelCollect& cl, int npre,
                                                 int n_items, int per) {
  SelState* s_cur = sm.cur;
  long long* s_cum = sm.cum;
  unsigned long long (*s_min)[SEL_THREADS / 32] = sm.red_min;
  unsigned long long (*s_lo)[SEL_THREADS / 32] = sm.red_lo;
  unsigned long long (*s_hi)[SEL_THREADS / 32] = sm.red_hi;
  const int item = blockIdx.y;
  const BpmItem it = items[item];
  for (int l = 0; l < nq; ++l) {
    const SelState before = states[st_idx(npre - 1, l, item, n_items)];
    if (before.active == 0) {
      if (threadIdx.x == 0) {
        s_cur[l].active = 0;
        if (blockIdx.x == 0) states[st_idx(npre, l, item, n_items)].active = 0;
      }
      __syncthreads();
      continue;
    }
    sel_advance(before, hist + hist_idx(item, (npre - 1 == 0) ? 0 : l, npre - 1), sel_bits(npre - 1), &s_cur[l], s_cum);
    if (threadIdx.x == 0) s_cur[l].active = (s_cur[l].count <= SEL_CAP) ? 2 : 1;     // 2: collected
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) states[st_idx(npre, l, item, n_items)] = s_cur[l];
  }
  // levels with the same prefix share a bucket (same count, same mode): one comparison per sample and group
  const SelGroups gr = sel_groups(s_cur, nq, false);
  const int64_t i0 = static_cast<int64_t>(blockIdx.x) * per * SEL_THREADS;
  if (gr.n == 0 || i0 >= it.m) return;
  const int up = sel_shift(npre - 1);                       // bits below the resolved digits
  unsigned long long best[SEL_MAXQ], blo[SEL_MAXQ], bhi[SEL_MAXQ], pref[SEL_MAXQ];
  bool collected[SEL_MAXQ];
#pragma unroll
  for (int g = 0; g < SEL_MAXQ; ++g) {
    best[g] = ~0ull; blo[g] = ~0ull; bhi[g] = 0ull;
    pref[g] = (g < gr.n) ? s_cur[gr.lvl[g]].prefix : 0ull;
    collected[g] = (g < gr.n) && s_cur[gr.lvl[g]].active == 2;
  }
  const double* __restrict__ xi = x + it.m_off;
  for (int k0 = 0; k0 < per; k0 += SEL_BATCH) {
    unsigned long long keys[SEL_BATCH];
    bool inb[SEL_BATCH];
#pragma unroll
    for (int u = 0; u < SEL_BATCH; ++u) {                   // independent loads first
      const int64_t i = i0 + static_cast<int64_t>(k0 + u) * SEL_THREADS + threadIdx.x;
      inb[u] = (k0 + u < per) && i < it.m;
      keys[u] = inb[u] ? f64_key(xi[i]) : 0ull;
    }
#pragma unroll
    for (int u = 0; u < SEL_BATCH; ++u) {
      if (!inb[u]) continue;
      const unsigned long long key = keys[u];
      const unsigned long long top = key >> up;
#pragma unroll
      for (int g = 0; g < SEL_MAXQ; ++g) {
        if (g >= gr.n) break;
        if (top == pref[g]) {
          if (collected[g]) {
#pragma unroll
            for (int l = 0; l < SEL_MAXQ; ++l) {              // a few thousand keys per recording at most
              if (l >= nq || gr.of[l] != g) continue;
              const unsigned int pos = atomicAdd(cl.count + item * SEL_MAXQ + l, 1u);
              if (pos < SEL_CAP) cl.buf[(static_cast<size_t>(item) * SEL_MAXQ + l) * SEL_CAP + pos] = key;
            }
          } else {
            blo[g] = key < blo[g] ? key : blo[g];
            bhi[g] = key > bhi[g] ? key : bhi[g];
          }
        } else if (top > pref[g] && key < best[g]) {
          best[g] = key;
        }
      }
    }
  }
#pragma unroll
  for (int g = 0; g < SEL_MAXQ; ++g) {
    if (g >= gr.n) break;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const unsigned long long t = __shfl_xor_sync(0xffffffffu, best[g], o);
      best[g] = t < best[g] ? t : best[g];
    }
    if (!collected[g]) {
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        const unsigned long long a = __shfl_xor_sync(0xffffffffu, blo[g], o), b = __shfl_xor_sync(0xffffffffu, bhi[g], o);
        blo[g] = a < blo[g] ? a : blo[g];
        bhi[g] = b > bhi[g] ? b : bhi[g];
      }
    }
    if ((threadIdx.x & 31) == 0) {
      s_min[g][threadIdx.x >> 5] = best[g];
      s_lo[g][threadIdx.x >> 5] = blo[g];
      s_hi[g][threadIdx.x >> 5] = bhi[g];
    }
  }
  __syncthreads();
  int my_group = -1;
#pragma unroll
  for (int l = 0; l < SEL_MAXQ; ++l) if (l == static_cast<int>(threadIdx.x) && l < nq) my_group = gr.of[l];
  if (my_group >= 0) {
    const int l = threadIdx.x, g = my_group;
    unsigned long long bm = s_min[g][0], lo = s_lo[g][0], hi = s_hi[g][0];
    for (int w = 1; w < SEL_THREADS / 32; ++w) {
      bm = s_min[g][w] < bm ? s_min[g][w] : bm;
      lo = s_lo[g][w] < lo ? s_lo[g][w] : lo;
      hi = s_hi[g][w] > hi ? s_hi[g][w] : hi;
    }
    if (bm != ~0ull) atomicMin(cl.next_above + item * SEL_MAXQ + l, bm);
    if (s_cur[l].active == 1) {
      if (lo != ~0ull) atomicMin(cl.bmin + item * SEL_MAXQ + l, lo);
      if (hi != 0ull || lo != ~0ull) atomicMax(cl.bmax + item * SEL_MAXQ + l, hi);
    }
  }
}

__global__ void __launch_bounds__(SEL_THREADS) k_select_collect(const double* __restrict__ x,
                                                                const BpmItem* __restrict__ items, int nq,
                                                                SelState* __restrict__ states,
                                                                const unsigned int* __restrict__ hist, SelCollect cl,
                                                                int npre, int n_items, int per) {
  __shared__ SelShared sm;
  sel_collect_body(sm, x, items, nq, states, hist, cl, npre, n_items, per);
}

// numpy's quantile from the two neighbouring order statistics (keys) and the fractional index
__device__ __forceinline__ double sel_lerp(unsigned long long ka, unsigned long long kb, double t) {
  const double a = key_f64(ka), b = key_f64(kb);
  const double diff = __dsub_rn(b, a);
  double r = __dadd_rn(a, __dmul_rn(diff, t));
  if (t >= 0.5) r = __dsub_rn(b, __dmul_rn(diff, __dsub_rn(1.0, t)));
  return r;
}

__global__ void __launch_bounds__(SEL_FIN_THREADS) k_select_finish(const double* __restrict__ x,
                                                                   const BpmItem* __restrict__ items, SelLevels lv,
                                                                   const SelState* __restrict__ states, SelCollect cl,
                                                                   int npre, int n_items) {
  __shared__ unsigned long long s_key[SEL_CAP];             // sorted bucket | digit histogram of the slow path
  __shared__ int s_scan[40];
  __shared__ unsigned long long s_pref;
  __shared__ long long s_rank, s_below, s_count;
  __shared__ unsigned long long s_red[SEL_FIN_THREADS / 32];
  const int item = blockIdx.x;
  const BpmItem it = items[item];
  const int tid = threadIdx.x;
  {
    const int l = blockIdx.y;                               // one CTA per (recording, level)
    const SelState s = states[st_idx(npre, l, item, n_items)];
    if (s.active == 0) return;                              // uniform per CTA
    unsigned long long ka, kb;
    if (s.active == 2) {
      const int nc = static_cast<int>(s.count);
      int P = 32;
      while (P < nc) P <<= 1;
      const unsigned long long* src = cl.buf + (static_cast<size_t>(item) * SEL_MAXQ + l) * SEL_CAP;
      for (int t = tid; t < P; t += SEL_FIN_THREADS) s_key[t] = (t < nc) ? src[t] : ~0ull;
      __syncthreads();
      if (nc <= SEL_RANK_MAX) {
        // only two order statistics of the bucket are wanted: every thread ranks its own key(s) against the
        // whole bucket (broadcast shared-memory reads, ties broken by position) instead of a bitonic sort
        // with ~50 block barriers -- a 60-minute recording leaves a few hundred keys here
        const int want = static_cast<int>(s.rank);
        if (tid == 0) { s_red[0] = ~0ull; s_red[1] = ~0ull; }
        __syncthreads();
        for (int t = tid; t < nc; t += SEL_FIN_THREADS) {
          const unsigned long long mine = s_key[t];
          int before = 0;
          for (int j = 0; j < nc; ++j) {
            const unsigned long long o = s_key[j];
            before += (o < mine || (o == mine && j < t)) ? 1 : 0;
          }
          if (before == want) s_red[0] = mine;
          if (before == want + 1) s_red[1] = mine;
        }
        __syncthreads();
        ka = s_red[0];
        if (s.rank + 1 < nc) kb = s_red[1];
        else {
          const unsigned long long na = cl.next_above[item * SEL_MAXQ + l];
          kb = (na != ~0ull) ? na : ka;
        }
        __syncthreads();
      } else {
      for (int k2 = 2; k2 <= P; k2 <<= 1) {
        for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
          for (int t = tid; t < P / 2; t += SEL_FIN_THREADS) {
            const int lo_i = ((t & ~(j2 - 1)) << 1) | (t & (j2 - 1));
            const int hi_i = lo_i + j2;
            const bool upw = ((lo_i & k2) == 0);
            const unsigned long long a = s_key[lo_i], b = s_key[hi_i];
            if ((a > b) == upw) { s_key[lo_i] = b; s_key[hi_i] = a; }
          }
          __syncthreads();
        }
      }
      ka = s_key[s.rank];
      // element k+1: inside the bucket, else the smallest key above it (none: k is the maximum)
      if (s.rank + 1 < nc) kb = s_key[s.rank + 1];
      else {
        const unsigned long long na = cl.next_above[item * SEL_MAXQ + l];
        kb = (na != ~0ull) ? na : ka;
      }
      __syncthreads();
      }
    } else if (cl.bmin[item * SEL_MAXQ + l] == cl.bmax[item * SEL_MAXQ + l]) {
      // the oversized bucket is ONE value repeated (digital silence, clipping): nothing to resolve
      ka = cl.bmin[item * SEL_MAXQ + l];
      kb = ka;
      if (s.rank + 1 >= s.count) {
        const unsigned long long na = cl.next_above[item * SEL_MAXQ + l];
        if (na != ~0ull) kb = na;
      }
    } else {
      // slow path (more than SEL_CAP distinct keys agreeing in every resolved digit): the remaining
      // digit passes by this one CTA over the whole recording
      unsigned int* hist = reinterpret_cast<unsigned int*>(s_key);
      if (tid == 0) { s_pref = s.prefix; s_rank = s.rank; s_below = s.below; s_count = s.count; }
      __syncthreads();
      const double* __restrict__ xi = x + it.m_off;
      for (int p = npre; p < SEL_PASSES; ++p) {
        const int sh = sel_shift(p), bits = sel_bits(p), upb = sh + bits;
        const unsigned int mask = (1u << bits) - 1u;
        for (int t = tid; t < SEL_BINS; t += SEL_FIN_THREADS) hist[t] = 0;
        __syncthreads();
        const unsigned long long pref = s_pref;
        for (int64_t i = tid; i < it.m; i += SEL_FIN_THREADS) {
          const unsigned long long key = f64_key(xi[i]);
          if ((key >> upb) == pref) atomicAdd(&hist[static_cast<unsigned int>(key >> sh) & mask], 1u);
        }
        __syncthreads();
        constexpr int PER = SEL_BINS / SEL_FIN_THREADS;
        int loc[PER], sum = 0;
#pragma unroll
        for (int u = 0; u < PER; ++u) { loc[u] = static_cast<int>(hist[tid * PER + u]); sum += loc[u]; }
        int total;
        int ex = block_exclusive_scan(sum, &total, s_scan);
        const long long rank = s_rank;
        __syncthreads();
        if (rank >= ex && rank < ex + sum) {
          long long c = ex;
#pragma unroll
          for (int u = 0; u < PER; ++u) {
            if (rank >= c && rank < c + loc[u]) {
              s_pref = (pref << bits) | static_cast<unsigned long long>(tid * PER + u);
              s_rank = rank - c;
              s_below = s_below + c;
              s_count = loc[u];
            }
            c += loc[u];
          }
        }
        __syncthreads();
      }
      ka = s_pref;                                          // all 64 bits resolved: the key itself
      kb = ka;
      if (s_below + s_count <= s.k + 1) {                   // duplicates do not cover element k+1
        unsigned long long best = ~0ull;
        for (int64_t i = tid; i < it.m; i += SEL_FIN_THREADS) {
          const unsigned long long key = f64_key(xi[i]);
          if (key > ka && key < best) best = key;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
          const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o);
          best = t < best ? t : best;
        }
        if ((tid & 31) == 0) s_red[tid >> 5] = best;
        __syncthreads();
        best = s_red[0];
        for (int w = 1; w < SEL_FIN_THREADS / 32; ++w) best = s_red[w] < best ? s_red[w] : best;
        if (best != ~0ull) kb = best;
      }
      __syncthreads();
    }
    if (tid == 0) lv.out[l][item] = sel_lerp(ka, kb, s.gamma);
  }
}

struct SelectBuffers {
  SelState* states;
  unsigned int* hist;
  unsigned int* count;
  unsigned long long* buf;
  unsigned long long* next_above;   // [2][MAXQ * n]: next_above | bmin (both preset to ~0)
  unsigned long long* bmax;         // [MAXQ * n] (preset to 0)
  size_t zero_bytes;
};

static int carve_select(Workspace& ws, int n_items, SelectBuffers* b) {
  b->states = ws.take<SelState>(static_cast<size_t>(SEL_PASSES + 1) * SEL_MAXQ * n_items);
  // the two digit histograms per level and the collect counters are zeroed together
  // one zeroed region: bmax | histograms | collect counters;  one 0xff region: next_above | bmin
  b->bmax = ws.take<unsigned long long>(static_cast<size_t>(SEL_MAXQ) * n_items);
  b->hist = ws.take<unsigned int>(static_cast<size_t>(n_items) * SEL_MAXQ * SEL_PASSES * SEL_BINS + SEL_MAXQ * n_items);
  b->count = b->hist + static_cast<size_t>(n_items) * SEL_MAXQ * SEL_PASSES * SEL_BINS;
  b->zero_bytes = ws.measuring() ? 0 : static_cast<size_t>(reinterpret_cast<char*>(b->count + SEL_MAXQ * n_items) -
                                                            reinterpret_cast<char*>(b->bmax));
  b->buf = ws.take<unsigned long long>(static_cast<size_t>(n_items) * SEL_MAXQ * SEL_CAP);
  b->next_above = ws.take<unsigned long long>(static_cast<size_t>(2) * SEL_MAXQ * n_items);
  return ws.overflow ? BPM_ERR_WORKSPACE : BPM_OK;
}

size_t quantile_workspace_bytes(int n_items) {
  Workspace ws(nullptr, 0);
  SelectBuffers b;
  carve_select(ws, n_items, &b);
  return ws.used;
}

// out[l][i] = np.quantile(x_i, q[l]) for every recording i with cond[l][i] != 0 (cond[l] may be
// null: all recordings); skipped entries are left untouched.
int quantile_multi_run(const double* x, const BpmItem* items, const BatchShape& sh, int nq, const double* q,
                       const int* const* cond, double* const* out, Workspace& ws, cudaStream_t st) {
  if (!x || !items || sh.n_items <= 0 || nq < 1 || nq > SEL_MAXQ) return BPM_ERR_ARG;
  SelLevels lv;
  lv.nq = nq;
  for (int l = 0; l < SEL_MAXQ; ++l) { lv.q[l] = 0.0; lv.cond[l] = nullptr; lv.out[l] = nullptr; }
  for (int l = 0; l < nq; ++l) {
    if (!(q[l] >= 0.0 && q[l] <= 1.0) || !out[l]) return BPM_ERR_ARG;
    lv.q[l] = q[l];
    lv.cond[l] = cond ? cond[l] : nullptr;
    lv.out[l] = out[l];
  }
  SelectBuffers b;
  BPM_TRY(carve_select(ws, sh.n_items, &b));
  const int n = sh.n_items;
  const size_t nlv = static_cast<size_t>(SEL_MAXQ) * n;
  if (cudaMemsetAsync(b.bmax, 0, b.zero_bytes, st) != cudaSuccess) return BPM_ERR_CUDA;
  if (cudaMemsetAsync(b.next_above, 0xff, sizeof(unsigned long long) * 2 * nlv, st) != cudaSuccess) return BPM_ERR_CUDA;
  // digit passes before the bucket is collected: two resolve 22 bits, enough below ~4 M samples; a third
  // (33 bits) keeps the bucket under SEL_CAP for the long streams (24 h at 333 Hz = 28.8 M samples)
  const int npre = sh.max_m > (1ll << 22) ? 3 : 2;
  SelCollect cl{b.buf, b.count, b.next_above, b.next_above + nlv, b.bmax};
  // samples per thread: every CTA pays for a 2048-bin histogram (zero, flush, the previous pass's scan),
  // so tiles stay large -- 16 per thread (266 CTAs for a 60-min recording; 4 per thread and 1062 CTAs
  // measured no faster), up to 64 for long streams and big batches
  long long per_ll = cdiv(static_cast<long long>(sh.max_m) * n, static_cast<long long>(SEL_THREADS) * 148 * 8);
  const int per = per_ll < 16 ? 16 : (per_ll > 64 ? 64 : static_cast<int>(per_ll));
  const dim3 grid(cdiv(sh.max_m, static_cast<long long>(per) * SEL_THREADS), n);
  for (int p = 0; p < npre; ++p) {
    BPM_KERNEL(k_select_pass);
    if (sel_shift(p) >= 32) k_select_pass<true><<<grid, SEL_THREADS, 0, st>>>(x, items, p, nq, lv, b.states, b.hist, n, per);
    else k_select_pass<false><<<grid, SEL_THREADS, 0, st>>>(x, items, p, nq, lv, b.states, b.hist, n, per);
    BPM_LAUNCH_OK();
  }
  BPM_KERNEL(k_select_collect);
  k_select_collect<<<grid, SEL_THREADS, 0, st>>>(x, items, nq, b.states, b.hist, cl, npre, n, per);
  BPM_LAUNCH_OK();
  BPM_KERNEL(k_select_finish);
  k_select_finish<<<dim3(n, nq), SEL_FIN_THREADS, 0, st>>>(x, items, lv, b.states, cl, npre, n);
  BPM_LAUNCH_OK();
  return BPM_OK;
}

int quantile_run(const double* x, const BpmItem* items, const BatchShape& sh, double q, const int* cond,
                 double* out, Workspace& ws, cudaStream_t st) {
  const int* conds[1] = {cond};
  double* outs[1] = {out};
  return quantile_multi_run(x, items, sh, 1, &q, conds, outs, ws, st);
}

}  // namespace bpm
