// K4: scipy.signal.find_peaks(sign*x, height=, prominence=, distance=), conditions in
// scipy's order (scipy/signal/_peak_finding.py:943-1010; internals restated in SURVEY.md
// Appendix A.3):  plateau-aware local maxima -> height -> distance -> prominence.
// Called by the reference at bpm_analysis.py:227 (peaks of the envelope above the noise
// floor), :1070 (troughs = peaks of -envelope) and :1496-1497 / :1529-1530 (extrema of the
// smoothed BPM series).  Three launches for any signal length:
//
//   k_localmax_compact    per sample: is it the midpoint of a strict local-maximum plateau and
//                         (optionally) at or above its height threshold; the candidates are written
//                         in order by a SINGLE-PASS compaction (decoupled look-back over tile counts).
//   k_distance_tiles      scipy's greedy "highest first removes neighbours closer than d" as the
//                         well-founded recursion  kept(k) <=> no kept higher-priority neighbour within d,
//                         evaluated depth-first per candidate with the results memoised in shared memory
//                         (no rounds, no barriers).  A CTA owns 512 consecutive candidates and stages
//                         them with a halo of 64 on either side; a candidate is settled when its
//                         dependency chain (strictly rising priority, each hop < d samples) stays inside
//                         the staged range, which is the case for all but pathological inputs.  Whatever
//                         is left is finished exactly by the last CTA of the recording over global memory.
//   k_prominence_compact  prominence >= threshold for every survivor: a warp walks the two sides of
//                         each of its survivors 32 samples per step with ballot early exit (a higher
//                         sample stops a side, a sample low enough passes it).  The survivors are
//                         written in order, again by single-pass compaction.
// Equal-height candidates within `distance`: the later index wins (what a stable argsort
// gives scipy); numpy's default sort is unstable, so the reference does not pin this case.
#include "common.cuh"

namespace bpm {

constexpr int PK_THREADS = 256;
constexpr int PK_PER = 8;
constexpr int PK_TILE = PK_THREADS * PK_PER;     // samples per tile of k_localmax_compact

// ---- plateaus (scipy _local_maxima_1d: the midpoint (left + right) // 2 of a flat run that is
// higher than both neighbours is the peak; runs touching either end are not peaks).
// Sample i is that midpoint iff  dr == dl or dr == dl + 1  (dl / dr = equal samples to its left /
// right).  Both sides are walked TOGETHER and the walk stops as soon as one side has ended and the
// other is already too long, i.e. after min(dl, dr) + 2 steps -- a flat run of length L costs
// O(L * min(L, cap)) in total instead of O(L^2), so digital silence does not stall the kernel.
// A sample more than `cap` away from both ends of its run (PLATEAU_DEEP) is settled per tile: one
// warp walks the run to its two edges, 32 samples per step, and flags the midpoint if the tile
// holds it.
constexpr int PL_CAP = 512;
enum { PLATEAU_NO = 0, PLATEAU_MID = 1, PLATEAU_DEEP = 2 };

template <class Val>
__device__ __forceinline__ int plateau_test(Val val, int64_t i, int64_t n, double c, int cap) {
  long long dl = -1, dr = -1;
  for (int s = 1; s <= cap + 1; ++s) {
    if (dl < 0 && (i - s < 0 || val(i - s) != c)) dl = s - 1;
    if (dr < 0 && (i + s > n - 1 || val(i + s) != c)) dr = s - 1;
    if (dl >= 0 && dr >= 0) break;
    if (dl >= 0 && s > dl + 1) break;        // right side longer than dl + 1: not the midpoint
    if (dr >= 0 && s > dr) break;            // left side longer than dr: not the midpoint
  }
  if (dl >= 0 && dr >= 0) {
    if (dr != dl && dr != dl + 1) return PLATEAU_NO;
    const int64_t L = i - dl, R = i + dr;
    return (L >= 1 && R <= n - 2 && val(L - 1) < c && val(R + 1) < c) ? PLATEAU_MID : PLATEAU_NO;
  }
  // still open on the right after cap + 1 steps with the left side at least cap long: the run is
  // longer than 2 * cap + 1 and this sample may be its midpoint (both sides open, or dl == cap)
  if (dr < 0 && (dl < 0 || dl == cap)) return PLATEAU_DEEP;
  return PLATEAU_NO;
}


// one end of the flat run around p (all samples == c): last index in direction dir that still equals c.
// A warp covers 256 samples per step (8 independent loads per lane): a million-sample run of digital
// silence is walked in ~4 k steps.
__device__ int64_t warp_run_edge(const double* __restrict__ xi, int sign, int64_t n, int64_t p, double c, int dir) {
  const int lane = threadIdx.x & 31;
  int64_t e = p;
  while (true) {
    bool eq[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int64_t j = e + dir * (1 + u * 32 + lane);
      eq[u] = (j >= 0 && j <= n - 1) && signed_val(xi[j], sign) == c;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const unsigned m = __ballot_sync(0xffffffffu, eq[u]);
      if (m != 0xffffffffu) return e + dir * (u * 32 + __ffs(~m) - 1);
    }
    e += dir * 256;
  }
}

// element e of the staged tile (one pad slot per 8 doubles: chunk reads are bank-conflict free)
__device__ __forceinline__ int pk_pad(int e) { return e + (e >> 3); }

__global__ void __launch_bounds__(PK_THREADS) k_localmax_compact(const double* __restrict__ x, int sign,
                                                                 const double* __restrict__ height,
                                                                 const BpmItem* __restrict__ items,
                                                                 unsigned long long* __restrict__ status,
                                                                 int64_t status_stride,
                                                                 int64_t* __restrict__ cand,
                                                                 int64_t* __restrict__ cand_count, ChunkInfo ci) {
  __shared__ double xs[PK_TILE + 2 + (PK_TILE + 2) / 8 + 1];     // samples i0 - 1 .. i0 + PK_TILE
  __shared__ unsigned char s_mask[PK_THREADS], s_deep[PK_THREADS];
  __shared__ int s_scan[34];
  __shared__ long long s_off;
  const int item = blockIdx.y;
  const BpmItem it = items[item];
  const int64_t n = it.m;
  const int64_t i0 = static_cast<int64_t>(blockIdx.x) * PK_TILE;
  if (i0 >= n && !(n == 0 && blockIdx.x == 0)) return;
  if (n == 0) { if (threadIdx.x == 0) cand_count[item] = 0; return; }
  const double* __restrict__ xi = x + it.m_off;
  const int tid = threadIdx.x;
  if (ci.edge_hits != nullptr && tid < 32 && n >= 2) {
    // a flat run that starts at an artificial end of a time chunk and reaches its core hides where the
    // run really begins: plateau midpoints there cannot be decided inside the chunk
    if (blockIdx.x == 0 && ci.open_left) {
      const int64_t R = warp_run_edge(xi, sign, n, 0, signed_val(xi[0], sign), +1);
      if (tid == 0 && R + 1 >= ci.core_lo) atomicAdd(ci.edge_hits, 1ull);
    }
    if (i0 + PK_TILE >= n && ci.open_right) {
      const int64_t L = warp_run_edge(xi, sign, n, n - 1, signed_val(xi[n - 1], sign), -1);
      if (tid == 0 && L - 1 < ci.core_hi) atomicAdd(ci.edge_hits, 1ull);
    }
  }
  for (int e = tid; e < PK_TILE + 2; e += PK_THREADS) {
    const int64_t i = i0 - 1 + e;
    xs[pk_pad(e)] = (i >= 0 && i < n) ? signed_val(xi[i], sign) : 0.0;
  }
  __syncthreads();
  unsigned mask = 0, deep = 0;
#pragma unroll
  for (int k = 0; k < PK_PER; ++k) {
    const int e = tid * PK_PER + k + 1;                          // staged index of sample i
    const int64_t i = i0 + tid * PK_PER + k;
    if (i < 1 || i > n - 2) continue;
    const double c = xs[pk_pad(e)], l = xs[pk_pad(e - 1)], r = xs[pk_pad(e + 1)];
    bool pk = false;
    if (l < c && r < c) {
      pk = true;
    } else if ((l == c || r == c) && l <= c && r <= c) {
      const int t = plateau_test([&](int64_t j) { return signed_val(xi[j], sign); }, i, n, c, PL_CAP);
      pk = (t == PLATEAU_MID);
      if (t == PLATEAU_DEEP) deep |= 1u << k;
    }
    if (pk && height != nullptr) pk = (height[it.m_off + i] <= c);
    if (pk) mask |= 1u << k;
  }
  s_mask[tid] = static_cast<unsigned char>(mask);
  s_deep[tid] = static_cast<unsigned char>(deep);
  if (__syncthreads_or(deep != 0)) {
    // very long flat runs crossing this tile (rare: digital silence): warp 0 settles them one by one
    if (tid < 32) {
      int64_t done_to = -1;                                      // samples <= done_to belong to runs already settled
      for (int t = 0; t < PK_THREADS; ++t) {
        const unsigned dm = s_deep[t];
        if (dm == 0) continue;
        for (int k = 0; k < PK_PER; ++k) {
          if (!((dm >> k) & 1)) continue;
          const int64_t p = i0 + t * PK_PER + k;
          if (p <= done_to) continue;
          const double c = signed_val(xi[p], sign);
          const int64_t L = warp_run_edge(xi, sign, n, p, c, -1), R = warp_run_edge(xi, sign, n, p, c, +1);
          done_to = R;
          if (L < 1 || R > n - 2) continue;
          if (!(signed_val(xi[L - 1], sign) < c) || !(signed_val(xi[R + 1], sign) < c)) continue;
          const int64_t mid = (L + R) / 2;
          if (mid < i0 || mid >= i0 + PK_TILE) continue;
          if (height != nullptr && !(height[it.m_off + mid] <= c)) continue;
          if ((tid & 31) == 0) s_mask[(mid - i0) / PK_PER] |= static_cast<unsigned char>(1u << ((mid - i0) % PK_PER));
        }
      }
    }
    __syncthreads();
    mask = s_mask[tid];
  }
  const int cnt = __popc(mask);
  int total;
  const int ex = block_exclusive_scan(cnt, &total, s_scan);
  if (tid < 32) {
    const long long off = lookback_exclusive(status + static_cast<int64_t>(item) * status_stride, blockIdx.x, total);
    if (tid == 0) s_off = off;
  }
  __syncthreads();
  int64_t o = it.m_off + s_off + ex;
#pragma unroll
  for (int k = 0; k < PK_PER; ++k)
    if ((mask >> k) & 1) cand[o++] = i0 + tid * PK_PER + k;
  if (tid == 0 && i0 + PK_TILE >= n) cand_count[item] = s_off + total;
}

#ifdef BPM_DEBUG_COUNTERS
// [0] tiles, [1] rounds summed, [2] max rounds of a tile, [3] candidates left pending by the tiles,
// [4] rounds of the global finish, [5] candidates, [6..8] thread-0 cycles: staging, rounds, write-back
__device__ unsigned long long g_dbg_pk[16];
#define PKD_ADD(i, v) atomicAdd(&g_dbg_pk[i], static_cast<unsigned long long>(v))
#define PKD_MAX(i, v) atomicMax(&g_dbg_pk[i], static_cast<unsigned long long>(v))
#else
#define PKD_ADD(i, v)
#define PKD_MAX(i, v)
#endif

// A recording that is one TIME CHUNK of a longer stream (bpm_analysis_b200/stream.py): its ends may be
// artificial.  Decisions for peaks inside [core_lo, core_hi) are final only if no dependency reached
// an open end; every case where that cannot be proven is counted in *edge_hits (the caller then
// falls back to the unchunked evaluation): candidates the distance tiles left to the global finish
// (all other chains are at most DT_DEPTH hops of < d samples long, far inside the halo) and prominence
// walks of core peaks that stopped at an open end.  Single-recording calls only.
//
// Memoised results make the TRUE dependency chain of a decision unbounded in principle (a long ramp of
// candidates alternates kept / removed all the way), so "not pending" alone proves nothing about how
// far an open end can reach.  What does: an ANCHOR, a candidate that outranks every candidate within
// the distance.  It is kept whatever happens elsewhere, it removes every candidate within d of it, and
// therefore no decision on its far side can depend on anything on its near side.  The kernel reports
// the anchors nearest to the core (ci.anchors); the caller accepts the core's decisions only if one of
// them lies between each open end's doubtful zone and the core.
// ------------------------------------------------------------------ distance
// Threads per tile.  A recording of a few M samples has a few hundred tiles: the kernel's time is ONE
// tile's latency (staging, neighbour masks, evaluation, write-back, each a dependent phase), so every
// candidate gets its own thread (512: 31 -> 21 us per launch at C2).  Long streams and big batches
// have thousands of tiles and are throughput-bound: 4 candidates per thread keep more tiles resident
// (128: 218 us at C4 against 249 with 512).
constexpr int DT_THREADS_SMALL = 512;
constexpr int DT_THREADS_LARGE = 128;
constexpr int64_t DT_LARGE_MIN = 4 << 20;         // total samples from which the throughput form is used
constexpr int DT_OWN = 512;                       // candidates a CTA settles per tile
constexpr int DT_HALO = 64;                       // staged on either side of them
constexpr int DT_STAGE = DT_OWN + 2 * DT_HALO;
constexpr int DT_DEPTH = 24;                      // dependency chains followed this deep, longer ones are left pending
enum : unsigned char { DST_REMOVED = 0, DST_KEPT = 1, DST_PENDING = 3 };   // global state of a candidate

__device__ __forceinline__ void note_anchor(const ChunkInfo& ci, int pk) {
  if (ci.anchors == nullptr) return;
  if (pk <= ci.core_lo) atomicMax(ci.anchors, static_cast<long long>(pk));
  if (pk >= ci.core_hi - 1) atomicMin(ci.anchors + 1, static_cast<long long>(pk));
}

__device__ __forceinline__ bool higher_priority(double va, int64_t ka, double vb, int64_t kb) {
  return va > vb || (va == vb && ka > kb);
}

// The rule  kept(k) <=> no kept higher-priority neighbour within d  is a well-founded recursion
// (priority strictly rises along every dependency), so each thread simply EVALUATES it for its own
// candidates: scan the higher-priority neighbours; one of them kept -> removed; one still unknown ->
// evaluate that one first (explicit stack, depth-first), then look again; none left -> kept.
// Results are memoised in shared memory (a state only ever goes from open to its final value, and
// every thread would write the same value), so no barriers are needed and threads never wait for
// each other.  Staged candidates whose neighbourhood is not completely staged cannot be kept here;
// an owned candidate that depends on one of those -- or on a chain deeper than DT_DEPTH -- is left
// pending for the exact global finish below (never seen on real envelopes: chains are a few long).
template <int DT_THREADS>
__global__ void __launch_bounds__(DT_THREADS) k_distance_tiles(const double* __restrict__ x, int sign,
                                                               const BpmItem* __restrict__ items,
                                                               const int64_t* __restrict__ cand,
                                                               const int64_t* __restrict__ cand_count, int distance,
                                                               unsigned char* __restrict__ state,
                                                               int* __restrict__ pending,
                                                               unsigned int* __restrict__ ticket, ChunkInfo ci) {
  __shared__ int s_pos[DT_STAGE];
  __shared__ double s_val[DT_STAGE];
  __shared__ unsigned char s_st[DT_STAGE];        // 0 open, 1 kept, 2 removed, 3 cannot be settled in this tile
  __shared__ unsigned short s_stack[DT_DEPTH][DT_THREADS];
  __shared__ unsigned int s_hi[DT_STAGE];         // d <= 32: which of the <= 16 + 16 neighbours within d have the higher priority
  __shared__ int s_last;
  const int item = blockIdx.y;
  const BpmItem it = items[item];
  const int64_t nc = cand_count[item];
  const int64_t* __restrict__ pos = cand + it.m_off;
  const double* __restrict__ xi = x + it.m_off;
  unsigned char* st_out = state + it.m_off;
  const int tid = threadIdx.x;
  const int d = distance;
  // Local maxima lie at least 2 samples apart, so at most (d - 1) / 2 <= 15 candidates per side are
  // within d <= 32: their "has the higher priority" relation fits one word per candidate, computed
  // once; evaluating a candidate is then a loop over the set bits (a couple of byte loads) instead of
  // a rescan of positions and values.
  const bool compact = d <= 32;
  int my_pending = 0;
  for (int64_t k0 = static_cast<int64_t>(blockIdx.x) * DT_OWN; k0 < nc; k0 += static_cast<int64_t>(gridDim.x) * DT_OWN) {
    const int64_t k1 = min(nc, k0 + DT_OWN);
    if (d <= 1) {
      for (int64_t k = k0 + tid; k < k1; k += DT_THREADS) st_out[k] = DST_KEPT;
      continue;
    }
    const int64_t s0 = max(static_cast<int64_t>(0), k0 - DT_HALO), s1 = min(nc, k1 + DT_HALO);
    const int L = static_cast<int>(s1 - s0);
#ifdef BPM_DEBUG_COUNTERS
    long long t_dbg = clock64();
    int rounds_dbg = 0;
#endif
    __syncthreads();                                              // previous tile's staging is no longer read
    {
      constexpr int DT_PER = (DT_STAGE + DT_THREADS - 1) / DT_THREADS;
      int64_t pp[DT_PER];
#pragma unroll
      for (int u = 0; u < DT_PER; ++u) {
        const int t = tid + u * DT_THREADS;
        pp[u] = (t < L) ? pos[s0 + t] : 0;
      }
#pragma unroll
      for (int u = 0; u < DT_PER; ++u) {
        const int t = tid + u * DT_THREADS;
        if (t < L) {
          s_pos[t] = static_cast<int>(pp[u]);
          s_val[t] = signed_val(xi[pp[u]], sign);
          s_st[t] = 0;
        }
      }
    }
    __syncthreads();
    const int p_first = s_pos[0], p_last = s_pos[L - 1];
    const bool open_left = s0 > 0, open_right = s1 < nc;
    volatile unsigned char* S = s_st;
#ifdef BPM_DEBUG_COUNTERS
    if (tid == 0) { const long long t2 = clock64(); PKD_ADD(6, t2 - t_dbg); t_dbg = t2; }
#endif
    const int own0 = static_cast<int>(k0 - s0), own1 = static_cast<int>(k1 - s0);
    if (compact) {
      for (int k = tid; k < L; k += DT_THREADS) {
        const int pk = s_pos[k];
        const double vk = s_val[k];
        unsigned int m = 0;
        for (int b = 0, k2 = k - 1; k2 >= 0 && pk - s_pos[k2] < d; ++b, --k2)
          if (s_val[k2] > vk) m |= 1u << b;                       // equal heights: the later index wins
        for (int b = 16, k2 = k + 1; k2 < L && s_pos[k2] - pk < d; ++b, ++k2)
          if (s_val[k2] >= vk) m |= 1u << b;
        const bool full = (!open_left || pk - p_first >= d) && (!open_right || p_last - pk >= d);
        s_hi[k] = m;
        if (m == 0) {                                             // no higher-priority neighbour at all
          s_st[k] = full ? 1 : 3;
          if (full) note_anchor(ci, pk);
        }
      }
      __syncthreads();
      for (int kk = own0 + tid; kk < own1; kk += DT_THREADS) {
        int sp = 0, cur = kk;
        while (true) {
          if (S[cur] == 0) {
#ifdef BPM_DEBUG_COUNTERS
            ++rounds_dbg;
#endif
            bool any_keep = false, any_unknown = false;
            int open_j = -1;
            unsigned int m = s_hi[cur];
            while (m) {
              const int b = __ffs(m) - 1;
              m &= m - 1;
              const int k2 = (b < 16) ? cur - 1 - b : cur + 1 + (b - 16);
              const unsigned char s2 = S[k2];
              if (s2 == 1) { any_keep = true; break; }
              if (s2 == 0) open_j = k2; else if (s2 == 3) any_unknown = true;
            }
            if (any_keep) {
              S[cur] = 2;
            } else if (open_j >= 0) {
              if (sp < DT_DEPTH) { s_stack[sp++][tid] = static_cast<unsigned short>(cur); cur = open_j; continue; }
              S[cur] = 3;                                         // chain too deep for this tile
            } else {
              const int pk = s_pos[cur];
              const bool full = (!open_left || pk - p_first >= d) && (!open_right || p_last - pk >= d);
              S[cur] = (any_unknown || !full) ? 3 : 1;
            }
          }
          if (sp == 0) break;
          cur = s_stack[--sp][tid];
        }
      }
    } else
    for (int kk = own0 + tid; kk < own1; kk += DT_THREADS) {
      int sp = 0, cur = kk;
      while (true) {
        if (S[cur] == 0) {
#ifdef BPM_DEBUG_COUNTERS
          ++rounds_dbg;
#endif
          const int pk = s_pos[cur];
          const double vk = s_val[cur];
          bool any_keep = false, any_unknown = false, any_hi = false;
          int open_j = -1;
          for (int k2 = cur - 1; k2 >= 0 && pk - s_pos[k2] < d; --k2) {
            if (s_val[k2] > vk) {                                 // equal heights: the later index wins
              any_hi = true;
              const unsigned char s2 = S[k2];
              if (s2 == 1) { any_keep = true; break; }
              if (s2 == 0) open_j = k2; else if (s2 == 3) any_unknown = true;
            }
          }
          if (!any_keep) {
            for (int k2 = cur + 1; k2 < L && s_pos[k2] - pk < d; ++k2) {
              if (s_val[k2] >= vk) {
                any_hi = true;
                const unsigned char s2 = S[k2];
                if (s2 == 1) { any_keep = true; break; }
                if (s2 == 0) open_j = k2; else if (s2 == 3) any_unknown = true;
              }
            }
          }
          if (any_keep) {
            S[cur] = 2;
          } else if (open_j >= 0) {
            if (sp < DT_DEPTH) { s_stack[sp++][tid] = static_cast<unsigned short>(cur); cur = open_j; continue; }
            S[cur] = 3;                                           // chain too deep for this tile
          } else {
            const bool full = (!open_left || pk - p_first >= d) && (!open_right || p_last - pk >= d);
            S[cur] = (any_unknown || !full) ? 3 : 1;
            if (!any_hi && full) note_anchor(ci, pk);
          }
        }
        if (sp == 0) break;
        cur = s_stack[--sp][tid];
      }
    }
    __syncthreads();
#ifdef BPM_DEBUG_COUNTERS
    rounds_dbg = __syncthreads_count(rounds_dbg) ? rounds_dbg : 0;
    PKD_ADD(1, rounds_dbg);
    if (tid == 0) {
      const long long t2 = clock64(); PKD_ADD(7, t2 - t_dbg); t_dbg = t2;
      PKD_ADD(0, 1);
    }
#endif
    for (int64_t k = k0 + tid; k < k1; k += DT_THREADS) {
      const unsigned char s = s_st[k - s0];
      st_out[k] = (s == 1) ? DST_KEPT : (s == 2 ? DST_REMOVED : DST_PENDING);
      my_pending += (s == 3 || s == 0) ? 1 : 0;
    }
#ifdef BPM_DEBUG_COUNTERS
    if (tid == 0) { const long long t2 = clock64(); PKD_ADD(8, t2 - t_dbg); }
#endif
  }
  PKD_ADD(3, my_pending);
  // ---- the last CTA of this recording finishes whatever chains escaped their halo (exact, slow, rare)
  my_pending = __syncthreads_count(my_pending != 0) ? 1 : 0;     // (only "any" matters)
  if (tid == 0) {
    if (my_pending) atomicAdd(pending + item, 1);
    __threadfence();
    s_last = (atomicAdd(ticket + item, 1u) == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (tid == 0) PKD_ADD(5, nc);
  if (atomicAdd(pending + item, 0) == 0) return;
  if (tid == 0 && ci.edge_hits != nullptr) atomicAdd(ci.edge_hits, 1ull);   // a chain longer than the tiles follow
  volatile unsigned char* stt = st_out;
  while (true) {
    int changed = 0;
    if (tid == 0) PKD_ADD(4, 1);
    for (int64_t k = tid; k < nc; k += DT_THREADS) {
      if (stt[k] != DST_PENDING) continue;
      const int64_t pk = pos[k];
      const double vk = signed_val(xi[pk], sign);
      bool any_keep = false, any_open = false;
      for (int64_t k2 = k - 1; k2 >= 0 && pk - pos[k2] < d; --k2) {
        if (higher_priority(signed_val(xi[pos[k2]], sign), k2, vk, k)) {
          const unsigned char s2 = stt[k2];
          any_keep |= (s2 == DST_KEPT); any_open |= (s2 == DST_PENDING);
        }
      }
      for (int64_t k2 = k + 1; k2 < nc && pos[k2] - pk < d; ++k2) {
        if (higher_priority(signed_val(xi[pos[k2]], sign), k2, vk, k)) {
          const unsigned char s2 = stt[k2];
          any_keep |= (s2 == DST_KEPT); any_open |= (s2 == DST_PENDING);
        }
      }
      if (any_keep) { stt[k] = DST_REMOVED; changed = 1; }
      else if (!any_open) { stt[k] = DST_KEPT; changed = 1; }
    }
    __threadfence_block();
    if (!__syncthreads_or(changed)) break;
  }
}

// ------------------------------------------------------------------ prominence
// does the prominence of the peak at p reach thr?  (scipy _peak_prominences, wlen=None)
// Warp-cooperative, both sides walked together 32 samples at a time: a side passes as soon as
// a sample low enough (x[p] - x[i] >= thr) is met before the walk would stop (a sample above
// the peak, or the end of the signal); it fails when the walk stops first.
// *edge (optional): set to 1 / 2 when the decision was "walk stopped at the left / right END of the
// array" -- for a time chunk of a longer stream that end is artificial and the decision is not final.
template <class Val>
__device__ __forceinline__ bool warp_prominence_ok(Val val, int64_t n, int64_t p, double thr, int* edge = nullptr) {
  const int lane = threadIdx.x & 31;
  const double xp = val(p);
  if (__dsub_rn(xp, xp) >= thr) return true;            // the peak itself is the running minimum
  int done_l = 0, done_r = 0;                             // 0 open, 1 passed
  int64_t bl = p - 1, br = p + 1;
  while (true) {
    const int64_t il = bl - lane, ir = br + lane;
    const bool vl = !done_l && il >= 0, vr = !done_r && ir < n;
    const double xl = vl ? val(il) : 0.0;
    const double xr = vr ? val(ir) : 0.0;
    if (!done_l) {
      const bool stop = !vl || (xl > xp);
      const bool pass = vl && !(xl > xp) && (__dsub_rn(xp, xl) >= thr);
      const unsigned bs = __ballot_sync(0xffffffffu, stop), bp = __ballot_sync(0xffffffffu, pass);
      const int fs = bs ? __ffs(bs) : 33, fp = bp ? __ffs(bp) : 33;
      if (fp < fs) done_l = 1;
      else if (bs) {
        if (edge != nullptr && bl - (fs - 1) < 0) *edge = 1;      // the stopping lane ran off the left end
        return false;
      } else bl -= 32;
    }
    if (!done_r) {
      const bool stop = !vr || (xr > xp);
      const bool pass = vr && !(xr > xp) && (__dsub_rn(xp, xr) >= thr);
      const unsigned bs = __ballot_sync(0xffffffffu, stop), bp = __ballot_sync(0xffffffffu, pass);
      const int fs = bs ? __ffs(bs) : 33, fp = bp ? __ffs(bp) : 33;
      if (fp < fs) done_r = 1;
      else if (bs) {
        if (edge != nullptr && br + (fs - 1) >= n) *edge = 2;     // ... off the right end
        return false;
      } else br += 32;
    }
    if (done_l && done_r) return true;
  }
}

constexpr int PC_THREADS = 256;                   // candidates per tile of k_prominence_compact

__global__ void __launch_bounds__(PC_THREADS) k_prominence_compact(const double* __restrict__ x, int sign,
                                                                   const BpmItem* __restrict__ items,
                                                                   const int64_t* __restrict__ cand,
                                                                   const int64_t* __restrict__ cand_count,
                                                                   const unsigned char* __restrict__ state,
                                                                   const double* __restrict__ prominence,
                                                                   unsigned long long* __restrict__ status,
                                                                   int64_t status_stride,
                                                                   int64_t* __restrict__ out_idx,
                                                                   int64_t* __restrict__ out_count, ChunkInfo ci) {
  __shared__ int s_scan[34];
  __shared__ long long s_off;
  const int item = blockIdx.y;
  const BpmItem it = items[item];
  const int64_t nc = cand_count[item];
  const int64_t k0 = static_cast<int64_t>(blockIdx.x) * PC_THREADS;
  if (k0 >= nc) {
    if (nc == 0 && blockIdx.x == 0 && threadIdx.x == 0) out_count[item] = 0;
    return;
  }
  const int64_t* __restrict__ pos = cand + it.m_off;
  const double* __restrict__ xi = x + it.m_off;
  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t k = k0 + tid;
  bool live = (k < nc) && state[it.m_off + k] == DST_KEPT;
  const int64_t p = (k < nc) ? pos[k] : 0;
  if (prominence != nullptr) {
    // the warp walks its survivors one after the other, both sides 32 samples per step (a lane-per-peak
    // walk was measured slower: 410 vs 237 us on the 24-h stream -- the slowest lane sets the pace)
    const double thr = prominence[item];
    unsigned todo = __ballot_sync(0xffffffffu, live);
    // (staging the signal around a warp's candidates in shared memory first was measured slower: 548 vs
    // 365 us on the 24-h stream -- the walks touch about as many samples as the staging itself)
    auto val = [&](int64_t i) -> double { return signed_val(xi[i], sign); };
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const int64_t ps = __shfl_sync(0xffffffffu, p, src);
      int edge = 0;
      const bool ok = warp_prominence_ok(val, it.m, ps, thr, ci.edge_hits ? &edge : nullptr);
      if (lane == src) {
        live = ok;
        if (edge != 0 && ps >= ci.core_lo && ps < ci.core_hi &&
            ((edge == 1 && ci.open_left) || (edge == 2 && ci.open_right)))
          atomicAdd(ci.edge_hits, 1ull);
      }
    }
  }
  int total;
  const int ex = block_exclusive_scan(live ? 1 : 0, &total, s_scan);
  if (tid < 32) {
    const long long off = lookback_exclusive(status + static_cast<int64_t>(item) * status_stride, blockIdx.x, total);
    if (tid == 0) s_off = off;
  }
  __syncthreads();
  if (live) out_idx[it.m_off + s_off + ex] = p;
  if (tid == 0 && k0 + PC_THREADS >= nc) out_count[item] = s_off + total;
}

// ------------------------------------------------------------------ host side
struct PeakBuffers {
  unsigned char* cstate;          // [total_m] state of every candidate after the distance rule
  int64_t* cand;                  // [total_m]
  int64_t* cand_count;            // [n_items]
  // one zeroed region: look-back words of the two compactions, pending counters, tickets
  unsigned long long* status_a;   // [n_items][stride_a]
  unsigned long long* status_c;   // [n_items][stride_c]
  int* pending;                   // [n_items]
  unsigned int* ticket;           // [n_items]
  int64_t stride_a, stride_c;
  size_t zero_bytes;
};

static int carve_peaks(Workspace& ws, int64_t total_m, int64_t max_m, int n_items, PeakBuffers* b) {
  b->cstate = ws.take<unsigned char>(total_m);
  b->cand = ws.take<int64_t>(total_m);
  b->cand_count = ws.take<int64_t>(n_items);
  b->stride_a = max_m / PK_TILE + 2;
  b->stride_c = (max_m / 2 + 1) / PC_THREADS + 2;
  const size_t words = static_cast<size_t>(n_items) * (b->stride_a + b->stride_c + 1);     // + pending | ticket
  unsigned long long* z = ws.take<unsigned long long>(words);
  b->status_a = z;
  b->status_c = z ? z + static_cast<size_t>(n_items) * b->stride_a : nullptr;
  unsigned long long* tail = z ? z + static_cast<size_t>(n_items) * (b->stride_a + b->stride_c) : nullptr;
  b->pending = reinterpret_cast<int*>(tail);
  b->ticket = reinterpret_cast<unsigned int*>(tail) + n_items;
  b->zero_bytes = words * sizeof(unsigned long long);
  return ws.overflow ? BPM_ERR_WORKSPACE : BPM_OK;
}

size_t find_peaks_workspace_bytes(int64_t total_m, int n_items) {
  Workspace ws(nullptr, 0);
  PeakBuffers b;
  carve_peaks(ws, total_m, total_m, n_items, &b);          // max_m <= total_m
  return ws.used;
}

// prominence_ready: event after which the prominence threshold is valid (it may be produced on
// another stream while the local-maximum / distance steps run here); nullptr = already valid.
int find_peaks_run(const double* x, int sign, const double* height, const double* prominence, int distance,
                   const BpmItem* items, const BatchShape& sh, int64_t* out_idx, int64_t* out_count,
                   Workspace& ws, cudaStream_t st, cudaEvent_t prominence_ready, const ChunkInfo* chunk) {
  if (!x || !items || !out_idx || !out_count || sh.n_items <= 0 || distance < 1) return BPM_ERR_ARG;
  if (chunk != nullptr && sh.n_items != 1) return BPM_ERR_ARG;
  const ChunkInfo ci = chunk ? *chunk : ChunkInfo{0, 0, 0, 0, nullptr, nullptr};
  PeakBuffers b;
  BPM_TRY(carve_peaks(ws, sh.total_m, sh.max_m, sh.n_items, &b));
  if (cudaMemsetAsync(b.status_a, 0, b.zero_bytes, st) != cudaSuccess) return BPM_ERR_CUDA;
  BPM_KERNEL(k_localmax_compact);
  k_localmax_compact<<<dim3(cdiv(sh.max_m > 0 ? sh.max_m : 1, PK_TILE), sh.n_items), PK_THREADS, 0, st>>>(
      x, sign, height, items, b.status_a, b.stride_a, b.cand, b.cand_count, ci);
  BPM_LAUNCH_OK();
  // a local maximum needs a lower neighbour on both sides: at most (m - 1) / 2 candidates
  const int64_t max_c = sh.max_m / 2 + 1;
  {
    int64_t gx = (max_c + DT_OWN - 1) / DT_OWN;
    const int64_t cap = (148 * 12 + sh.n_items - 1) / sh.n_items;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    BPM_KERNEL(k_distance_tiles);
    if (sh.total_m >= DT_LARGE_MIN)
      k_distance_tiles<DT_THREADS_LARGE><<<dim3(static_cast<unsigned>(gx), sh.n_items), DT_THREADS_LARGE, 0, st>>>(
          x, sign, items, b.cand, b.cand_count, distance, b.cstate, b.pending, b.ticket, ci);
    else
      k_distance_tiles<DT_THREADS_SMALL><<<dim3(static_cast<unsigned>(gx), sh.n_items), DT_THREADS_SMALL, 0, st>>>(
          x, sign, items, b.cand, b.cand_count, distance, b.cstate, b.pending, b.ticket, ci);
    BPM_LAUNCH_OK();
  }
  if (prominence_ready && cudaStreamWaitEvent(st, prominence_ready, 0) != cudaSuccess) return BPM_ERR_CUDA;
  BPM_KERNEL(k_prominence_compact);
  k_prominence_compact<<<dim3(cdiv(max_c, PC_THREADS), sh.n_items), PC_THREADS, 0, st>>>(
      x, sign, items, b.cand, b.cand_count, b.cstate, prominence, b.status_c, b.stride_c, out_idx, out_count, ci);
  BPM_LAUNCH_OK();
  return BPM_OK;
}

}  // namespace bpm

#ifdef BPM_DEBUG_COUNTERS
extern "C" int bpm_debug_counters_peaks(unsigned long long* out_host, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out_host, bpm::g_dbg_pk, sizeof(unsigned long long) * 16);
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(bpm::g_dbg_pk, z, sizeof(z)); }
  return 0;
}
#endif
