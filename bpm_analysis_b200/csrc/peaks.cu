// K4: scipy.signal.find_peaks(sign*x, height=, prominence=, distance=), conditions in
// scipy's order (scipy/signal/_peak_finding.py:943-1010; internals restated in SURVEY.md
// Appendix A.3):  plateau-aware local maxima -> height -> distance -> prominence.
// Called by the reference at bpm_analysis.py:227 (peaks of the envelope above the noise
// floor) and :1070 (troughs = peaks of -envelope).
//
//   k_localmax_flags   per sample: is it the midpoint of a strict local maximum plateau,
//                      and (optionally) at or above its height threshold; per-tile counts.
//   k_tile_scan        exclusive scan of tile counts per recording -> offsets, totals.
//   k_scatter          ordered compaction of flagged positions (int64).
//   k_distance         greedy "highest first removes neighbours closer than d" resolved as a
//                      fix-point on clusters of candidates (a cluster = run of candidates with
//                      gaps < d; clusters never interact), staged in shared memory.
//   k_prominence       warp-cooperative prominence walk (both sides at once, 32 samples per
//                      step, ballot early exit) for every survivor.
// Equal-height candidates within `distance`: the later index wins (what a stable argsort
// gives scipy); numpy's default sort is unstable, so the reference does not pin this case.
#include "common.cuh"

namespace bpm {

constexpr int PK_THREADS = 256;
constexpr int PK_PER = 8;
constexpr int PK_TILE = PK_THREADS * PK_PER;     // domain elements per tile

__device__ __forceinline__ int64_t pk_slot0(int64_t dom_off, int item) { return dom_off / PK_TILE + item; }

// ---- plateaus (scipy _local_maxima_1d: the midpoint (left + right) // 2 of a flat run that is
// higher than both neighbours is the peak; runs touching either end are not peaks).
// Sample i is that midpoint iff  dr == dl or dr == dl + 1  (dl / dr = equal samples to its left /
// right).  Both sides are walked TOGETHER and the walk stops as soon as one side has ended and the
// other is already too long, i.e. after min(dl, dr) + 2 steps -- a flat run of length L costs
// O(L * min(L, cap)) in total instead of O(L^2), so digital silence does not stall the kernel.
// Runs longer than 2 * cap + 1 cannot be settled by their midpoint within `cap` steps: the sample
// exactly `cap` to the right of such a run's left edge registers the run (PLATEAU_REGISTER) and a
// warp finishes it afterwards (k_long_plateaus).
constexpr int PL_CAP = 512;
constexpr int PL_PER_TILE = 8;          // registering samples lie > PL_CAP apart: at most 4 per 2048-sample tile
enum { PLATEAU_NO = 0, PLATEAU_MID = 1, PLATEAU_REGISTER = 2 };

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
  // the run is longer than cap on the right; register it from the one sample `cap` past its left edge
  if (dl == cap && dr < 0 && i - dl >= 1 && val(i - dl - 1) < c) return PLATEAU_REGISTER;
  return PLATEAU_NO;
}

__global__ void __launch_bounds__(PK_THREADS) k_localmax_flags(const double* __restrict__ x, int sign,
                                                               const double* __restrict__ height,
                                                               const BpmItem* __restrict__ items,
                                                               unsigned char* __restrict__ flags,
                                                               int* __restrict__ tile_counts,
                                                               int64_t* __restrict__ long_runs,
                                                               int* __restrict__ long_count) {
  __shared__ int s_cnt[PK_THREADS / 32];
  __shared__ int s_nreg;                      // long flat runs registered by this tile (<= PL_PER_TILE)
  __shared__ long long s_reg[PL_PER_TILE];
  if (threadIdx.x == 0) s_nreg = 0;
  __syncthreads();
  const int item = blockIdx.y;
  const BpmItem it = items[item];
  const int64_t i0 = static_cast<int64_t>(blockIdx.x) * PK_TILE;
  if (i0 >= it.m) return;
  const double* __restrict__ xi = x + it.m_off;
  const int64_t n = it.m;
  int cnt = 0;
#pragma unroll
  for (int k = 0; k < PK_PER; ++k) {
    const int64_t i = i0 + k * PK_THREADS + threadIdx.x;
    if (i >= n) continue;
    bool pk = false;
    if (i >= 1 && i <= n - 2) {
      const double c = signed_val(xi[i], sign);
      const double l = signed_val(xi[i - 1], sign), r = signed_val(xi[i + 1], sign);
      if (l < c && r < c) {
        pk = true;
      } else if ((l == c || r == c) && l <= c && r <= c) {
        const int t = plateau_test([&](int64_t j) { return signed_val(xi[j], sign); }, i, n, c, PL_CAP);
        pk = (t == PLATEAU_MID);
        if (t == PLATEAU_REGISTER) {
          const int slot = atomicAdd(&s_nreg, 1);
          if (slot < PL_PER_TILE) s_reg[slot] = i - PL_CAP;              // left edge of the run
        }
      }
      if (pk && height != nullptr) pk = (height[it.m_off + i] <= c);
    }
    flags[it.m_off + i] = pk ? 1 : 0;
    cnt += pk ? 1 : 0;
  }
  cnt = warp_sum(cnt);
  if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < PK_THREADS / 32; ++w) t += s_cnt[w];
    const int64_t slot = pk_slot0(it.m_off, item) + blockIdx.x;
    tile_counts[slot] = t;
    // every tile writes its (almost always zero) registration count: no memset, no global atomics
    const int nr = s_nreg < PL_PER_TILE ? s_nreg : PL_PER_TILE;
    long_count[slot] = nr;
    for (int e = 0; e < nr; ++e) long_runs[slot * PL_PER_TILE + e] = s_reg[e];
  }
}

// flat runs longer than 2 * PL_CAP + 1 (registered by k_localmax_flags): one warp walks each run
// 32 samples per step, flags its midpoint if the run is a strict local maximum and bumps the
// count of the tile the midpoint lies in.  Almost always there is nothing to do.  Runs at the top
// of k_tile_scan (one CTA per recording, after every k_localmax_flags CTA has finished).
struct LongRuns {
  const double* x;          // nullptr: no plateau stage in this compaction
  const double* height;
  const int64_t* runs;
  const int* count;
  unsigned char* flags;
  int sign;
};

__device__ void finish_long_plateaus(const LongRuns& lr, const BpmItem& it, int item, int* __restrict__ tile_counts) {
  const double* __restrict__ xi = lr.x + it.m_off;
  const int64_t n = it.m;
  const int lane = threadIdx.x & 31;
  const int64_t nt = (n + 2047) / 2048, slot0 = it.m_off / 2048 + item;        // sample tiles (PK_TILE)
  // the warps look at 32 tiles at a time; tiles with registered runs are rare
  for (int64_t tb = static_cast<int64_t>(threadIdx.x >> 5) * 32; tb < nt; tb += static_cast<int64_t>(blockDim.x >> 5) * 32) {
    const int mine = (tb + lane < nt) ? lr.count[slot0 + tb + lane] : 0;
    unsigned pending = __ballot_sync(0xffffffffu, mine > 0);
    while (pending) {
      const int src = __ffs(pending) - 1;
      pending &= pending - 1;
      const int cnt = __shfl_sync(0xffffffffu, mine, src);
      for (int e = 0; e < cnt; ++e) {
    const int64_t L = lr.runs[(slot0 + tb + src) * PL_PER_TILE + e];
    const double c = signed_val(xi[L], lr.sign);
    int64_t R = L;
    while (true) {
      const int64_t j = R + 1 + lane;
      const bool eq = (j <= n - 1) && signed_val(xi[j], lr.sign) == c;
      const unsigned m = __ballot_sync(0xffffffffu, eq);
      if (m == 0xffffffffu) { R += 32; continue; }
      R += __ffs(~m) - 1;
      break;
    }
    if (R - L + 1 <= 2 * PL_CAP + 1) continue;          // short enough: its midpoint settled it already
    if (R > n - 2 || !(signed_val(xi[R + 1], lr.sign) < c)) continue;
    const int64_t mid = (L + R) / 2;
    if (lr.height != nullptr && !(lr.height[it.m_off + mid] <= c)) continue;
    if (lane == 0) {
      lr.flags[it.m_off + mid] = 1;
      atomicAdd(tile_counts + pk_slot0(it.m_off, item) + mid / PK_TILE, 1);
    }
      }
    }
  }
}

// per recording: tile_counts -> exclusive offsets (in place), total -> totals[item]
// dom_len == nullptr: the domain is the recording's m samples.
__global__ void __launch_bounds__(256) k_tile_scan(const BpmItem* __restrict__ items,
                                                   const int64_t* __restrict__ dom_len,
                                                   int* __restrict__ tile_counts, int64_t* __restrict__ totals,
                                                   LongRuns lr) {
  __shared__ int s_scan[34];
  __shared__ int s_carry;
  const int item = blockIdx.x;
  const BpmItem it = items[item];
  if (lr.x != nullptr) {
    finish_long_plateaus(lr, it, item, tile_counts);
    __threadfence();
    __syncthreads();
  }
  const int64_t len = dom_len ? dom_len[item] : it.m;
  const int64_t nt = (len + PK_TILE - 1) / PK_TILE;
  int* tc = tile_counts + pk_slot0(it.m_off, item);
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < nt; base += blockDim.x) {
    const int64_t t = base + threadIdx.x;
    const int v = (t < nt) ? tc[t] : 0;
    int total;
    const int ex = block_exclusive_scan(v, &total, s_scan);
    const int carry = s_carry;
    if (t < nt) tc[t] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) s_carry = carry + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) totals[item] = s_carry;
}

// ordered compaction: out[m_off + rank] = value of every flagged domain element.
// src == nullptr: the value is the element's own index; else value = src[m_off + index].
__global__ void __launch_bounds__(PK_THREADS) k_scatter(const unsigned char* __restrict__ flags,
                                                        const int64_t* __restrict__ src,
                                                        const BpmItem* __restrict__ items,
                                                        const int64_t* __restrict__ dom_len,
                                                        const int* __restrict__ tile_offsets,
                                                        int64_t* __restrict__ out) {
  __shared__ unsigned char s_f[PK_TILE];
  __shared__ int s_scan[34];
  const int item = blockIdx.y;
  const BpmItem it = items[item];
  const int64_t len = dom_len ? dom_len[item] : it.m;
  const int64_t i0 = static_cast<int64_t>(blockIdx.x) * PK_TILE;
  if (i0 >= len) return;
  for (int t = threadIdx.x; t < PK_TILE; t += PK_THREADS) {
    const int64_t i = i0 + t;
    s_f[t] = (i < len) ? flags[it.m_off + i] : 0;
  }
  __syncthreads();
  int c = 0;
#pragma unroll
  for (int k = 0; k < PK_PER; ++k) c += s_f[threadIdx.x * PK_PER + k];
  int total;
  int ex = block_exclusive_scan(c, &total, s_scan);
  int64_t o = it.m_off + tile_offsets[pk_slot0(it.m_off, item) + blockIdx.x] + ex;
#pragma unroll
  for (int k = 0; k < PK_PER; ++k) {
    if (s_f[threadIdx.x * PK_PER + k]) {
      const int64_t i = i0 + threadIdx.x * PK_PER + k;
      out[o++] = src ? src[it.m_off + i] : i;
    }
  }
}

// ------------------------------------------------------------------ distance + prominence
__device__ __forceinline__ bool higher_priority(double va, int64_t ka, double vb, int64_t kb) {
  return va > vb || (va == vb && ka > kb);
}

// does the prominence of the peak at p reach thr?  (scipy _peak_prominences, wlen=None)
// Warp-cooperative, both sides walked together 32 samples at a time: a side passes as soon as
// a sample low enough (x[p] - x[i] >= thr) is met before the walk would stop (a sample above
// the peak, or the end of the signal); it fails when the walk stops first.
__device__ bool warp_prominence_ok(const double* __restrict__ xi, int sign, int64_t n, int64_t p, double thr) {
  const int lane = threadIdx.x & 31;
  const double xp = signed_val(xi[p], sign);
  if (__dsub_rn(xp, xp) >= thr) return true;            // the peak itself is the running minimum
  int done_l = 0, done_r = 0;                             // 0 open, 1 passed
  int64_t bl = p - 1, br = p + 1;
  while (true) {
    const int64_t il = bl - lane, ir = br + lane;
    const bool vl = !done_l && il >= 0, vr = !done_r && ir < n;
    const double xl = vl ? signed_val(xi[il], sign) : 0.0;
    const double xr = vr ? signed_val(xi[ir], sign) : 0.0;
    if (!done_l) {
      const bool stop = !vl || (xl > xp);
      const bool pass = vl && !(xl > xp) && (__dsub_rn(xp, xl) >= thr);
      const unsigned bs = __ballot_sync(0xffffffffu, stop), bp = __ballot_sync(0xffffffffu, pass);
      const int fs = bs ? __ffs(bs) : 33, fp = bp ? __ffs(bp) : 33;
      if (fp < fs) done_l = 1;
      else if (bs) return false;
      else bl -= 32;
    }
    if (!done_r) {
      const bool stop = !vr || (xr > xp);
      const bool pass = vr && !(xr > xp) && (__dsub_rn(xp, xr) >= thr);
      const unsigned bs = __ballot_sync(0xffffffffu, stop), bp = __ballot_sync(0xffffffffu, pass);
      const int fs = bs ? __ffs(bs) : 33, fp = bp ? __ffs(bp) : 33;
      if (fp < fs) done_r = 1;
      else if (bs) return false;
      else br += 32;
    }
    if (done_l && done_r) return true;
  }
}

constexpr int DP_THREADS = 256;
constexpr int DP_TILE = 256;       // nominal candidates per CTA
constexpr int DP_CAP = 1536;       // candidates a CTA can stage in shared memory

// Distance rule: state[k] = 1 kept / 0 removed for every candidate.
__global__ void __launch_bounds__(DP_THREADS) k_distance(const double* __restrict__ x, int sign,
                                                         const BpmItem* __restrict__ items,
                                                         const int64_t* __restrict__ cand,
                                                         const int64_t* __restrict__ cand_count, int distance,
                                                         unsigned char* __restrict__ state) {
  __shared__ long long s_edge[2];
  __shared__ int s_pos[DP_CAP];
  __shared__ double s_val[DP_CAP];
  __shared__ unsigned char s_st[DP_CAP];
  const int item = blockIdx.y;
  const BpmItem it = items[item];
  const int64_t nc = cand_count[item];
  const int64_t* __restrict__ pos = cand + it.m_off;
  const double* __restrict__ xi = x + it.m_off;
  unsigned char* st_out = state + it.m_off;
  const int64_t d = distance;
  // the grid is bounded (the candidate count is only known on the device): tiles are block-strided
  for (int64_t k0 = static_cast<int64_t>(blockIdx.x) * DP_TILE; k0 < nc; k0 += static_cast<int64_t>(gridDim.x) * DP_TILE) {
  const int64_t k1 = min(nc, k0 + DP_TILE);
  __syncthreads();

  if (distance <= 1) {
    for (int64_t k = k0 + threadIdx.x; k < k1; k += DP_THREADS) st_out[k] = 1;
    continue;
  }
  // own the clusters whose head lies in [k0, k1): ks = first head >= k0, ke = first head >= k1
  // (a cluster = run of candidates with gaps < d; a head is a candidate >= d after its predecessor)
  if (threadIdx.x < 2) s_edge[threadIdx.x] = nc;
  __syncthreads();
  for (int e = 0; e < 2; ++e) {
    const int64_t from = e == 0 ? k0 : k1;
    for (int64_t base = from; base < nc; base += DP_THREADS) {
      const int64_t k = base + threadIdx.x;
      const bool head = (k < nc) && (k == 0 || pos[k] - pos[k - 1] >= d);
      if (head) atomicMin(reinterpret_cast<long long*>(&s_edge[e]), static_cast<long long>(k));
      __syncthreads();
      const bool found = s_edge[e] < nc;
      __syncthreads();
      if (found) break;
    }
  }
  const int64_t ks = s_edge[0], ke = s_edge[1];
  const int64_t len = ke - ks;
  if (len <= 0) continue;

  if (len <= DP_CAP) {
    const int L = static_cast<int>(len);
    for (int t = threadIdx.x; t < L; t += DP_THREADS) {
      const int64_t pp = pos[ks + t];
      s_pos[t] = static_cast<int>(pp);
      s_val[t] = signed_val(xi[pp], sign);
      s_st[t] = 0;
    }
    __syncthreads();
    const int di = distance;
    while (true) {
      int changed = 0;
      for (int k = threadIdx.x; k < L; k += DP_THREADS) {
        if (s_st[k] != 0) continue;
        const int pk = s_pos[k];
        const double vk = s_val[k];
        bool any_keep = false, any_open = false;
        for (int k2 = k - 1; k2 >= 0 && pk - s_pos[k2] < di; --k2) {
          if (s_val[k2] > vk) {                            // equal heights: the later index wins
            const unsigned char s2 = s_st[k2];
            any_keep |= (s2 == 1); any_open |= (s2 == 0);
          }
        }
        for (int k2 = k + 1; k2 < L && s_pos[k2] - pk < di; ++k2) {
          if (s_val[k2] >= vk) {
            const unsigned char s2 = s_st[k2];
            any_keep |= (s2 == 1); any_open |= (s2 == 0);
          }
        }
        if (any_keep) { s_st[k] = 2; changed = 1; }
        else if (!any_open) { s_st[k] = 1; changed = 1; }
      }
      if (!__syncthreads_or(changed)) break;
    }
    for (int t = threadIdx.x; t < L; t += DP_THREADS) st_out[ks + t] = (s_st[t] == 1) ? 1 : 0;
    continue;
  }

  // oversized cluster run: same fix-point straight on global memory
  volatile unsigned char* stt = st_out;
  for (int64_t k = ks + threadIdx.x; k < ke; k += DP_THREADS) stt[k] = 0;
  __syncthreads();
  while (true) {
    int changed = 0;
    for (int64_t k = ks + threadIdx.x; k < ke; k += DP_THREADS) {
      if (stt[k] != 0) continue;
      const int64_t pk = pos[k];
      const double vk = signed_val(xi[pk], sign);
      bool any_keep = false, any_open = false;
      for (int64_t k2 = k - 1; k2 >= ks && pk - pos[k2] < d; --k2) {
        if (higher_priority(signed_val(xi[pos[k2]], sign), k2, vk, k)) {
          const unsigned char s2 = stt[k2];
          any_keep |= (s2 == 1); any_open |= (s2 == 0);
        }
      }
      for (int64_t k2 = k + 1; k2 < ke && pos[k2] - pk < d; ++k2) {
        if (higher_priority(signed_val(xi[pos[k2]], sign), k2, vk, k)) {
          const unsigned char s2 = stt[k2];
          any_keep |= (s2 == 1); any_open |= (s2 == 0);
        }
      }
      if (any_keep) { stt[k] = 2; changed = 1; }
      else if (!any_open) { stt[k] = 1; changed = 1; }
    }
    if (!__syncthreads_or(changed)) break;
  }
  for (int64_t k = ks + threadIdx.x; k < ke; k += DP_THREADS) stt[k] = (stt[k] == 1) ? 1 : 0;
  }  // tile loop
}

// Prominence rule on the survivors of the distance rule: one warp per candidate, grid-stride.
constexpr int PR_THREADS = 256;
constexpr int PR_BLOCKS = 148 * 8;

__global__ void __launch_bounds__(PR_THREADS) k_prominence(const double* __restrict__ x, int sign,
                                                           const BpmItem* __restrict__ items,
                                                           const int64_t* __restrict__ cand,
                                                           const int64_t* __restrict__ cand_count,
                                                           const double* __restrict__ prominence,
                                                           unsigned char* __restrict__ state) {
  const int item = blockIdx.y;
  const BpmItem it = items[item];
  const int64_t nc = cand_count[item];
  const double thr = prominence[item];
  const int64_t* __restrict__ pos = cand + it.m_off;
  const double* __restrict__ xi = x + it.m_off;
  unsigned char* stt = state + it.m_off;
  const int lane = threadIdx.x & 31;
  const int64_t warps = static_cast<int64_t>(gridDim.x) * (PR_THREADS / 32);
  // a warp takes 32 consecutive candidates at a time: state and position come in with one coalesced
  // load each, then the survivors of the distance step are walked one after the other
  for (int64_t c0 = (static_cast<int64_t>(blockIdx.x) * (PR_THREADS / 32) + (threadIdx.x >> 5)) * 32; c0 < nc;
       c0 += warps * 32) {
    const int64_t k = c0 + lane;
    const bool live = (k < nc) && stt[k] == 1;
    const int64_t p = live ? pos[k] : 0;
    unsigned todo = __ballot_sync(0xffffffffu, live);
    bool drop = false;
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const int64_t ps = __shfl_sync(0xffffffffu, p, src);
      const bool ok = warp_prominence_ok(xi, sign, it.m, ps, thr);
      if (lane == src && !ok) drop = true;
    }
    if (drop) stt[k] = 0;
  }
}

// counts of set flags per PK_TILE of a (device-length) domain
__global__ void __launch_bounds__(PK_THREADS) k_count_flags(const unsigned char* __restrict__ flags,
                                                            const BpmItem* __restrict__ items,
                                                            const int64_t* __restrict__ dom_len,
                                                            int* __restrict__ tile_counts) {
  __shared__ int s_cnt[PK_THREADS / 32];
  const int item = blockIdx.y;
  const BpmItem it = items[item];
  const int64_t len = dom_len ? dom_len[item] : it.m;
  const int64_t i0 = static_cast<int64_t>(blockIdx.x) * PK_TILE;
  if (i0 >= len) return;
  int cnt = 0;
  for (int k = 0; k < PK_PER; ++k) {
    const int64_t i = i0 + k * PK_THREADS + threadIdx.x;
    if (i < len) cnt += flags[it.m_off + i];
  }
  cnt = warp_sum(cnt);
  if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < PK_THREADS / 32; ++w) t += s_cnt[w];
    tile_counts[pk_slot0(it.m_off, item) + blockIdx.x] = t;
  }
}

// Prominence test with per-32-sample block minima / maxima (shared memory): whole blocks that
// contain neither a stop (a sample above the peak) nor a pass (a sample low enough) are skipped
// 32 at a time.  Same decision as warp_prominence_ok.
__device__ bool warp_prominence_ok_blocks(const double* __restrict__ xs, const double* __restrict__ bmin,
                                          const double* __restrict__ bmax, int n, int p, double thr) {
  const int lane = threadIdx.x & 31;
  const double xp = xs[p];
  if (__dsub_rn(xp, xp) >= thr) return true;
  const int nblk = (n + 31) >> 5;
#pragma unroll 1
  for (int side = 0; side < 2; ++side) {
    const int dir = side == 0 ? -1 : +1;
    int pos = p + dir;                                     // next sample to examine
    bool decided = false, ok = false;
    while (!decided) {
      if (pos < 0 || pos >= n) { ok = false; break; }      // ran off the signal: walk stops
      // sample-wise over the rest of the current block (in walking direction)
      const int blk = pos >> 5;
      const int idx = pos + dir * lane;
      const bool valid = idx >= 0 && idx < n && (idx >> 5) == blk;
      const double v = valid ? xs[idx] : 0.0;
      const bool stop = valid && (v > xp);
      const bool pass = valid && !(v > xp) && (__dsub_rn(xp, v) >= thr);
      const unsigned bs = __ballot_sync(0xffffffffu, stop), bp = __ballot_sync(0xffffffffu, pass);
      const int fs = bs ? __ffs(bs) : 33, fp = bp ? __ffs(bp) : 33;
      if (fp < fs) { ok = true; decided = true; break; }
      if (bs) { ok = false; decided = true; break; }
      // move to the edge of the next block and skip whole quiet blocks 32 at a time
      int nb = blk + dir;
      while (true) {
        if (nb < 0 || nb >= nblk) { pos = dir < 0 ? -1 : n; break; }
        const int b = nb + dir * lane;
        const bool bvalid = b >= 0 && b < nblk;
        const bool hot = bvalid && ((bmax[b] > xp) || (__dsub_rn(xp, bmin[b]) >= thr));
        const unsigned bh = __ballot_sync(0xffffffffu, hot || !bvalid);
        if (bh == 0) { nb += dir * 32; continue; }
        const int first = __ffs(bh) - 1;                   // first hot (or out-of-range) block on the way
        const int bsel = nb + dir * first;
        if (bsel < 0 || bsel >= nblk) pos = dir < 0 ? -1 : n;
        else pos = dir < 0 ? (bsel << 5) + 31 : (bsel << 5);
        if (pos >= n) pos = n - 1;                          // ragged last block, walking left into it
        break;
      }
    }
    if (!ok) return false;
  }
  return true;
}

// ------------------------------------------------------------------ one-CTA find_peaks
// Short signals (the BPM series of the beat-list reductions, short recordings): the whole
// pipeline -- local maxima, height, ordered compaction, distance fix-point, prominence,
// final compaction -- in ONE launch, signal and candidate list staged in shared memory.
constexpr int FPS_THREADS = 512;
constexpr int FPS_MAXN = 8192;

__global__ void __launch_bounds__(FPS_THREADS) k_find_peaks_small(const double* __restrict__ x, int sign,
                                                                  const double* __restrict__ height,
                                                                  const double* __restrict__ prominence, int distance,
                                                                  const BpmItem* __restrict__ items,
                                                                  int64_t* __restrict__ out_idx,
                                                                  int64_t* __restrict__ out_count) {
  extern __shared__ __align__(16) unsigned char fps_raw[];
  double* xs = reinterpret_cast<double*>(fps_raw);                       // [FPS_MAXN]
  int* cpos = reinterpret_cast<int*>(xs + FPS_MAXN);                     // [FPS_MAXN / 2 + 1]
  unsigned char* cst = reinterpret_cast<unsigned char*>(cpos + FPS_MAXN / 2 + 1);
  __shared__ double s_bmin[FPS_MAXN / 32], s_bmax[FPS_MAXN / 32];
  __shared__ int s_scan[34];
  __shared__ int s_base;
  const int item = blockIdx.x;
  const BpmItem it = items[item];
  const int n = static_cast<int>(it.m);
  const int tid = threadIdx.x;
  const double* __restrict__ xi = x + it.m_off;
  for (int i = tid; i < n; i += FPS_THREADS) xs[i] = signed_val(xi[i], sign);
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int b = tid; b < (n + 31) / 32; b += FPS_THREADS) {
    double mn = INFINITY, mx = -INFINITY;
    for (int i = b * 32; i < min(n, b * 32 + 32); ++i) { mn = fmin(mn, xs[i]); mx = fmax(mx, xs[i]); }
    s_bmin[b] = mn; s_bmax[b] = mx;
  }
  // ---- local maxima (plateau midpoints) + height, ordered: every thread owns a contiguous run of
  //      samples, so ONE block scan places all candidates
  {
    const int per = (n + FPS_THREADS - 1) / FPS_THREADS;
    const int i0 = tid * per, i1 = min(n, i0 + per);
    auto is_peak = [&](int i) -> bool {
      if (i < 1 || i > n - 2) return false;
      bool pk = false;
      const double c = xs[i], l = xs[i - 1], r = xs[i + 1];
      if (l < c && r < c) {
        pk = true;
      } else if ((l == c || r == c) && l <= c && r <= c) {
        pk = plateau_test([&](int64_t j) { return xs[j]; }, i, n, c, FPS_MAXN) == PLATEAU_MID;
      }
      if (pk && height != nullptr) pk = (height[it.m_off + i] <= c);
      return pk;
    };
    int cnt = 0;
    for (int i = i0; i < i1; ++i) cnt += is_peak(i) ? 1 : 0;
    int total;
    int ex = block_exclusive_scan(cnt, &total, s_scan);
    for (int i = i0; i < i1; ++i)
      if (is_peak(i)) { cpos[ex] = i; cst[ex] = 0; ++ex; }
    if (tid == 0) s_base = total;
    __syncthreads();
  }
  const int nc = s_base;
  // ---- distance: fix-point of "kept iff no kept higher-priority candidate closer than d"
  // (lock-step rounds; polling without barriers was measured slower)
  if (distance > 1) {
    while (true) {
      int changed = 0;
      for (int k = tid; k < nc; k += FPS_THREADS) {
        if (cst[k] != 0) continue;
        const int pk = cpos[k];
        const double vk = xs[pk];
        bool any_keep = false, any_open = false;
        for (int k2 = k - 1; k2 >= 0 && pk - cpos[k2] < distance; --k2) {
          if (xs[cpos[k2]] > vk) { const unsigned char s2 = cst[k2]; any_keep |= (s2 == 1); any_open |= (s2 == 0); }
        }
        for (int k2 = k + 1; k2 < nc && cpos[k2] - pk < distance; ++k2) {
          if (xs[cpos[k2]] >= vk) { const unsigned char s2 = cst[k2]; any_keep |= (s2 == 1); any_open |= (s2 == 0); }
        }
        if (any_keep) { cst[k] = 2; changed = 1; }
        else if (!any_open) { cst[k] = 1; changed = 1; }
      }
      if (!__syncthreads_or(changed)) break;
    }
  } else {
    for (int k = tid; k < nc; k += FPS_THREADS) cst[k] = 1;
    __syncthreads();
  }
  // ---- prominence: one warp per survivor
  if (prominence != nullptr) {
    const double thr = prominence[item];
    const int warp = tid >> 5, lane = tid & 31;
    for (int k = warp; k < nc; k += FPS_THREADS / 32) {
      if (cst[k] != 1) continue;
      const bool ok = warp_prominence_ok_blocks(xs, s_bmin, s_bmax, n, cpos[k], thr);
      if (lane == 0 && !ok) cst[k] = 2;
    }
    __syncthreads();
  }
  // ---- ordered output (contiguous runs of candidates per thread, one block scan)
  int64_t* oi = out_idx + it.m_off;
  {
    const int per = (nc + FPS_THREADS - 1) / FPS_THREADS;
    const int k0 = tid * per, k1 = min(nc, k0 + per);
    int cnt = 0;
    for (int k = k0; k < k1; ++k) cnt += (cst[k] == 1) ? 1 : 0;
    int total;
    int ex = block_exclusive_scan(cnt, &total, s_scan);
    for (int k = k0; k < k1; ++k)
      if (cst[k] == 1) oi[ex++] = cpos[k];
    if (tid == 0) s_base = total;
    __syncthreads();
  }
  if (tid == 0) out_count[item] = s_base;
}

// ------------------------------------------------------------------ host side
struct PeakBuffers {
  unsigned char* flags;     // [total_m]  sample-domain flags, then candidate-domain flags
  unsigned char* cstate;    // [total_m]
  int* tile_counts;         // [total_m / PK_TILE + n_items + 1]
  int64_t* cand;            // [total_m]
  int64_t* cand_count;      // [n_items]
  int64_t* long_runs;       // [tiles][PL_PER_TILE] left edges of very long flat runs, per sample tile
  int* long_count;          // [tiles]
};

static int carve_peaks(Workspace& ws, int64_t total_m, int n_items, PeakBuffers* b) {
  b->flags = ws.take<unsigned char>(total_m);
  b->cstate = ws.take<unsigned char>(total_m);
  b->tile_counts = ws.take<int>(total_m / PK_TILE + n_items + 1);
  b->cand = ws.take<int64_t>(total_m);
  b->cand_count = ws.take<int64_t>(n_items);
  b->long_runs = ws.take<int64_t>((total_m / PK_TILE + n_items + 1) * PL_PER_TILE);
  b->long_count = ws.take<int>(total_m / PK_TILE + n_items + 1);
  return ws.overflow ? BPM_ERR_WORKSPACE : BPM_OK;
}

size_t find_peaks_workspace_bytes(int64_t total_m, int n_items) {
  Workspace ws(nullptr, 0);
  PeakBuffers b;
  carve_peaks(ws, total_m, n_items, &b);
  return ws.used;
}

// compaction of a flag array over a domain (samples or a device-length list)
static int compact_run_lr(const unsigned char* flags, const int64_t* src, const BpmItem* items, const BatchShape& sh,
                          const int64_t* dom_len, int64_t max_len, bool counts_ready, int* tile_counts,
                          int64_t* out, int64_t* out_count, const LongRuns& lr, cudaStream_t st) {
  const dim3 grid(cdiv(max_len > 0 ? max_len : 1, PK_TILE), sh.n_items);
  if (!counts_ready) {
    BPM_KERNEL(k_count_flags);
    k_count_flags<<<grid, PK_THREADS, 0, st>>>(flags, items, dom_len, tile_counts);
    BPM_LAUNCH_OK();
  }
  BPM_KERNEL(k_tile_scan);
  k_tile_scan<<<sh.n_items, 256, 0, st>>>(items, dom_len, tile_counts, out_count, lr);
  BPM_LAUNCH_OK();
  BPM_KERNEL(k_scatter);
  k_scatter<<<grid, PK_THREADS, 0, st>>>(flags, src, items, dom_len, tile_counts, out);
  BPM_LAUNCH_OK();
  return BPM_OK;
}

int compact_run(const unsigned char* flags, const int64_t* src, const BpmItem* items, const BatchShape& sh,
                const int64_t* dom_len, int64_t max_len, bool counts_ready, int* tile_counts,
                int64_t* out, int64_t* out_count, cudaStream_t st) {
  return compact_run_lr(flags, src, items, sh, dom_len, max_len, counts_ready, tile_counts, out, out_count,
                        LongRuns{nullptr, nullptr, nullptr, nullptr, nullptr, 1}, st);
}

// prominence_ready: event after which the prominence threshold is valid (it may be produced on
// another stream while the local-maximum / distance steps run here); nullptr = already valid.
int find_peaks_run(const double* x, int sign, const double* height, const double* prominence, int distance,
                   const BpmItem* items, const BatchShape& sh, int64_t* out_idx, int64_t* out_count,
                   Workspace& ws, cudaStream_t st, cudaEvent_t prominence_ready) {
  if (!x || !items || !out_idx || !out_count || sh.n_items <= 0 || distance < 1) return BPM_ERR_ARG;
  if (sh.max_m <= FPS_MAXN) {
    if (prominence_ready && cudaStreamWaitEvent(st, prominence_ready, 0) != cudaSuccess) return BPM_ERR_CUDA;
    const size_t smem = sizeof(double) * FPS_MAXN + sizeof(int) * (FPS_MAXN / 2 + 1) + (FPS_MAXN / 2 + 1);
    cudaFuncSetAttribute(k_find_peaks_small, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    BPM_KERNEL(k_find_peaks_small);
    k_find_peaks_small<<<sh.n_items, FPS_THREADS, smem, st>>>(x, sign, height, prominence, distance, items, out_idx,
                                                            out_count);
    BPM_LAUNCH_OK();
    return BPM_OK;
  }
  PeakBuffers b;
  BPM_TRY(carve_peaks(ws, sh.total_m, sh.n_items, &b));
  const dim3 grid(cdiv(sh.max_m, PK_TILE), sh.n_items);
  BPM_KERNEL(k_localmax_flags);
  k_localmax_flags<<<grid, PK_THREADS, 0, st>>>(x, sign, height, items, b.flags, b.tile_counts, b.long_runs, b.long_count);
  BPM_LAUNCH_OK();
  BPM_TRY(compact_run_lr(b.flags, nullptr, items, sh, nullptr, sh.max_m, true, b.tile_counts, b.cand, b.cand_count,
                         LongRuns{x, height, b.long_runs, b.long_count, b.flags, sign}, st));
  // a local maximum needs a lower neighbour on both sides: at most (m-1)/2 candidates
  const int64_t max_c = sh.max_m / 2 + 1;
  BPM_KERNEL(k_distance);
  {
    int64_t gx = (max_c + DP_TILE - 1) / DP_TILE;
    const int64_t cap = (148 * 8 + sh.n_items - 1) / sh.n_items;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    k_distance<<<dim3(static_cast<unsigned>(gx), sh.n_items), DP_THREADS, 0, st>>>(x, sign, items, b.cand,
                                                                                  b.cand_count, distance, b.cstate);
  }
  BPM_LAUNCH_OK();
  if (prominence_ready && cudaStreamWaitEvent(st, prominence_ready, 0) != cudaSuccess) return BPM_ERR_CUDA;
  if (prominence != nullptr) {
    const int64_t want = (max_c + PR_THREADS / 32 - 1) / (PR_THREADS / 32);
    const unsigned gx = static_cast<unsigned>(want < PR_BLOCKS ? (want > 0 ? want : 1) : PR_BLOCKS);
    BPM_KERNEL(k_prominence);
    k_prominence<<<dim3(gx, sh.n_items), PR_THREADS, 0, st>>>(x, sign, items, b.cand, b.cand_count, prominence,
                                                              b.cstate);
    BPM_LAUNCH_OK();
  }
  BPM_TRY(compact_run(b.cstate, b.cand, items, sh, b.cand_count, max_c, false, b.tile_counts, out_idx, out_count, st));
  return BPM_OK;
}

}  // namespace bpm
