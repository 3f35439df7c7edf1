// Pieces the time-chunked (multi-GPU) evaluation of ONE long recording needs beside the regular
// operators (bpm_analysis_b200/stream.py):
//
//   bpm_key_histogram / bpm_key_collect   one digit pass / the final bucket of an EXACT np.quantile
//       (bpm_analysis.py:1067, :225) over a stream whose samples are spread over several ranks: every
//       rank histograms its own samples, the histograms are summed over NVLink (ncclAllReduce of 2048
//       counters), all ranks pick the same bucket, and after two or three digits the few keys left in
//       the bucket are gathered and ordered.  Same order-preserving 64-bit keys as select.cu.
// The chunk-mode operators themselves live beside their unchunked forms (pipeline.cu, peaks.cu):
//   bpm_noise_floor_chunk   _calculate_dynamic_noise_floor (:1064-1117) on one chunk + halo with the
//       stream-wide thresholds GIVEN and without the count-based fall-backs (those are decided on the
//       stream's totals by the caller).
//   bpm_find_peaks_chunk    find_peaks on a chunk + halo, reporting every decision that could depend on
//       samples outside the chunk (edge_hits, anchors).
//   bpm_deviation_series    the deviation / smoothed-deviation series (:96-100) from a GIVEN strength list
//       (the smoothing window is 5 % of the stream's total peak count, so it runs after the gather).
#include "common.cuh"

namespace bpm {

constexpr int KH_THREADS = 256;
constexpr int KH_PER = 16;
constexpr int KH_MAXBINS = 2048;

// hist[bin] += #{ i : (key(x_i) >> up) == prefix (or any, if up >= 64) and ((key(x_i) >> shift) & mask) == bin }
__global__ void __launch_bounds__(KH_THREADS) k_key_hist(const double* __restrict__ x, int64_t n, int shift, int bits,
                                                         unsigned long long prefix,
                                                         const unsigned long long* __restrict__ state,
                                                         unsigned long long* __restrict__ hist) {
  __shared__ unsigned int s_hist[KH_MAXBINS];
  if (state != nullptr) prefix = state[0];                      // the descent's state lives on the device
  const int nb = 1 << bits;
  for (int t = threadIdx.x; t < nb; t += KH_THREADS) s_hist[t] = 0;
  __syncthreads();
  const int up = shift + bits;
  const unsigned int mask = static_cast<unsigned int>(nb - 1);
  const int64_t i0 = static_cast<int64_t>(blockIdx.x) * KH_THREADS * KH_PER;
#pragma unroll 4
  for (int k = 0; k < KH_PER; ++k) {
    const int64_t i = i0 + k * KH_THREADS + threadIdx.x;
    const bool in = i < n;
    const unsigned long long key = in ? f64_key(x[i]) : 0ull;
    const bool match = in && (up >= 64 || (key >> up) == prefix);
    const unsigned int bin = static_cast<unsigned int>(key >> shift) & mask;
    // one shared-memory atomic per distinct bin per warp (envelope values share their leading digits)
    const unsigned act = __ballot_sync(0xffffffffu, match);
    if (match) {
      const unsigned peers = __match_any_sync(act, bin);
      if ((threadIdx.x & 31) == (__ffs(peers) - 1)) atomicAdd(&s_hist[bin], static_cast<unsigned int>(__popc(peers)));
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < nb; t += KH_THREADS) {
    const unsigned int c = s_hist[t];
    if (c) atomicAdd(hist + t, static_cast<unsigned long long>(c));
  }
}

// keys whose leading digits equal `prefix` (key >> up == prefix) -> out_keys[0 .. min(count, cap)), count_min[0] += count,
// count_min[1] = min(count_min[1], smallest key ABOVE the bucket), count_min[2] / [3] = smallest / largest key IN the
// bucket (a bucket of one repeated value -- digital silence, clipping -- needs no ordering however large it is)
__global__ void __launch_bounds__(KH_THREADS) k_key_collect(const double* __restrict__ x, int64_t n, int up,
                                                            unsigned long long prefix,
                                                            const unsigned long long* __restrict__ state, int64_t cap,
                                                            unsigned long long* __restrict__ out_keys,
                                                            unsigned long long* __restrict__ count_min) {
  __shared__ unsigned long long s_min[KH_THREADS / 32], s_lo[KH_THREADS / 32], s_hi[KH_THREADS / 32];
  if (state != nullptr) prefix = state[0];
  unsigned long long best = ~0ull, lo = ~0ull, hi = 0ull;
  const int64_t i0 = static_cast<int64_t>(blockIdx.x) * KH_THREADS * KH_PER;
  for (int k = 0; k < KH_PER; ++k) {
    const int64_t i = i0 + k * KH_THREADS + threadIdx.x;
    if (i >= n) continue;
    const unsigned long long key = f64_key(x[i]);
    const unsigned long long top = key >> up;
    if (top == prefix) {
      const unsigned long long pos = atomicAdd(count_min, 1ull);
      if (static_cast<int64_t>(pos) < cap) out_keys[pos] = key;
      lo = key < lo ? key : lo;
      hi = key > hi ? key : hi;
    } else if (top > prefix && key < best) {
      best = key;
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o);
    best = t < best ? t : best;
    const unsigned long long tl = __shfl_xor_sync(0xffffffffu, lo, o);
    lo = tl < lo ? tl : lo;
    const unsigned long long th = __shfl_xor_sync(0xffffffffu, hi, o);
    hi = th > hi ? th : hi;
  }
  if ((threadIdx.x & 31) == 0) { s_min[threadIdx.x >> 5] = best; s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < KH_THREADS / 32; ++w) {
      best = s_min[w] < best ? s_min[w] : best;
      lo = s_lo[w] < lo ? s_lo[w] : lo;
      hi = s_hi[w] > hi ? s_hi[w] : hi;
    }
    if (best != ~0ull) atomicMin(count_min + 1, best);
    if (lo != ~0ull) { atomicMin(count_min + 2, lo); atomicMax(count_min + 3, hi); }
  }
}

// One step of the descent on the device: the bin of the (summed) histogram that holds element `rank`
// becomes the next digit.  state = {prefix, rank inside the bucket, bucket size}.
constexpr int KP_THREADS = 256;
__global__ void __launch_bounds__(KP_THREADS) k_key_pick(const unsigned long long* __restrict__ hist, int bits,
                                                         unsigned long long* __restrict__ state) {
  __shared__ unsigned long long s_part[KP_THREADS];
  const int nb = 1 << bits;
  const int per = (nb + KP_THREADS - 1) / KP_THREADS;
  const int b0 = threadIdx.x * per;
  unsigned long long sum = 0;
  for (int u = 0; u < per; ++u) sum += (b0 + u < nb) ? hist[b0 + u] : 0ull;
  s_part[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long run = 0;
    for (int t = 0; t < KP_THREADS; ++t) { const unsigned long long v = s_part[t]; s_part[t] = run; run += v; }
  }
  __syncthreads();
  const unsigned long long rank = state[1];
  unsigned long long c = s_part[threadIdx.x];
  __syncthreads();
  if (rank >= c && rank < c + sum) {
    for (int u = 0; u < per && b0 + u < nb; ++u) {
      const unsigned long long h = hist[b0 + u];
      if (rank < c + h) {
        state[0] = (state[0] << bits) | static_cast<unsigned long long>(b0 + u);
        state[1] = rank - c;
        state[2] = h;
        break;
      }
      c += h;
    }
  }
}

// The end of the descent: `rows` = per rank {count, smallest key above the bucket, smallest / largest key in it,
// keys[cap]} (all-gathered);
// the bucket's keys are ordered and the two order statistics around the virtual index interpolated the
// way numpy does.  status: 0 fine, 1 = the bucket does not fit (the caller resolves more digits).
constexpr int KF_THREADS = 512;
constexpr int KF_CAP = 4096;
__global__ void __launch_bounds__(KF_THREADS) k_key_finish(const unsigned long long* __restrict__ rows, int world, int64_t cap,
                                                           const unsigned long long* __restrict__ state, double gamma,
                                                           double* __restrict__ out, long long* __restrict__ status) {
  __shared__ unsigned long long s_key[KF_CAP];
  __shared__ int s_off[65];
  const int tid = threadIdx.x;
  const int64_t pitch = cap + 4;
  unsigned long long above = ~0ull;
  for (int r = 0; r < world; ++r) above = rows[r * pitch + 1] < above ? rows[r * pitch + 1] : above;
  if (tid == 0) {
    unsigned long long total64 = 0, lo = ~0ull, hi = 0ull;
    for (int r = 0; r < world; ++r) {
      const unsigned long long* row = rows + r * pitch;
      total64 += row[0];
      if (row[0]) { lo = row[2] < lo ? row[2] : lo; hi = row[3] > hi ? row[3] : hi; }
    }
    int verdict = 0;                                          // 0: order the bucket, -1: cannot, -2: one repeated value
    if (total64 == 0 || total64 != state[2]) verdict = -1;
    else if (lo == hi) {
      verdict = -2;
      const unsigned long long kb = (state[1] + 1 < total64) ? lo : (above != ~0ull ? above : lo);
      s_key[0] = lo; s_key[1] = kb;
    } else if (total64 > static_cast<unsigned long long>(KF_CAP) || world > 64) verdict = -1;
    int run = 0;
    for (int r = 0; r < world && verdict == 0; ++r) {
      s_off[r] = run;
      const unsigned long long c = rows[r * pitch];
      if (c > static_cast<unsigned long long>(cap)) verdict = -1; else run += static_cast<int>(c);
    }
    s_off[64] = verdict == 0 ? run : verdict;
  }
  __syncthreads();
  const int total = s_off[64];
  if (total == -1) { if (tid == 0) *status = 1; return; }
  if (total == -2) {
    if (tid == 0) {
      const double a = key_f64(s_key[0]), b = key_f64(s_key[1]);
      const double diff = __dsub_rn(b, a);
      double r = __dadd_rn(a, __dmul_rn(diff, gamma));
      if (gamma >= 0.5) r = __dsub_rn(b, __dmul_rn(diff, __dsub_rn(1.0, gamma)));
      *out = r;
      *status = 0;
    }
    return;
  }
  for (int r = 0; r < world; ++r) {
    const unsigned long long* row = rows + r * pitch;
    const int c = static_cast<int>(row[0]), o = s_off[r];
    for (int t = tid; t < c; t += KF_THREADS) s_key[o + t] = row[4 + t];
  }
  int P = 32;
  while (P < total) P <<= 1;
  for (int t = total + tid; t < P; t += KF_THREADS) s_key[t] = ~0ull;
  __syncthreads();
  for (int k2 = 2; k2 <= P; k2 <<= 1) {
    for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
      for (int t = tid; t < P / 2; t += KF_THREADS) {
        const int lo_i = ((t & ~(j2 - 1)) << 1) | (t & (j2 - 1));
        const int hi_i = lo_i + j2;
        const bool upw = ((lo_i & k2) == 0);
        const unsigned long long a = s_key[lo_i], b = s_key[hi_i];
        if ((a > b) == upw) { s_key[lo_i] = b; s_key[hi_i] = a; }
      }
      __syncthreads();
    }
  }
  if (tid == 0) {
    const int rank = static_cast<int>(state[1]);
    const unsigned long long ka = s_key[rank];
    const unsigned long long kb = (rank + 1 < total) ? s_key[rank + 1] : (above != ~0ull ? above : ka);
    const double a = key_f64(ka), b = key_f64(kb);
    const double diff = __dsub_rn(b, a);
    double r = __dadd_rn(a, __dmul_rn(diff, gamma));
    if (gamma >= 0.5) r = __dsub_rn(b, __dmul_rn(diff, __dsub_rn(1.0, gamma)));
    *out = r;
    *status = 0;
  }
}

// ---------------------------------------------------------------- the chunk's proof, on the device
// (same obligations as ShardedFrontEnd._prove on the host; see bpm_b200.h)
struct ProofGeom {
  long long n, core_lo, core_hi, t_lo, t_hi;
  int at_start, at_end;
  long long filter_halo, distance, window;
};

__device__ long long lower_bound_ll(const int64_t* a, long long n, long long v) {      // first index with a[i] >= v
  long long lo = 0, hi = n;
  while (lo < hi) { const long long mid = (lo + hi) >> 1; if (a[mid] < v) lo = mid + 1; else hi = mid; }
  return lo;
}

// Outputs of a rolling quantile over np.interp(knots) that equal the stream's, given that the knot list is
// the stream's inside [lo, hi): i0 / i1 = lower bounds of lo / hi in the list.
__device__ void proven_range(const int64_t* knots, long long i0, long long i1, const ProofGeom& g, long long* x_lo,
                             long long* x_hi) {
  const long long off = (g.window - 1) / 2, left = g.window - 1 - off;
  if (i1 <= i0) {
    const bool whole = g.at_start && g.at_end;
    *x_lo = 0; *x_hi = whole ? g.n : 0;
    return;
  }
  *x_lo = g.at_start ? 0 : knots[i0] + left;
  *x_hi = g.at_end ? g.n : knots[i1 - 1] - off + 1;
}

// flags: {edge hits, trough anchors (left, right), peak anchors (left, right)}; counts: {kept, all, peaks}
// out: {bad, all troughs in core, kept in core, peaks in core, first kept in core, first peak in core, x3 lo, x3 hi}
// One warp: the binary searches (dependent global loads, ~5 us each) run side by side on its lanes.
__global__ void k_chunk_proof(const int64_t* __restrict__ every, const int64_t* __restrict__ kept,
                              const int64_t* __restrict__ peaks, const long long* __restrict__ counts,
                              const long long* __restrict__ flags, const long long* __restrict__ q_status, ProofGeom g,
                              long long* __restrict__ out) {
  if (blockIdx.x != 0 || threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  const long long n_kept = counts[0], n_all = counts[1], n_peaks = counts[2];
  // round 1 -- lane: 0 every@t_lo, 1 every@t_hi, 2 every@core_lo, 3 every@core_hi, 4 kept@core_lo, 5 kept@core_hi,
  //                  6 peaks@core_lo, 7 peaks@core_hi
  const int64_t* list = lane < 4 ? every : (lane < 6 ? kept : peaks);
  const long long len = lane < 4 ? n_all : (lane < 6 ? n_kept : n_peaks);
  const long long key = lane == 0 ? g.t_lo : lane == 1 ? g.t_hi : (lane & 1) ? g.core_hi : g.core_lo;
  const long long r1 = lane < 8 ? lower_bound_ll(list, len, key) : 0;
  long long x2lo = 0, x2hi = 0;
  if (lane == 0) {
    const long long i1 = __shfl_sync(0xffffffffu, r1, 1);
    proven_range(every, r1, i1, g, &x2lo, &x2hi);
    if (x2lo < g.t_lo) x2lo = g.t_lo;
    if (x2hi > g.t_hi) x2hi = g.t_hi;
  } else {
    (void)__shfl_sync(0xffffffffu, r1, 1);
  }
  x2lo = __shfl_sync(0xffffffffu, x2lo, 0);
  x2hi = __shfl_sync(0xffffffffu, x2hi, 0);
  // round 2 -- lane 0: kept@x2lo, lane 1: kept@x2hi
  const long long r2 = lane < 2 ? lower_bound_ll(kept, n_kept, lane == 0 ? x2lo : x2hi) : 0;
  const long long k1 = __shfl_sync(0xffffffffu, r2, 1);
  long long v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = __shfl_sync(0xffffffffu, r1, j);
  if (lane != 0) return;
  long long x3lo, x3hi;
  proven_range(kept, r2, k1, g, &x3lo, &x3hi);
  const long long d = g.distance;
  bool ok = flags[0] == 0;
  if (q_status != nullptr) ok = ok && q_status[0] == 0 && q_status[1] == 0;
  if (d > 1) {
    ok = ok && (g.at_start || flags[1] - d >= g.filter_halo);
    ok = ok && (g.at_end || flags[2] + d <= g.n - g.filter_halo);
  }
  ok = ok && x3lo <= g.core_lo && x3hi >= g.core_hi;
  if (d > 1) {
    ok = ok && (g.at_start || (flags[3] >= 0 && flags[3] - d >= x3lo));
    ok = ok && (g.at_end || (flags[4] < g.n && flags[4] + d < x3hi));
  }
  out[0] = ok ? 0 : 1;
  out[1] = v[3] - v[2]; out[2] = v[5] - v[4]; out[3] = v[7] - v[6];
  out[4] = v[4]; out[5] = v[6]; out[6] = x3lo; out[7] = x3hi;
}

// ---------------------------------------------------------------- the one list exchange
// A rank's contribution: [kept troughs of its core | peaks of its core | their strengths (bit patterns)],
// indices moved to stream coordinates, each part padded to the longest among the ranks (cap_t, cap_p).
// `proof` is the rank's own k_chunk_proof output (counts and first indices).
__global__ void __launch_bounds__(256) k_chunk_pack(const int64_t* __restrict__ kept, const int64_t* __restrict__ peaks,
                                                    const double* __restrict__ strength,
                                                    const long long* __restrict__ proof, long long e0, long long cap_t,
                                                    long long cap_p, long long* __restrict__ out) {
  const long long nk = proof[2], np = proof[3], lk = proof[4], lp = proof[5];
  const long long total = cap_t + 2 * cap_p;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long v = 0;
    if (i < cap_t) { if (i < nk) v = kept[lk + i] + e0; }
    else if (i < cap_t + cap_p) { const long long k = i - cap_t; if (k < np) v = peaks[lp + k] + e0; }
    else { const long long k = i - cap_t - cap_p; if (k < np) v = __double_as_longlong(strength[lp + k]); }
    out[i] = v;
  }
}

// rows = the all-gathered contributions, table = the all-gathered proofs: the stream's lists in rank order
__global__ void __launch_bounds__(256) k_chunk_unpack(const long long* __restrict__ rows, const long long* __restrict__ table,
                                                      int world, long long cap_t, long long cap_p,
                                                      int64_t* __restrict__ troughs, int64_t* __restrict__ peaks,
                                                      double* __restrict__ strength) {
  __shared__ long long s_t[65], s_p[65];
  if (threadIdx.x == 0) {
    long long t = 0, p = 0;
    for (int r = 0; r < world; ++r) { s_t[r] = t; s_p[r] = p; t += table[r * 8 + 2]; p += table[r * 8 + 3]; }
    s_t[world] = t; s_p[world] = p;
  }
  __syncthreads();
  const long long pitch = cap_t + 2 * cap_p;
  const int r = blockIdx.y;
  const long long nk = s_t[r + 1] - s_t[r], np = s_p[r + 1] - s_p[r];
  const long long* row = rows + r * pitch;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nk + 2 * np;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    if (i < nk) troughs[s_t[r] + i] = row[i];
    else if (i < nk + np) peaks[s_p[r] + (i - nk)] = row[cap_t + (i - nk)];
    else strength[s_p[r] + (i - nk - np)] = __longlong_as_double(row[cap_t + cap_p + (i - nk - np)]);
  }
}

}  // namespace bpm

using namespace bpm;

extern "C" {

int bpm_key_histogram(const double* x, int64_t n, int shift, int bits, uint64_t prefix, const uint64_t* state,
                      uint64_t* hist, void* stream) {
  if (!hist || n < 0 || bits < 1 || bits > 11 || shift < 0 || shift + bits > 64) return BPM_ERR_ARG;
  if (n == 0) return BPM_OK;                      // a rank without samples contributes nothing
  if (!x) return BPM_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BPM_KERNEL(k_key_hist);
  k_key_hist<<<cdiv(n, KH_THREADS * KH_PER), KH_THREADS, 0, st>>>(
      x, n, shift, bits, prefix, reinterpret_cast<const unsigned long long*>(state),
      reinterpret_cast<unsigned long long*>(hist));
  BPM_LAUNCH_OK();
  return BPM_OK;
}

int bpm_key_collect(const double* x, int64_t n, int up_shift, uint64_t prefix, const uint64_t* state, int64_t cap,
                    uint64_t* out_keys, uint64_t* count_min, void* stream) {
  if (!out_keys || !count_min || n < 0 || cap < 0 || up_shift < 0 || up_shift > 63) return BPM_ERR_ARG;
  if (n == 0) return BPM_OK;
  if (!x) return BPM_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BPM_KERNEL(k_key_collect);
  k_key_collect<<<cdiv(n, KH_THREADS * KH_PER), KH_THREADS, 0, st>>>(
      x, n, up_shift, prefix, reinterpret_cast<const unsigned long long*>(state), cap,
      reinterpret_cast<unsigned long long*>(out_keys), reinterpret_cast<unsigned long long*>(count_min));
  BPM_LAUNCH_OK();
  return BPM_OK;
}

int bpm_key_pick(const uint64_t* hist, int bits, uint64_t* state, void* stream) {
  if (!hist || !state || bits < 1 || bits > 11) return BPM_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BPM_KERNEL(k_key_pick);
  k_key_pick<<<1, KP_THREADS, 0, st>>>(
      reinterpret_cast<const unsigned long long*>(hist), bits, reinterpret_cast<unsigned long long*>(state));
  BPM_LAUNCH_OK();
  return BPM_OK;
}

int bpm_key_finish(const uint64_t* rows, int world, int64_t cap, const uint64_t* state, double gamma, double* out,
                   int64_t* status, void* stream) {
  if (!rows || !state || !out || !status || world < 1 || cap < 1) return BPM_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BPM_KERNEL(k_key_finish);
  k_key_finish<<<1, KF_THREADS, 0, st>>>(
      reinterpret_cast<const unsigned long long*>(rows), world, cap, reinterpret_cast<const unsigned long long*>(state),
      gamma, out, reinterpret_cast<long long*>(status));
  BPM_LAUNCH_OK();
  return BPM_OK;
}

int bpm_chunk_proof(const int64_t* all_troughs, const int64_t* kept_troughs, const int64_t* peaks, const int64_t* counts,
                    const int64_t* flags, const int64_t* quantile_status, int64_t n, int64_t core_lo, int64_t core_hi,
                    int64_t trough_lo, int64_t trough_hi, int at_start, int at_end, int64_t filter_halo, int distance,
                    int window, int64_t* out, void* stream) {
  if (!all_troughs || !kept_troughs || !peaks || !counts || !flags || !out) return BPM_ERR_ARG;
  const ProofGeom g{n, core_lo, core_hi, trough_lo, trough_hi, at_start, at_end, filter_halo, distance, window};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BPM_KERNEL(k_chunk_proof);
  k_chunk_proof<<<1, 32, 0, st>>>(
      all_troughs, kept_troughs, peaks, reinterpret_cast<const long long*>(counts),
      reinterpret_cast<const long long*>(flags), reinterpret_cast<const long long*>(quantile_status), g,
      reinterpret_cast<long long*>(out));
  BPM_LAUNCH_OK();
  return BPM_OK;
}

int bpm_chunk_pack(const int64_t* kept, const int64_t* peaks, const double* strength, const int64_t* proof,
                   int64_t chunk_origin, int64_t cap_troughs, int64_t cap_peaks, int64_t* out, void* stream) {
  if (!kept || !peaks || !strength || !proof || !out || cap_troughs < 0 || cap_peaks < 0) return BPM_ERR_ARG;
  const int64_t total = cap_troughs + 2 * cap_peaks;
  if (total == 0) return BPM_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int64_t g = cdiv(total, 256);
  if (g > 148 * 4) g = 148 * 4;
  BPM_KERNEL(k_chunk_pack);
  k_chunk_pack<<<static_cast<unsigned>(g), 256, 0, st>>>(kept, peaks, strength, reinterpret_cast<const long long*>(proof),
                                                         chunk_origin, cap_troughs, cap_peaks,
                                                         reinterpret_cast<long long*>(out));
  BPM_LAUNCH_OK();
  return BPM_OK;
}

int bpm_chunk_unpack(const int64_t* rows, const int64_t* table, int world, int64_t cap_troughs, int64_t cap_peaks,
                     int64_t* troughs, int64_t* peaks, double* strength, void* stream) {
  if (!rows || !table || !troughs || !peaks || !strength || world < 1 || world > 64) return BPM_ERR_ARG;
  const int64_t total = cap_troughs + 2 * cap_peaks;
  if (total == 0) return BPM_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int64_t g = cdiv(total, 256);
  if (g > 64) g = 64;
  BPM_KERNEL(k_chunk_unpack);
  k_chunk_unpack<<<dim3(static_cast<unsigned>(g), world), 256, 0, st>>>(
      reinterpret_cast<const long long*>(rows), reinterpret_cast<const long long*>(table), world, cap_troughs, cap_peaks,
      troughs, peaks, strength);
  BPM_LAUNCH_OK();
  return BPM_OK;
}

}  // extern "C"
