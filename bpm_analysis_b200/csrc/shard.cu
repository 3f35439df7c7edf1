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
                                                         unsigned long long prefix, unsigned long long* __restrict__ hist) {
  __shared__ unsigned int s_hist[KH_MAXBINS];
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
// count_min[1] = min(count_min[1], smallest key ABOVE the bucket)
__global__ void __launch_bounds__(KH_THREADS) k_key_collect(const double* __restrict__ x, int64_t n, int up,
                                                            unsigned long long prefix, int64_t cap,
                                                            unsigned long long* __restrict__ out_keys,
                                                            unsigned long long* __restrict__ count_min) {
  __shared__ unsigned long long s_min[KH_THREADS / 32];
  unsigned long long best = ~0ull;
  const int64_t i0 = static_cast<int64_t>(blockIdx.x) * KH_THREADS * KH_PER;
  for (int k = 0; k < KH_PER; ++k) {
    const int64_t i = i0 + k * KH_THREADS + threadIdx.x;
    if (i >= n) continue;
    const unsigned long long key = f64_key(x[i]);
    const unsigned long long top = key >> up;
    if (top == prefix) {
      const unsigned long long pos = atomicAdd(count_min, 1ull);
      if (static_cast<int64_t>(pos) < cap) out_keys[pos] = key;
    } else if (top > prefix && key < best) {
      best = key;
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o);
    best = t < best ? t : best;
  }
  if ((threadIdx.x & 31) == 0) s_min[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < KH_THREADS / 32; ++w) best = s_min[w] < best ? s_min[w] : best;
    if (best != ~0ull) atomicMin(count_min + 1, best);
  }
}

}  // namespace bpm

using namespace bpm;

extern "C" {

int bpm_key_histogram(const double* x, int64_t n, int shift, int bits, uint64_t prefix, uint64_t* hist, void* stream) {
  if (!hist || n < 0 || bits < 1 || bits > 11 || shift < 0 || shift + bits > 64) return BPM_ERR_ARG;
  if (n == 0) return BPM_OK;                      // a rank without samples contributes nothing
  if (!x) return BPM_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BPM_KERNEL(k_key_hist);
  k_key_hist<<<cdiv(n, KH_THREADS * KH_PER), KH_THREADS, 0, st>>>(x, n, shift, bits, prefix,
                                                                  reinterpret_cast<unsigned long long*>(hist));
  BPM_LAUNCH_OK();
  return BPM_OK;
}

int bpm_key_collect(const double* x, int64_t n, int up_shift, uint64_t prefix, int64_t cap, uint64_t* out_keys,
                    uint64_t* count_min, void* stream) {
  if (!out_keys || !count_min || n < 0 || cap < 0 || up_shift < 0 || up_shift > 63) return BPM_ERR_ARG;
  if (n == 0) return BPM_OK;
  if (!x) return BPM_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BPM_KERNEL(k_key_collect);
  k_key_collect<<<cdiv(n, KH_THREADS * KH_PER), KH_THREADS, 0, st>>>(
      x, n, up_shift, prefix, cap, reinterpret_cast<unsigned long long*>(out_keys),
      reinterpret_cast<unsigned long long*>(count_min));
  BPM_LAUNCH_OK();
  return BPM_OK;
}

}  // extern "C"
