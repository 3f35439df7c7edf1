// Shared helpers for libbpm_b200 (sm_100a).  See include/bpm_b200.h for the ABI.
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <math.h>
#include <string.h>
#include "../../include/bpm_b200.h"

namespace bpm {

constexpr int PADLEN = 15;            // scipy filtfilt default pad for this filter (Appendix A.1)
constexpr int SCAN_CHUNK = 8;         // == design.py SCAN_CHUNK
constexpr int SCAN_THREADS = 256;     // == design.py SCAN_THREADS
constexpr int SCAN_TILE = SCAN_CHUNK * SCAN_THREADS;
constexpr int N_POW = 16;
constexpr int MIN_PERIODS = 3;       // rolling(..., min_periods=3), bpm_analysis.py:1085

// The library is driven from several host threads (GUI / server workers, the ranks of a thread world):
// the launch counter is atomic, the "kernel being launched" label is per thread.  The per-kernel event
// profile (bpm_profile_begin / end) is a single-threaded measuring mode by contract.
extern std::atomic<int64_t> g_launches;          // counted by BPM_LAUNCH_OK
extern thread_local const char* g_cur_kernel;    // set by BPM_KERNEL just before a launch
extern bool g_profiling;                         // bpm_profile_begin / bpm_profile_end
void profile_mark(const char* name, cudaStream_t st);

// BPM_KERNEL(name); name<<<...>>>(...); BPM_LAUNCH_OK();   -- `st` is the stream in scope
#define BPM_KERNEL(name) (::bpm::g_cur_kernel = #name)

#define BPM_LAUNCH_OK()                                                   \
  do {                                                                    \
    ::bpm::g_launches.fetch_add(1, std::memory_order_relaxed);            \
    if (cudaGetLastError() != cudaSuccess) return BPM_ERR_CUDA;           \
    if (::bpm::g_profiling) ::bpm::profile_mark(::bpm::g_cur_kernel, st); \
  } while (0)

#define BPM_TRY(expr)                 \
  do {                                \
    int _rc = (expr);                 \
    if (_rc != BPM_OK) return _rc;    \
  } while (0)

// ---------------------------------------------------------------- fork / join inside one call
// Two sub-steps of a stage that do not depend on each other (the envelope quantiles and the
// local-maximum / distance steps of the trough search) are enqueued on the caller's stream and on
// a per-thread auxiliary stream, tied together with events.  Under stream capture this becomes two
// parallel branches of the caller's CUDA graph; eagerly the two streams simply overlap.  The
// auxiliary stream is the library's only persistent object (one per host thread and device,
// created on first use); the events live for the duration of the call.
struct ForkJoin {
  cudaStream_t main = nullptr, aux = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool active = false;
  // after begin(): work enqueued on `aux` runs after everything already on `st`
  int begin(cudaStream_t st) {
    main = st;
    aux = st;
    if (g_profiling) return BPM_OK;                       // per-kernel timing wants one serial stream
    static thread_local cudaStream_t t_aux[16] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return BPM_OK;   // fall back to serial
    if (!t_aux[dev] && cudaStreamCreateWithFlags(&t_aux[dev], cudaStreamNonBlocking) != cudaSuccess) {
      t_aux[dev] = nullptr;
      return BPM_OK;
    }
    if (cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming) != cudaSuccess) return BPM_ERR_CUDA;
    if (cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming) != cudaSuccess) return BPM_ERR_CUDA;
    if (cudaEventRecord(ev_fork, st) != cudaSuccess) return BPM_ERR_CUDA;
    if (cudaStreamWaitEvent(t_aux[dev], ev_fork, 0) != cudaSuccess) return BPM_ERR_CUDA;
    aux = t_aux[dev];
    active = true;
    return BPM_OK;
  }
  // the aux branch is complete up to here
  int end_aux() {
    if (active && cudaEventRecord(ev_join, aux) != cudaSuccess) return BPM_ERR_CUDA;
    return BPM_OK;
  }
  // event the main stream has to wait for before using the aux branch's results (nullptr: serial)
  cudaEvent_t join_event() const { return active ? ev_join : nullptr; }
  ~ForkJoin() {
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
  }
};

// ---------------------------------------------------------------- workspace
// Bump allocator over the caller's workspace.  With base == nullptr it only
// measures, so *_workspace_bytes() and the real call share one code path.
struct Workspace {
  char* base;
  size_t cap;
  size_t used = 0;
  bool overflow = false;
  Workspace(void* b, size_t c) : base(static_cast<char*>(b)), cap(c) {}
  template <class T>
  T* take(size_t n) {
    size_t bytes = (n * sizeof(T) + 255) & ~size_t(255);
    size_t at = used;
    used += bytes;
    if (base == nullptr) return reinterpret_cast<T*>(uintptr_t(256));  // measuring
    if (used > cap) { overflow = true; return nullptr; }
    return reinterpret_cast<T*>(base + at);
  }
  bool measuring() const { return base == nullptr; }
};

struct BatchShape {
  int n_items = 0;
  int64_t total_m = 0;   // sum of m  (assumes items are packed back to back)
  int64_t max_m = 0;
  int64_t max_n_in = 0;
};

inline BatchShape batch_shape(const BpmItem* items_host, int n_items) {
  BatchShape s;
  s.n_items = n_items;
  for (int i = 0; i < n_items; ++i) {
    if (items_host[i].m > s.max_m) s.max_m = items_host[i].m;
    if (items_host[i].n_in > s.max_n_in) s.max_n_in = items_host[i].n_in;
    int64_t end = items_host[i].m_off + items_host[i].m;
    if (end > s.total_m) s.total_m = end;
  }
  return s;
}

inline unsigned cdiv(int64_t a, int64_t b) { return static_cast<unsigned>((a + b - 1) / b); }

// ------------------------------------------------------------- device helpers
// total order on doubles as unsigned keys (negative values flipped)
__host__ __device__ __forceinline__ unsigned long long f64_key(double v) {
#ifdef __CUDA_ARCH__
  unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(v));
#else
  unsigned long long b;
  memcpy(&b, &v, 8);
#endif
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_f64(unsigned long long k) {
  unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double(static_cast<long long>(b));
}

__device__ __forceinline__ double shfl_up_f64(double v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
__device__ __forceinline__ double shfl_f64(double v, int l) { return __shfl_sync(0xffffffffu, v, l); }

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// exclusive prefix of `v` across the block (blockDim.x <= 1024); total in *total
__device__ __forceinline__ int block_exclusive_scan(int v, int* total, int* smem /* >= 33 ints */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) smem[w] = inc;
  __syncthreads();
  if (w == 0) {
    int nw = (blockDim.x + 31) >> 5;
    int x = lane < nw ? smem[lane] : 0;
    int xi = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, xi, o);
      if (lane >= o) xi += t;
    }
    smem[lane] = xi - x;
    if (lane == 31) smem[32] = xi;
  }
  __syncthreads();
  int res = smem[w] + inc - v;
  *total = smem[32];
  __syncthreads();
  return res;
}

// A recording that is one TIME CHUNK of a longer stream (see peaks.cu)
struct ChunkInfo {
  int64_t core_lo, core_hi;
  int open_left, open_right;
  unsigned long long* edge_hits;     // device counter, nullptr: not a chunk
  long long* anchors;                // [0]: max position <= core_lo, [1]: min position >= core_hi - 1 of a candidate
                                     // that outranks EVERY candidate within the distance (see k_distance_tiles)
};

// ------------------------------------------------------------- single-pass ordered compaction
// Decoupled look-back over per-tile counts: a tile publishes (status | count) as ONE 64-bit word
// -- status 1: this tile's own count, status 2: inclusive prefix up to and including this tile --
// and then reads its predecessors' words, newest first, until it meets an inclusive prefix.  Tiles
// only ever wait for tiles with a smaller index (already resident or finished).  `status` points
// at the recording's first word; the words must be zero before the launch.
constexpr unsigned long long LB_VALUE_MASK = (1ull << 62) - 1ull;

// Called by ALL 32 lanes of one warp; returns the number of elements in tiles before `tile`.
// (A wider window -- 8 predecessors per lane -- was measured slower on both the 1 M and the 28.8 M
// sample streams: the nearest inclusive prefix is almost always within the first 32 tiles.)
__device__ __forceinline__ long long lookback_exclusive(unsigned long long* status, long long tile, long long my_count) {
  const int lane = threadIdx.x & 31;
  volatile unsigned long long* st = status;
  if (tile == 0) {
    if (lane == 0) st[0] = (2ull << 62) | static_cast<unsigned long long>(my_count);
    return 0;
  }
  if (lane == 0) st[tile] = (1ull << 62) | static_cast<unsigned long long>(my_count);
  long long excl = 0;
  long long base = tile - 1;
  while (true) {
    const long long t = base - lane;
    unsigned long long w = (2ull << 62);                        // before the first tile: inclusive prefix 0
    if (t >= 0) {
      do { w = st[t]; } while ((w >> 62) == 0ull);
    }
    const unsigned incl = __ballot_sync(0xffffffffu, (w >> 62) == 2ull);
    long long v = static_cast<long long>(w & LB_VALUE_MASK);
    if (incl) {
      const int first = __ffs(incl) - 1;                        // nearest tile holding an inclusive prefix
      if (lane > first) v = 0;
#pragma unroll
      for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      excl += v;
      break;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    excl += v;
    base -= 32;
  }
  if (lane == 0) st[tile] = (2ull << 62) | static_cast<unsigned long long>(excl + my_count);
  return excl;
}

// x' = sign * x with sign in {+1, -1} (exact)
__device__ __forceinline__ double signed_val(double v, int sign) { return sign < 0 ? -v : v; }

}  // namespace bpm
