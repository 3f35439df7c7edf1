// K3: np.quantile(x, q), method 'linear' (bpm_analysis.py:225, :1067, :1075, :1114;
// numpy/lib/_function_base_impl.py:126-129, 4657-4678).
//
// Exact order statistics by most-significant-digit radix select on the order-preserving
// 64-bit image of each float64 (six passes: 11,11,11,11,11,9 bits), then the next larger
// element, then numpy's two-branch lerp evaluated with unfused float64 operations so the
// result is bit-identical to numpy's.  All items of the batch advance together; a pass
// whose `cond` says the quantile is not needed for an item skips that item.
#include "common.cuh"

namespace bpm {

constexpr int SEL_PASSES = 6;
constexpr int SEL_BINS = 2048;
constexpr int SEL_THREADS = 256;
constexpr int SEL_PER_THREAD = 16;
constexpr int SEL_TILE = SEL_THREADS * SEL_PER_THREAD;

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

// state after pass p-1 from the state after pass p-2 and hist[p-1]; every block computes it
// redundantly, block 0 of the item stores it.
__device__ void sel_advance(const SelState* __restrict__ prev, const unsigned int* __restrict__ hist, int bins,
                            SelState* out_shared, int* scratch /* >= 34 ints */) {
  // block-wide: find bucket b with cum[b] <= rank < cum[b+1]
  __shared__ long long s_cum[SEL_THREADS + 1];
  const int per = (bins + SEL_THREADS - 1) / SEL_THREADS;
  const int b0 = threadIdx.x * per;
  long long local = 0;
  for (int b = b0; b < min(b0 + per, bins); ++b) local += hist[b];
  s_cum[threadIdx.x + 1] = local;
  if (threadIdx.x == 0) s_cum[0] = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int t = 1; t <= SEL_THREADS; ++t) s_cum[t] += s_cum[t - 1];
  }
  __syncthreads();
  const long long rank = prev->rank;
  const long long lo = s_cum[threadIdx.x], hi = s_cum[threadIdx.x + 1];
  if (rank >= lo && rank < hi) {
    long long c = lo;
    for (int b = b0; b < min(b0 + per, bins); ++b) {
      const long long h = hist[b];
      if (rank < c + h) {
        out_shared->prefix = (prev->prefix << (bins == 512 ? 9 : 11)) | static_cast<unsigned long long>(b);
        out_shared->rank = rank - c;
        out_shared->below = prev->below + c;
        out_shared->count = h;
        break;
      }
      c += h;
    }
    out_shared->k = prev->k;
    out_shared->gamma = prev->gamma;
    out_shared->active = prev->active;
    out_shared->next_key = ~0ull;
  }
  __syncthreads();
  (void)scratch;
}

__global__ void k_select_init(const BpmItem* __restrict__ items, int n_items, double q,
                              const int* __restrict__ cond, SelState* __restrict__ st) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_items) return;
  const long long n = items[i].m;
  SelState s;
  const double v = __dmul_rn(static_cast<double>(n - 1), q);   // (n - 1) * q
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
  s.active = (n > 0 && (cond == nullptr || cond[i] != 0)) ? 1 : 0;
  st[i] = s;
}

// pass p: (p > 0) resolve pass p-1, then histogram digit p of the elements under the prefix
__global__ void __launch_bounds__(SEL_THREADS) k_select_pass(const double* __restrict__ x,
                                                             const BpmItem* __restrict__ items, int p,
                                                             SelState* __restrict__ states /* [SEL_PASSES+1][n_items] */,
                                                             unsigned int* __restrict__ hist /* [n_items][SEL_PASSES][SEL_BINS] */,
                                                             int n_items) {
  __shared__ unsigned int s_hist[SEL_BINS];
  __shared__ SelState s_cur;
  const int item = blockIdx.y;
  const BpmItem it = items[item];
  const SelState* prev = states + static_cast<size_t>(p) * n_items + item;        // state before pass p-1 ... see below
  // states[0] = initial; states[p] = after resolving pass p-1
  if (prev[0].active == 0 && p == 0) return;
  if (p > 0) {
    const SelState* before = states + static_cast<size_t>(p - 1) * n_items + item;
    if (before->active == 0) return;
    sel_advance(before, hist + (static_cast<size_t>(item) * SEL_PASSES + (p - 1)) * SEL_BINS,
                1 << sel_bits(p - 1), &s_cur, nullptr);
    if (blockIdx.x == 0 && threadIdx.x == 0) states[static_cast<size_t>(p) * n_items + item] = s_cur;
  } else {
    if (threadIdx.x == 0) s_cur = prev[0];
    __syncthreads();
  }
  const int64_t i0 = static_cast<int64_t>(blockIdx.x) * SEL_TILE;
  if (i0 >= it.m) return;
  for (int t = threadIdx.x; t < SEL_BINS; t += SEL_THREADS) s_hist[t] = 0;
  __syncthreads();
  const unsigned long long prefix = s_cur.prefix;
  const int sh = sel_shift(p), bits = sel_bits(p);
  const int up = sh + bits;                       // bits above this digit
  const unsigned int mask = (1u << bits) - 1u;
  const double* __restrict__ xi = x + it.m_off;
#pragma unroll 4
  for (int k = 0; k < SEL_PER_THREAD; ++k) {
    const int64_t i = i0 + k * SEL_THREADS + threadIdx.x;
    if (i < it.m) {
      const unsigned long long key = f64_key(xi[i]);
      const bool match = (up >= 64) ? true : ((key >> up) == prefix);
      if (match) atomicAdd(&s_hist[static_cast<unsigned int>(key >> sh) & mask], 1u);
    }
  }
  __syncthreads();
  unsigned int* gh = hist + (static_cast<size_t>(item) * SEL_PASSES + p) * SEL_BINS;
  for (int t = threadIdx.x; t < (1 << bits); t += SEL_THREADS) {
    const unsigned int c = s_hist[t];
    if (c) atomicAdd(gh + t, c);
  }
}

// resolve the last pass, then find the smallest key above the selected one
__global__ void __launch_bounds__(SEL_THREADS) k_select_next(const double* __restrict__ x,
                                                             const BpmItem* __restrict__ items,
                                                             SelState* __restrict__ states,
                                                             const unsigned int* __restrict__ hist, int n_items) {
  __shared__ SelState s_cur;
  __shared__ unsigned long long s_min[SEL_THREADS / 32];
  const int item = blockIdx.y;
  const BpmItem it = items[item];
  const SelState* before = states + static_cast<size_t>(SEL_PASSES - 1) * n_items + item;
  if (before->active == 0) return;
  sel_advance(before, hist + (static_cast<size_t>(item) * SEL_PASSES + (SEL_PASSES - 1)) * SEL_BINS,
              1 << sel_bits(SEL_PASSES - 1), &s_cur, nullptr);
  SelState* fin = states + static_cast<size_t>(SEL_PASSES) * n_items + item;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    // next_key is accumulated with atomicMin by all blocks; it was preset to ~0 by the memset-free init below
    fin->prefix = s_cur.prefix; fin->rank = s_cur.rank; fin->below = s_cur.below; fin->count = s_cur.count;
    fin->k = s_cur.k; fin->gamma = s_cur.gamma; fin->active = 1;
  }
  const int64_t i0 = static_cast<int64_t>(blockIdx.x) * SEL_TILE;
  if (i0 >= it.m) return;
  const unsigned long long sel = s_cur.prefix;
  unsigned long long best = ~0ull;
  const double* __restrict__ xi = x + it.m_off;
  for (int k = 0; k < SEL_PER_THREAD; ++k) {
    const int64_t i = i0 + k * SEL_THREADS + threadIdx.x;
    if (i < it.m) {
      const unsigned long long key = f64_key(xi[i]);
      if (key > sel && key < best) best = key;
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
    for (int w = 1; w < SEL_THREADS / 32; ++w) best = s_min[w] < best ? s_min[w] : best;
    if (best != ~0ull) atomicMin(&fin->next_key, best);
  }
}

__global__ void k_select_finish(const SelState* __restrict__ fin, int n_items, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_items) return;
  const SelState s = fin[i];
  if (!s.active) return;
  const double a = key_f64(s.prefix);
  // element k+1: the same value when duplicates cover it, else the next larger element
  double b = a;
  if (s.below + s.count <= s.k + 1 && s.next_key != ~0ull) b = key_f64(s.next_key);
  const double t = s.gamma;
  const double diff = __dsub_rn(b, a);
  double r = __dadd_rn(a, __dmul_rn(diff, t));
  if (t >= 0.5) r = __dsub_rn(b, __dmul_rn(diff, __dsub_rn(1.0, t)));
  out[i] = r;
}

__global__ void k_select_preset(SelState* __restrict__ fin, int n_items) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_items) { fin[i].next_key = ~0ull; fin[i].active = 0; }
}

struct SelectBuffers {
  SelState* states;
  unsigned int* hist;
};

static int carve_select(Workspace& ws, int n_items, SelectBuffers* b) {
  b->states = ws.take<SelState>(static_cast<size_t>(SEL_PASSES + 1) * n_items);
  b->hist = ws.take<unsigned int>(static_cast<size_t>(n_items) * SEL_PASSES * SEL_BINS);
  return ws.overflow ? BPM_ERR_WORKSPACE : BPM_OK;
}

size_t quantile_workspace_bytes(int n_items) {
  Workspace ws(nullptr, 0);
  SelectBuffers b;
  carve_select(ws, n_items, &b);
  return ws.used;
}

// out[i] = np.quantile(x_i, q) for every item with cond[i] != 0 (cond may be null: all items);
// out[i] is left untouched for skipped items.
int quantile_run(const double* x, const BpmItem* items, const BatchShape& sh, double q, const int* cond,
                 double* out, Workspace& ws, cudaStream_t st) {
  if (!x || !items || !out || sh.n_items <= 0 || !(q >= 0.0 && q <= 1.0)) return BPM_ERR_ARG;
  SelectBuffers b;
  BPM_TRY(carve_select(ws, sh.n_items, &b));
  const int n = sh.n_items;
  if (cudaMemsetAsync(b.hist, 0, sizeof(unsigned int) * static_cast<size_t>(n) * SEL_PASSES * SEL_BINS, st) != cudaSuccess)
    return BPM_ERR_CUDA;
  BPM_KERNEL(k_select_init);
  k_select_init<<<cdiv(n, 128), 128, 0, st>>>(items, n, q, cond, b.states);
  BPM_LAUNCH_OK();
  BPM_KERNEL(k_select_preset);
  k_select_preset<<<cdiv(n, 128), 128, 0, st>>>(b.states + static_cast<size_t>(SEL_PASSES) * n, n);
  BPM_LAUNCH_OK();
  const dim3 grid(cdiv(sh.max_m, SEL_TILE), n);
  for (int p = 0; p < SEL_PASSES; ++p) {
    BPM_KERNEL(k_select_pass);
    k_select_pass<<<grid, SEL_THREADS, 0, st>>>(x, items, p, b.states, b.hist, n);
    BPM_LAUNCH_OK();
  }
  BPM_KERNEL(k_select_next);
  k_select_next<<<grid, SEL_THREADS, 0, st>>>(x, items, b.states, b.hist, n);
  BPM_LAUNCH_OK();
  BPM_KERNEL(k_select_finish);
  k_select_finish<<<cdiv(n, 128), 128, 0, st>>>(b.states + static_cast<size_t>(SEL_PASSES) * n, n, out);
  BPM_LAUNCH_OK();
  return BPM_OK;
}

}  // namespace bpm
