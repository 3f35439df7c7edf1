// K0 + K1 + K2 (+K2b): pcm -> zero-phase band-pass at the kept samples -> envelope.
//
// Replaces, on the device, what preprocess_audio does with numpy/scipy/pandas
// (bpm_analysis.py:1015-1054): stereo mean, x[::ds], butter(2,'band') + filtfilt,
// abs, centred rolling mean.  The mathematics of the blocked forward-backward
// evaluation is in bpm_analysis_b200/design.py; a serial numpy model of exactly
// these kernels is oracle/kernel_models.py::blocked_filtfilt.
//
//   k_contract_*   per kept sample j:  uf[j] = sum_l wf[l] x[E_j+l]   (4-vector)
//                                      ub0[j] = sum_l q[l] x[E_j+l]   (4-vector)
//                  HBM-bound when block == ds (filter at the original rate).
//   k_filter_init  s_f[E_0] from the 15 padded samples and the steady state zi*x[0].
//   k_scan<FWD>    s_f[j+1] = Ad s_f[j] + uf[j]              (chunked scan, warp-shuffle carries)
//   k_filter_tail  serial forward over the last block + right padding, then backward
//                  down to E_{m-1}, seeded with zi*y_f[last] (scipy's second pass).
//   k_scan<BWD>    s_b[j] = Ad s_b[j+1] + P s_f[j] + ub0[j];  y[j] = C s_b[j] + D y_f[j]
//   k_envelope     |y| -> centred rolling mean (pandas FixedWindowIndexer semantics).
#include "filter_common.cuh"

namespace bpm {

std::atomic<int64_t> g_launches{0};
thread_local const char* g_cur_kernel = "?";
bool g_profiling = false;

// sosfilt.cu
bool sosfilt_fused_envelope_ok(int env_window);
int sosfilt_run(const void* pcm, int pcm_dtype, int channels, const BpmItem* items, const BatchShape& sh,
                int64_t stride, const double* design, const double* design_host, int env_window, double* filtered,
                double* envelope, double* absmax, Workspace& ws, cudaStream_t st);
size_t sosfilt_workspace_bytes(int64_t total_m, int n_items);
// contract.cu
int contract_run(PcmView pv, const BpmItem* items, int n_items, const BatchShape& sh, int64_t stride,
                 const double* design, const double* design_host, int block, double* uf, double* ub0, double* xe,
                 cudaStream_t st);

// ------------------------------------------------------------------ init / tail
// One warp per recording: the lanes fetch the (strided, possibly reflected) samples with
// independent loads into shared memory, then lane 0 runs the short serial recurrence from there.
constexpr int FE_WARPS = 4;

__global__ void __launch_bounds__(32 * FE_WARPS) k_filter_init(PcmView pcm, const BpmItem* __restrict__ items,
                                                              int n_items, int64_t stride,
                                                              const double* __restrict__ design,
                                                              double* __restrict__ s0) {
  __shared__ double s_x[FE_WARPS][PADLEN + 1];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * FE_WARPS + w;
  if (item >= n_items) return;
  const BpmItem it = items[item];
  const DesignView d{design};
  const ExtSignal x = make_ext(pcm, it, stride);
  if (lane < PADLEN) s_x[w][lane] = x.at(lane);
  __syncwarp();
  if (lane != 0) return;
  double s[4];
  const double x0 = s_x[w][0];
#pragma unroll
  for (int c = 0; c < 4; ++c) s[c] = d.zi()[c] * x0;
  for (int e = 0; e < PADLEN; ++e) df2t_step(d.sos(), s, s_x[w][e]);
#pragma unroll
  for (int c = 0; c < 4; ++c) s0[4 * item + c] = s[c];
}

__global__ void __launch_bounds__(32 * FE_WARPS) k_filter_tail(PcmView pcm, const BpmItem* __restrict__ items,
                                                              int n_items, int64_t stride,
                                                              const double* __restrict__ design,
                                                              const double* __restrict__ sf, int tail_cap,
                                                              double* __restrict__ sb_last) {
  extern __shared__ double s_tail[];                       // [FE_WARPS][tail_cap]
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * FE_WARPS + w;
  if (item >= n_items) return;
  const BpmItem it = items[item];
  const DesignView d{design};
  const ExtSignal x = make_ext(pcm, it, stride);
  const int blk = d.block();
  const int64_t le = x.n_dec + 2 * PADLEN;
  const int64_t e_last = PADLEN + (it.m - 1) * blk;
  const int lt = static_cast<int>(le - e_last);
  double* yf = s_tail + static_cast<size_t>(w) * tail_cap;
  for (int k = lane; k < lt; k += 32) yf[k] = x.at(e_last + k);
  __syncwarp();
  if (lane != 0) return;
  double s[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) s[c] = sf[4 * (it.m_off + it.m - 1) + c];
  for (int k = 0; k < lt; ++k) yf[k] = df2t_step(d.sos(), s, yf[k]);
  const double ylast = yf[lt - 1];
#pragma unroll
  for (int c = 0; c < 4; ++c) s[c] = d.zi()[c] * ylast;
  for (int k = lt - 1; k >= 1; --k) df2t_step(d.sos(), s, yf[k]);
#pragma unroll
  for (int c = 0; c < 4; ++c) sb_last[4 * item + c] = s[c];
}

// ------------------------------------------------------------------ low-rate scans
// One pass per direction.  A CTA scans its tile locally, PUBLISHES the tile aggregate (4 doubles
// + a flag), then looks back over the K preceding tiles' aggregates -- the recurrence contracts,
// so K = 1..4 tiles reach 1e-22 and older history is dropped -- and applies the resulting start
// state.  Tiles only ever wait for tiles with a smaller block index (already resident or done),
// so the spin cannot deadlock.  SRC == 1 is the reference's decimate-then-filter order (block
// == 1): the 4-vectors uf / ub0 are two multiplies per sample and are formed on the fly -- the
// forward pass gathers the strided PCM frames itself (device memory or mapped pinned host
// memory) and leaves them in xe for the backward pass; nothing else is staged in HBM.
struct ScanParams {
  const BpmItem* items;
  const double* design;
  const double* u;        // SRC 0 -- FWD: uf, BWD: ub0
  double* sf;             // FWD: out      BWD: in
  double* xe;             // kept input samples: read (SRC 0, BWD) or written by FWD (SRC 1)
  const double* s_init;   // FWD: s0[item] BWD: sb_last[item]
  double* agg;            // tile aggregates [tile_slot][4]
  int* flags;             // [tile_slot]: 1 once agg is visible
  double* y;              // BWD: filtered signal out
  unsigned long long* absmax_bits;  // BWD: max |y| per item (bit pattern of a non-negative double)
  PcmView pcm;            // SRC 1, FWD
  int64_t stride;
};

__device__ __forceinline__ int64_t tile_slot0(const BpmItem& it, int item) {
  return it.m_off / SCAN_TILE + item;
}

// Shared-memory staging of a warp's 256 samples: 4-vector input terms as double2 pairs, one pad
// slot per 16 so that both the coalesced fill (consecutive lanes -> consecutive slots) and the
// per-thread reads (lane stride 16 slots + 1) are bank-conflict free.
constexpr int SCAN_WARP_SAMPLES = 32 * SCAN_CHUNK;
constexpr int SCAN_BUF_SLOTS = 2 * SCAN_WARP_SAMPLES + (2 * SCAN_WARP_SAMPLES) / 16;
__device__ __forceinline__ int scan_slot(int e) { return e + (e >> 4); }

#ifdef BPM_DEBUG_COUNTERS
__device__ unsigned long long g_dbg_scan[16];
#define SC_TICK(i) do { if (threadIdx.x == 0) { long long t_ = clock64(); atomicAdd(&g_dbg_scan[(i) + 8 * DIR], (unsigned long long)(t_ - t_phase)); t_phase = t_; } } while (0)
#else
#define SC_TICK(i)
#endif

template <int DIR /*0 fwd, 1 bwd*/, int SRC /*0 staged uf/ub0, 1 block == 1 on the fly*/>
__global__ void __launch_bounds__(SCAN_THREADS, DIR == 0 ? 3 : 2) k_scan(ScanParams p) {
  __shared__ double sm_pow[9 * 16];      // Ad^(CHUNK*2^k), k = 0..8
  __shared__ double sm_Ad[16], sm_P[16], sm_C[4];
  __shared__ double sm_tot[SCAN_THREADS / 32][4];
  __shared__ double sm_pre[SCAN_THREADS / 32][4];

  const int item = blockIdx.y;
  const BpmItem it = p.items[item];
  const DesignView d{p.design};
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t ns = it.m - 1;
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * SCAN_TILE;
  const double Dd = d.D();

  if (blockIdx.x == 0 && tid == 0) {
    // the state that needs no step: s_f[0], or y[m-1] from s_b[m-1]
    const double* si = p.s_init + 4 * item;
    if (DIR == 0) {
#pragma unroll
      for (int c = 0; c < 4; ++c) p.sf[4 * it.m_off + c] = si[c];
    } else {
      const double* sfl = p.sf + 4 * (it.m_off + it.m - 1);
      double yf = Dd * p.xe[it.m_off + it.m - 1], yb = 0.0;
#pragma unroll
      for (int c = 0; c < 4; ++c) { yf += d.C()[c] * sfl[c]; yb += d.C()[c] * si[c]; }
      const double yy = yb + Dd * yf;
      p.y[it.m_off + it.m - 1] = yy;
      atomicMax(p.absmax_bits + item, static_cast<unsigned long long>(__double_as_longlong(fabs(yy))));
    }
  }
  if (r0 >= ns) return;
#ifdef BPM_DEBUG_COUNTERS
  long long t_phase = clock64();
#endif

  for (int t = tid; t < 9 * 16; t += SCAN_THREADS) sm_pow[t] = d.pw(0)[t];
  if (tid < 16) { sm_Ad[tid] = d.Ad()[tid]; sm_P[tid] = d.P()[tid]; }
  if (tid < 4) sm_C[tid] = d.C()[tid];
  __syncthreads();

  SC_TICK(0);
  const int64_t rbeg = r0 + static_cast<int64_t>(tid) * SCAN_CHUNK;
  extern __shared__ __align__(16) unsigned char scan_dyn[];
  double2* buf = reinterpret_cast<double2*>(scan_dyn) + warp * SCAN_BUF_SLOTS;          // this warp's input terms
  double* yfs = reinterpret_cast<double*>(scan_dyn + sizeof(double2) * SCAN_BUF_SLOTS * (SCAN_THREADS / 32)) +
                warp * SCAN_WARP_SAMPLES;                                             // BWD: forward-pass outputs y_f
  // The warp's samples are one contiguous ascending run of j: [j_lo, j_lo + 256).  Local sample sl
  // belongs to scan position r = rw + sl (FWD) or r = rw + 255 - sl (BWD).
  const int64_t rw = r0 + static_cast<int64_t>(warp) * SCAN_WARP_SAMPLES;
  const int64_t j_lo = (DIR == 0) ? rw : (it.m - 2 - rw - (SCAN_WARP_SAMPLES - 1));

  // ---- fill: coalesced loads, all of a thread's loads independent and in flight together
  if (SRC == 1 && DIR == 0) {
    // decimate-first forward pass: gather the strided PCM frames (device memory or mapped pinned
    // host memory), leave them in xe for the backward pass, u = wf[0] * x
    const ExtSignal x = make_ext(p.pcm, it, p.stride);
    const double* __restrict__ wf = d.wf();
    const double w0 = wf[0], w1 = wf[1], w2 = wf[2], w3 = wf[3];
    double xv[SCAN_CHUNK];
#pragma unroll
    for (int i = 0; i < SCAN_CHUNK; ++i) {
      const int64_t j = j_lo + i * 32 + lane;
      xv[i] = (j < ns) ? x.at(PADLEN + j) : 0.0;
    }
#pragma unroll
    for (int i = 0; i < SCAN_CHUNK; ++i) {
      const int sl = i * 32 + lane;
      const int64_t j = j_lo + sl;
      if (j < ns) {
        p.xe[it.m_off + j] = xv[i];
        buf[scan_slot(2 * sl)] = make_double2(w0 * xv[i], w1 * xv[i]);
        buf[scan_slot(2 * sl + 1)] = make_double2(w2 * xv[i], w3 * xv[i]);
        if (j == ns - 1) p.xe[it.m_off + it.m - 1] = x.at(PADLEN + it.m - 1);
      }
    }
  } else {
    const int h = lane & 1;                                     // which half of the 4-vector this lane carries
    double q0a = 0, q0b = 0, q1a = 0, q1b = 0;
    if (SRC == 1) { q0a = d.q()[2 * h]; q0b = d.q()[2 * h + 1]; q1a = d.q()[4 + 2 * h]; q1b = d.q()[4 + 2 * h + 1]; }
    const double2* gu = reinterpret_cast<const double2*>(p.u + 4 * (it.m_off + j_lo));
    const double2* gs = reinterpret_cast<const double2*>(p.sf + 4 * (it.m_off + j_lo));
    constexpr int FB = 4;                                       // loads in flight per array and thread
#pragma unroll 1
    for (int ib = 0; ib < 2 * SCAN_CHUNK; ib += FB) {
      double2 vu[FB], vs[FB];
      double xj[FB], xn[FB];
#pragma unroll
      for (int k = 0; k < FB; ++k) {
        const int e = (ib + k) * 32 + lane;
        const int64_t j = j_lo + (e >> 1);
        const bool ok = (j >= 0 && j < ns);
        vu[k] = make_double2(0.0, 0.0); vs[k] = vu[k]; xj[k] = 0.0; xn[k] = 0.0;
        if (ok) {
          if (SRC == 0) vu[k] = gu[e];
          if (DIR == 1) { vs[k] = gs[e]; xj[k] = p.xe[it.m_off + j]; }
          if (SRC == 1) xn[k] = p.xe[it.m_off + j + 1];
        }
      }
#pragma unroll
      for (int k = 0; k < FB; ++k) {
        const int e = (ib + k) * 32 + lane;
        const int64_t j = j_lo + (e >> 1);
        const bool ok = (j >= 0 && j < ns);
        double2 v = vu[k];
        if (SRC == 1) v = make_double2(q0a * xj[k] + q1a * xn[k], q0b * xj[k] + q1b * xn[k]);
        if (DIR == 1) {
          // the partner lane holds the other half of s_f[j]
          const double ox = __shfl_xor_sync(0xffffffffu, vs[k].x, 1), oy = __shfl_xor_sync(0xffffffffu, vs[k].y, 1);
          const double sv[4] = {h ? ox : vs[k].x, h ? oy : vs[k].y, h ? vs[k].x : ox, h ? vs[k].y : oy};
          const double* Pr = sm_P + 8 * h;                      // rows 2h, 2h+1 of P
          v.x += Pr[0] * sv[0] + Pr[1] * sv[1] + Pr[2] * sv[2] + Pr[3] * sv[3];
          v.y += Pr[4] * sv[0] + Pr[5] * sv[1] + Pr[6] * sv[2] + Pr[7] * sv[3];
          if (ok && h == 0)
            yfs[e >> 1] = sm_C[0] * sv[0] + sm_C[1] * sv[1] + sm_C[2] * sv[2] + sm_C[3] * sv[3] + Dd * xj[k];
        }
        if (ok) buf[scan_slot(e)] = v;
      }
    }
  }
  __syncwarp();
  SC_TICK(1);

  // ---- sweep 1: aggregate of this thread's chunk from a zero state
  double z[4] = {0, 0, 0, 0};
#pragma unroll
  for (int c = 0; c < SCAN_CHUNK; ++c) {
    const int64_t r = rbeg + c;
    if (r < ns) {
      const int sl = (DIR == 0) ? (lane * SCAN_CHUNK + c) : (SCAN_WARP_SAMPLES - 1 - (lane * SCAN_CHUNK + c));
      const double2 a = buf[scan_slot(2 * sl)], bb = buf[scan_slot(2 * sl + 1)];
      const double u[4] = {a.x, a.y, bb.x, bb.y};
      affine4(sm_Ad, z, u);
    }
  }

  // inclusive scan of chunk aggregates across the warp
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int o = 1 << k;
    double zo[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) zo[c] = shfl_up_f64(z[c], o);
    if (lane >= o) {
      double t4[4];
      matvec4(sm_pow + 16 * k, zo, t4);
#pragma unroll
      for (int c = 0; c < 4; ++c) z[c] += t4[c];
    }
  }
  if (lane == 31) {
#pragma unroll
    for (int c = 0; c < 4; ++c) sm_tot[warp][c] = z[c];
  }
  __syncthreads();
  SC_TICK(2);
  if (warp == 0) {
    const int64_t slot0 = tile_slot0(it, item);
    const int64_t b = blockIdx.x;
    double* ag = p.agg + 4 * slot0;
    if (lane == 0) {
      // tile aggregate (zero start): 8 serial steps with Ad^(32*CHUNK); publish it
      double a[4] = {0, 0, 0, 0};
      for (int w = 0; w < SCAN_THREADS / 32; ++w) affine4(sm_pow + 16 * 5, a, sm_tot[w]);
#pragma unroll
      for (int c = 0; c < 4; ++c) __stcg(ag + 4 * b + c, a[c]);
      __threadfence();
      atomicExch(p.flags + slot0 + b, 1);
    }
    SC_TICK(3);
    // look back over the preceding tiles' aggregates (contraction makes older ones vanish).
    // The lanes wait for / fetch 32 tiles at a time concurrently; lane 0 then folds them in
    // oldest first:  start <- Ad^tile start + agg[t].
    const int64_t K = d.lookback();
    const int64_t k0 = (b > K) ? b - K : 0;
    double start[4] = {0, 0, 0, 0};
    if (k0 == 0) {
      const double* si = p.s_init + 4 * item;
#pragma unroll
      for (int c = 0; c < 4; ++c) start[c] = si[c];
    }
    for (int64_t t0 = k0; t0 < b; t0 += 32) {
      const int64_t t = t0 + lane;
      double g[4] = {0, 0, 0, 0};
      if (t < b) {
        volatile int* f = p.flags + slot0 + t;
        while (*f == 0) { }
        __threadfence();
#pragma unroll
        for (int c = 0; c < 4; ++c) g[c] = __ldcg(ag + 4 * t + c);
      }
      const int cnt = static_cast<int>((b - t0) < 32 ? (b - t0) : 32);
      for (int l = 0; l < cnt; ++l) {
        double gl[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) gl[c] = shfl_f64(g[c], l);
        affine4(sm_pow + 8 * 16, start, gl);
      }
    }
    SC_TICK(4);
    if (lane == 0) {
      // state at the start of every warp's stretch
      for (int w = 0; w < SCAN_THREADS / 32; ++w) {
#pragma unroll
        for (int c = 0; c < 4; ++c) sm_pre[w][c] = start[c];
        affine4(sm_pow + 16 * 5, start, sm_tot[w]);
      }
    }
  }
  __syncthreads();

  SC_TICK(5);
  // state before this thread's chunk = Ad^(CHUNK*lane) * warp_start + (inclusive state of lane-1)
  double st[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const double prev = shfl_up_f64(z[c], 1);
    st[c] = (lane == 0) ? 0.0 : prev;
  }
  {
    double pre[4] = {sm_pre[warp][0], sm_pre[warp][1], sm_pre[warp][2], sm_pre[warp][3]};
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      if (lane & (1 << k)) {
        double t4[4];
        matvec4(sm_pow + 16 * k, pre, t4);
#pragma unroll
        for (int c = 0; c < 4; ++c) pre[c] = t4[c];
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) st[c] += pre[c];
  }

  // ---- sweep 2: apply (input terms come back out of shared memory)
  double amax = 0.0;
#pragma unroll
  for (int c = 0; c < SCAN_CHUNK; ++c) {
    const int64_t r = rbeg + c;
    if (r < ns) {
      const int sl = (DIR == 0) ? (lane * SCAN_CHUNK + c) : (SCAN_WARP_SAMPLES - 1 - (lane * SCAN_CHUNK + c));
      const double2 a = buf[scan_slot(2 * sl)], bb = buf[scan_slot(2 * sl + 1)];
      const double u[4] = {a.x, a.y, bb.x, bb.y};
      affine4(sm_Ad, st, u);
      if (DIR == 0) {
        double2* po = reinterpret_cast<double2*>(p.sf + 4 * (it.m_off + r + 1));
        po[0] = make_double2(st[0], st[1]);
        po[1] = make_double2(st[2], st[3]);
      } else {
        const int64_t j = it.m - 2 - r;
        const double yb = sm_C[0] * st[0] + sm_C[1] * st[1] + sm_C[2] * st[2] + sm_C[3] * st[3];
        const double yy = yb + Dd * yfs[sl];
        p.y[it.m_off + j] = yy;
        amax = fmax(amax, fabs(yy));
      }
    }
  }
  SC_TICK(6);
  if (DIR == 1) {
#pragma unroll
    for (int o = 16; o; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if (lane == 0) atomicMax(p.absmax_bits + item, static_cast<unsigned long long>(__double_as_longlong(amax)));
  }
}

// ------------------------------------------------------------------ envelope (K2)
// pandas rolling(window=w, min_periods=1, center=True).mean() of |y|
// (bpm_analysis.py:1052-1054): window [i+1+off-w, i+off] clipped, off=(w-1)//2.
constexpr int ENV_THREADS = 256;
constexpr int ENV_R = 4;             // consecutive outputs per thread (they share the staged window)
constexpr int ENV_TILE = ENV_THREADS * ENV_R;
constexpr int ENV_MAX_W = 16384;     // rate // 10 of an undecimated 96 kHz recording still fits (174 KB of smem)

// staged element e lives at e + e/4: the fill (consecutive e) and the per-thread reads (element
// 4 t + k, i.e. 5 doubles between neighbouring lanes) are both bank-conflict free
__device__ __forceinline__ int env_slot(int e) { return e + (e >> 2); }
static size_t env_smem_bytes(int w) { return sizeof(double) * (static_cast<size_t>(ENV_TILE + w - 1) * 5 / 4 + 2); }

// A thread forms ENV_R consecutive means from ONE pass over the w + ENV_R - 1 staged values they
// share (9 shared-memory loads per output instead of 33 at the default window: the per-output
// version was bound by shared-memory bandwidth, not HBM; eight outputs per thread with the means
// staged back through shared memory for coalesced stores measured SLOWER: 178 vs 152 us at C4).
// Every sum still adds its window in
// ascending index order from 0.0, and positions outside the recording are staged as +0.0
// (x + 0.0 == x for the non-negative partial sums), so the results are bit-identical to the
// one-output-per-thread evaluation.
__global__ void __launch_bounds__(ENV_THREADS) k_envelope(const double* __restrict__ y,
                                                          const BpmItem* __restrict__ items, int w,
                                                          double* __restrict__ env) {
  extern __shared__ double s_abs[];          // env_slot(ENV_TILE + w - 1) + 1
  const BpmItem it = items[blockIdx.y];
  const int64_t i0 = static_cast<int64_t>(blockIdx.x) * ENV_TILE;
  if (i0 >= it.m) return;
  const int off = (w - 1) / 2;
  const int left = w - 1 - off;
  const int64_t lo = i0 - left;
  const int n = ENV_TILE + w - 1;
  for (int t = threadIdx.x; t < n; t += ENV_THREADS) {
    const int64_t k = lo + t;
    s_abs[env_slot(t)] = (k >= 0 && k < it.m) ? fabs(y[it.m_off + k]) : 0.0;
  }
  __syncthreads();
  const int64_t ib = i0 + static_cast<int64_t>(threadIdx.x) * ENV_R;
  if (ib >= it.m) return;
  // output r (r = 0..3) sums the thread's elements k = r .. r + w - 1; element k is staged at
  // 5 * threadIdx.x + k + k / 4
  const double* __restrict__ e0 = s_abs + 5 * threadIdx.x;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int k = 0;
#pragma unroll
  for (; k < 3; ++k) {
    const double v = e0[k];
    if (k < w) s0 = __dadd_rn(s0, v);
    if (k >= 1 && k <= w) s1 = __dadd_rn(s1, v);
    if (k >= 2 && k <= w + 1) s2 = __dadd_rn(s2, v);
  }
#pragma unroll 4
  for (; k < w; ++k) {
    const double v = e0[k + (k >> 2)];
    s0 = __dadd_rn(s0, v); s1 = __dadd_rn(s1, v); s2 = __dadd_rn(s2, v); s3 = __dadd_rn(s3, v);
  }
  for (; k < w + 3; ++k) {
    const double v = e0[k + (k >> 2)];
    if (k < w) s0 = __dadd_rn(s0, v);
    if (k >= 1 && k <= w) s1 = __dadd_rn(s1, v);
    if (k >= 2 && k <= w + 1) s2 = __dadd_rn(s2, v);
    if (k >= 3) s3 = __dadd_rn(s3, v);
  }
  const double sums[ENV_R] = {s0, s1, s2, s3};
#pragma unroll
  for (int r = 0; r < ENV_R; ++r) {
    const int64_t i = ib + r;
    if (i < it.m) {
      const int64_t a = max(static_cast<int64_t>(0), i - left), b = min(it.m - 1, i + off);
      env[it.m_off + i] = __ddiv_rn(sums[r], static_cast<double>(b - a + 1));
    }
  }
}

// K2b: np.int16(y / max|y| * 32767)
__global__ void k_debug_wav(const double* __restrict__ y, const double* __restrict__ absmax,
                            const BpmItem* __restrict__ items, int16_t* __restrict__ out) {
  const BpmItem it = items[blockIdx.y];
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= it.m) return;
  const double v = __dmul_rn(__ddiv_rn(y[it.m_off + i], absmax[blockIdx.y]), 32767.0);
  out[it.m_off + i] = static_cast<int16_t>(static_cast<int>(v));
}

// ------------------------------------------------------------------ host side
struct FrontendBuffers {
  double *uf, *ub0, *xe, *sf, *agg, *s0, *sb_last, *tail;
  int* flags;
  int64_t tiles;
  int tail_cap;
};

static int carve_frontend(Workspace& ws, int64_t total_m, int n_items, int block, FrontendBuffers* b) {
  const int64_t tiles = total_m / SCAN_TILE + n_items + 1;
  b->uf = ws.take<double>(4 * total_m);
  b->ub0 = ws.take<double>(4 * total_m);
  b->xe = ws.take<double>(total_m);
  b->sf = ws.take<double>(4 * total_m);
  b->agg = ws.take<double>(2 * 4 * tiles);      // forward | backward
  b->flags = ws.take<int>(2 * tiles);
  b->tiles = tiles;
  b->s0 = ws.take<double>(4 * n_items);
  b->sb_last = ws.take<double>(4 * n_items);
  b->tail_cap = block + PADLEN + 1;
  b->tail = ws.take<double>(static_cast<size_t>(n_items) * b->tail_cap);
  return ws.overflow ? BPM_ERR_WORKSPACE : BPM_OK;
}

// worst-case block for workspace sizing (the tail scratch is the only part that depends on it)
constexpr int MAX_BLOCK = 4096;

size_t frontend_workspace_bytes(int64_t total_m, int n_items) {
  Workspace ws(nullptr, 0);
  FrontendBuffers b;
  carve_frontend(ws, total_m, n_items, MAX_BLOCK, &b);
  const size_t sos = sosfilt_workspace_bytes(total_m, n_items);   // block == 1 takes sosfilt.cu's (smaller) scratch
  return ws.used > sos ? ws.used : sos;
}

int frontend_run(const void* pcm, int pcm_dtype, int channels, const BpmItem* items,
                 const BpmItem* items_host, int n_items, int64_t stride, const double* design,
                 const double* design_host, int64_t design_words, int block, int env_window, double* filtered,
                 double* envelope, double* absmax, Workspace& ws, cudaStream_t st) {
  if (!pcm || !items || !items_host || !design || !envelope || !absmax) return BPM_ERR_ARG;
  if (!filtered && (block > 1 || !sosfilt_fused_envelope_ok(env_window))) return BPM_ERR_ARG;
  if (n_items <= 0 || stride < 1 || channels < 1 || pcm_dtype < 0 || pcm_dtype > BPM_PCM_F64) return BPM_ERR_ARG;
  if (block < 1 || block > MAX_BLOCK || design_words < BPM_DESIGN_HEADER_WORDS + 16 * block + 12 + BPM_DESIGN_LANE_WORDS) return BPM_ERR_ARG;
  if (env_window < 1 || env_window > ENV_MAX_W) return BPM_ERR_ARG;
  const BatchShape sh = batch_shape(items_host, n_items);
  for (int i = 0; i < n_items; ++i) {
    const int64_t n_dec = (items_host[i].n_in + stride - 1) / stride;
    if (n_dec <= PADLEN) return BPM_ERR_TOO_SHORT;           // scipy: len(x) must be > padlen
    if (items_host[i].m != (n_dec + block - 1) / block) return BPM_ERR_ARG;
  }
  if (cudaMemsetAsync(absmax, 0, sizeof(double) * n_items, st) != cudaSuccess) return BPM_ERR_CUDA;
  if (block == 1) {
    // the reference's decimate-then-filter order: forward + backward cascade scans over the extended
    // signal with the envelope fused into the backward epilogue (sosfilt.cu), two launches
    if (!design_host) return BPM_ERR_ARG;
    BPM_TRY(sosfilt_run(pcm, pcm_dtype, channels, items, sh, stride, design, design_host, env_window, filtered,
                        envelope, absmax, ws, st));
    if (!sosfilt_fused_envelope_ok(env_window)) {
      const size_t smem = env_smem_bytes(env_window);
      if (smem > 48 * 1024)
        cudaFuncSetAttribute(k_envelope, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      BPM_KERNEL(k_envelope);
      k_envelope<<<dim3(cdiv(sh.max_m, ENV_TILE), n_items), ENV_THREADS, smem, st>>>(filtered, items, env_window,
                                                                                     envelope);
      BPM_LAUNCH_OK();
    }
    return BPM_OK;
  }
  FrontendBuffers b;
  BPM_TRY(carve_frontend(ws, sh.total_m, n_items, block, &b));
  PcmView pv{pcm, pcm_dtype, channels};

  BPM_TRY(contract_run(pv, items, n_items, sh, stride, design, design_host, block, b.uf, b.ub0, b.xe, st));
  BPM_KERNEL(k_filter_init);
  k_filter_init<<<cdiv(n_items, FE_WARPS), 32 * FE_WARPS, 0, st>>>(pv, items, n_items, stride, design, b.s0);
  BPM_LAUNCH_OK();

  if (cudaMemsetAsync(b.flags, 0, sizeof(int) * 2 * b.tiles, st) != cudaSuccess) return BPM_ERR_CUDA;
  const dim3 sgrid(cdiv(sh.max_m > 1 ? sh.max_m - 1 : 1, SCAN_TILE), n_items);
  ScanParams sp{items, design, b.uf, b.sf, b.xe, b.s0, b.agg, b.flags, filtered,
                reinterpret_cast<unsigned long long*>(absmax), pv, stride};
  const size_t scan_smem_f = sizeof(double2) * SCAN_BUF_SLOTS * (SCAN_THREADS / 32);
  const size_t scan_smem_b = scan_smem_f + sizeof(double) * SCAN_TILE;
  BPM_KERNEL(k_scan);
  cudaFuncSetAttribute(k_scan<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(scan_smem_f));
  k_scan<0, 0><<<sgrid, SCAN_THREADS, scan_smem_f, st>>>(sp);
  BPM_LAUNCH_OK();
  {
    const size_t smem = sizeof(double) * FE_WARPS * b.tail_cap;
    cudaFuncSetAttribute(k_filter_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    BPM_KERNEL(k_filter_tail);
    k_filter_tail<<<cdiv(n_items, FE_WARPS), 32 * FE_WARPS, smem, st>>>(pv, items, n_items, stride, design, b.sf,
                                                                        b.tail_cap, b.sb_last);
  }
  BPM_LAUNCH_OK();
  sp.u = b.ub0;
  sp.s_init = b.sb_last;
  sp.agg = b.agg + 4 * b.tiles;
  sp.flags = b.flags + b.tiles;
  BPM_KERNEL(k_scan);
  cudaFuncSetAttribute(k_scan<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(scan_smem_b));
  k_scan<1, 0><<<sgrid, SCAN_THREADS, scan_smem_b, st>>>(sp);
  BPM_LAUNCH_OK();
  {
    const size_t smem = env_smem_bytes(env_window);
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(k_envelope, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    BPM_KERNEL(k_envelope);
    k_envelope<<<dim3(cdiv(sh.max_m, ENV_TILE), n_items), ENV_THREADS, smem, st>>>(filtered, items, env_window,
                                                                                   envelope);
  }
  BPM_LAUNCH_OK();
  return BPM_OK;
}

int debug_wav_run(const double* filtered, const double* absmax, const BpmItem* items,
                  const BpmItem* items_host, int n_items, int16_t* out, cudaStream_t st) {
  if (!filtered || !absmax || !items || !items_host || !out || n_items <= 0) return BPM_ERR_ARG;
  const BatchShape sh = batch_shape(items_host, n_items);
  BPM_KERNEL(k_debug_wav);
  k_debug_wav<<<dim3(cdiv(sh.max_m, 256), n_items), 256, 0, st>>>(filtered, absmax, items, out);
  BPM_LAUNCH_OK();
  return BPM_OK;
}

}  // namespace bpm

#ifdef BPM_DEBUG_COUNTERS
extern "C" int bpm_debug_counters_scan(unsigned long long* out_host, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out_host, bpm::g_dbg_scan, sizeof(unsigned long long) * 16);
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(bpm::g_dbg_scan, z, sizeof(z)); }
  return 0;
}
#endif
