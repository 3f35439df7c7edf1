// K0 + K1 + K2 in the reference's decimate-then-filter order (bpm_analysis.py:1031-1054):
//   x = audio[::ds];  y = filtfilt(b, a, x);  env = rolling_mean(|y|, w, center=True, min_periods=1)
// as TWO launches per batch.
//
// scipy's filtfilt (Appendix A.1 of SURVEY.md) is: odd-extend x by 15 samples on both sides, run the
// filter forward over the extended signal starting from the steady state zi * x_ext[0], run the same
// filter backward over that output starting from zi * y_f[last], drop the padding.  Both passes are
// the same linear recurrence  s' = A s + B u  (4 states: two biquads in direct form II transposed),
// so both are ONE kernel, k_sos_scan<DIR>, a chunked parallel scan over the EXTENDED signal:
//
//   * a thread owns SS_CHUNK consecutive samples.  Sweep 1 runs the cascade over them from a zero
//     state (the chunk's zero-state response z; the chunk as an affine map is s -> A^8 s + z);
//   * the z of a warp's 32 chunks are combined by an inclusive shuffle scan with the tabulated
//     powers A^(8 2^k); the eight warp totals by one more 3-step scan in warp 0;
//   * the CTA publishes its tile aggregate (+ flag) and looks back over the K preceding tiles'
//     aggregates -- the recurrence contracts, K = 1..5 tiles reach 1e-22 -- which gives the tile's
//     start state with no separate reduce launch;
//   * sweep 2 re-runs the cascade from the now known state and produces the outputs.
// The filter coefficients and power matrices travel as kernel parameters, i.e. they sit in the
// constant bank and reach DFMA as operands: no shared-memory or register traffic for them.
//
// Forward (DIR 0) gathers the strided PCM frames itself (any wavfile dtype, interleaved channels
// averaged) and writes y_f for the whole extended signal.  Backward (DIR 1) reads y_f and, in its
// epilogue, forms |y| and the centred rolling mean for the samples it owns: a backward tile scans
// SS_TILE samples but only the first SS_TILE - SS_HALO of them enter its aggregate; the remaining
// SS_HALO samples are scanned again by the next tile, so every tile holds all |y| its envelope
// windows need (w - 1 <= SS_HALO; wider windows take the separate k_envelope kernel of filter.cu).
#include <stdlib.h>
#include "filter_common.cuh"

namespace bpm {

constexpr int SS_CHUNK = 8;
constexpr int SS_THREADS = 256;
constexpr int SS_WARPS = SS_THREADS / 32;
constexpr int SS_TILE = SS_CHUNK * SS_THREADS;        // 2048 samples scanned per CTA
constexpr int SS_HALO = 64;
constexpr int SS_WARP_SAMPLES = 32 * SS_CHUNK;        // 256
constexpr int SS_WBUF = SS_WARP_SAMPLES + SS_WARP_SAMPLES / 8;   // padded: one slot per 8
constexpr int SS_BUF = SS_WARPS * SS_WBUF;            // == padded size of a whole tile

// element e of a (warp or tile) buffer: one pad slot per 8 doubles, so that lane-strided accesses
// (e = i*32 + lane) and chunk accesses (e = 8*lane + c) are both bank-conflict free
__device__ __forceinline__ int ss_pad(int e) { return e + (e >> 3); }

struct SosScanArgs {
  const BpmItem* items;
  double sos[12];          // two sections: b0 b1 b2 1 a1 a2
  double zi[4];            // steady state of the cascade for a unit step
  double pw[8][16];        // A^(8 2^k), k = 0..7
  double ppart[16];        // A^part: one published aggregate to the next
  const double* lane_pow;  // device: A^(8 l), l = 0..31 (design image), one 4x4 per lane
  int part;                // samples per aggregate (SS_TILE, or SS_TILE - SS_HALO with the fused envelope)
  int lookback;            // preceding aggregates that still matter
  int env_window;          // > 0: fused envelope (DIR 1)
  int64_t stride;          // DIR 0: PCM frames per kept sample
  PcmView pcm;             // DIR 0
  double* yf;              // forward output / backward input: item i at m_off + 30 i, length m + 30
  double* agg;             // [slot][4]
  int* flags;              // [slot]
  double* y;               // DIR 1: filtered signal (may be null)
  double* env;             // DIR 1: envelope (fused)
  unsigned long long* absmax_bits;   // DIR 1: max |y| per item
};

__device__ __forceinline__ int64_t ss_slot0(const BpmItem& it, int item, int part) {
  return it.m_off / part + 2 * static_cast<int64_t>(item);
}

// out = M v   (M in the constant bank)
__device__ __forceinline__ void ss_matvec(const double (&M)[16], const double v[4], double out[4]) {
#pragma unroll
  for (int r = 0; r < 4; ++r) out[r] = M[4 * r] * v[0] + M[4 * r + 1] * v[1] + M[4 * r + 2] * v[2] + M[4 * r + 3] * v[3];
}

// One int16 frame of a STRIDED gather.  A plain load makes L2 fetch 128 bytes of DRAM for the 2 bytes
// wanted; the 64-byte prefetch-size hint halves that (measured on B200, tools/gather_probe.cu: 139 MB
// -> 69.6 MB of DRAM reads and 25 -> 18.7 us for the 1.09 M frames of a 60-minute 48 kHz recording;
// cudaLimitMaxL2FetchGranularity = 32 changes nothing).
__device__ __forceinline__ double ss_load_i16_strided(const int16_t* p) {
  short v;
  asm volatile("ld.global.L2::64B.s16 %0, [%1];" : "=h"(v) : "l"(p));
  return static_cast<double>(v);
}

// one sample through the cascade (direct form II transposed), coefficients from the constant bank
__device__ __forceinline__ double ss_step(const double (&sos)[12], double s[4], double x) {
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const double y = sos[6 * k] * x + s[2 * k];
    s[2 * k] = sos[6 * k + 1] * x - sos[6 * k + 4] * y + s[2 * k + 1];
    s[2 * k + 1] = sos[6 * k + 2] * x - sos[6 * k + 5] * y;
    x = y;
  }
  return x;
}

// fill of an edge tile (odd extension) or of any tile of the other sample formats: one frame at a
// time through the accessor.  Out of line: it keeps the dtype dispatch and the reflection branches
// out of the scan kernel's instruction stream.
__device__ __noinline__ void ss_fill_generic(const PcmView pcm, const BpmItem it, int64_t stride, int64_t t0, int nvalid,
                                             double* wb) {
  const ExtSignal x = make_ext(pcm, it, stride);
  const int lane = threadIdx.x & 31;
  const int sw = (threadIdx.x >> 5) * SS_WARP_SAMPLES + lane;
#pragma unroll 1
  for (int i = 0; i < SS_CHUNK; ++i)
    wb[ss_pad(i * 32 + lane)] = (sw + i * 32 < nvalid) ? x.at(t0 + sw + i * 32) : 0.0;
}

__device__ __noinline__ double ss_first_input(const PcmView pcm, const BpmItem it, int64_t stride) {
  return make_ext(pcm, it, stride).at(0);
}

template <int DIR /*0 forward from PCM, 1 backward from y_f*/, int MONO16 /*DIR 0: mono int16 fast path*/,
          int OCC /*CTAs per SM the register budget is cut for*/>
__global__ void __launch_bounds__(SS_THREADS, OCC) k_sos_scan(const __grid_constant__ SosScanArgs a) {
  __shared__ double sm_buf[SS_BUF];            // warp-private transposition buffers == the padded tile
  __shared__ double sm_env[DIR == 1 ? SS_BUF + SS_HALO + SS_HALO / 8 + 2 : 1];   // |y| of the tile (+ zeros), then the outputs
  __shared__ double sm_tot[SS_WARPS][4];
  __shared__ double sm_pre[SS_WARPS][4];
  __shared__ double sm_part[4];

  const int item = blockIdx.y;
  const BpmItem it = a.items[item];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n_ext = it.m + 2 * PADLEN;
  // scan positions: forward r = extended index; backward r counts down from the last extended
  // sample and stops at the first real one (nothing is needed from the left padding)
  const int64_t n_scan = (DIR == 0) ? n_ext : it.m + PADLEN;
  const int64_t t0 = static_cast<int64_t>(blockIdx.x) * a.part;
  if (t0 >= n_scan) return;
  const int nvalid = static_cast<int>(min(static_cast<int64_t>(SS_TILE), n_scan - t0));   // scan positions of this tile
  double* __restrict__ yf = a.yf + it.m_off + 2 * PADLEN * static_cast<int64_t>(item);
  double* wb = sm_buf + warp * SS_WBUF;
  const double2* __restrict__ lane_rows = reinterpret_cast<const double2*>(a.lane_pow) + 8 * lane;

  // ---- fill: lane-strided loads (coalesced for y_f; independent gathers for the PCM), then the
  //      thread's own 8 consecutive samples come back out of the warp's buffer
  {
    const int sw = warp * SS_WARP_SAMPLES + lane;          // tile-local position of this lane's first load
    bool filled = false;
    if (DIR == 0) {
      // interior tiles of a mono int16 recording: plain strided loads, no reflection / dtype dispatch
      const bool interior = MONO16 && t0 >= PADLEN && t0 + SS_TILE <= PADLEN + it.m;
      if (!interior) {
        ss_fill_generic(a.pcm, it, a.stride, t0, nvalid, wb);
        filled = true;
      }
    }
    if (!filled) {
      double v[SS_CHUNK];
      if (DIR == 0) {
        const int16_t* __restrict__ p16 =
            static_cast<const int16_t*>(a.pcm.base) + it.in_off + (t0 - PADLEN + sw) * a.stride;
        const int64_t step = 32 * a.stride;
#pragma unroll
        for (int i = 0; i < SS_CHUNK; ++i) v[i] = (a.stride > 1) ? ss_load_i16_strided(p16 + i * step) : static_cast<double>(p16[i * step]);
      } else {
        const double* __restrict__ src = yf + (n_ext - 1 - t0 - sw);
#pragma unroll
        for (int i = 0; i < SS_CHUNK; ++i) v[i] = (sw + i * 32 < nvalid) ? src[-(i * 32)] : 0.0;
      }
#pragma unroll
      for (int i = 0; i < SS_CHUNK; ++i) wb[ss_pad(i * 32 + lane)] = v[i];
    }
  }
  __syncwarp();
  const double* __restrict__ xin = wb + ss_pad(lane * SS_CHUNK);      // this thread's 8 consecutive inputs (no pad inside)

  // ---- sweep 1: zero-state response of the chunk
  double z[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int c = 0; c < SS_CHUNK; ++c) ss_step(a.sos, z, xin[c]);

  // inclusive scan across the warp:  z <- A^(8 o) z(lane - o) + z
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int o = 1 << k;
    double zo[4], t4[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) zo[c] = shfl_up_f64(z[c], o);
    ss_matvec(a.pw[k], zo, t4);
    if (lane >= o) {
#pragma unroll
      for (int c = 0; c < 4; ++c) z[c] += t4[c];
    }
  }
  if (lane == 31) {
#pragma unroll
    for (int c = 0; c < 4; ++c) sm_tot[warp][c] = z[c];
  }
  const int part_chunks = a.part / SS_CHUNK;              // chunks inside the published aggregate
  if (a.part < SS_TILE && tid == part_chunks - 1) {
#pragma unroll
    for (int c = 0; c < 4; ++c) sm_part[c] = z[c];
  }
  __syncthreads();

  if (warp == 0) {
    // zero-start state at the END of every warp (lanes 0..7): 3-step scan with A^256, A^512, A^1024
    double T[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) T[c] = (lane < SS_WARPS) ? sm_tot[lane][c] : 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int o = 1 << k;
      double To[4], t4[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) To[c] = shfl_up_f64(T[c], o);
      ss_matvec(a.pw[5 + k], To, t4);
      if (lane >= o) {
#pragma unroll
        for (int c = 0; c < 4; ++c) T[c] += t4[c];
      }
    }
    // zero-start state at the START of warp `lane`
    double pre0[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const double up = shfl_up_f64(T[c], 1);
      pre0[c] = (lane == 0) ? 0.0 : up;
    }
    // the tile aggregate: state after `part` samples from a zero start
    double ag[4];
    if (a.part == SS_TILE) {
#pragma unroll
      for (int c = 0; c < 4; ++c) ag[c] = shfl_f64(T[c], SS_WARPS - 1);
    } else {
      const int wq = (part_chunks - 1) >> 5, nq = ((part_chunks - 1) & 31) + 1;   // boundary thread: warp, chunks into it
      double v[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) v[c] = shfl_f64(pre0[c], wq);
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        if ((nq >> k) & 1) {
          double t4[4];
          ss_matvec(a.pw[k], v, t4);
#pragma unroll
          for (int c = 0; c < 4; ++c) v[c] = t4[c];
        }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) ag[c] = v[c] + sm_part[c];
    }
    const int64_t slot0 = ss_slot0(it, item, a.part);
    const int64_t b = blockIdx.x;
    double* agp = a.agg + 4 * slot0;
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < 4; ++c) __stcg(agp + 4 * b + c, ag[c]);
      __threadfence();
      atomicExch(a.flags + slot0 + b, 1);
    }
    // look back over the preceding aggregates, oldest first:  start <- A^part start + agg[t]
    const int64_t K = a.lookback;
    const int64_t k0 = (b > K) ? b - K : 0;
    double start[4] = {0.0, 0.0, 0.0, 0.0};
    if (k0 == 0) {
      // the true initial state: zi * (first input of this pass)
      double u0;
      if (DIR == 0) u0 = ss_first_input(a.pcm, it, a.stride);
      else u0 = yf[n_ext - 1];
#pragma unroll
      for (int c = 0; c < 4; ++c) start[c] = a.zi[c] * u0;
    }
    for (int64_t tb = k0; tb < b; tb += 32) {
      const int64_t t = tb + lane;
      double g[4] = {0.0, 0.0, 0.0, 0.0};
      if (t < b) {
        volatile int* f = a.flags + slot0 + t;
        while (*f == 0) { }
        __threadfence();
#pragma unroll
        for (int c = 0; c < 4; ++c) g[c] = __ldcg(agp + 4 * t + c);
      }
      const int cnt = static_cast<int>((b - tb) < 32 ? (b - tb) : 32);
      for (int l = 0; l < cnt; ++l) {
        double gl[4], t4[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) gl[c] = shfl_f64(g[c], l);
        ss_matvec(a.ppart, start, t4);
#pragma unroll
        for (int c = 0; c < 4; ++c) start[c] = t4[c] + gl[c];
      }
    }
    // state at the start of warp `lane`:  A^(256 lane) start + pre0
    double sp[4] = {start[0], start[1], start[2], start[3]};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if ((lane >> k) & 1) {
        double t4[4];
        ss_matvec(a.pw[5 + k], sp, t4);
#pragma unroll
        for (int c = 0; c < 4; ++c) sp[c] = t4[c];
      }
    }
    if (lane < SS_WARPS) {
#pragma unroll
      for (int c = 0; c < 4; ++c) sm_pre[lane][c] = sp[c] + pre0[c];
    }
  }
  __syncthreads();

  // ---- state before this thread's chunk = A^(8 lane) warp_start + (inclusive state of lane - 1)
  double st[4];
  {
    const double pre[4] = {sm_pre[warp][0], sm_pre[warp][1], sm_pre[warp][2], sm_pre[warp][3]};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const double prev = shfl_up_f64(z[c], 1);
      const double2 m01 = __ldg(lane_rows + 2 * c), m23 = __ldg(lane_rows + 2 * c + 1);   // row c of A^(8 lane)
      st[c] = ((lane == 0) ? 0.0 : prev) + (m01.x * pre[0] + m01.y * pre[1] + m23.x * pre[2] + m23.y * pre[3]);
    }
  }

  // ---- sweep 2: the outputs
  double x[SS_CHUNK];
#pragma unroll
  for (int c = 0; c < SS_CHUNK; ++c) x[c] = ss_step(a.sos, st, xin[c]);

  if (DIR == 0) {
    // y_f of the whole extended signal, coalesced through the warp's buffer (in place of the inputs)
    double* xout = wb + ss_pad(lane * SS_CHUNK);
#pragma unroll
    for (int c = 0; c < SS_CHUNK; ++c) xout[c] = x[c];
    __syncwarp();
    const int sw = warp * SS_WARP_SAMPLES + lane;
    double* __restrict__ dst = yf + t0 + sw;
#pragma unroll
    for (int i = 0; i < SS_CHUNK; ++i)
      if (sw + i * 32 < nvalid) dst[i * 32] = wb[ss_pad(i * 32 + lane)];
    return;
  }

  // ---- backward epilogue.  Tile-local ASCENDING index la = SS_TILE - 1 - sl (sl = position in scan
  //      order); signal index j = jlo + la.  Samples outside the recording are staged as 0.
  const int64_t jhi = it.m + (PADLEN - 1) - t0;           // signal index of sl == 0
  const int64_t jlo = jhi - (SS_TILE - 1);
  // la range that lies inside the recording: [la_min, la_max]
  const int la_min = static_cast<int>(max(static_cast<int64_t>(0), -jlo));
  const int la_max = static_cast<int>(min(static_cast<int64_t>(SS_TILE - 1), it.m - 1 - jlo));
  __syncthreads();                                        // every warp has consumed its inputs: the buffers become the tile
#pragma unroll
  for (int c = 0; c < SS_CHUNK; ++c) {
    const int la = SS_TILE - 1 - (tid * SS_CHUNK + c);
    const double v = (la >= la_min && la <= la_max) ? x[c] : 0.0;
    sm_buf[ss_pad(la)] = v;
    sm_env[ss_pad(la)] = fabs(v);
  }
  if (tid < SS_HALO + SS_HALO / 8 + 1) sm_env[ss_pad(SS_TILE) + tid] = 0.0;    // windows of tile 0 may reach past the tile
  __syncthreads();
  if (a.y != nullptr) {
    // filtered signal + max |y| over the samples of this tile's partition (the debug WAV needs both)
    double amax = 0.0;
    const int la_lo = max(la_min, SS_TILE - a.part);
    for (int la = la_max - tid; la >= la_lo; la -= SS_THREADS) {
      const double yy = sm_buf[ss_pad(la)];
      a.y[it.m_off + jlo + la] = yy;
      amax = fmax(amax, fabs(yy));
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if (lane == 0 && amax > 0.0)
      atomicMax(a.absmax_bits + item, static_cast<unsigned long long>(__double_as_longlong(amax)));
  }
  if (a.env_window <= 0) return;
  // envelope: pandas window [i - left, i + off] clipped to the recording (bpm_analysis.py:1053-1054).
  // Tile t owns the outputs j in (jhi - off - part, jhi - off] (tile 0: up to m - 1); thread `tid`
  // forms the 8 consecutive ones starting at la0 = (part offset) + 8 tid with a sliding sum over the
  // staged |y| (zeros outside the recording, so clipped windows need no special case in the sum).
  const int w = a.env_window;
  const int off = (w - 1) / 2, left = w - 1 - off;
  // tile-local index of output k = 0 of thread 0:  first - jlo  with  first = jhi - off - part + 1
  const int q0 = SS_TILE - a.part - off;                  // >= left  because  w - 1 <= SS_TILE - part
  const double inv_w = 1.0 / static_cast<double>(w);
  // A thread forms its 8 consecutive outputs with a sliding sum that is re-started per thread: the
  // partial sums never mix magnitudes further apart than 38 samples.  (Window sums as differences of a
  // tile-wide prefix sum were tried and REJECTED: after a loud passage the differences in a quiet one
  // -- a dropout decaying to 1e-17 -- are rounding noise of the loud one; the envelope still matched to
  // 1e-11 of its maximum but the troughs found in the quiet stretch did not.)
  const int own_lo = max(la_min, q0);                     // owned outputs, tile-local, inclusive
  const int own_hi = (blockIdx.x == 0) ? la_max : min(la_max, SS_TILE - 1 - off);
  double e[SS_CHUNK];
  const int qb = q0 + tid * SS_CHUNK;                     // tile-local index of this thread's first output
  const bool any = (qb + SS_CHUNK - 1 >= own_lo) && (qb <= own_hi);
  if (any) {
    const double* __restrict__ av = sm_env;
    const int lb = qb - left;                             // first window element of output k = 0
    double sacc = 0.0;
    for (int q = 0; q < w; ++q) sacc += av[ss_pad(lb + q)];
#pragma unroll
    for (int k = 0; k < SS_CHUNK; ++k) {
      const int la = qb + k;
      const int wa = max(la_min, la - left), wb2 = min(la_max, la + off);
      const int cnt = wb2 - wa + 1;
      e[k] = (cnt == w) ? sacc * inv_w : __ddiv_rn(sacc, static_cast<double>(cnt));
      sacc = (sacc + av[ss_pad(lb + k + w)]) - av[ss_pad(lb + k)];
    }
  }
  __syncthreads();                                        // all windows are read: the sums can be overwritten
  if (any) {
#pragma unroll
    for (int k = 0; k < SS_CHUNK; ++k) sm_env[ss_pad(qb + k)] = e[k];
  }
  __syncthreads();
  {
    double* __restrict__ dst = a.env + it.m_off + jlo;
    for (int la = own_lo + tid; la <= own_hi; la += SS_THREADS) dst[la] = sm_env[ss_pad(la)];
  }
}

// ------------------------------------------------------------------ host side
static void matmul4(const long double* A, const long double* B, long double* C) {
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      long double s = 0.0L;
      for (int k = 0; k < 4; ++k) s += A[4 * r + k] * B[4 * k + c];
      C[4 * r + c] = s;
    }
}

struct SosBuffers {
  double* yf;
  double* agg;
  int* flags;
  int64_t slots;       // per direction
};

static int carve_sosfilt(Workspace& ws, int64_t total_m, int n_items, SosBuffers* b) {
  const int64_t slots = total_m / (SS_TILE - SS_HALO) + 2 * static_cast<int64_t>(n_items) + 2;
  b->yf = ws.take<double>(total_m + 2 * PADLEN * static_cast<int64_t>(n_items));
  b->agg = ws.take<double>(2 * 4 * slots);
  b->flags = ws.take<int>(2 * slots);
  b->slots = slots;
  return ws.overflow ? BPM_ERR_WORKSPACE : BPM_OK;
}

size_t sosfilt_workspace_bytes(int64_t total_m, int n_items) {
  Workspace ws(nullptr, 0);
  SosBuffers b;
  carve_sosfilt(ws, total_m, n_items, &b);
  return ws.used;
}

bool sosfilt_fused_envelope_ok(int env_window) { return env_window >= 1 && env_window - 1 <= SS_HALO; }

// a1 for block == 1.  design_host: the packed design image in HOST memory (kernel parameters).
// With a window the fused epilogue cannot take, `envelope` is left untouched and the caller runs
// k_envelope on `filtered` (which must then be non-null).
int sosfilt_run(const void* pcm, int pcm_dtype, int channels, const BpmItem* items, const BatchShape& sh,
                int64_t stride, const double* design, const double* design_host, int env_window, double* filtered,
                double* envelope, double* absmax, Workspace& ws, cudaStream_t st) {
  SosBuffers b;
  BPM_TRY(carve_sosfilt(ws, sh.total_m, sh.n_items, &b));
  const bool fused = sosfilt_fused_envelope_ok(env_window) && envelope != nullptr;
  if (!fused && filtered == nullptr) return BPM_ERR_ARG;
  if (cudaMemsetAsync(b.flags, 0, sizeof(int) * 2 * b.slots, st) != cudaSuccess) return BPM_ERR_CUDA;

  SosScanArgs a;
  memset(&a, 0, sizeof(a));
  a.items = items;
  const double* d = design_host;
  for (int i = 0; i < 12; ++i) a.sos[i] = d[4 + i];
  for (int i = 0; i < 4; ++i) a.zi[i] = d[16 + i];
  for (int k = 0; k < 8; ++k)
    for (int i = 0; i < 16; ++i) a.pw[k][i] = d[56 + 16 * k + i];
  const double look = d[1];
  const int K = look > 1.0e9 ? 1000000000 : static_cast<int>(look);
  a.lane_pow = design + BPM_DESIGN_HEADER_WORDS + 16 * 1 + 12;   // block == 1: after wf, q, wq8
  a.stride = stride;
  a.pcm = PcmView{pcm, pcm_dtype, channels};
  a.yf = b.yf;
  a.y = filtered;
  a.env = envelope;
  a.absmax_bits = reinterpret_cast<unsigned long long*>(absmax);

  // forward: aggregates one full tile apart (A^2048 = pw index 8 of the design image)
  a.part = SS_TILE;
  for (int i = 0; i < 16; ++i) a.ppart[i] = d[56 + 16 * 8 + i];
  a.lookback = K;
  a.env_window = 0;
  a.agg = b.agg;
  a.flags = b.flags;
  const int64_t max_ext = sh.max_m + 2 * PADLEN;
  const dim3 gf(cdiv(max_ext, SS_TILE), sh.n_items);
  BPM_KERNEL(k_sos_fwd);
  // register budget: 80 registers (3 CTAs per SM) or 64 (4 per SM, a few spills in the warp scan);
  // BPM_SOS_OCC selects for experiments, the default is what measured faster on B200
  static const int occ = [] { const char* e = getenv("BPM_SOS_OCC"); return (e && atoi(e) == 3) ? 3 : 4; }();
  const bool mono16 = (pcm_dtype == BPM_PCM_I16 && channels == 1);
  if (occ == 4) {
    if (mono16) k_sos_scan<0, 1, 4><<<gf, SS_THREADS, 0, st>>>(a);
    else k_sos_scan<0, 0, 4><<<gf, SS_THREADS, 0, st>>>(a);
  } else {
    if (mono16) k_sos_scan<0, 1, 3><<<gf, SS_THREADS, 0, st>>>(a);
    else k_sos_scan<0, 0, 3><<<gf, SS_THREADS, 0, st>>>(a);
  }
  BPM_LAUNCH_OK();

  // backward (+ envelope): aggregates SS_TILE - SS_HALO apart when the envelope is fused
  a.agg = b.agg + 4 * b.slots;
  a.flags = b.flags + b.slots;
  if (fused) {
    a.part = SS_TILE - SS_HALO;                       // 1984 = 1024 + 512 + 256 + 128 + 64
    long double P[16], Q[16], R[16];
    for (int i = 0; i < 16; ++i) P[i] = d[56 + 16 * 7 + i];
    for (int k = 6; k >= 3; --k) {
      for (int i = 0; i < 16; ++i) Q[i] = d[56 + 16 * k + i];
      matmul4(P, Q, R);
      for (int i = 0; i < 16; ++i) P[i] = R[i];
    }
    for (int i = 0; i < 16; ++i) a.ppart[i] = static_cast<double>(P[i]);
    a.lookback = K > 900000000 ? K : K + K / 16 + 1;  // the aggregates lie a little closer together
    a.env_window = env_window;
  }
  const dim3 gb(cdiv(sh.max_m + PADLEN, a.part), sh.n_items);
  BPM_KERNEL(k_sos_bwd);
  if (occ == 4) k_sos_scan<1, 0, 4><<<gb, SS_THREADS, 0, st>>>(a);
  else k_sos_scan<1, 0, 3><<<gb, SS_THREADS, 0, st>>>(a);
  BPM_LAUNCH_OK();
  return BPM_OK;
}

}  // namespace bpm
