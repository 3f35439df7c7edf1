// K0 / K1 input side: the per-kept-sample contractions of the blocked band-pass (design.py) and
// the stand-alone decimation x[::ds] (bpm_analysis.py:1016, :1033).
//
//   k_contract_*   per kept sample j:  uf[j] = sum_l wf[l] x[E_j+l]   (4-vector)
//                                      ub0[j] = sum_l q[l] x[E_j+l]   (4-vector)
//                  HBM-bound when block == ds (filter at the original rate).
//   k_gather_frames  K0 alone as float64 (zero-copy ingest from mapped pinned host memory).
#include "filter_common.cuh"

namespace bpm {

// ------------------------------------------------------------ contraction (generic)
// Any dtype / channel count / stride / block.  One thread per kept sample.
__global__ void __launch_bounds__(256) k_contract_generic(PcmView pcm, const BpmItem* __restrict__ items,
                                                          int64_t stride, const double* __restrict__ design,
                                                          double* __restrict__ uf, double* __restrict__ ub0,
                                                          double* __restrict__ xe) {
  const BpmItem it = items[blockIdx.y];
  const int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= it.m) return;
  const DesignView d{design};
  const int blk = d.block();
  const ExtSignal x = make_ext(pcm, it, stride);
  const int64_t E = PADLEN + j * blk;
  const double x0 = x.at(E);
  xe[it.m_off + j] = x0;
  if (j >= it.m - 1) return;
  const double* __restrict__ wf = d.wf();
  const double* __restrict__ q = d.q();
  double f[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
  double v = x0;
  for (int l = 0; l < blk; ++l) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      f[c] += wf[4 * l + c] * v;
      b[c] += q[4 * l + c] * v;
    }
    v = x.at(E + l + 1);
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) b[c] += q[4 * blk + c] * v;
  double* pf = uf + 4 * (it.m_off + j);
  double* pb = ub0 + 4 * (it.m_off + j);
  reinterpret_cast<double2*>(pf)[0] = make_double2(f[0], f[1]);
  reinterpret_cast<double2*>(pf)[1] = make_double2(f[2], f[3]);
  reinterpret_cast<double2*>(pb)[0] = make_double2(b[0], b[1]);
  reinterpret_cast<double2*>(pb)[1] = make_double2(b[2], b[3]);
}

// ------------------------------------------------------------ K0 alone: x[::stride] as float64
// np.mean(axis=1) + audio_data[::downsample_factor] (bpm_analysis.py:1016, :1033).  Lets a host
// pipeline overlap the PCIe-bound ingest of the NEXT recording (the PCM may be mapped pinned host
// memory: one 32-byte sector per kept frame crosses the bus) with the compute of the current one:
// few CTAs, eight independent loads in flight per thread, coalesced float64 stores.
constexpr int GF_UNROLL = 8;
__global__ void __launch_bounds__(256) k_gather_frames(PcmView pcm, const BpmItem* __restrict__ items, int64_t stride,
                                                       double* __restrict__ out) {
  const BpmItem it = items[blockIdx.y];
  const int64_t n_dec = (it.n_in + stride - 1) / stride;
  const int64_t T = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t j0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; j0 < n_dec; j0 += T * GF_UNROLL) {
    double v[GF_UNROLL];
#pragma unroll
    for (int k = 0; k < GF_UNROLL; ++k) {
      const int64_t j = j0 + k * T;
      v[k] = (j < n_dec) ? pcm_frame(pcm, it.in_off + j * stride) : 0.0;
    }
#pragma unroll
    for (int k = 0; k < GF_UNROLL; ++k) {
      const int64_t j = j0 + k * T;
      if (j < n_dec) out[it.m_off + j] = v[k];
    }
  }
}

// ------------------------------------------------------ contraction (int16, full rate)
// The HBM-bound kernel: mono int16 at the original rate (stride 1), block = ds.
//   * a CTA owns CT_BLOCKS consecutive blocks; one elected thread brings their PCM span into
//     shared memory with a single TMA bulk copy (cp.async.bulk + mbarrier), 16-byte aligned;
//   * each thread accumulates the 8 dot products of CT_J blocks; the weight row of sample l is
//     the same for every thread, so it lives in constant memory and reaches the FP64 pipe as a
//     uniform-register operand (LDCU + DFMA R,R,UR,R): no shared-memory traffic for weights;
//   * the first / last CTA of a recording (odd-extension samples) read global memory directly.
// 8 DFMA per input sample: at 64 FP64 lanes per SM that is about the time HBM needs to
// deliver the 2 bytes, so the kernel sits where the FP64 and HBM rooflines meet.
constexpr int CW_MAX_BLOCK = 767;              // weight image limit: 8*(block+1) doubles of constant memory

__constant__ double c_contract_w[8 * (CW_MAX_BLOCK + 1)];

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// int16 -> float64 without the (slow) I2F.F64 conversion: 2^52 + 2^31 + x assembled from bits,
// then one DADD
__device__ __forceinline__ double i16_to_f64(int x) {
  return __hiloint2double(0x43300000, x ^ 0x80000000) - 4503601774854144.0;
}

__device__ __forceinline__ void mbar_init(uint32_t bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_TMA:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_TMA;\n"
      "bra WAIT_TMA;\n"
      "DONE_TMA:\n"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
// one elected thread: arm the barrier with the byte count and start the bulk copy global -> shared
__device__ __forceinline__ void tma_load_1d(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// Persistent: a CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... of its recording with a
// two-stage TMA pipeline (the copy of tile k+2 is issued as soon as tile k's buffer is free), so
// the FP64 pipe never waits for HBM after the first tile.
template <int CT_THREADS, int CT_J>
__global__ void __launch_bounds__(CT_THREADS) k_contract_i16(const int16_t* __restrict__ pcm,
                                                             const BpmItem* __restrict__ items,
                                                             const double* __restrict__ design,
                                                             double* __restrict__ uf, double* __restrict__ ub0,
                                                             double* __restrict__ xe, int stage_bytes) {
  constexpr int CT_BLOCKS = CT_THREADS * CT_J;             // kept samples per tile
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long s_bar[2];
  const BpmItem it = items[blockIdx.y];
  const DesignView d{design};
  const int blk = d.block();
  const int64_t n = it.n_in;                               // stride 1: n_dec == n_in
  const int64_t num_tiles = (it.m + CT_BLOCKS - 1) / CT_BLOCKS;
  const int16_t* __restrict__ src = pcm + it.in_off;
  const int tid = threadIdx.x;
  const uint32_t bar0 = smem_u32(&s_bar[0]), bar1 = smem_u32(&s_bar[1]);

  // geometry of a tile: first staged data index, samples, whether TMA can fetch it whole
  auto tile_geom = [&](int64_t tile, int64_t& i0, int& nb, int& span, bool& interior) {
    const int64_t j0 = tile * CT_BLOCKS;
    nb = static_cast<int>(min(static_cast<int64_t>(CT_BLOCKS), it.m - j0));
    span = nb * blk + 1;                                   // samples E_j0 .. E_j0 + nb*blk
    i0 = j0 * blk;
    interior = (i0 >= 8) && (i0 + span + 8 <= n);
  };
  auto issue = [&](int64_t tile, int stage) {              // thread 0 only
    int64_t i0; int nb, span; bool interior;
    tile_geom(tile, i0, nb, span, interior);
    if (!interior) return;
    const uintptr_t addr = reinterpret_cast<uintptr_t>(src + i0);
    const int head = static_cast<int>((addr & 15) >> 1);
    const uint32_t bytes = static_cast<uint32_t>(((head + span + 7) >> 3) << 4);
    tma_load_1d(smem_u32(smem_raw + static_cast<size_t>(stage) * stage_bytes), src + i0 - head, bytes,
                stage ? bar1 : bar0);
  };

  if (tid == 0) {
    mbar_init(bar0);
    mbar_init(bar1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (blockIdx.x < num_tiles) issue(blockIdx.x, 0);
    if (blockIdx.x + gridDim.x < num_tiles) issue(blockIdx.x + gridDim.x, 1);
  }
  __syncthreads();                                         // barriers initialised before anyone polls

  uint32_t uses0 = 0, uses1 = 0;                           // completed TMA uses per stage -> wait parity
  int k = 0;
  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++k) {
    const int stage = k & 1;
    int64_t i0; int nb, span; bool interior;
    tile_geom(tile, i0, nb, span, interior);
    const int64_t j0 = tile * CT_BLOCKS;
    double acc[CT_J][8];
#pragma unroll
    for (int jj = 0; jj < CT_J; ++jj)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[jj][c] = 0.0;
    double x0[CT_J];
    if (interior) {
      if (stage) { mbar_wait(bar1, uses1 & 1); ++uses1; } else { mbar_wait(bar0, uses0 & 1); ++uses0; }
      const uintptr_t addr = reinterpret_cast<uintptr_t>(src + i0);
      const int head = static_cast<int>((addr & 15) >> 1);
      const int16_t* __restrict__ xs =
          reinterpret_cast<const int16_t*>(smem_raw + static_cast<size_t>(stage) * stage_bytes) + head + tid * blk;
      const int jstride = CT_THREADS * blk;
      // blocks beyond nb (last tile only) read staged-but-unused or stale shared memory; their
      // results are discarded below, the reads stay inside the stage buffer
#pragma unroll
      for (int jj = 0; jj < CT_J; ++jj) x0[jj] = static_cast<double>(xs[jj * jstride]);
#pragma unroll 4
      for (int l = 0; l <= blk; ++l) {
        double v[CT_J];
#pragma unroll
        for (int jj = 0; jj < CT_J; ++jj) v[jj] = i16_to_f64(xs[jj * jstride + l]);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const double w = c_contract_w[8 * l + c];
#pragma unroll
          for (int jj = 0; jj < CT_J; ++jj) acc[jj][c] += w * v[jj];
        }
      }
    } else {
      PcmView pv{pcm, BPM_PCM_I16, 1};
      const ExtSignal x = make_ext(pv, it, 1);
#pragma unroll
      for (int jj = 0; jj < CT_J; ++jj) {
        const int jl = jj * CT_THREADS + tid;
        x0[jj] = 0.0;
        if (jl >= nb) continue;
        const int64_t E = PADLEN + (j0 + jl) * blk;
        x0[jj] = x.at(E);
        if (j0 + jl >= it.m - 1) continue;
        for (int l = 0; l <= blk; ++l) {
          const double v = x.at(E + l);
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[jj][c] += c_contract_w[8 * l + c] * v;
        }
      }
    }
    __syncthreads();                                       // everyone is done with this stage's buffer
    if (tid == 0 && tile + 2 * static_cast<int64_t>(gridDim.x) < num_tiles)
      issue(tile + 2 * static_cast<int64_t>(gridDim.x), stage);
#pragma unroll
    for (int jj = 0; jj < CT_J; ++jj) {
      const int jl = jj * CT_THREADS + tid;
      if (jl >= nb) continue;
      const int64_t j = j0 + jl;
      xe[it.m_off + j] = x0[jj];
      if (j >= it.m - 1) continue;
      double2* pf = reinterpret_cast<double2*>(uf + 4 * (it.m_off + j));
      double2* pb = reinterpret_cast<double2*>(ub0 + 4 * (it.m_off + j));
      pf[0] = make_double2(acc[jj][0], acc[jj][1]);
      pf[1] = make_double2(acc[jj][2], acc[jj][3]);
      pb[0] = make_double2(acc[jj][4], acc[jj][5]);
      pb[1] = make_double2(acc[jj][6], acc[jj][7]);
    }
  }
}

// ------------------------------------------------------------------ host side
int contract_run(PcmView pv, const BpmItem* items, int n_items, const BatchShape& sh, int64_t stride,
                 const double* design, int block, double* uf, double* ub0, double* xe, cudaStream_t st) {
  const void* pcm = pv.base;
  const int pcm_dtype = pv.dtype, channels = pv.channels;
  struct { double *uf, *ub0, *xe; } b{uf, ub0, xe};
  const bool fast = (pcm_dtype == BPM_PCM_I16 && channels == 1 && stride == 1 && block >= 8 &&
                     block <= CW_MAX_BLOCK && (reinterpret_cast<uintptr_t>(pcm) & 1) == 0);
  if (fast) {
    if (cudaMemcpyToSymbolAsync(c_contract_w, design + BPM_DESIGN_HEADER_WORDS + 4 * (2 * block + 1),
                                sizeof(double) * 8 * (block + 1), 0, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
      return BPM_ERR_CUDA;
    // tile shape: small CTAs for long blocks (more resident warps per SM measured fastest),
    // wide CTAs for short blocks (keeps each bulk copy above a few KB)
    // (tile shapes were measured on C2: 32 threads x 2 blocks per thread for long blocks, 128 x 2 for short)
    BPM_KERNEL(k_contract_i16);
    // PCM span of a CTA (threads * J blocks), rounded up to whole 16-byte words on both sides
#define BPM_LAUNCH_CONTRACT(T, J, CTAS_PER_SM)                                                           \
    do {                                                                                                 \
      const size_t stage = ((2 * (static_cast<size_t>(T) * J * block + 1 + 16) + 32) + 127) & ~size_t(127); \
      const int64_t tiles = (sh.max_m + T * J - 1) / (T * J);                                            \
      int64_t gx = tiles;                                                                                \
      size_t smem = stage;                    /* CTAS_PER_SM == 0: one tile per CTA, single buffer */    \
      if (CTAS_PER_SM > 0) {                                                                             \
        gx = (static_cast<int64_t>(148) * CTAS_PER_SM + n_items - 1) / n_items;                          \
        if (gx > tiles) gx = tiles;                                                                      \
        smem = 2 * stage;                                                                                \
      }                                                                                                  \
      if (gx < 1) gx = 1;                                                                                \
      cudaFuncSetAttribute(k_contract_i16<T, J>, cudaFuncAttributeMaxDynamicSharedMemorySize,            \
                           static_cast<int>(smem));                                                      \
      k_contract_i16<T, J><<<dim3(static_cast<unsigned>(gx), n_items), T, smem, st>>>(                   \
          static_cast<const int16_t*>(pcm), items, design, b.uf, b.ub0, b.xe, static_cast<int>(stage));  \
    } while (0)
    if (block >= 96) BPM_LAUNCH_CONTRACT(32, 2, 0);
    else BPM_LAUNCH_CONTRACT(128, 2, 0);
#undef BPM_LAUNCH_CONTRACT
  } else if (block > 1) {
    BPM_KERNEL(k_contract_generic);
    k_contract_generic<<<dim3(cdiv(sh.max_m, 256), n_items), 256, 0, st>>>(pv, items, stride, design, b.uf, b.ub0, b.xe);
  }
  BPM_LAUNCH_OK();
  return BPM_OK;
}

int gather_frames_run(const void* pcm, int pcm_dtype, int channels, const BpmItem* items, const BpmItem* items_host,
                      int n_items, int64_t stride, double* out, cudaStream_t st) {
  if (!pcm || !items || !items_host || !out || n_items <= 0 || stride < 1 || channels < 1 || pcm_dtype < 0 ||
      pcm_dtype > BPM_PCM_F64)
    return BPM_ERR_ARG;
  for (int i = 0; i < n_items; ++i)
    if (items_host[i].m != (items_host[i].n_in + stride - 1) / stride) return BPM_ERR_ARG;   // out is laid out by m_off
  PcmView pv{pcm, pcm_dtype, channels};
  int gx = (148 + n_items - 1) / n_items;
  BPM_KERNEL(k_gather_frames);
  k_gather_frames<<<dim3(gx, n_items), 256, 0, st>>>(pv, items, stride, out);
  BPM_LAUNCH_OK();
  return BPM_OK;
}

}  // namespace bpm
