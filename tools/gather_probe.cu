// Probe: what does a strided gather of one int16 frame in `stride` (the reference's x[::ds],
// bpm_analysis.py:1033) cost in DRAM traffic and time on B200, and does any load flavour or the
// L2 fetch-granularity limit change it?   built here as tools/libgather_probe.so (the .so travels to the GPU box), driven by tools/gather_probe.py
// Run plain for times; under `ncu --metrics dram__bytes_read.sum` for the bytes per variant.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k_gather(const int16_t* __restrict__ x, int64_t stride, int64_t m, double* __restrict__ out) {
  const int64_t T = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t j0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; j0 < m; j0 += 8 * T) {
    short v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int64_t j = j0 + k * T;
      v[k] = 0;
      if (j < m) {
        const int16_t* p = x + j * stride;
        if (MODE == 0) v[k] = *p;
        if (MODE == 1) asm volatile("ld.global.nc.L1::no_allocate.s16 %0, [%1];" : "=h"(v[k]) : "l"(p));
        if (MODE == 2) asm volatile("ld.global.cs.s16 %0, [%1];" : "=h"(v[k]) : "l"(p));
        if (MODE == 3) asm volatile("ld.global.cv.s16 %0, [%1];" : "=h"(v[k]) : "l"(p));
        if (MODE == 4) asm volatile("ld.global.L2::64B.s16 %0, [%1];" : "=h"(v[k]) : "l"(p));
        if (MODE == 5) asm volatile("ld.global.L1::evict_first.s16 %0, [%1];" : "=h"(v[k]) : "l"(p));
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int64_t j = j0 + k * T;
      if (j < m) out[j] = static_cast<double>(v[k]);
    }
  }
}

template <int MODE>
static void run(const char* name, const int16_t* x, int64_t stride, int64_t m, double* out, int grid) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) k_gather<MODE><<<grid, 256>>>(x, stride, m, out);
  cudaEventRecord(e0);
  for (int i = 0; i < 10; ++i) k_gather<MODE><<<grid, 256>>>(x, stride, m, out);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("  %-34s grid %5d  %8.2f us per launch  (%s)\n", name, grid, ms * 100.0, cudaGetErrorString(cudaGetLastError()));
}

extern "C" int gather_probe_main() {
  const int64_t configs[3][2] = {{172800000, 159}, {345600000, 12}, {26460000LL * 8, 146}};
  for (int g = 0; g < 3; ++g) {                       // fetch granularity: default, 32, 128
    if (g == 1) printf("cudaLimitMaxL2FetchGranularity=32 -> %d\n", (int)cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32));
    if (g == 2) printf("cudaLimitMaxL2FetchGranularity=128 -> %d\n", (int)cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 128));
    size_t lim = 0;
    cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity);
    printf("== L2 fetch granularity limit %zu\n", lim);
    for (auto& c : configs) {
      const int64_t n = c[0], stride = c[1], m = (n + stride - 1) / stride;
      int16_t* x; double* out;
      cudaMalloc(&x, n * 2); cudaMalloc(&out, m * 8);
      cudaMemset(x, 1, n * 2);
      printf(" N %lld stride %lld M %lld\n", (long long)n, (long long)stride, (long long)m);
      int grid = static_cast<int>((m + 2047) / 2048);
      if (grid > 148 * 8) grid = 148 * 8;
      run<0>("ld.global", x, stride, m, out, grid);
      if (g == 0) {
        run<1>("ld.global.nc.L1::no_allocate", x, stride, m, out, grid);
        run<2>("ld.global.cs", x, stride, m, out, grid);
        run<3>("ld.global.cv", x, stride, m, out, grid);
        run<4>("ld.global.L2::64B", x, stride, m, out, grid);
        run<5>("ld.global.L1::evict_first", x, stride, m, out, grid);
      }
      cudaFree(x); cudaFree(out);
    }
  }
  return 0;
}
