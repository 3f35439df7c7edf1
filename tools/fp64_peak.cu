// Micro-benchmark: sustained DFMA rate per SM on this GPU (independent chains, no memory).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/fp64_peak tools/fp64_peak.cu && /tmp/fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_dfma(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = x[i] * a + b;
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount, threads = 1024, iters = 20000;
  constexpr int ILP = 8;
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * threads * 2);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k_dfma<ILP><<<sms * 2, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fma = double(sms) * 2 * threads * iters * ILP;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("DFMA: %.2f T FMA/s = %.2f TFLOP/s; per SM per clock (at %d MHz): %.1f\n", fma / ms * 1e-9, 2 * fma / ms * 1e-9,
           clk / 1000, fma / (ms * 1e-3) / sms / (clk * 1e3));
  }
  return 0;
}
