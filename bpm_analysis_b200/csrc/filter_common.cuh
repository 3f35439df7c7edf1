// Device-side views shared by the filter kernels (filter.cu) and the full-rate contraction
// (contract.cu): the filter design image, PCM access in every wavfile dtype, scipy's odd extension.
#pragma once
#include "common.cuh"

namespace bpm {

// ------------------------------------------------------------------ design view
struct DesignView {
  const double* w;
  __device__ __forceinline__ int block() const { return static_cast<int>(w[0]); }
  __device__ __forceinline__ int lookback() const {
    double v = w[1];
    return v > 1.0e9 ? 1000000000 : static_cast<int>(v);
  }
  __device__ __forceinline__ double D() const { return w[2]; }
  __device__ __forceinline__ const double* sos() const { return w + 4; }
  __device__ __forceinline__ const double* zi() const { return w + 16; }
  __device__ __forceinline__ const double* C() const { return w + 20; }
  __device__ __forceinline__ const double* Ad() const { return w + 24; }
  __device__ __forceinline__ const double* P() const { return w + 40; }
  __device__ __forceinline__ const double* pw(int k) const { return w + 56 + 16 * k; }
  __device__ __forceinline__ const double* wf() const { return w + BPM_DESIGN_HEADER_WORDS; }
  __device__ __forceinline__ const double* q() const { return w + BPM_DESIGN_HEADER_WORDS + 4 * block(); }
};

// ------------------------------------------------------------------ pcm access
struct PcmView {
  const void* base;
  int dtype;
  int channels;
};

// one frame as float64; multi-channel frames are averaged the way np.mean(axis=1) does
// (bpm_analysis.py:1016): integer sums are exact, float32 accumulates in float32.
__device__ __forceinline__ double pcm_frame(const PcmView& p, int64_t f) {
  const int ch = p.channels;
  switch (p.dtype) {
    case BPM_PCM_I16: {
      const int16_t* b = static_cast<const int16_t*>(p.base);
      if (ch == 1) return static_cast<double>(b[f]);
      long long s = 0;
      for (int c = 0; c < ch; ++c) s += b[f * ch + c];
      return static_cast<double>(s) / static_cast<double>(ch);
    }
    case BPM_PCM_I32: {
      const int32_t* b = static_cast<const int32_t*>(p.base);
      if (ch == 1) return static_cast<double>(b[f]);
      long long s = 0;
      for (int c = 0; c < ch; ++c) s += b[f * ch + c];
      return static_cast<double>(s) / static_cast<double>(ch);
    }
    case BPM_PCM_U8: {
      const uint8_t* b = static_cast<const uint8_t*>(p.base);
      if (ch == 1) return static_cast<double>(b[f]);
      long long s = 0;
      for (int c = 0; c < ch; ++c) s += b[f * ch + c];
      return static_cast<double>(s) / static_cast<double>(ch);
    }
    case BPM_PCM_F32: {
      const float* b = static_cast<const float*>(p.base);
      if (ch == 1) return static_cast<double>(b[f]);
      float s = b[f * ch];
      for (int c = 1; c < ch; ++c) s = __fadd_rn(s, b[f * ch + c]);
      return static_cast<double>(__fdiv_rn(s, static_cast<float>(ch)));
    }
    default: {
      const double* b = static_cast<const double*>(p.base);
      if (ch == 1) return b[f];
      double s = b[f * ch];
      for (int c = 1; c < ch; ++c) s = __dadd_rn(s, b[f * ch + c]);
      return __ddiv_rn(s, static_cast<double>(ch));
    }
  }
}

// the filter's input with scipy's odd extension (Appendix A.1): e in [0, n_dec + 30)
struct ExtSignal {
  PcmView pcm;
  int64_t in_off, n_dec, stride;
  __device__ __forceinline__ double s(int64_t i) const { return pcm_frame(pcm, in_off + i * stride); }
  __device__ __forceinline__ double at(int64_t e) const {
    int64_t i = e - PADLEN;
    if (i < 0) return __dsub_rn(__dmul_rn(2.0, s(0)), s(-i));
    if (i >= n_dec) return __dsub_rn(__dmul_rn(2.0, s(n_dec - 1)), s(2 * (n_dec - 1) - i));
    return s(i);
  }
};

__device__ __forceinline__ ExtSignal make_ext(const PcmView& pcm, const BpmItem& it, int64_t stride) {
  ExtSignal x;
  x.pcm = pcm;
  x.in_off = it.in_off;
  x.stride = stride;
  x.n_dec = (it.n_in + stride - 1) / stride;
  return x;
}

// one sample through the two-section cascade, direct form II transposed
__device__ __forceinline__ double df2t_step(const double* __restrict__ sos, double s[4], double x) {
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const double b0 = sos[6 * k], b1 = sos[6 * k + 1], b2 = sos[6 * k + 2];
    const double a1 = sos[6 * k + 4], a2 = sos[6 * k + 5];
    const double y = b0 * x + s[2 * k];
    s[2 * k] = b1 * x - a1 * y + s[2 * k + 1];
    s[2 * k + 1] = b2 * x - a2 * y;
    x = y;
  }
  return x;
}

__device__ __forceinline__ void matvec4(const double* __restrict__ M, const double v[4], double out[4]) {
#pragma unroll
  for (int r = 0; r < 4; ++r)
    out[r] = M[4 * r] * v[0] + M[4 * r + 1] * v[1] + M[4 * r + 2] * v[2] + M[4 * r + 3] * v[3];
}
// v = M v + add
__device__ __forceinline__ void affine4(const double* __restrict__ M, double v[4], const double add[4]) {
  double t[4];
  matvec4(M, v, t);
#pragma unroll
  for (int r = 0; r < 4; ++r) v[r] = t[r] + add[r];
}

}  // namespace bpm
