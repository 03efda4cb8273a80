// SFU-free GELU / GELU' for the bf16 epilogues (shared by the GEMM epilogue and the fused adapter kernels).
#pragma once
#include "common.cuh"

// exact-erf GELU for the bf16 epilogue without the SFU: odd minimax polynomial of erf on |z| <= 3.3 (|err| < 7e-5,
// two orders below bf16 resolution), clamped outside.  fp32 mode never runs this (SIMT kernel, erff).
__device__ __forceinline__ float erf_poly(float z) {
  z = fminf(fmaxf(z, -3.3f), 3.3f);
  const float u = z * z;
  float p = 2.0112330506e-08f;
  p = fmaf(p, u, -1.1352825549e-06f);
  p = fmaf(p, u, 2.7990254297e-05f);
  p = fmaf(p, u, -3.9921021419e-04f);
  p = fmaf(p, u, 3.6915675290e-03f);
  p = fmaf(p, u, -2.3606465426e-02f);
  p = fmaf(p, u, 1.0890402398e-01f);
  p = fmaf(p, u, -3.7390216815e-01f);
  p = fmaf(p, u, 1.1280026288e+00f);
  return z * p;
}
__device__ __forceinline__ float gelu_fast(float x) {
  const float h = 0.5f * x;
  return fmaf(h, erf_poly(x * 0.70710678118654752f), h);
}
__device__ __forceinline__ float dgelu_fast(float x) {
  const float cdf = fmaf(0.5f, erf_poly(x * 0.70710678118654752f), 0.5f);
  return fmaf(x * 0.39894228040143268f, __expf(-0.5f * x * x), cdf);
}

// the same on two values per instruction (sm_100 packed fp32 FMA/MUL): the epilogue of the GELU GEMMs is bound by
// instruction issue of the two epilogue warps each SM sub-partition has, and these halve its FMA count
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 erf_poly2(float2 x) {        // erf(x / sqrt 2)
  float2 z = __fmul2_rn(x, f2(0.70710678118654752f));
  z.x = fminf(fmaxf(z.x, -3.3f), 3.3f);
  z.y = fminf(fmaxf(z.y, -3.3f), 3.3f);
  const float2 u = __fmul2_rn(z, z);
  float2 p = f2(2.0112330506e-08f);
  p = __ffma2_rn(p, u, f2(-1.1352825549e-06f));
  p = __ffma2_rn(p, u, f2(2.7990254297e-05f));
  p = __ffma2_rn(p, u, f2(-3.9921021419e-04f));
  p = __ffma2_rn(p, u, f2(3.6915675290e-03f));
  p = __ffma2_rn(p, u, f2(-2.3606465426e-02f));
  p = __ffma2_rn(p, u, f2(1.0890402398e-01f));
  p = __ffma2_rn(p, u, f2(-3.7390216815e-01f));
  p = __ffma2_rn(p, u, f2(1.1280026288e+00f));
  return __fmul2_rn(z, p);
}
__device__ __forceinline__ float2 gelu_fast2(float2 x) {
  const float2 h = __fmul2_rn(x, f2(0.5f));
  return __ffma2_rn(h, erf_poly2(x), h);
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// g = gelu(x), d = gelu'(x) = cdf + x * pdf from one erf evaluation
__device__ __forceinline__ void gelu_dgelu2(float2 x, float2& g, float2& d) {
  const float2 cdf = __ffma2_rn(erf_poly2(x), f2(0.5f), f2(0.5f));
  g = __fmul2_rn(x, cdf);
  const float2 t = __fmul2_rn(__fmul2_rn(x, x), f2(-0.5f * 1.4426950408889634f));
  const float2 e = make_float2(ex2f(t.x), ex2f(t.y));
  d = __ffma2_rn(__fmul2_rn(x, f2(0.39894228040143268f)), e, cdf);
}
