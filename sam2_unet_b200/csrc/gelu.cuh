// SFU-light GELU / GELU' for the bf16 epilogues (shared by the GEMM epilogue and the fused adapter kernels).
#pragma once
#include "common.cuh"

// exact-erf GELU for the bf16 epilogue: the normal cdf as an odd polynomial,
//   Phi(x) - 1/2 = x * P(x^2)  on |x| <= 4  (7 coefficients, |err| <= 1.1e-4 = 1/18 of a bf16 half-ulp at 1),
// clamped outside (Phi(-4) = 3e-5).  The GEMM epilogues that evaluate it are bound by the FMA pipe of the two epilogue
// warps each SM sub-partition has (ncu source view: the polynomial region holds ~60 % of the samples, stall reason
// "wait" = pipe issue interval), so the form is chosen for the fewest FMA-pipe instructions per value: the 1/sqrt(2)
// and 1/2 of 0.5 * (1 + erf(x / sqrt 2)) are folded into the coefficients, the clamp runs on the ALU pipe, and the
// 1/sqrt(2 pi) of the density is folded into the exponent of the one ex2.  Per PAIR of values (packed fp32 x 2):
// GELU 2 FMUL2 + 7 FFMA2, GELU and GELU' together 2 FMUL2 + 9 FFMA2 + 2 MUFU (was 7 FMUL2 + 10 FFMA2 with the
// 9-coefficient erf polynomial on |z| <= 3.3).  fp32 mode never runs this (SIMT kernel, erff).
constexpr float GELU_CLAMP = 4.0f;
constexpr float GELU_C0 = 3.9852691573e-01f, GELU_C1 = -6.5388362295e-02f, GELU_C2 = 9.1129552496e-03f,
                GELU_C3 = -8.7898177956e-04f, GELU_C4 = 5.4190489554e-05f, GELU_C5 = -1.8918942160e-06f,
                GELU_C6 = 2.8161092398e-08f;

__device__ __forceinline__ float ncdf_fast(float x) {           // Phi(x)
  const float xc = fminf(fmaxf(x, -GELU_CLAMP), GELU_CLAMP);
  const float u = xc * xc;
  float p = fmaf(GELU_C6, u, GELU_C5);
  p = fmaf(p, u, GELU_C4);
  p = fmaf(p, u, GELU_C3);
  p = fmaf(p, u, GELU_C2);
  p = fmaf(p, u, GELU_C1);
  p = fmaf(p, u, GELU_C0);
  return fmaf(xc, p, 0.5f);
}
__device__ __forceinline__ float gelu_fast(float x) { return x * ncdf_fast(x); }
__device__ __forceinline__ float dgelu_fast(float x) {
  const float xc = fminf(fmaxf(x, -GELU_CLAMP), GELU_CLAMP);
  return fmaf(xc * 0.39894228040143268f, __expf(-0.5f * xc * xc), ncdf_fast(x));
}

// the same on two values per instruction (sm_100 packed fp32 FMA / MUL)
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 gelu_clamp2(float2 x) {
  return make_float2(fminf(fmaxf(x.x, -GELU_CLAMP), GELU_CLAMP), fminf(fmaxf(x.y, -GELU_CLAMP), GELU_CLAMP));
}
__device__ __forceinline__ float2 ncdf2(float2 xc, float2 u) {   // Phi(xc), u = xc * xc
  float2 p = __ffma2_rn(f2(GELU_C6), u, f2(GELU_C5));
  p = __ffma2_rn(p, u, f2(GELU_C4));
  p = __ffma2_rn(p, u, f2(GELU_C3));
  p = __ffma2_rn(p, u, f2(GELU_C2));
  p = __ffma2_rn(p, u, f2(GELU_C1));
  p = __ffma2_rn(p, u, f2(GELU_C0));
  return __ffma2_rn(xc, p, f2(0.5f));
}
__device__ __forceinline__ float2 gelu_fast2(float2 x) {
  const float2 xc = gelu_clamp2(x);
  return __fmul2_rn(x, ncdf2(xc, __fmul2_rn(xc, xc)));
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// g = gelu(x), d = gelu'(x) = Phi(x) + x * phi(x) from one polynomial evaluation; phi(x) = 2^(-x^2 log2(e) / 2 +
// log2(1 / sqrt(2 pi))).  Beyond the clamp the derivative is held at its value at +-4 (1.0005 / -0.0005).
__device__ __forceinline__ void gelu_dgelu2(float2 x, float2& g, float2& d) {
  const float2 xc = gelu_clamp2(x);
  const float2 u = __fmul2_rn(xc, xc);
  const float2 cdf = ncdf2(xc, u);
  g = __fmul2_rn(x, cdf);
  const float2 t = __ffma2_rn(u, f2(-0.5f * 1.4426950408889634f), f2(-1.3257480647361593f));
  d = __ffma2_rn(xc, make_float2(ex2f(t.x), ex2f(t.y)), cdf);
}
