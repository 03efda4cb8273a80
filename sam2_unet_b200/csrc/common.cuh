// Shared device/host helpers for the SAM2-UNet sm_100a kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "sam2unet_b200.h"   // the C ABI: definitions below must match these declarations

typedef __nv_bfloat16 bf16;

enum S2uDtype { S2U_F32 = 0, S2U_BF16 = 1 };

// error convention of the C-ABI: 0 = ok, >0 = cudaError_t of the launch, <0 = argument error
#define S2U_EINVAL (-1)
#define S2U_EUNSUPPORTED (-2)
#define S2U_LAUNCH_CHECK()                      \
  do {                                          \
    cudaError_t e__ = cudaGetLastError();       \
    if (e__ != cudaSuccess) return (int)e__;    \
  } while (0)

// ---- programmatic dependent launch ---------------------------------------------------------------------------
// A train step is ~1,700 short kernels in dependency order; the launch + scheduling latency between two of them is
// comparable to the run time of the small ones.  Every kernel of this library starts with pdl_sync(): it lets the NEXT
// kernel of the stream be scheduled (griddepcontrol.launch_dependents) and then waits until the PREVIOUS one has
// completed and flushed (griddepcontrol.wait) before touching memory, and every launch goes through S2U_LAUNCH,
// which marks the kernel as programmatically serialised - so the launch latency of kernel N+1 overlaps kernel N
// (CUDA graphs record this as programmatic edges).  S2U_PDL=0 in the environment restores plain stream order.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() {
  pdl_launch_dependents();
  pdl_wait();
}
#ifdef __CUDACC__
#include <cstdlib>
static inline bool s2u_pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("S2U_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}
template <typename... KArgs, typename... Args>
static inline void s2u_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = s2u_pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);     // errors surface in S2U_LAUNCH_CHECK
}
#define S2U_LAUNCH(kernel, grid, block, smem, stream, ...) s2u_launch(kernel, grid, block, smem, stream, __VA_ARGS__)
#endif

// opt a kernel into > 48 KB of dynamic shared memory, once per instantiation
#define S2U_ALLOW_SMEM(kernel)                                                                          \
  do {                                                                                                  \
    static bool done__ = false;                                                                         \
    if (!done__) {                                                                                      \
      cudaError_t e__ = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
      if (e__ != cudaSuccess) return (int)e__;                                                          \
      done__ = true;                                                                                    \
    }                                                                                                   \
  } while (0)

#define S2U_DISPATCH_T(dtype, ...)              \
  if ((dtype) == S2U_F32) {                     \
    typedef float T;                            \
    __VA_ARGS__                                 \
  } else if ((dtype) == S2U_BF16) {             \
    typedef bf16 T;                             \
    __VA_ARGS__                                 \
  } else {                                      \
    return S2U_EINVAL;                          \
  }

__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(bf16* p, float v) { *p = __float2bfloat16(v); }

// round-trip through the storage type (what a consumer kernel would read back)
__device__ __forceinline__ float rnd(float v, const float*) { return v; }
__device__ __forceinline__ float rnd(float v, const bf16*) { return __bfloat162float(__float2bfloat16(v)); }

// exact (erf) GELU, nn.GELU() default (SAM2UNet.py:58, hieradet.py:93)
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float dgelu_f(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// 8-element vector access (16 B for bf16, 2x16 B for fp32); pointers must be 16-byte aligned.
struct F8 { float v[8]; };
__device__ __forceinline__ F8 ld8(const float* p) {
  F8 r;
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ F8 ld8(const bf16* p) {
  F8 r;
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    r.v[2 * i] = f.x; r.v[2 * i + 1] = f.y;
  }
  return r;
}
__device__ __forceinline__ void st8(float* p, const F8& r) {
  *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
__device__ __forceinline__ void st8(bf16* p, const F8& r) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(r.v[2 * i], r.v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// GEMM epilogue description shared by the SIMT and the tcgen05 GEMM (see gemm.cu)
// GEMM_OUT_F32 / GEMM_RESID_F32: C / resid are fp32 although the operands are bf16 (the residual stream of the trunk is
// kept in fp32 in bf16 mode); GEMM_PRE_FINAL: pre_out receives the FINAL value (a compute-dtype copy of C) instead of
// the pre-activation one.
// GEMM_SAVE_DGELU: pre_out receives gelu'(pre-activation) instead of the pre-activation, so that the backward GEMM
// only multiplies by it (GEMM_MULAUX: result *= aux) and the erf/exp are evaluated once, in the forward epilogue.
// GEMM_RELU: max(v, 0) last (after the residual): eval-mode conv + folded BatchNorm (+ residual) + ReLU in one GEMM.
enum GemmFlags { GEMM_GELU = 1, GEMM_DGELU = 2, GEMM_RESID = 4, GEMM_OUT_F32 = 16, GEMM_RESID_F32 = 32,
                 GEMM_PRE_FINAL = 64, GEMM_SAVE_DGELU = 128, GEMM_MULAUX = 256, GEMM_RELU = 512 };
struct GemmEpi {
  const float* bias;   // [N] or null
  void* pre_out;       // T [M, ld_pre]: value before the activation (saved for GELU'), or null
  const void* aux;     // T [M, ld_aux]: pre-activation whose GELU' multiplies the result (GEMM_DGELU)
  const void* resid;   // T [M, ld_res]: added last (GEMM_RESID)
  int ld_pre, ld_aux, ld_res;
  int flags;
};
