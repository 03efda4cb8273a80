// Fused adapter + LayerNorm kernels of a trunk block, bf16 mode (fp32 residual stream).
//
//   forward   xa = x + gelu(gelu(x W1^T + b1) W2^T + b2)          /root/reference/SAM2UNet.py:57-63
//             n1 = LayerNorm(xa)                                   sam2/modeling/backbones/hieradet.py:134
//   backward  dxa = LayerNorm'(dn1) + dres ; dh2 = dxa * gelu'(.) ; dh1 = (dh2 W2) * gelu'(.) ; dx = dxa + dh1 W1
//             db2 += colsum(dh2) ; db1 += colsum(dh1)   (dW2 = dh2^T u and dW1 = dh1^T x stay on the tcgen05
//             weight-gradient kernel, which reads the dh2 / dh1 this kernel writes)
//
// One launch replaces GEMM(K=C,N=32) + GEMM(K=32,N=C, fp32 stream epilogue) + ln_fwd in the forward and
// ln_bwd + fold + GEMM + colsum + GEMM in the backward.  Both are HBM-bound: per row of C channels the forward
// moves 4C (x) + 4C (xa) + 2C (gelu') + 2C (n1) bytes, the backward 2C + 4C + 2C + 2C in and 2C + 2C out; the two
// projections are 4 x 2*C*32 FLOP per row (< 1 % of the tensor peak at the HBM rate), so they run on warp-level
// mma.sync with the 32-wide hidden activation in registers: nothing between x and n1 touches memory.
//
// Work split: ONE persistent CTA per SM of RG x NW warps (<= 20).  A row group is 8 rows shared by NW warps, each
// owning a contiguous slice of 16*NG columns; a lane (g = lane / 4, q = lane % 4) holds row g and, per group of 16
// columns, the 4 consecutive columns 4q .. 4q+3 (16-byte loads of x, 8-byte stores of bf16).  That is the upper half
// of the m16n8 accumulator layout of two n-tiles with the column order permuted (n-tile 0 <- columns 4q+{0,1},
// n-tile 1 <- columns 4q+{2,3}; rows 8..15 of the MMA stay zero: the tensor pipe is idle anyway, and 8-row groups
// give 20 warps per SM a share of a 5808-row stage); the weights are staged once per CTA in shared memory with the
// matching permutation, so the same registers serve as A operand (down-projection), residual and LayerNorm row.
// Row statistics and the K-split partial sums of the down-projection cross the NW warps of a row group through smem
// and a named barrier per row group, so row groups drift apart and loads of one overlap the arithmetic of another.
// RG is chosen on the host so that the row count splits evenly over the SMs (5808 rows -> 146 CTAs x 40 rows).
#include "common.cuh"
#include "gelu.cuh"

namespace adp {

constexpr int HID = 32;                    // adapter width (SAM2UNet.py:56)
constexpr int MAX_WARPS = 20;              // 640 threads x <= 102 registers

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack2(uint32_t u) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
}
// rows 8..15 of A are zero (a1 = a3 = 0); c[2], c[3] stay zero
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src)
               : "memory");
}
__device__ __forceinline__ void cp_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
template <int NW>
__device__ __forceinline__ void group_sync(int rg) {       // the NW warps of one row group
  if (NW > 1) asm volatile("bar.sync %0, %1;" ::"r"(rg + 1), "n"(NW * 32) : "memory");
}

// shared-memory image of the two weight matrices of one direction:
//   WD [HID][ldd]  "down" operand, row j = the C input weights of hidden unit j (ldd = C + pad: conflict-free 8-byte
//                  reads by (row g, column 4q) lanes)
//   WU [C][HID]    "up" operand, row c = the HID weights of output channel c, hidden index permuted so that lane q's
//                  eight values (both k-steps of both B registers) are one 16-byte read: WU[c][8q + 2t + e] = W[c][8t + 2q + e]
template <int C>
__device__ __forceinline__ void stage_weights(bf16* WD, bf16* WU, int ldd, const bf16* __restrict__ down,
                                              const bf16* __restrict__ up) {
  for (int i = threadIdx.x; i < HID * (C / 8); i += blockDim.x) {
    const int j = i / (C / 8), c8 = i % (C / 8);
    cp16(WD + j * ldd + c8 * 8, down + (size_t)j * C + c8 * 8);
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const uint4* src = reinterpret_cast<const uint4*>(up + (size_t)c * HID);
    const uint4 s0 = __ldg(src), s1 = __ldg(src + 1), s2 = __ldg(src + 2), s3 = __ldg(src + 3);
    uint4* dst = reinterpret_cast<uint4*>(WU + c * HID);
    dst[0] = make_uint4(s0.x, s1.x, s2.x, s3.x);
    dst[1] = make_uint4(s0.y, s1.y, s2.y, s3.y);
    dst[2] = make_uint4(s0.z, s1.z, s2.z, s3.z);
    dst[3] = make_uint4(s0.w, s1.w, s2.w, s3.w);
  }
}

// acc[t] (hidden n-tile t; [0], [1] = row g, hidden 8t+2q+{0,1}) += A . WD[:, slice]^T over this warp's column slice;
// a[s] = the two A registers of column group s (columns 4q+{0,1} and 4q+{2,3} of row g)
template <int NG>
__device__ __forceinline__ void down_proj(const uint32_t (&a)[NG][2], const bf16* WD, int ldd, int cbase, int g, int q,
                                          float (&acc)[4][4]) {
#pragma unroll
  for (int s = 0; s < NG; ++s) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const uint2 b = *reinterpret_cast<const uint2*>(WD + (8 * t + g) * ldd + cbase + 16 * s + 4 * q);
      mma16816(acc[t], a[s][0], a[s][1], b.x, b.y);
    }
  }
}

// o[0..3] = row g, columns 4q..4q+3 of the group at col0:  hidden (A registers ha[k-step][tile]) . WU^T
__device__ __forceinline__ void up_proj_group(const uint32_t (&ha)[2][2], const bf16* WU, int col0, int g, int q,
                                              float (&o)[4]) {
  const int cA = col0 + 4 * (g >> 1) + (g & 1);
  const uint4 wA = *reinterpret_cast<const uint4*>(WU + cA * HID + 8 * q);
  const uint4 wB = *reinterpret_cast<const uint4*>(WU + (cA + 2) * HID + 8 * q);
  float accA[4] = {0.f, 0.f, 0.f, 0.f}, accB[4] = {0.f, 0.f, 0.f, 0.f};
  mma16816(accA, ha[0][0], ha[0][1], wA.x, wA.y);
  mma16816(accA, ha[1][0], ha[1][1], wA.z, wA.w);
  mma16816(accB, ha[0][0], ha[0][1], wB.x, wB.y);
  mma16816(accB, ha[1][0], ha[1][1], wB.z, wB.w);
  o[0] = accA[0]; o[1] = accA[1]; o[2] = accB[0]; o[3] = accB[1];
}

// K-split partial sums of the hidden activation -> full sums in every warp of the row group; hp = [NW][8][HID] floats
template <int NW>
__device__ __forceinline__ void reduce_hidden(float (&h)[4][4], float* hp, int rg, int w, int g, int q) {
  if (NW == 1) return;
#pragma unroll
  for (int t = 0; t < 4; ++t)
    *reinterpret_cast<float2*>(hp + w * 256 + g * HID + 8 * t + 2 * q) = make_float2(h[t][0], h[t][1]);
  group_sync<NW>(rg);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    float2 a = make_float2(0.f, 0.f);
#pragma unroll
    for (int ww = 0; ww < NW; ++ww) {
      const float2 x = *reinterpret_cast<const float2*>(hp + ww * 256 + g * HID + 8 * t + 2 * q);
      a.x += x.x; a.y += x.y;
    }
    h[t][0] = a.x; h[t][1] = a.y;
  }
}

// two partial sums of row g -> totals over the whole row; st = [NW][8][2] floats of this row group
template <int NW>
__device__ __forceinline__ void reduce_row(float& a, float& b, float* st, int rg, int w, int g, int q) {
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if (NW == 1) return;
  if (q == 0) *reinterpret_cast<float2*>(st + (w * 8 + g) * 2) = make_float2(a, b);
  group_sync<NW>(rg);
  a = b = 0.f;
#pragma unroll
  for (int ww = 0; ww < NW; ++ww) {
    const float2 x = *reinterpret_cast<const float2*>(st + (ww * 8 + g) * 2);
    a += x.x; b += x.y;
  }
}

static int pad_ldd(int C) {                               // row pitch (elements) with pitch/2 words = 8 or 24 mod 32
  int ldd = C;
  while (((ldd / 2) % 32) != 8 && ((ldd / 2) % 32) != 24) ldd += 8;
  return ldd;
}
static size_t smem_bytes(int C, int warps) {
  return (size_t)HID * pad_ldd(C) * 2 + (size_t)C * HID * 2 + (size_t)warps * 256 * 4 + 2 * (size_t)warps * 16 * 4 +
         (size_t)(C + HID) * 4;
}

// ------------------------------------------------------------------------------------------------- forward
template <int NG, int NW>
__global__ void __launch_bounds__(32 * MAX_WARPS, 1)
adapter_ln_fwd_kernel(const float* __restrict__ xs, const bf16* __restrict__ W1, const float* __restrict__ b1,
                      const bf16* __restrict__ W2, const float* __restrict__ b2, const float* __restrict__ gamma,
                      const float* __restrict__ beta, float eps, float* __restrict__ xa, bf16* __restrict__ n1,
                      float* __restrict__ mean_out, float* __restrict__ rstd_out, bf16* __restrict__ u_out,
                      bf16* __restrict__ g1_out, bf16* __restrict__ g2_out, int R, int ldd, int RG) {
  constexpr int C = 16 * NG * NW;
  extern __shared__ __align__(16) uint8_t smem[];
  bf16* WD = reinterpret_cast<bf16*>(smem);
  bf16* WU = WD + HID * ldd;
  float* hpart = reinterpret_cast<float*>(WU + C * HID);
  float* stat = hpart + RG * NW * 256;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const int rg = warp / NW, w = warp % NW;
  const int cbase = w * 16 * NG;
  float* hp = hpart + rg * NW * 256;
  float* st0 = stat + rg * NW * 16;
  float* st1 = stat + RG * NW * 16 + rg * NW * 16;
  const int rows_cta = 8 * RG;
  const int ntiles = (R + rows_cta - 1) / rows_cta;

  pdl_launch_dependents();
  pdl_wait();
  bool staged = false;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long r = (long long)tile * rows_cta + rg * 8 + g;
    const bool ok = r < R;
    // this row group's rows of x first (HBM latency), the weights (L2, once per CTA) behind them
    float v[NG][4];
#pragma unroll
    for (int s = 0; s < NG; ++s) {
      const float4 a = ok ? __ldcs(reinterpret_cast<const float4*>(xs + r * C + cbase + 16 * s + 4 * q))
                          : make_float4(0.f, 0.f, 0.f, 0.f);
      v[s][0] = a.x; v[s][1] = a.y; v[s][2] = a.z; v[s][3] = a.w;
    }
    if (!staged) {
      stage_weights<C>(WD, WU, ldd, W1, W2);
      cp_wait_all();
      __syncthreads();
      staged = true;
    }
    // h = x W1^T (this warp's K slice), reduced over the warps of the row group
    float h[4][4];
#pragma unroll
    for (int t = 0; t < 4; ++t) h[t][0] = h[t][1] = h[t][2] = h[t][3] = 0.f;
    {
      uint32_t a[NG][2];
#pragma unroll
      for (int s = 0; s < NG; ++s) {
        a[s][0] = pack2(v[s][0], v[s][1]);
        a[s][1] = pack2(v[s][2], v[s][3]);
      }
      down_proj<NG>(a, WD, ldd, cbase, g, q, h);
    }
    reduce_hidden<NW>(h, hp, rg, w, g, q);
    // u = gelu(h + b1), saved with gelu'(h + b1)
    uint32_t ha[2][2];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 bb = __ldg(reinterpret_cast<const float2*>(b1 + 8 * t + 2 * q));
      float2 gu, du;
      gelu_dgelu2(make_float2(h[t][0] + bb.x, h[t][1] + bb.y), gu, du);
      const uint32_t pu = pack2(gu.x, gu.y);
      ha[t >> 1][t & 1] = pu;
      if (w == 0 && u_out && ok) {
        *reinterpret_cast<uint32_t*>(u_out + r * HID + 8 * t + 2 * q) = pu;
        *reinterpret_cast<uint32_t*>(g1_out + r * HID + 8 * t + 2 * q) = pack2(du.x, du.y);
      }
    }
    // xa = x + gelu(u W2^T + b2), gelu' saved; row sum on the fly
    float s0 = 0.f;
#pragma unroll
    for (int s = 0; s < NG; ++s) {
      const int c = cbase + 16 * s + 4 * q;
      float o[4];
      up_proj_group(ha, WU, cbase + 16 * s, g, q, o);
      const float4 bb = __ldg(reinterpret_cast<const float4*>(b2 + c));
      float2 g0, d0, g1, d1;
      gelu_dgelu2(make_float2(o[0] + bb.x, o[1] + bb.y), g0, d0);
      gelu_dgelu2(make_float2(o[2] + bb.z, o[3] + bb.w), g1, d1);
      v[s][0] += g0.x; v[s][1] += g0.y; v[s][2] += g1.x; v[s][3] += g1.y;
      s0 += (v[s][0] + v[s][1]) + (v[s][2] + v[s][3]);
      if (ok) {
        *reinterpret_cast<float4*>(xa + r * C + c) = make_float4(v[s][0], v[s][1], v[s][2], v[s][3]);
        if (g2_out) *reinterpret_cast<uint2*>(g2_out + r * C + c) = make_uint2(pack2(d0.x, d0.y), pack2(d1.x, d1.y));
      }
    }
    // LayerNorm over the row: two passes over the registers
    float z = 0.f;
    reduce_row<NW>(s0, z, st0, rg, w, g, q);
    const float mu = s0 * (1.f / C);
    float q0 = 0.f;
#pragma unroll
    for (int s = 0; s < NG; ++s) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float d = v[s][e] - mu;
        q0 = fmaf(d, d, q0);
      }
    }
    z = 0.f;
    reduce_row<NW>(q0, z, st1, rg, w, g, q);
    const float rs = rsqrtf(q0 * (1.f / C) + eps);
#pragma unroll
    for (int s = 0; s < NG; ++s) {
      const int c = cbase + 16 * s + 4 * q;
      const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float4 bt = __ldg(reinterpret_cast<const float4*>(beta + c));
      if (ok)
        *reinterpret_cast<uint2*>(n1 + r * C + c) =
            make_uint2(pack2(fmaf((v[s][0] - mu) * rs, gm.x, bt.x), fmaf((v[s][1] - mu) * rs, gm.y, bt.y)),
                       pack2(fmaf((v[s][2] - mu) * rs, gm.z, bt.z), fmaf((v[s][3] - mu) * rs, gm.w, bt.w)));
    }
    if (w == 0 && q == 0 && ok && mean_out) {
      mean_out[r] = mu;
      rstd_out[r] = rs;
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward
template <int NG, int NW>
__global__ void __launch_bounds__(32 * MAX_WARPS, 1)
adapter_ln_bwd_kernel(const bf16* __restrict__ dn1, const float* __restrict__ xa, const float* __restrict__ mean,
                      const float* __restrict__ rstd, const float* __restrict__ gamma, const bf16* __restrict__ dres,
                      const bf16* __restrict__ g2, const bf16* __restrict__ g1, const bf16* __restrict__ W2t,
                      const bf16* __restrict__ W1t, bf16* __restrict__ dh2_out, bf16* __restrict__ dh1_out,
                      bf16* __restrict__ dx_out, float* __restrict__ db1, float* __restrict__ db2, int R, int ldd,
                      int RG) {
  constexpr int C = 16 * NG * NW;
  extern __shared__ __align__(16) uint8_t smem[];
  bf16* WD = reinterpret_cast<bf16*>(smem);
  bf16* WU = WD + HID * ldd;
  float* hpart = reinterpret_cast<float*>(WU + C * HID);
  float* stat = hpart + RG * NW * 256;
  float* cs2 = stat + 2 * RG * NW * 16;        // [C] column sums of dh2 of this CTA, then [HID] of dh1
  float* cs1 = cs2 + C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const int rg = warp / NW, w = warp % NW;
  const int cbase = w * 16 * NG;
  float* hp = hpart + rg * NW * 256;
  float* st0 = stat + rg * NW * 16;
  const int rows_cta = 8 * RG;
  const int ntiles = (R + rows_cta - 1) / rows_cta;
  const bool hi = lane & 16, mid = lane & 8, low = lane & 4;

  pdl_launch_dependents();
  pdl_wait();
  bool staged = false;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long r = (long long)tile * rows_cta + rg * 8 + g;
    const bool ok = r < R;
    float v[NG][4];                 // xa -> xhat -> dxa
    uint2 dn[NG];                   // dn1 (packed bf16)
#pragma unroll
    for (int s = 0; s < NG; ++s) {
      const int c = cbase + 16 * s + 4 * q;
      const float4 a = ok ? __ldcs(reinterpret_cast<const float4*>(xa + r * C + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      v[s][0] = a.x; v[s][1] = a.y; v[s][2] = a.z; v[s][3] = a.w;
      dn[s] = ok ? __ldcs(reinterpret_cast<const uint2*>(dn1 + r * C + c)) : make_uint2(0u, 0u);
    }
    const float mu = ok ? mean[r] : 0.f, rs = ok ? rstd[r] : 0.f;
    if (!staged) {
      stage_weights<C>(WD, WU, ldd, W2t, W1t);
      for (int j = threadIdx.x; j < C + HID; j += blockDim.x) cs2[j] = 0.f;
      cp_wait_all();
      __syncthreads();
      staged = true;
    }
    // LayerNorm backward: gd = dn1 * gamma, s1 = sum gd, s2 = sum gd * xhat
    float a0 = 0.f, b0 = 0.f;
#pragma unroll
    for (int s = 0; s < NG; ++s) {
      const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + cbase + 16 * s + 4 * q));
      const float2 p0 = unpack2(dn[s].x), p1 = unpack2(dn[s].y);
      const float gd[4] = {p0.x * gm.x, p0.y * gm.y, p1.x * gm.z, p1.y * gm.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v[s][e] = (v[s][e] - mu) * rs;
        a0 += gd[e];
        b0 = fmaf(gd[e], v[s][e], b0);
      }
    }
    reduce_row<NW>(a0, b0, st0, rg, w, g, q);
    const float m1 = a0 * (1.f / C), m2 = b0 * (1.f / C);
    // dxa = rstd * (gd - m1 - xhat * m2) + dres ; dh2 = dxa * gelu'
    uint32_t a[NG][2];
#pragma unroll
    for (int s = 0; s < NG; ++s) {
      const int c = cbase + 16 * s + 4 * q;
      const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float2 p0 = unpack2(dn[s].x), p1 = unpack2(dn[s].y);
      const float gd[4] = {p0.x * gm.x, p0.y * gm.y, p1.x * gm.z, p1.y * gm.w};
      uint2 e0 = make_uint2(0u, 0u), f0 = make_uint2(0u, 0u);
      if (ok) {
        if (dres) e0 = __ldcs(reinterpret_cast<const uint2*>(dres + r * C + c));
        f0 = __ldcs(reinterpret_cast<const uint2*>(g2 + r * C + c));
      }
      const float2 r0 = unpack2(e0.x), r1 = unpack2(e0.y), k0 = unpack2(f0.x), k1 = unpack2(f0.y);
      const float rr[4] = {r0.x, r0.y, r1.x, r1.y}, kk[4] = {k0.x, k0.y, k1.x, k1.y};
      float t[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v[s][e] = fmaf(rs, gd[e] - m1 - v[s][e] * m2, rr[e]);
        t[e] = v[s][e] * kk[e];
      }
      a[s][0] = pack2(t[0], t[1]);
      a[s][1] = pack2(t[2], t[3]);
      if (ok) *reinterpret_cast<uint2*>(dh2_out + r * C + c) = make_uint2(a[s][0], a[s][1]);
      // db2: column sums over the 8 rows by a halving butterfly over g (4 shuffles), one shared-memory add per column
      const float c0 = (hi ? t[2] : t[0]) + __shfl_xor_sync(0xffffffffu, hi ? t[0] : t[2], 16);
      const float c1 = (hi ? t[3] : t[1]) + __shfl_xor_sync(0xffffffffu, hi ? t[1] : t[3], 16);
      float k = (mid ? c1 : c0) + __shfl_xor_sync(0xffffffffu, mid ? c0 : c1, 8);
      k += __shfl_xor_sync(0xffffffffu, k, 4);
      if (!low) atomicAdd(cs2 + c + (hi ? 2 : 0) + (mid ? 1 : 0), k);
    }
    // dh1 = (dh2 W2) * gelu'(h1)
    float h[4][4];
#pragma unroll
    for (int t = 0; t < 4; ++t) h[t][0] = h[t][1] = h[t][2] = h[t][3] = 0.f;
    down_proj<NG>(a, WD, ldd, cbase, g, q, h);
    reduce_hidden<NW>(h, hp, rg, w, g, q);
    uint32_t ha[2][2];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 f = unpack2(ok ? *reinterpret_cast<const uint32_t*>(g1 + r * HID + 8 * t + 2 * q) : 0u);
      h[t][0] *= f.x;
      h[t][1] *= f.y;
      const uint32_t pu = pack2(h[t][0], h[t][1]);
      ha[t >> 1][t & 1] = pu;
      if (w == 0 && ok) *reinterpret_cast<uint32_t*>(dh1_out + r * HID + 8 * t + 2 * q) = pu;
    }
    if (w == 0) {
      // db1: 8 columns per lane -> one column per lane (7 shuffles), one shared-memory add per column and tile
      float c4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {               // i = 2 * (t & 1) + e; a lane keeps t >> 1 == hi
        const float lo2 = h[i >> 1][i & 1], hi2 = h[2 + (i >> 1)][i & 1];
        c4[i] = (hi ? hi2 : lo2) + __shfl_xor_sync(0xffffffffu, hi ? lo2 : hi2, 16);
      }
      float c2[2];
#pragma unroll
      for (int e = 0; e < 2; ++e)
        c2[e] = (mid ? c4[2 + e] : c4[e]) + __shfl_xor_sync(0xffffffffu, mid ? c4[e] : c4[2 + e], 8);
      const float c1 = (low ? c2[1] : c2[0]) + __shfl_xor_sync(0xffffffffu, low ? c2[0] : c2[1], 4);
      atomicAdd(cs1 + 8 * ((hi ? 2 : 0) + (mid ? 1 : 0)) + 2 * q + (low ? 1 : 0), c1);
    }
    // dx = dxa + dh1 W1
#pragma unroll
    for (int s = 0; s < NG; ++s) {
      float o[4];
      up_proj_group(ha, WU, cbase + 16 * s, g, q, o);
      if (ok)
        *reinterpret_cast<uint2*>(dx_out + r * C + cbase + 16 * s + 4 * q) =
            make_uint2(pack2(v[s][0] + o[0], v[s][1] + o[1]), pack2(v[s][2] + o[2], v[s][3] + o[3]));
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < C; j += blockDim.x) atomicAdd(db2 + j, cs2[j]);
  if (threadIdx.x < HID) atomicAdd(db1 + threadIdx.x, cs1[threadIdx.x]);
}

// (NG, NW) for a channel count: slices of 16*NG columns, NG in {2, 6, 7, 9} (embed dims 32 / 96 / 112 / 144), NW warps
static bool pick(int C, int* ng, int* nw) {
  for (int w : {1, 2, 4, 8}) {
    if (C % (16 * w)) continue;
    const int n = C / (16 * w);
    if (n == 2 || n == 6 || n == 7 || n == 9) { *ng = n; *nw = w; return true; }
  }
  return false;
}
static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}
// row groups per CTA: fewest rounds x (rows + a fixed per-round cost), larger CTAs on ties
static int pick_rg(long long R, int nw, int* grid) {
  const int sms = num_sms();
  int rg_max = MAX_WARPS / nw;
  if (nw > 1 && rg_max > 15) rg_max = 15;                 // one named barrier per row group
  int best = 0;
  long long best_cost = 0;
  for (int rg = rg_max; rg >= 1; --rg) {
    const long long tiles = (R + 8 * rg - 1) / (8 * rg);
    const long long rounds = (tiles + sms - 1) / sms;
    const long long cost = rounds * (8 * rg + 8);
    if (!best || cost < best_cost) { best = rg; best_cost = cost; }
  }
  const long long tiles = (R + 8 * best - 1) / (8 * best);
  *grid = (int)(tiles < sms ? tiles : sms);
  return best;
}

#define ADP_CASE(NGv, NWv, ...)                          \
  if (ng == NGv && nw == NWv) {                          \
    constexpr int NG = NGv, NW = NWv;                    \
    __VA_ARGS__                                          \
  } else
#define ADP_DISPATCH(...)                                                                                       \
  ADP_CASE(2, 1, __VA_ARGS__) ADP_CASE(2, 2, __VA_ARGS__) ADP_CASE(2, 4, __VA_ARGS__) ADP_CASE(2, 8, __VA_ARGS__)  \
  ADP_CASE(6, 1, __VA_ARGS__) ADP_CASE(6, 2, __VA_ARGS__) ADP_CASE(6, 4, __VA_ARGS__) ADP_CASE(6, 8, __VA_ARGS__)  \
  ADP_CASE(7, 1, __VA_ARGS__) ADP_CASE(7, 2, __VA_ARGS__) ADP_CASE(7, 4, __VA_ARGS__) ADP_CASE(7, 8, __VA_ARGS__)  \
  ADP_CASE(9, 1, __VA_ARGS__) ADP_CASE(9, 2, __VA_ARGS__) ADP_CASE(9, 4, __VA_ARGS__) ADP_CASE(9, 8, __VA_ARGS__)  \
  { return S2U_EUNSUPPORTED; }

}  // namespace adp

extern "C" {

int s2u_adapter_supported(int C) {
  int ng, nw;
  return adp::pick(C, &ng, &nw) ? 1 : 0;
}

// bf16 operands, fp32 stream.  u / g1 / g2 (saved for the backward) may all be NULL (inference).
int s2u_adapter_ln_fwd(const float* x, const void* W1, const float* b1, const void* W2, const float* b2,
                       const float* gamma, const float* beta, float eps, float* xa, void* n1, float* mean, float* rstd,
                       void* u, void* g1, void* g2, long long R, int C, void* stream) {
  if (R <= 0 || R > 0x7fffffffLL || C <= 0) return S2U_EINVAL;
  if ((u == nullptr) != (g1 == nullptr) || (u == nullptr) != (g2 == nullptr)) return S2U_EINVAL;
  int ng, nw, grid;
  if (!adp::pick(C, &ng, &nw)) return S2U_EUNSUPPORTED;
  const int ldd = adp::pad_ldd(C);
  const int rg = adp::pick_rg(R, nw, &grid);
  const size_t smem = adp::smem_bytes(C, rg * nw);
  cudaStream_t st = (cudaStream_t)stream;
  ADP_DISPATCH({
    S2U_ALLOW_SMEM((adp::adapter_ln_fwd_kernel<NG, NW>));
    S2U_LAUNCH((adp::adapter_ln_fwd_kernel<NG, NW>), grid, 32 * rg * nw, smem, st, x, (const bf16*)W1, b1, (const bf16*)W2, b2,
               gamma, beta, eps, xa, (bf16*)n1, mean, rstd, (bf16*)u, (bf16*)g1, (bf16*)g2, (int)R, ldd, rg);
  })
  S2U_LAUNCH_CHECK();
  return 0;
}

// W2t [32, C] and W1t [C, 32]: the transposed operands (the engine's ".wt" shadows).  db1[32] += colsum(dh1),
// db2[C] += colsum(dh2).
int s2u_adapter_ln_bwd(const void* dn1, const float* xa, const float* mean, const float* rstd, const float* gamma,
                       const void* dres, const void* g2, const void* g1, const void* W2t, const void* W1t, void* dh2,
                       void* dh1, void* dx, float* db1, float* db2, long long R, int C, void* stream) {
  if (R <= 0 || R > 0x7fffffffLL || C <= 0 || !db1 || !db2) return S2U_EINVAL;
  int ng, nw, grid;
  if (!adp::pick(C, &ng, &nw)) return S2U_EUNSUPPORTED;
  const int ldd = adp::pad_ldd(C);
  const int rg = adp::pick_rg(R, nw, &grid);
  const size_t smem = adp::smem_bytes(C, rg * nw);
  cudaStream_t st = (cudaStream_t)stream;
  ADP_DISPATCH({
    S2U_ALLOW_SMEM((adp::adapter_ln_bwd_kernel<NG, NW>));
    S2U_LAUNCH((adp::adapter_ln_bwd_kernel<NG, NW>), grid, 32 * rg * nw, smem, st, (const bf16*)dn1, xa, mean, rstd, gamma,
               (const bf16*)dres, (const bf16*)g2, (const bf16*)g1, (const bf16*)W2t, (const bf16*)W1t, (bf16*)dh2,
               (bf16*)dh1, (bf16*)dx, db1, db2, (int)R, ldd, rg);
  })
  S2U_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
