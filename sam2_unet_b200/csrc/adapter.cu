// Fused adapter + LayerNorm kernels of a trunk block, bf16 mode (fp32 residual stream).
//
//   forward   xa = x + gelu(gelu(x W1^T + b1) W2^T + b2)          /root/reference/SAM2UNet.py:57-63
//             n1 = LayerNorm(xa)                                   sam2/modeling/backbones/hieradet.py:134
//   backward  dxa = LayerNorm'(dn1) + dres ; dh2 = dxa * gelu'(.) ; dh1 = (dh2 W2) * gelu'(.) ; dx = dxa + dh1 W1
//             db2 += colsum(dh2) ; db1 += colsum(dh1)   (dW2 = dh2^T u and dW1 = dh1^T x stay on the tcgen05
//             weight-gradient kernel, which reads the dh2 / dh1 this kernel writes)
//
// One launch replaces GEMM(K=C,N=32) + GEMM(K=32,N=C, fp32 stream epilogue) + ln_fwd in the forward and
// ln_bwd + GEMM + colsum + GEMM (+ fold) in the backward.  Both are HBM-bound: per row of C channels the forward
// moves 4C (x) + 4C (xa) + 2C (gelu') + 2C (n1) bytes, the backward 2C + 4C + 2C + 2C in and 2C + 2C out; the two
// projections are 4 x 2*C*32 FLOP per row (< 1 % of the tensor peak at the HBM rate), so they run on warp-level
// mma.sync with the 32-wide hidden activation in registers: nothing between x and n1 touches memory.
//
// Work split: a tile is 16 rows (the MMA M); NW warps share a tile, each owning a contiguous slice of 16*NG columns.
// A lane (g = lane / 4, q = lane % 4) holds rows g and g + 8 and, per group of 16 columns, the 4 consecutive columns
// 4q .. 4q+3 (16-byte loads of x, 8-byte stores of bf16).  That is the m16n8 accumulator layout of two n-tiles with
// the column order permuted (n-tile 0 <- columns 4q+{0,1}, n-tile 1 <- columns 4q+{2,3}); the weights are staged in
// shared memory with the matching permutation, so the same registers serve as A operand (down-projection), residual
// and LayerNorm row.  Row statistics and the K-split partial sums of the down-projection cross warps through smem.
#include "common.cuh"
#include "gelu.cuh"

namespace adp {

constexpr int HID = 32;                    // adapter width (SAM2UNet.py:56)
constexpr int NREP = 32;                   // replicated column-sum accumulators (same scheme as norm.cu)

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack2(uint32_t u) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
}
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src)
               : "memory");
}
__device__ __forceinline__ void cp_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// shared-memory image of the two weight matrices of one direction:
//   WD [HID][ldd]  "down" operand, row j = the C input weights of hidden unit j (ldd = C + pad: conflict-free 8-byte
//                  reads by (row g, column 4q) lanes)
//   WU [C][HID]    "up" operand, row c = the HID weights of output channel c, hidden index permuted so that lane q's
//                  eight values (both k-steps of both B registers) are one 16-byte read: WU[c][8q + 2t + e] = W[c][8t + 2q + e]
template <int C, int THREADS>
__device__ __forceinline__ void stage_weights(bf16* WD, bf16* WU, int ldd, const bf16* __restrict__ down,
                                              const bf16* __restrict__ up) {
  for (int i = threadIdx.x; i < HID * (C / 8); i += THREADS) {
    const int j = i / (C / 8), c8 = i % (C / 8);
    cp16(WD + j * ldd + c8 * 8, down + (size_t)j * C + c8 * 8);
  }
  for (int c = threadIdx.x; c < C; c += THREADS) {
    const uint4* src = reinterpret_cast<const uint4*>(up + (size_t)c * HID);
    const uint4 s0 = __ldg(src), s1 = __ldg(src + 1), s2 = __ldg(src + 2), s3 = __ldg(src + 3);
    uint4* dst = reinterpret_cast<uint4*>(WU + c * HID);
    dst[0] = make_uint4(s0.x, s1.x, s2.x, s3.x);
    dst[1] = make_uint4(s0.y, s1.y, s2.y, s3.y);
    dst[2] = make_uint4(s0.z, s1.z, s2.z, s3.z);
    dst[3] = make_uint4(s0.w, s1.w, s2.w, s3.w);
  }
}

// acc[t] (hidden n-tile t) += A . WD[:, slice]^T over this warp's column slice; a[s] = A fragments of column group s
template <int NG>
__device__ __forceinline__ void down_proj(const uint32_t (&a)[NG][4], const bf16* WD, int ldd, int cbase, int g, int q,
                                          float (&acc)[4][4]) {
#pragma unroll
  for (int s = 0; s < NG; ++s) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const uint2 b = *reinterpret_cast<const uint2*>(WD + (8 * t + g) * ldd + cbase + 16 * s + 4 * q);
      mma16816(acc[t], a[s][0], a[s][1], a[s][2], a[s][3], b.x, b.y);
    }
  }
}

// o[0..3] = row g, columns 4q..4q+3 of group s; o[4..7] = row g + 8:  hidden (A fragments ha, two k-steps) . WU^T
__device__ __forceinline__ void up_proj_group(const uint32_t (&ha)[2][4], const bf16* WU, int col0, int g, int q,
                                              float (&o)[8]) {
  const int cA = col0 + 4 * (g >> 1) + (g & 1);
  const uint4 wA = *reinterpret_cast<const uint4*>(WU + cA * HID + 8 * q);
  const uint4 wB = *reinterpret_cast<const uint4*>(WU + (cA + 2) * HID + 8 * q);
  float accA[4] = {0.f, 0.f, 0.f, 0.f}, accB[4] = {0.f, 0.f, 0.f, 0.f};
  mma16816(accA, ha[0][0], ha[0][1], ha[0][2], ha[0][3], wA.x, wA.y);
  mma16816(accA, ha[1][0], ha[1][1], ha[1][2], ha[1][3], wA.z, wA.w);
  mma16816(accB, ha[0][0], ha[0][1], ha[0][2], ha[0][3], wB.x, wB.y);
  mma16816(accB, ha[1][0], ha[1][1], ha[1][2], ha[1][3], wB.z, wB.w);
  o[0] = accA[0]; o[1] = accA[1]; o[2] = accB[0]; o[3] = accB[1];
  o[4] = accA[2]; o[5] = accA[3]; o[6] = accB[2]; o[7] = accB[3];
}

// K-split partial sums of the hidden activation -> full sums in every warp of the row group (fragment layout:
// h[t][0,1] = row g, hidden 8t+2q+{0,1}; h[t][2,3] = row g + 8)
template <int NW>
__device__ __forceinline__ void reduce_hidden(float (&h)[4][4], float* hp, int w, int g, int q) {
  if (NW == 1) return;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    *reinterpret_cast<float2*>(hp + w * 512 + g * HID + 8 * t + 2 * q) = make_float2(h[t][0], h[t][1]);
    *reinterpret_cast<float2*>(hp + w * 512 + (g + 8) * HID + 8 * t + 2 * q) = make_float2(h[t][2], h[t][3]);
  }
  __syncthreads();
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    float2 a = make_float2(0.f, 0.f), b = make_float2(0.f, 0.f);
#pragma unroll
    for (int ww = 0; ww < NW; ++ww) {
      const float2 x = *reinterpret_cast<const float2*>(hp + ww * 512 + g * HID + 8 * t + 2 * q);
      const float2 y = *reinterpret_cast<const float2*>(hp + ww * 512 + (g + 8) * HID + 8 * t + 2 * q);
      a.x += x.x; a.y += x.y; b.x += y.x; b.y += y.y;
    }
    h[t][0] = a.x; h[t][1] = a.y; h[t][2] = b.x; h[t][3] = b.y;
  }
}

// two per-row partial sums (rows g and g + 8) -> totals over the whole row; `st` = [NW][16][2] floats of this row group
template <int NW>
__device__ __forceinline__ void reduce_rows(float& a0, float& b0, float& a1, float& b1, float* st, int w, int g, int q) {
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    b0 += __shfl_xor_sync(0xffffffffu, b0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    b1 += __shfl_xor_sync(0xffffffffu, b1, o);
  }
  if (NW == 1) return;
  if (q == 0) {
    *reinterpret_cast<float2*>(st + (w * 16 + g) * 2) = make_float2(a0, b0);
    *reinterpret_cast<float2*>(st + (w * 16 + g + 8) * 2) = make_float2(a1, b1);
  }
  __syncthreads();
  a0 = b0 = a1 = b1 = 0.f;
#pragma unroll
  for (int ww = 0; ww < NW; ++ww) {
    const float2 x = *reinterpret_cast<const float2*>(st + (ww * 16 + g) * 2);
    const float2 y = *reinterpret_cast<const float2*>(st + (ww * 16 + g + 8) * 2);
    a0 += x.x; b0 += x.y; a1 += y.x; b1 += y.y;
  }
}

template <int NG, int NW>
struct Geo {
  static constexpr int C = 16 * NG * NW;
  static constexpr int WARPS = NW >= 8 ? NW : 4;
  static constexpr int THREADS = 32 * WARPS;
  static constexpr int RG = WARPS / NW;                  // row groups (16 rows each) per CTA
  static constexpr int ROWS = 16 * RG;
  static constexpr int MINB = NW >= 8 ? 1 : 2;
};
static int pad_ldd(int C) {                               // row pitch (elements) with pitch/2 words = 8 or 24 mod 32
  int ldd = C;
  while (((ldd / 2) % 32) != 8 && ((ldd / 2) % 32) != 24) ldd += 8;
  return ldd;
}
static size_t smem_bytes(int C, int warps) {
  return (size_t)HID * pad_ldd(C) * 2 + (size_t)C * HID * 2 + (size_t)warps * 512 * 4 + 2 * (size_t)warps * 16 * 2 * 4;
}

// ------------------------------------------------------------------------------------------------- forward
template <int NG, int NW>
__global__ void __launch_bounds__(Geo<NG, NW>::THREADS, Geo<NG, NW>::MINB)
adapter_ln_fwd_kernel(const float* __restrict__ xs, const bf16* __restrict__ W1, const float* __restrict__ b1,
                      const bf16* __restrict__ W2, const float* __restrict__ b2, const float* __restrict__ gamma,
                      const float* __restrict__ beta, float eps, float* __restrict__ xa, bf16* __restrict__ n1,
                      float* __restrict__ mean_out, float* __restrict__ rstd_out, bf16* __restrict__ u_out,
                      bf16* __restrict__ g1_out, bf16* __restrict__ g2_out, int R, int ldd) {
  using G = Geo<NG, NW>;
  constexpr int C = G::C;
  extern __shared__ __align__(16) uint8_t smem[];
  bf16* WD = reinterpret_cast<bf16*>(smem);
  bf16* WU = WD + HID * ldd;
  float* hpart = reinterpret_cast<float*>(WU + C * HID);
  float* stat = hpart + G::WARPS * 512;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const int rg = warp / NW, w = warp % NW;
  const int cbase = w * 16 * NG;
  const long long r0 = (long long)blockIdx.x * G::ROWS + rg * 16 + g, r1 = r0 + 8;
  const bool ok0 = r0 < R, ok1 = r1 < R;
  float* hp = hpart + rg * NW * 512;
  float* st0 = stat + rg * NW * 32;
  float* st1 = stat + G::WARPS * 32 + rg * NW * 32;

  pdl_launch_dependents();
  pdl_wait();
  // the tile's rows of x first (HBM latency), the weights (L2) behind them
  float v[NG][8];
#pragma unroll
  for (int s = 0; s < NG; ++s) {
    const int c = cbase + 16 * s + 4 * q;
    const float4 a = ok0 ? __ldcs(reinterpret_cast<const float4*>(xs + r0 * C + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 b = ok1 ? __ldcs(reinterpret_cast<const float4*>(xs + r1 * C + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    v[s][0] = a.x; v[s][1] = a.y; v[s][2] = a.z; v[s][3] = a.w;
    v[s][4] = b.x; v[s][5] = b.y; v[s][6] = b.z; v[s][7] = b.w;
  }
  stage_weights<C, G::THREADS>(WD, WU, ldd, W1, W2);
  cp_wait_all();
  __syncthreads();

  // h = x W1^T (this warp's K slice), reduced over the warps of the row group
  float h[4][4];
#pragma unroll
  for (int t = 0; t < 4; ++t) h[t][0] = h[t][1] = h[t][2] = h[t][3] = 0.f;
  {
    uint32_t a[NG][4];
#pragma unroll
    for (int s = 0; s < NG; ++s) {
      a[s][0] = pack2(v[s][0], v[s][1]); a[s][1] = pack2(v[s][4], v[s][5]);
      a[s][2] = pack2(v[s][2], v[s][3]); a[s][3] = pack2(v[s][6], v[s][7]);
    }
    down_proj<NG>(a, WD, ldd, cbase, g, q, h);
  }
  reduce_hidden<NW>(h, hp, w, g, q);
  // u = gelu(h + b1), saved with gelu'(h + b1)
  uint32_t ha[2][4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 bb = __ldg(reinterpret_cast<const float2*>(b1 + 8 * t + 2 * q));
    float2 gu, du, gl, dl;
    gelu_dgelu2(make_float2(h[t][0] + bb.x, h[t][1] + bb.y), gu, du);
    gelu_dgelu2(make_float2(h[t][2] + bb.x, h[t][3] + bb.y), gl, dl);
    const uint32_t pu = pack2(gu.x, gu.y), pl = pack2(gl.x, gl.y);
    ha[t >> 1][(t & 1) * 2] = pu;
    ha[t >> 1][(t & 1) * 2 + 1] = pl;
    if (w == 0 && u_out) {
      if (ok0) {
        *reinterpret_cast<uint32_t*>(u_out + r0 * HID + 8 * t + 2 * q) = pu;
        *reinterpret_cast<uint32_t*>(g1_out + r0 * HID + 8 * t + 2 * q) = pack2(du.x, du.y);
      }
      if (ok1) {
        *reinterpret_cast<uint32_t*>(u_out + r1 * HID + 8 * t + 2 * q) = pl;
        *reinterpret_cast<uint32_t*>(g1_out + r1 * HID + 8 * t + 2 * q) = pack2(dl.x, dl.y);
      }
    }
  }
  // xa = x + gelu(u W2^T + b2), gelu' saved; row sums on the fly
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int s = 0; s < NG; ++s) {
    const int c = cbase + 16 * s + 4 * q;
    float o[8];
    up_proj_group(ha, WU, cbase + 16 * s, g, q, o);
    const float4 bb = __ldg(reinterpret_cast<const float4*>(b2 + c));
    const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
    float d[8];
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      float2 gg, dd;
      gelu_dgelu2(make_float2(o[e] + bv[e & 3], o[e + 1] + bv[(e & 3) + 1]), gg, dd);
      v[s][e] += gg.x; v[s][e + 1] += gg.y;
      d[e] = dd.x; d[e + 1] = dd.y;
    }
    s0 += (v[s][0] + v[s][1]) + (v[s][2] + v[s][3]);
    s1 += (v[s][4] + v[s][5]) + (v[s][6] + v[s][7]);
    if (ok0) {
      *reinterpret_cast<float4*>(xa + r0 * C + c) = make_float4(v[s][0], v[s][1], v[s][2], v[s][3]);
      if (g2_out) *reinterpret_cast<uint2*>(g2_out + r0 * C + c) = make_uint2(pack2(d[0], d[1]), pack2(d[2], d[3]));
    }
    if (ok1) {
      *reinterpret_cast<float4*>(xa + r1 * C + c) = make_float4(v[s][4], v[s][5], v[s][6], v[s][7]);
      if (g2_out) *reinterpret_cast<uint2*>(g2_out + r1 * C + c) = make_uint2(pack2(d[4], d[5]), pack2(d[6], d[7]));
    }
  }
  // LayerNorm over the row: two passes over the registers
  float z0 = 0.f, z1 = 0.f;
  reduce_rows<NW>(s0, z0, s1, z1, st0, w, g, q);
  const float mu0 = s0 * (1.f / C), mu1 = s1 * (1.f / C);
  float q0 = 0.f, q1 = 0.f;
#pragma unroll
  for (int s = 0; s < NG; ++s) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float d0 = v[s][e] - mu0, d1 = v[s][4 + e] - mu1;
      q0 = fmaf(d0, d0, q0);
      q1 = fmaf(d1, d1, q1);
    }
  }
  z0 = z1 = 0.f;
  reduce_rows<NW>(q0, z0, q1, z1, st1, w, g, q);
  const float rs0 = rsqrtf(q0 * (1.f / C) + eps), rs1 = rsqrtf(q1 * (1.f / C) + eps);
#pragma unroll
  for (int s = 0; s < NG; ++s) {
    const int c = cbase + 16 * s + 4 * q;
    const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 bt = __ldg(reinterpret_cast<const float4*>(beta + c));
    if (ok0)
      *reinterpret_cast<uint2*>(n1 + r0 * C + c) =
          make_uint2(pack2(fmaf((v[s][0] - mu0) * rs0, gm.x, bt.x), fmaf((v[s][1] - mu0) * rs0, gm.y, bt.y)),
                     pack2(fmaf((v[s][2] - mu0) * rs0, gm.z, bt.z), fmaf((v[s][3] - mu0) * rs0, gm.w, bt.w)));
    if (ok1)
      *reinterpret_cast<uint2*>(n1 + r1 * C + c) =
          make_uint2(pack2(fmaf((v[s][4] - mu1) * rs1, gm.x, bt.x), fmaf((v[s][5] - mu1) * rs1, gm.y, bt.y)),
                     pack2(fmaf((v[s][6] - mu1) * rs1, gm.z, bt.z), fmaf((v[s][7] - mu1) * rs1, gm.w, bt.w)));
  }
  if (w == 0 && q == 0 && mean_out) {
    if (ok0) { mean_out[r0] = mu0; rstd_out[r0] = rs0; }
    if (ok1) { mean_out[r1] = mu1; rstd_out[r1] = rs1; }
  }
}

// ------------------------------------------------------------------------------------------------ backward
// ws2: NREP x C replicated accumulators of db2 (zero on entry, folded and re-zeroed by colsum_fold_kernel)
template <int NG, int NW>
__global__ void __launch_bounds__(Geo<NG, NW>::THREADS, Geo<NG, NW>::MINB)
adapter_ln_bwd_kernel(const bf16* __restrict__ dn1, const float* __restrict__ xa, const float* __restrict__ mean,
                      const float* __restrict__ rstd, const float* __restrict__ gamma, const bf16* __restrict__ dres,
                      const bf16* __restrict__ g2, const bf16* __restrict__ g1, const bf16* __restrict__ W2t,
                      const bf16* __restrict__ W1t, bf16* __restrict__ dh2_out, bf16* __restrict__ dh1_out,
                      bf16* __restrict__ dx_out, float* __restrict__ db1, float* __restrict__ ws2, int R, int ldd) {
  using G = Geo<NG, NW>;
  constexpr int C = G::C;
  extern __shared__ __align__(16) uint8_t smem[];
  bf16* WD = reinterpret_cast<bf16*>(smem);
  bf16* WU = WD + HID * ldd;
  float* hpart = reinterpret_cast<float*>(WU + C * HID);
  float* stat = hpart + G::WARPS * 512;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const int rg = warp / NW, w = warp % NW;
  const int cbase = w * 16 * NG;
  const long long r0 = (long long)blockIdx.x * G::ROWS + rg * 16 + g, r1 = r0 + 8;
  const bool ok0 = r0 < R, ok1 = r1 < R;
  float* hp = hpart + rg * NW * 512;
  float* st0 = stat + rg * NW * 32;

  pdl_launch_dependents();
  pdl_wait();
  float v[NG][8];                 // xa -> xhat -> dxa
  uint2 d0[NG], d1[NG];           // dn1 rows g / g + 8 (packed bf16), later dres
  const uint2 zero2 = make_uint2(0u, 0u);
#pragma unroll
  for (int s = 0; s < NG; ++s) {
    const int c = cbase + 16 * s + 4 * q;
    const float4 a = ok0 ? __ldcs(reinterpret_cast<const float4*>(xa + r0 * C + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 b = ok1 ? __ldcs(reinterpret_cast<const float4*>(xa + r1 * C + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    v[s][0] = a.x; v[s][1] = a.y; v[s][2] = a.z; v[s][3] = a.w;
    v[s][4] = b.x; v[s][5] = b.y; v[s][6] = b.z; v[s][7] = b.w;
    d0[s] = ok0 ? __ldcs(reinterpret_cast<const uint2*>(dn1 + r0 * C + c)) : zero2;
    d1[s] = ok1 ? __ldcs(reinterpret_cast<const uint2*>(dn1 + r1 * C + c)) : zero2;
  }
  const float mu0 = ok0 ? mean[r0] : 0.f, rs0 = ok0 ? rstd[r0] : 0.f;
  const float mu1 = ok1 ? mean[r1] : 0.f, rs1 = ok1 ? rstd[r1] : 0.f;
  stage_weights<C, G::THREADS>(WD, WU, ldd, W2t, W1t);

  // LayerNorm backward: gd = dn1 * gamma, s1 = sum gd, s2 = sum gd * xhat
  float a0 = 0.f, b0 = 0.f, a1 = 0.f, b1 = 0.f;
#pragma unroll
  for (int s = 0; s < NG; ++s) {
    const int c = cbase + 16 * s + 4 * q;
    const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float2 p0 = unpack2(d0[s].x), p1 = unpack2(d0[s].y), p2 = unpack2(d1[s].x), p3 = unpack2(d1[s].y);
    const float gd[8] = {p0.x * gm.x, p0.y * gm.y, p1.x * gm.z, p1.y * gm.w, p2.x * gm.x, p2.y * gm.y, p3.x * gm.z, p3.y * gm.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[s][e] = (v[s][e] - mu0) * rs0;
      v[s][4 + e] = (v[s][4 + e] - mu1) * rs1;
      a0 += gd[e]; b0 = fmaf(gd[e], v[s][e], b0);
      a1 += gd[4 + e]; b1 = fmaf(gd[4 + e], v[s][4 + e], b1);
    }
  }
  cp_wait_all();
  reduce_rows<NW>(a0, b0, a1, b1, st0, w, g, q);       // (also orders the staged weights when NW > 1)
  if (NW == 1) __syncthreads();
  const float m10 = a0 * (1.f / C), m20 = b0 * (1.f / C), m11 = a1 * (1.f / C), m21 = b1 * (1.f / C);
  // dxa = rstd * (gd - m1 - xhat * m2) + dres ; dh2 = dxa * gelu'
  uint32_t a[NG][4];
#pragma unroll
  for (int s = 0; s < NG; ++s) {
    const int c = cbase + 16 * s + 4 * q;
    const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float2 p0 = unpack2(d0[s].x), p1 = unpack2(d0[s].y), p2 = unpack2(d1[s].x), p3 = unpack2(d1[s].y);
    const float gd[8] = {p0.x * gm.x, p0.y * gm.y, p1.x * gm.z, p1.y * gm.w, p2.x * gm.x, p2.y * gm.y, p3.x * gm.z, p3.y * gm.w};
    uint2 e0 = zero2, e1 = zero2, f0 = zero2, f1 = zero2;
    if (ok0) {
      if (dres) e0 = __ldcs(reinterpret_cast<const uint2*>(dres + r0 * C + c));
      f0 = __ldcs(reinterpret_cast<const uint2*>(g2 + r0 * C + c));
    }
    if (ok1) {
      if (dres) e1 = __ldcs(reinterpret_cast<const uint2*>(dres + r1 * C + c));
      f1 = __ldcs(reinterpret_cast<const uint2*>(g2 + r1 * C + c));
    }
    const float2 r00 = unpack2(e0.x), r01 = unpack2(e0.y), r10 = unpack2(e1.x), r11 = unpack2(e1.y);
    const float rr[8] = {r00.x, r00.y, r01.x, r01.y, r10.x, r10.y, r11.x, r11.y};
    const float2 k00 = unpack2(f0.x), k01 = unpack2(f0.y), k10 = unpack2(f1.x), k11 = unpack2(f1.y);
    const float kk[8] = {k00.x, k00.y, k01.x, k01.y, k10.x, k10.y, k11.x, k11.y};
    float t[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[s][e] = fmaf(rs0, gd[e] - m10 - v[s][e] * m20, rr[e]);
      v[s][4 + e] = fmaf(rs1, gd[4 + e] - m11 - v[s][4 + e] * m21, rr[4 + e]);
      t[e] = v[s][e] * kk[e];
      t[4 + e] = v[s][4 + e] * kk[4 + e];
    }
    a[s][0] = pack2(t[0], t[1]); a[s][1] = pack2(t[4], t[5]);
    a[s][2] = pack2(t[2], t[3]); a[s][3] = pack2(t[6], t[7]);
    if (ok0) *reinterpret_cast<uint2*>(dh2_out + r0 * C + c) = make_uint2(a[s][0], a[s][2]);
    if (ok1) *reinterpret_cast<uint2*>(dh2_out + r1 * C + c) = make_uint2(a[s][1], a[s][3]);
    // db2: column sums over the 16 rows by a halving butterfly over g (4 shuffles), one atomic per column
    const bool hi = lane & 16, mid = lane & 8;
    const float w0 = t[0] + t[4], w1 = t[1] + t[5], w2 = t[2] + t[6], w3 = t[3] + t[7];
    float k0 = (hi ? w2 : w0) + __shfl_xor_sync(0xffffffffu, hi ? w0 : w2, 16);
    float k1 = (hi ? w3 : w1) + __shfl_xor_sync(0xffffffffu, hi ? w1 : w3, 16);
    float k = (mid ? k1 : k0) + __shfl_xor_sync(0xffffffffu, mid ? k0 : k1, 8);
    k += __shfl_xor_sync(0xffffffffu, k, 4);
    if (!(lane & 4)) atomicAdd(ws2 + (size_t)(blockIdx.x % NREP) * C + c + (hi ? 2 : 0) + (mid ? 1 : 0), k);
  }
  // dh1 = (dh2 W2) * gelu'(h1)
  float h[4][4];
#pragma unroll
  for (int t = 0; t < 4; ++t) h[t][0] = h[t][1] = h[t][2] = h[t][3] = 0.f;
  down_proj<NG>(a, WD, ldd, cbase, g, q, h);
  reduce_hidden<NW>(h, hp, w, g, q);
  uint32_t ha[2][4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const uint32_t k0 = ok0 ? *reinterpret_cast<const uint32_t*>(g1 + r0 * HID + 8 * t + 2 * q) : 0u;
    const uint32_t k1 = ok1 ? *reinterpret_cast<const uint32_t*>(g1 + r1 * HID + 8 * t + 2 * q) : 0u;
    const float2 f0 = unpack2(k0), f1 = unpack2(k1);
    h[t][0] *= f0.x; h[t][1] *= f0.y; h[t][2] *= f1.x; h[t][3] *= f1.y;
    const uint32_t pu = pack2(h[t][0], h[t][1]), pl = pack2(h[t][2], h[t][3]);
    ha[t >> 1][(t & 1) * 2] = pu;
    ha[t >> 1][(t & 1) * 2 + 1] = pl;
    if (w == 0) {
      if (ok0) *reinterpret_cast<uint32_t*>(dh1_out + r0 * HID + 8 * t + 2 * q) = pu;
      if (ok1) *reinterpret_cast<uint32_t*>(dh1_out + r1 * HID + 8 * t + 2 * q) = pl;
    }
  }
  if (w == 0) {
    // db1: 8 columns x 2 rows per lane -> one column per lane (7 shuffles), one atomic per column and tile
    const bool hi = lane & 16, mid = lane & 8, low = lane & 4;
    float c4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {               // i = 2 * (t & 1) + e, keep t >> 1 == hi
      const int tl = i >> 1, e = i & 1;
      const float lo2 = h[tl][e] + h[tl][2 + e], hi2 = h[2 + tl][e] + h[2 + tl][2 + e];
      const float keep = hi ? hi2 : lo2, send = hi ? lo2 : hi2;
      c4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    float c2[2];
#pragma unroll
    for (int e = 0; e < 2; ++e)
      c2[e] = (mid ? c4[2 + e] : c4[e]) + __shfl_xor_sync(0xffffffffu, mid ? c4[e] : c4[2 + e], 8);
    const float c1 = (low ? c2[1] : c2[0]) + __shfl_xor_sync(0xffffffffu, low ? c2[0] : c2[1], 4);
    atomicAdd(db1 + 8 * ((hi ? 2 : 0) + (mid ? 1 : 0)) + 2 * q + (low ? 1 : 0), c1);
  }
  // dx = dxa + dh1 W1
#pragma unroll
  for (int s = 0; s < NG; ++s) {
    const int c = cbase + 16 * s + 4 * q;
    float o[8];
    up_proj_group(ha, WU, cbase + 16 * s, g, q, o);
    if (ok0)
      *reinterpret_cast<uint2*>(dx_out + r0 * C + c) =
          make_uint2(pack2(v[s][0] + o[0], v[s][1] + o[1]), pack2(v[s][2] + o[2], v[s][3] + o[3]));
    if (ok1)
      *reinterpret_cast<uint2*>(dx_out + r1 * C + c) =
          make_uint2(pack2(v[s][4] + o[4], v[s][5] + o[5]), pack2(v[s][6] + o[6], v[s][7] + o[7]));
  }
}

// folds the replicated column sums into the gradient and leaves the workspace zeroed for the next call
__global__ void __launch_bounds__(256) colsum_fold_kernel(float* __restrict__ ws, float* __restrict__ colsum, int C) {
  pdl_sync();
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= C) return;
  float v[NREP];
#pragma unroll
  for (int r = 0; r < NREP; ++r) v[r] = ws[(size_t)r * C + j];
  float s = 0.f;
#pragma unroll
  for (int r = 0; r < NREP; ++r) {
    s += v[r];
    ws[(size_t)r * C + j] = 0.f;
  }
  colsum[j] += s;
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
int s2u_adapter_ws_floats(int C) { return adp::NREP * C; }

// bf16 operands, fp32 stream.  u / g1 / g2 (saved for the backward) may all be NULL (inference).
int s2u_adapter_ln_fwd(const float* x, const void* W1, const float* b1, const void* W2, const float* b2,
                       const float* gamma, const float* beta, float eps, float* xa, void* n1, float* mean, float* rstd,
                       void* u, void* g1, void* g2, long long R, int C, void* stream) {
  if (R <= 0 || R > 0x7fffffffLL || C <= 0) return S2U_EINVAL;
  if ((u == nullptr) != (g1 == nullptr) || (u == nullptr) != (g2 == nullptr)) return S2U_EINVAL;
  int ng, nw;
  if (!adp::pick(C, &ng, &nw)) return S2U_EUNSUPPORTED;
  const int ldd = adp::pad_ldd(C);
  cudaStream_t st = (cudaStream_t)stream;
  ADP_DISPATCH({
    using G = adp::Geo<NG, NW>;
    const size_t smem = adp::smem_bytes(C, G::WARPS);
    S2U_ALLOW_SMEM((adp::adapter_ln_fwd_kernel<NG, NW>));
    S2U_LAUNCH((adp::adapter_ln_fwd_kernel<NG, NW>), ceil_div(R, G::ROWS), G::THREADS, smem, st, x, (const bf16*)W1, b1,
               (const bf16*)W2, b2, gamma, beta, eps, xa, (bf16*)n1, mean, rstd, (bf16*)u, (bf16*)g1, (bf16*)g2, (int)R, ldd);
  })
  S2U_LAUNCH_CHECK();
  return 0;
}

// W2t [32, C] and W1t [C, 32]: the transposed operands (the engine's ".wt" shadows).  db1[32] += colsum(dh1),
// db2[C] += colsum(dh2) through ws (s2u_adapter_ws_floats(C) floats, zero before the first use, left zeroed).
int s2u_adapter_ln_bwd(const void* dn1, const float* xa, const float* mean, const float* rstd, const float* gamma,
                       const void* dres, const void* g2, const void* g1, const void* W2t, const void* W1t, void* dh2,
                       void* dh1, void* dx, float* db1, float* db2, float* ws, long long R, int C, void* stream) {
  if (R <= 0 || R > 0x7fffffffLL || C <= 0 || !ws || !db1 || !db2) return S2U_EINVAL;
  int ng, nw;
  if (!adp::pick(C, &ng, &nw)) return S2U_EUNSUPPORTED;
  const int ldd = adp::pad_ldd(C);
  cudaStream_t st = (cudaStream_t)stream;
  ADP_DISPATCH({
    using G = adp::Geo<NG, NW>;
    const size_t smem = adp::smem_bytes(C, G::WARPS);
    S2U_ALLOW_SMEM((adp::adapter_ln_bwd_kernel<NG, NW>));
    S2U_LAUNCH((adp::adapter_ln_bwd_kernel<NG, NW>), ceil_div(R, G::ROWS), G::THREADS, smem, st, (const bf16*)dn1, xa, mean,
               rstd, gamma, (const bf16*)dres, (const bf16*)g2, (const bf16*)g1, (const bf16*)W2t, (const bf16*)W1t,
               (bf16*)dh2, (bf16*)dh1, (bf16*)dx, db1, ws, (int)R, ldd);
  })
  S2U_LAUNCH_CHECK();
  S2U_LAUNCH((adp::colsum_fold_kernel), ceil_div(C, 256), 256, 0, st, ws, db2, C);
  S2U_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
