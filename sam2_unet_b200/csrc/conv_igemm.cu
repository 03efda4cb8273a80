// Implicit-GEMM convolutions of the RFB / decoder (stride-1 "same" kh x kw convolutions with dilation on NHWC maps,
// /root/reference/SAM2UNet.py:68-125 RFB_modified / BasicConv2d, :9-49 DoubleConv / Up), bf16 on tcgen05.
//
//   conv_igemm_kernel   out[p, n] (+)= sum_{tap, ci} x[p + off(tap), ci] * Wm[n, tap * Cin + ci]
//                       forward (Wm = the [64][(ky,kx,ci)] operand) and input gradient (x = d(raw), Wm = the flipped /
//                       transposed [Cin][(ky,kx,co)] operand) of a convolution WITHOUT an im2col matrix: the A tile of
//                       k-block (tap, 64-channel chunk) is ONE 4-D TMA box {64 ch, 16 px, 8 px, 1 image} of the NHWC
//                       map at the tap's (dilated) offset - out-of-image taps are zero-filled by the TMA unit, which
//                       is exactly the convolution's zero padding - landing in shared memory as the 128-row K-major
//                       128-byte-swizzled operand tcgen05.mma reads.  Roles as in gemm.cu's persistent kernel (TMA
//                       producer warp, MMA warp, 8 epilogue warps, two TMEM accumulators).  The forward epilogue also
//                       accumulates the BatchNorm batch statistics (sum, sum of squares per channel of the bf16-rounded
//                       output, SAM2UNet.py:80-86) so no separate statistics pass reads the output again.
//   conv_wgrad_kernel   dW[co, ci, tap] += sum_p d(raw)[p, co] * x[p + off(tap), ci]: both operands MN-major straight
//                       from [64 pixels x 64 channels] TMA boxes (8 x 8 pixel squares; the x box shifted by the tap),
//                       UMMA M = 128 = two (tap, channel-chunk) slices, N = 64 output channels, split over pixel tiles
//                       with fp32 atomics into the state-dict layout [Cout][Cin][kh][kw].
//
// Algorithmic traffic per pixel: forward reads Cin and writes 64 bf16 (the taps hit L2), against (taps + 1) x Cin
// written and read by the im2col formulation it replaces.
#include <cuda.h>

#include "bn_common.cuh"
#include "common.cuh"
#include "umma.cuh"

namespace cig {
using namespace umma;

constexpr int TW = 16, TH = 8;                 // pixel rectangle of one 128-row tile
constexpr int BM = 128, BK = 64;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 64 + 32 * EPI_WARPS;

struct Geo {
  int B, H, W, tiles_x, tiles_y;
  int taps, KW, dil, ph, pw, kchunks;          // kchunks = Cin / 64
  // "tall" mode (KH > 1, N = 64): ONE activation box per (kx, channel chunk) covers the rows of all KH taps of that
  // column - TH + (KH-1)*dil image rows of TW pixels - and tap ky reads it at row offset ky*dil (whole image rows =
  // multiples of the 1024-byte swizzle atom, so the operand descriptor just moves its start address): KH x fewer
  // activation bytes through the TMA / L2 -> SM path, which is what bounds these kernels (ncu: 4.8 TB/s of TMA
  // traffic, tensor pipe 17 %)
  int tall, KH, a_bytes, stage_bytes, nstages;
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

constexpr int WRES_MAX_KB = 9;                 // resident-weight variant: up to 9 k-blocks of [64 x 64] (3x3 taps, Cin = 64)
// WRES: the whole weight operand (<= WRES_MAX_KB k-blocks, N = 64) is loaded ONCE per CTA and stays in shared memory,
// so the ring carries only the activation boxes: a third less L2 -> SM traffic for the 64 -> 64 convolutions, which are
// bound by exactly that (every tap re-reads its shifted box from L2)
template <int BN, int STAGES, bool WRES>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = WRES ? A_BYTES : A_BYTES + B_BYTES;
  static constexpr int W_BYTES = WRES ? WRES_MAX_KB * B_BYTES : 0;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + W_BYTES + 1024;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                    ((uint32_t)(BM >> 4) << 24);
};

template <int BN, int STAGES, bool WRES>
__global__ void __launch_bounds__(THREADS, 1) conv_igemm_kernel(const __grid_constant__ CUtensorMap tma_x,
                                                               const __grid_constant__ CUtensorMap tma_w,
                                                               bf16* __restrict__ C, int ldc, Geo g,
                                                               const float* __restrict__ bias,
                                                               const bf16* __restrict__ resid, int ld_res, int relu,
                                                               double* __restrict__ sums, BnFin fin) {
  using cfg = Cfg<BN, STAGES, WRES>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * STAGES + 5];
  __shared__ uint32_t tmem_holder;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t w_base = (smem_u32(smem_raw) + 1023u) & ~1023u;        // resident weights (WRES), then the ring
  const uint32_t stage_base = w_base + (uint32_t)cfg::W_BYTES;
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t w_bar = bar0 + 8u * (2 * STAGES + 4);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + 2 + a); };

  const int tiles_img = g.tiles_x * g.tiles_y;
  const int num_tiles = g.B * tiles_img;
  const int num_kb = g.taps * g.kchunks;                    // weight k-blocks
  const int units = g.tall ? g.KW * g.kchunks : num_kb;     // load units (ring slots) per tile
  const int nst = g.nstages;
  const int ntap = g.tall ? g.KH : 1;                       // taps served by one load unit
  const int n_outer = g.tall ? g.KW : g.taps;               // load units per tile = n_outer x kchunks

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_x);
    tma_prefetch_desc(&tma_w);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), EPI_WARPS);
    }
    mbar_init(w_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)),
                 "r"((uint32_t)cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      if (WRES) {                                            // the whole weight operand, once
        mbar_expect_tx(w_bar, (uint32_t)(num_kb * cfg::B_BYTES));
        for (int kb = 0; kb < num_kb; ++kb) tma_load_2d(w_base + (uint32_t)(kb * cfg::B_BYTES), &tma_w, w_bar, kb * BK, 0);
      }
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int b = tile / tiles_img, rem = tile - b * tiles_img;
        const int y0 = (rem / g.tiles_x) * TH, x0 = (rem % g.tiles_x) * TW;
        // (no integer divisions in these loops: one elected thread issues every load of the CTA)
        int s = (int)(it % (uint32_t)nst);
        uint32_t ph = (it / (uint32_t)nst) & 1u;
        for (int t0 = 0, ty = 0, tx = 0; t0 < n_outer; ++t0) {        // t0: tap (plain: ty, tx = its row, column) or kx (tall)
         const int oy = g.tall ? -g.ph : ty * g.dil - g.ph;
         const int ox = (g.tall ? t0 : tx) * g.dil - g.pw;
         if (++tx == g.KW) { tx = 0; ++ty; }
         for (int kc = 0; kc < g.kchunks; ++kc, ++it) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          mbar_expect_tx(full_bar(s), (uint32_t)g.stage_bytes);
          const uint32_t sa = stage_base + (uint32_t)(s * g.stage_bytes);
          tma_load_4d(sa, &tma_x, full_bar(s), kc * BK, x0 + ox, y0 + oy, b);
          if (!WRES) {
            for (int t = 0; t < ntap; ++t) {
              const int tap = g.tall ? t * g.KW + t0 : t0;
              tma_load_2d(sa + (uint32_t)(g.a_bytes + t * cfg::B_BYTES), &tma_w, full_bar(s), (tap * g.kchunks + kc) * BK, 0);
            }
          }
          if (++s == nst) { s = 0; ph ^= 1u; }
         }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t it = 0, lt = 0;
      if (WRES) mbar_wait(w_bar, 0);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
        const int acc = lt & 1;
        mbar_wait(tempty_bar(acc), ((lt >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(acc * BN);
        int s = (int)(it % (uint32_t)nst);
        uint32_t ph = (it / (uint32_t)nst) & 1u;
        for (int t0 = 0, u = 0; t0 < n_outer; ++t0) {
         for (int kc = 0; kc < g.kchunks; ++kc, ++u, ++it) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t sa = stage_base + (uint32_t)(s * g.stage_bytes);
          for (int t = 0; t < ntap; ++t) {
            const int kb = g.tall ? ((t * g.KW + t0) * g.kchunks + kc) : u;          // weight k-block of this tap
            // tall: tap ky = t starts t*dil image rows (of TW pixels x 128 B) into the box
            const uint64_t adesc = smem_desc_sw128(sa + (uint32_t)(g.tall ? t * g.dil * (TW * 128) : 0));
            const uint64_t bdesc = smem_desc_sw128(WRES ? w_base + (uint32_t)(kb * cfg::B_BYTES)
                                                        : sa + (uint32_t)(g.a_bytes + t * cfg::B_BYTES));
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              tc_mma_bf16(tacc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), cfg::IDESC,
                          (u > 0 || t > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit(empty_bar(s));
          if (++s == nst) { s = 0; ph ^= 1u; }
         }
        }
        tc_commit(tfull_bar(acc));
      }
    }
  } else {
    // epilogue: thread = tile row = one pixel; the warp's TMEM lane quarter is warp % 4, the two warps of a quarter
    // take alternate 32-column chunks.  Each lane converts and stores its own 64-byte row segment.
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    float ssum = 0.f, ssq = 0.f;                 // BatchNorm partial sums of column half * 32 + lane (BN == 64 only)
    uint32_t lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const int acc = lt & 1;
      const int b = tile / tiles_img, rem = tile - b * tiles_img;
      const int y = (rem / g.tiles_x) * TH + r / TW, x = (rem % g.tiles_x) * TW + r % TW;
      const bool valid = y < g.H && x < g.W;
      const long long pix = ((long long)b * g.H + y) * g.W + x;
      bf16* crow = C + pix * ldc;
      mbar_wait(tfull_bar(acc), (lt >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int c = half; c < BN / 32; c += 2) {
        uint32_t rr[32];
        tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * 32), rr);
        tc_wait_ld();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rr[j]);
        uint4* dst = reinterpret_cast<uint4*>(crow + c * 32);
        if (bias) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + c * 32) + j);
            v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
          }
        }
        if (resid && valid) {
          const uint4* rs = reinterpret_cast<const uint4*>(resid + pix * ld_res + c * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 u = rs[j];
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __bfloat1622float2(h[e]);
              v[8 * j + 2 * e] += f.x;
              v[8 * j + 2 * e + 1] += f.y;
            }
          }
        }
        if (relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        uint4 o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o[j]);
#pragma unroll
          for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[8 * j + 2 * e], v[8 * j + 2 * e + 1]);
        }
        if (valid) {
#pragma unroll
          for (int j = 0; j < 4; ++j) dst[j] = o[j];
        }
        if (sums) {
          // statistics of the ROUNDED output (what bn_apply normalises); rows outside the map contribute nothing.
          // Transposing butterfly: 32 values x 32 lanes -> lane l holds the column-l sum over the warp's 32 rows.
          float s[32], t[32];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&o[j]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __bfloat1622float2(h[e]);
              s[8 * j + 2 * e] = valid ? f.x : 0.f;
              s[8 * j + 2 * e + 1] = valid ? f.y : 0.f;
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) t[j] = s[j] * s[j];
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) {
            const bool up = lane & off;
#pragma unroll
            for (int j = 0; j < off; ++j) {
              const float ks = up ? s[j + off] : s[j], ss = up ? s[j] : s[j + off];
              const float kt = up ? t[j + off] : t[j], st = up ? t[j] : t[j + off];
              s[j] = ks + __shfl_xor_sync(0xffffffffu, ss, off);
              t[j] = kt + __shfl_xor_sync(0xffffffffu, st, off);
            }
          }
          ssum += s[0];
          ssq += t[0];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_bar(acc));
    }
    if (sums) {
      double* rep = sums + (size_t)(blockIdx.x % BN_NREP) * 2 * BN;
      atomicAdd(rep + half * 32 + lane, (double)ssum);
      atomicAdd(rep + BN + half * 32 + lane, (double)ssq);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)cfg::TMEM_COLS)
                 : "memory");
  }
  // BatchNorm finalisation (scale / shift, saved mean / rstd, running statistics) by whichever CTA finishes last
  if (sums && fin.gamma && last_block(reinterpret_cast<unsigned int*>(sums + BN_NREP * 2 * BN)))
    bn_finalize_block(sums, fin, (long long)g.B * g.H * g.W, BN, 1);
}

// ---------------------------------------------------------------------------------------------- weight gradient
constexpr int WG_STAGES = 4;
constexpr int WG_TPH = 16;                    // pixel tile: 8 px wide x 16 rows = two 64-pixel K groups per TMA box (the
                                              // kernels are bound by the number of boxes in flight, not by their bytes)
constexpr int WG_PX = 8 * WG_TPH;
constexpr int WG_X_BYTES = 2 * WG_PX * 128;   // two [128 px x 64 ch] boxes = the 128 rows of D
constexpr int WG_Y_BYTES = WG_PX * 128;
constexpr int WG_STAGE_BYTES = WG_X_BYTES + WG_Y_BYTES;
constexpr int WG_SMEM_BYTES = WG_STAGES * WG_STAGE_BYTES + 1024;
constexpr uint32_t WG_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) |
                              ((uint32_t)(128 >> 4) << 24);

struct WGeo {
  int B, H, W, tiles_x, tiles_y;               // 8 x WG_TPH pixel tiles
  int taps, KW, dil, ph, pw, kchunks, Cin, Cout;
};

__global__ void __launch_bounds__(128) conv_wgrad_kernel(const __grid_constant__ CUtensorMap tma_x,
                                                        const __grid_constant__ CUtensorMap tma_dy,
                                                        float* __restrict__ G, WGeo g, int tiles_per_split) {
  pdl_sync();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * WG_STAGES + 1];
  __shared__ uint32_t tmem_holder;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (WG_STAGES + s); };
  const uint32_t tfull = bar0 + 8u * (2 * WG_STAGES);
  const int nq = g.taps * g.kchunks;                      // 64-wide slices of the (tap, ci) axis
  const int tiles_img = g.tiles_x * g.tiles_y;
  const int total = g.B * tiles_img;
  const int t_beg = blockIdx.z * tiles_per_split;
  const int t_end = min(total, t_beg + tiles_per_split);
  const int n = t_end - t_beg;                            // >= 1 by construction of the grid

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_x);
    tma_prefetch_desc(&tma_dy);
    for (int s = 0; s < WG_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(64u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;

  if (warp == 0 && lane == 0) {
    // the two slices of this CTA: j -> (tap, channel chunk); a slice beyond the axis loads zeros (image index B)
    int oy[2], ox[2], c0[2], bsel[2];
    for (int h = 0; h < 2; ++h) {
      const int j = 2 * blockIdx.x + h;
      const int tap = j / g.kchunks;
      bsel[h] = j < nq ? 0 : g.B;
      oy[h] = (tap / g.KW) * g.dil - g.ph;
      ox[h] = (tap % g.KW) * g.dil - g.pw;
      c0[h] = (j % g.kchunks) * 64;
    }
    for (int i = 0; i < n; ++i) {
      const int s = i % WG_STAGES;
      const uint32_t ph = (uint32_t)(i / WG_STAGES) & 1u;
      const int tile = t_beg + i;
      const int b = tile / tiles_img, rem = tile - b * tiles_img;
      const int y0 = (rem / g.tiles_x) * WG_TPH, x0 = (rem % g.tiles_x) * 8;
      mbar_wait(empty_bar(s), ph ^ 1u);
      mbar_expect_tx(full_bar(s), (uint32_t)WG_STAGE_BYTES);
      const uint32_t sa = smem_base + (uint32_t)s * WG_STAGE_BYTES;
      tma_load_4d(sa, &tma_x, full_bar(s), c0[0], x0 + ox[0], y0 + oy[0], b + bsel[0]);
      tma_load_4d(sa + WG_PX * 128, &tma_x, full_bar(s), c0[1], x0 + ox[1], y0 + oy[1], b + bsel[1]);
      tma_load_4d(sa + WG_X_BYTES, &tma_dy, full_bar(s), 0, x0, y0, b);
    }
  } else if (warp == 1 && lane == 0) {
    for (int i = 0; i < n; ++i) {
      const int s = i % WG_STAGES;
      const uint32_t ph = (uint32_t)(i / WG_STAGES) & 1u;
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      const uint32_t sa = smem_base + (uint32_t)s * WG_STAGE_BYTES;
#pragma unroll
      for (int k = 0; k < WG_PX / 16; ++k) {               // 16 pixels (two 8-row groups = 2048 B) per instruction
        const uint64_t adesc = smem_desc_mn_sw128(sa + (uint32_t)k * 2048u, (uint32_t)(WG_PX * 128));
        const uint64_t bdesc = smem_desc_mn_sw128(sa + WG_X_BYTES + (uint32_t)k * 2048u, (uint32_t)(WG_PX * 128));
        tc_mma_bf16(tmem_base, adesc, bdesc, WG_IDESC, (i > 0 || k > 0) ? 1u : 0u);
      }
      tc_commit(empty_bar(s));
    }
    tc_commit(tfull);
  }
  __syncwarp();
  mbar_wait(tfull, 0);
  tc_fence_after();
  // D row = warp * 32 + lane -> slice j = 2 * blockIdx.x + row / 64, channel ci = chunk * 64 + row % 64; column = co
  const int row = warp * 32 + lane;
  const int j = 2 * blockIdx.x + (row >> 6);
  const int tap = j / g.kchunks, ci = (j % g.kchunks) * 64 + (row & 63);
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {
    uint32_t rr[32];
    tc_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32), rr);
    tc_wait_ld();
    if (j >= nq) continue;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const int co = c * 32 + k;
      if (co < g.Cout) atomicAdd(G + ((long long)co * g.Cin + ci) * g.taps + tap, __uint_as_float(rr[k]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
  }
}

// ---- host side
// NHWC map [B, H, W, C] with pixel pitch ld (elements) as a 4-D tensor {C, W, H, B}; box = {64 channels, bw, bh, 1}
static int make_map_nhwc(CUtensorMap* out, const void* ptr, int B, int H, int W, int C, long long ld, int bw, int bh) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return S2U_EUNSUPPORTED;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * W, (cuuint64_t)ld * 2 * W * H};
  cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -100 - (int)r;
}
// weight operand [N, K] row-major, box = {64 k, rows}
static int make_map_w(CUtensorMap* out, const void* ptr, int N, long long K, int rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return S2U_EUNSUPPORTED;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -100 - (int)r;
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
static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

template <int BN, int STAGES, bool WRES>
static int launch_igemm(const CUtensorMap& mx, const CUtensorMap& mw, bf16* C, int ldc, const Geo& g, const float* bias,
                        const bf16* resid, int ld_res, int relu, double* sums, const BnFin& fin, cudaStream_t st) {
  using cfg = Cfg<BN, STAGES, WRES>;
  static_assert(cfg::SMEM_BYTES <= 227 * 1024, "shared memory budget");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(conv_igemm_kernel<BN, STAGES, WRES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          226 * 1024);          // + < 1 KB of static shared memory
    if (ce != cudaSuccess) return (int)ce;
    attr_set = true;
  }
  const int tiles = g.B * g.tiles_x * g.tiles_y;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  const size_t smem = (size_t)g.nstages * g.stage_bytes + cfg::W_BYTES + 1024;
  S2U_LAUNCH((conv_igemm_kernel<BN, STAGES, WRES>), grid, THREADS, smem, st, mx, mw, C, ldc, g, bias, resid, ld_res,
             relu, sums, fin);
  S2U_LAUNCH_CHECK();
  return 0;
}

}  // namespace cig

extern "C" {

// 1 when s2u_conv_igemm / s2u_conv_wgrad take this convolution (else the caller uses s2u_im2col + s2u_gemm)
int s2u_conv_igemm_supported(int Cin, int N, int ldx, int ld_out) {
  return (Cin % 64 == 0 && (N == 64 || N == 128 || N == 256) && ldx % 8 == 0 && ld_out % 8 == 0) ? 1 : 0;
}

// out[B*H*W, N] (pitch ld_out) = epi(conv(x [B,H,W,Cin] pitch ldx, Wm [N][KH*KW*Cin])) with "same" zero padding and
// dilation; bf16.  epi: + bias[N] (NULL to skip), + resid [.., N] pitch ld_res (NULL to skip; resid == out accumulates),
// relu.  sums (N == 64 only, may be NULL): bn.cu's fp64 statistics workspace, receives the per-channel sum and sum of
// squares of the rounded output (finalise with s2u_bn_finalize).
static int conv_igemm_impl(const void* x, int ldx, int B, int H, int W, int Cin, const void* Wm, int N, int KH, int KW,
                           int dil, void* out, int ld_out, const float* bias, const void* resid, int ld_res, int relu,
                           double* sums, const BnFin& fin, void* stream) {
  if (B <= 0 || H <= 0 || W <= 0 || KH <= 0 || KW <= 0 || dil <= 0) return S2U_EINVAL;
  if (!s2u_conv_igemm_supported(Cin, N, ldx, ld_out) || !cig::aligned16(x) || !cig::aligned16(Wm) || !cig::aligned16(out))
    return S2U_EUNSUPPORTED;
  if ((sums && N != 64) || (resid && (ld_res % 8 || !cig::aligned16(resid)))) return S2U_EINVAL;
  cig::Geo g;
  g.B = B; g.H = H; g.W = W;
  g.tiles_x = ceil_div(W, cig::TW); g.tiles_y = ceil_div(H, cig::TH);
  g.taps = KH * KW; g.KW = KW; g.dil = dil;
  g.ph = dil * (KH - 1) / 2; g.pw = dil * (KW - 1) / 2;
  g.kchunks = Cin / 64;
  // ring geometry: plain = one [128 px x 64 ch] box (+ one weight k-block) per (tap, chunk); tall (KH > 1, N = 64) = one
  // box of TH + (KH-1)*dil image rows (+ KH weight k-blocks) per (kx, chunk)
  const bool wres = N == 64 && g.taps * g.kchunks <= cig::WRES_MAX_KB;
  const int b_bytes = N * cig::BK * 2, budget = 226 * 1024 - 1024 - (wres ? cig::WRES_MAX_KB * b_bytes : 0);
  static int tall_on = -1;
  if (tall_on < 0) { const char* e = getenv("S2U_CONV_TALL"); tall_on = (e && e[0] == '0') ? 0 : 1; }
  g.KH = KH;
  g.tall = 0;
  g.a_bytes = cig::BM * cig::BK * 2;
  g.stage_bytes = g.a_bytes + (wres ? 0 : b_bytes);
  int box_rows = cig::TH;
  if (tall_on && N == 64 && KH > 1) {
    const int rows = cig::TH + (KH - 1) * dil;
    const int stage = rows * cig::TW * 128 + (wres ? 0 : KH * b_bytes);
    if (rows <= 256 && budget / stage >= 2) {
      g.tall = 1;
      g.a_bytes = rows * cig::TW * 128;
      g.stage_bytes = stage;
      box_rows = rows;
    }
  }
  const int max_st = N == 64 ? 8 : (N == 128 ? 6 : 4);
  g.nstages = budget / g.stage_bytes < max_st ? budget / g.stage_bytes : max_st;
  CUtensorMap mx, mw;
  int rc = cig::make_map_nhwc(&mx, x, B, H, W, Cin, ldx, cig::TW, box_rows);
  if (rc) return rc;
  rc = cig::make_map_w(&mw, Wm, N, (long long)g.taps * Cin, N);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  switch (N) {
    case 64:
      if (wres)
        return cig::launch_igemm<64, 8, true>(mx, mw, (bf16*)out, ld_out, g, bias, (const bf16*)resid, ld_res, relu, sums, fin, st);
      return cig::launch_igemm<64, 8, false>(mx, mw, (bf16*)out, ld_out, g, bias, (const bf16*)resid, ld_res, relu, sums, fin, st);
    case 128: return cig::launch_igemm<128, 6, false>(mx, mw, (bf16*)out, ld_out, g, bias, (const bf16*)resid, ld_res, relu, sums, fin, st);
    case 256: return cig::launch_igemm<256, 4, false>(mx, mw, (bf16*)out, ld_out, g, bias, (const bf16*)resid, ld_res, relu, sums, fin, st);
  }
  return S2U_EUNSUPPORTED;
}

int s2u_conv_igemm(const void* x, int ldx, int B, int H, int W, int Cin, const void* Wm, int N, int KH, int KW, int dil,
                   void* out, int ld_out, const float* bias, const void* resid, int ld_res, int relu, double* sums,
                   void* stream) {
  return conv_igemm_impl(x, ldx, B, H, W, Cin, Wm, N, KH, KW, dil, out, ld_out, bias, resid, ld_res, relu, sums, BnFin{},
                         stream);
}

// Training forward of conv + BatchNorm statistics in ONE launch (N = 64): out = conv(x, Wm) (bf16, pitch ld_out), and
// the CTA that finishes last turns the batch statistics of the rounded output into scale / shift for s2u_bn_apply,
// the saved mean / rstd for the backward and the running-statistics update (SAM2UNet.py:80-86).  sums: the fp64
// workspace of s2u_bn_ws_doubles(64), zero on entry and on exit.
int s2u_conv_igemm_bn(const void* x, int ldx, int B, int H, int W, int Cin, const void* Wm, int KH, int KW, int dil,
                      void* out, int ld_out, double* sums, const float* gamma, const float* beta, float* running_mean,
                      float* running_var, long long* num_batches, float* scale, float* shift, float* save_mean,
                      float* save_rstd, float eps, float momentum, void* stream) {
  if (!sums || !gamma || !beta || !running_mean || !running_var || !scale || !shift) return S2U_EINVAL;
  const BnFin fin{gamma, beta, running_mean, running_var, num_batches, scale, shift, save_mean, save_rstd, eps, momentum};
  return conv_igemm_impl(x, ldx, B, H, W, Cin, Wm, 64, KH, KW, dil, out, ld_out, nullptr, nullptr, 0, 0, sums, fin, stream);
}

// G [Cout][Cin][KH][KW] (fp32, the state-dict layout) += d(raw)^T (*) x: the convolution's weight gradient from the
// un-expanded activations.  dy [B,H,W,Cout <= 64] pitch ld_dy, x [B,H,W,Cin] pitch ldx, bf16.
int s2u_conv_wgrad(const void* dy, int ld_dy, const void* x, int ldx, float* G, int B, int H, int W, int Cin, int Cout,
                   int KH, int KW, int dil, void* stream) {
  if (B <= 0 || H <= 0 || W <= 0 || KH <= 0 || KW <= 0 || dil <= 0) return S2U_EINVAL;
  if (Cin % 64 || Cout > 64 || Cout % 8 || ldx % 8 || ld_dy % 8 || !cig::aligned16(dy) || !cig::aligned16(x))
    return S2U_EUNSUPPORTED;
  cig::WGeo g;
  g.B = B; g.H = H; g.W = W;
  g.tiles_x = ceil_div(W, 8); g.tiles_y = ceil_div(H, cig::WG_TPH);
  g.taps = KH * KW; g.KW = KW; g.dil = dil;
  g.ph = dil * (KH - 1) / 2; g.pw = dil * (KW - 1) / 2;
  g.kchunks = Cin / 64; g.Cin = Cin; g.Cout = Cout;
  CUtensorMap mx, my;
  int rc = cig::make_map_nhwc(&mx, x, B, H, W, Cin, ldx, 8, cig::WG_TPH);
  if (rc) return rc;
  rc = cig::make_map_nhwc(&my, dy, B, H, W, Cout, ld_dy, 8, cig::WG_TPH);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(cig::conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          cig::WG_SMEM_BYTES);
    if (ce != cudaSuccess) return (int)ce;
    attr_set = true;
  }
  const int xt = ceil_div(g.taps * g.kchunks, 2);
  const int total = B * g.tiles_x * g.tiles_y;
  static int mult = 0;                                      // CTAs per SM (S2U_WG_CTAS_PER_SM): more splits = more
  if (mult == 0) {                                          // parallelism but one fp32 atomic per weight per split
    const char* e = getenv("S2U_WG_CTAS_PER_SM");
    mult = e ? atoi(e) : 1;                                 // measured (scripts/wg_bench.py): 1 <= 2 < 3 < 4
    if (mult < 1) mult = 1;
  }
  int splits = (mult * cig::num_sms() + xt - 1) / xt;
  if (splits > total / 2) splits = total / 2;               // at least 2 pixel tiles per CTA
  if (splits < 1) splits = 1;
  const int per = (total + splits - 1) / splits;
  splits = (total + per - 1) / per;
  dim3 grid(xt, 1, splits);
  S2U_LAUNCH((cig::conv_wgrad_kernel), grid, 128, cig::WG_SMEM_BYTES, (cudaStream_t)stream, mx, my, G, g, per);
  S2U_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
