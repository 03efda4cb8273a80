// Tensor-core (bf16 mma.sync m16n8k16, fp32 accumulate) flash-style windowed / global attention, forward and
// backward, same semantics and layouts as attention.cu (window partition, zero-pad-as-bias tokens, q max-pool,
// un-partition + crop all folded into the addressing; /root/reference/sam2/modeling/backbones/hieradet.py:56-81).
//
// One CTA = 4 warps = 64 query rows (forward, dQ) or 64 key rows (dK/dV) of one (window, head); the other
// operand streams through shared memory in 64-row tiles.  Head dims are zero-padded to HDP in {32,64,80,96}
// inside shared memory only (Hiera-L's 72 -> 80), so global memory keeps the reference layout.
// S/P never leave registers: the QK^T accumulator fragments are re-packed as the A operand of the PV product.
#include "common.cuh"

namespace amma {

struct Geom {
  int B, H, W, nh, hd, wh, ww, nwy, nwx, pool, Ho, Wo, qh, qw;
  float scale;
};

constexpr int BM = 64, BN = 64, NT = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 8 bf16 (16 B) of the qkv row at token (y, x), columns col..col+7; padded tokens hold the (bf16-rounded) bias
__device__ __forceinline__ uint4 tok8(const bf16* __restrict__ qkv, const float* __restrict__ bias, const Geom& g,
                                      int b, int y, int x, int col) {
  if (y < g.H && x < g.W)
    return *reinterpret_cast<const uint4*>(qkv + (((long long)b * g.H + y) * g.W + x) * (3LL * g.nh * g.hd) + col);
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(bias[col + 2 * i], bias[col + 2 * i + 1]);
  return u;
}
__device__ __forceinline__ void cp_async16(bf16* dst, const bf16* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// asynchronous variant: real tokens are copied global -> shared by cp.async (no registers, no stall), padded tokens
// are written directly
__device__ __forceinline__ void tok8_async(bf16* dst, const bf16* __restrict__ qkv, const float* __restrict__ bias,
                                           const Geom& g, int b, int y, int x, int col) {
  if (y < g.H && x < g.W) {
    cp_async16(dst, qkv + (((long long)b * g.H + y) * g.W + x) * (3LL * g.nh * g.hd) + col);
  } else {
    *reinterpret_cast<uint4*>(dst) = tok8(qkv, bias, g, b, y, x, col);
  }
}
__device__ __forceinline__ uint4 max8(uint4 a, uint4 b) {
  uint4 r;
  const __nv_bfloat162* x = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* y = reinterpret_cast<const __nv_bfloat162*>(&b);
  __nv_bfloat162* z = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) z[i] = __hmax2(x[i], y[i]);
  return r;
}

// Per-CTA view of one window.  Only REAL tokens are enumerated (row-major inside the real rh x rw part of the
// window); the zero-padded tokens of a border window all carry the same k = v = bias, so they are folded into ONE
// virtual key (index n_real) whose score gets +ln(n_pad): identical softmax, 2.5x less work on Hiera-L @ 352.
struct Win {
  int b, wy, wx;
  int rh, rw;          // real extent of the window
  int n_real, n_pad;   // real / padded key tokens
  int nk;              // keys to visit: n_real (+1 virtual pad key)
  int qrh, qrw, nq;    // real query extent (pooled grid when g.pool) and count
  float bonus;         // ln(n_pad) / scale, added to the raw score of the virtual key
};
__device__ __forceinline__ Win make_win(const Geom& g, int win) {
  Win w;
  w.b = win / (g.nwy * g.nwx);
  w.wy = (win / g.nwx) % g.nwy;
  w.wx = win % g.nwx;
  w.rh = min(g.wh, g.H - w.wy * g.wh);
  w.rw = min(g.ww, g.W - w.wx * g.ww);
  w.n_real = w.rh * w.rw;
  w.n_pad = g.wh * g.ww - w.n_real;
  w.nk = w.n_real + (w.n_pad > 0 ? 1 : 0);
  w.qrh = g.pool ? w.rh / 2 : w.rh;
  w.qrw = g.pool ? w.rw / 2 : w.rw;
  w.nq = w.qrh * w.qrw;
  w.bonus = w.n_pad > 0 ? __logf((float)w.n_pad) / g.scale : 0.f;
  return w;
}

// one 16-byte chunk (8 head-dim columns starting at c*8) of key `idx` of window w; `which` = 1 (k) or 2 (v)
__device__ __forceinline__ void put_kv(bf16* d, const bf16* qkv, const float* bias, const Geom& g, const Win& w,
                                       int head, int which, int idx, int c) {
  if (idx < w.nk && c * 8 < g.hd) {
    const int col = which * g.nh * g.hd + head * g.hd + c * 8;
    if (idx < w.n_real) {
      const int ty = idx / w.rw, tx = idx - ty * w.rw;
      tok8_async(d, qkv, bias, g, w.b, w.wy * g.wh + ty, w.wx * g.ww + tx, col);
    } else {
      *reinterpret_cast<uint4*>(d) = tok8(qkv, bias, g, w.b, g.H, g.W, col);   // the virtual pad key: bias
    }
  } else {
    *reinterpret_cast<uint4*>(d) = make_uint4(0, 0, 0, 0);
  }
}
// chunk c of real query `idx` (2x2 max-pooled when g.pool)
__device__ __forceinline__ void put_q(bf16* d, const bf16* qkv, const float* bias, const Geom& g, const Win& w,
                                      int head, int idx, int c) {
  if (idx < w.nq && c * 8 < g.hd) {
    const int py = idx / w.qrw, px = idx - py * w.qrw;
    const int col = head * g.hd + c * 8;
    if (!g.pool) {
      tok8_async(d, qkv, bias, g, w.b, w.wy * g.wh + py, w.wx * g.ww + px, col);
    } else {
      const int y = w.wy * g.wh + 2 * py, x = w.wx * g.ww + 2 * px;
      *reinterpret_cast<uint4*>(d) =
          max8(max8(tok8(qkv, bias, g, w.b, y, x, col), tok8(qkv, bias, g, w.b, y, x + 1, col)),
               max8(tok8(qkv, bias, g, w.b, y + 1, x, col), tok8(qkv, bias, g, w.b, y + 1, x + 1, col)));
    }
  } else {
    *reinterpret_cast<uint4*>(d) = make_uint4(0, 0, 0, 0);
  }
}

// output token of real query idx (always inside the cropped output grid), -1 beyond nq
__device__ __forceinline__ long long out_token(const Geom& g, const Win& w, int idx) {
  if (idx >= w.nq) return -1;
  const int py = idx / w.qrw, px = idx - py * w.qrw;
  return ((long long)w.b * g.Ho + (w.wy * g.qh + py)) * g.Wo + (w.wx * g.qw + px);
}
// chunk c of the [B,Ho,Wo,nh*hd] tensor `src` at the output token of query idx (zero beyond nq)
__device__ __forceinline__ void put_o(bf16* d, const bf16* src, const Geom& g, const Win& w, int head, int idx,
                                      int c) {
  const long long tok = out_token(g, w, idx);
  if (tok >= 0 && c * 8 < g.hd)
    cp_async16(d, src + tok * (g.nh * g.hd) + head * g.hd + c * 8);
  else
    *reinterpret_cast<uint4*>(d) = make_uint4(0, 0, 0, 0);
}

// 64-row tiles: rows = key / query indices r0..r0+63 of ONE window.  Two threads per row (NT = 2 * 64): the token
// address is worked out once per row and each thread then moves every other 16-byte chunk of it - the per-chunk
// index arithmetic of a chunk-per-thread loop was 40 % of the forward kernel's instructions (ncu), and these kernels
// are bound by instruction issue.
static_assert(NT == 2 * BM && BM == BN, "tile loaders assume two threads per row");
template <int HDP>
__device__ __forceinline__ void load_kv(bf16* Kdst, bf16* Vdst, const bf16* qkv, const float* bias, const Geom& g,
                                        const Win& w, int head, int r0) {
  constexpr int LD = HDP + 8, CH = HDP / 8;
  const int r = threadIdx.x >> 1, par = threadIdx.x & 1;
  const int idx = r0 + r;
  const int C = g.nh * g.hd;
  bf16* kd = Kdst + r * LD;
  bf16* vd = Vdst + r * LD;
  if (idx < w.n_real) {
    const int ty = idx / w.rw, tx = idx - ty * w.rw;
    const bf16* src = qkv + (((long long)w.b * g.H + (w.wy * g.wh + ty)) * g.W + (w.wx * g.ww + tx)) * (3LL * C) +
                      head * g.hd;
#pragma unroll
    for (int c = par; c < CH; c += 2) {
      if (c * 8 < g.hd) {
        cp_async16(kd + c * 8, src + C + c * 8);
        cp_async16(vd + c * 8, src + 2 * C + c * 8);
      } else {
        *reinterpret_cast<uint4*>(kd + c * 8) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(vd + c * 8) = make_uint4(0, 0, 0, 0);
      }
    }
  } else {
    const bool pad_key = idx < w.nk;                           // the virtual pad key: k = v = bias
#pragma unroll
    for (int c = par; c < CH; c += 2) {
      const bool live = pad_key && c * 8 < g.hd;
      *reinterpret_cast<uint4*>(kd + c * 8) =
          live ? tok8(qkv, bias, g, w.b, g.H, g.W, C + head * g.hd + c * 8) : make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(vd + c * 8) =
          live ? tok8(qkv, bias, g, w.b, g.H, g.W, 2 * C + head * g.hd + c * 8) : make_uint4(0, 0, 0, 0);
    }
  }
}
template <int HDP>
__device__ __forceinline__ void load_q(bf16* dst, const bf16* qkv, const float* bias, const Geom& g, const Win& w,
                                       int head, int r0) {
  constexpr int LD = HDP + 8, CH = HDP / 8;
  const int r = threadIdx.x >> 1, par = threadIdx.x & 1;
  const int idx = r0 + r;
  bf16* d = dst + r * LD;
  if (idx < w.nq && !g.pool) {
    const int py = idx / w.qrw, px = idx - py * w.qrw;
    const bf16* src = qkv + (((long long)w.b * g.H + (w.wy * g.wh + py)) * g.W + (w.wx * g.ww + px)) *
                                (3LL * g.nh * g.hd) + head * g.hd;
#pragma unroll
    for (int c = par; c < CH; c += 2) {
      if (c * 8 < g.hd) cp_async16(d + c * 8, src + c * 8);
      else *reinterpret_cast<uint4*>(d + c * 8) = make_uint4(0, 0, 0, 0);
    }
  } else {
#pragma unroll
    for (int c = par; c < CH; c += 2) put_q(d + c * 8, qkv, bias, g, w, head, idx, c);   // pooled queries / padding
  }
}
template <int HDP>
__device__ __forceinline__ void load_o(bf16* dst, const bf16* src, const Geom& g, const Win& w, int head, int r0) {
  constexpr int LD = HDP + 8, CH = HDP / 8;
  const int r = threadIdx.x >> 1, par = threadIdx.x & 1;
  const long long tok = out_token(g, w, r0 + r);
  bf16* d = dst + r * LD;
  const bf16* s = src + tok * (g.nh * g.hd) + head * g.hd;
#pragma unroll
  for (int c = par; c < CH; c += 2) {
    if (tok >= 0 && c * 8 < g.hd) cp_async16(d + c * 8, s + c * 8);
    else *reinterpret_cast<uint4*>(d + c * 8) = make_uint4(0, 0, 0, 0);
  }
}

// (packed kernel) write this thread's dq fragment rows (query idx, 2 adjacent head-dim columns per 8-column tile j) into the q third
// of dqkv, routing through the arg-max of the 2x2 q max-pool when g.pool (first maximum in scan order, like ATen)
template <int NJ>
__device__ __forceinline__ void scatter_dq(bf16* __restrict__ dqkv, const bf16* __restrict__ qkv,
                                           const float* __restrict__ bias, const Geom& g, const Win& w, int head,
                                           int idx, const float (*dq)[4], int half, int lane) {
  if (idx >= w.nq) return;
  const int C = g.nh * g.hd;
  const long long row3 = 3LL * C;
  const int py = idx / w.qrw, px = idx - py * w.qrw;
  if (!g.pool) {
    const int y = w.wy * g.wh + py, x = w.wx * g.ww + px;
    bf16* dst = dqkv + (((long long)w.b * g.H + y) * g.W + x) * row3 + head * g.hd;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int d = j * 8 + 2 * (lane & 3);
      if (d < g.hd)
        *reinterpret_cast<__nv_bfloat162*>(dst + d) = __floats2bfloat162_rn(dq[j][half * 2], dq[j][half * 2 + 1]);
    }
    return;
  }
  const int y = w.wy * g.wh + 2 * py, x = w.wx * g.ww + 2 * px;
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int d = j * 8 + 2 * (lane & 3) + e;
      if (d >= g.hd) continue;
      const int col = head * g.hd + d;
      float v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int yy = y + (k >> 1), xx = x + (k & 1);
        v[k] = (yy < g.H && xx < g.W) ? __bfloat162float(qkv[(((long long)w.b * g.H + yy) * g.W + xx) * row3 + col])
                                      : __bfloat162float(__float2bfloat16(bias[col]));
      }
      int best = 0;
#pragma unroll
      for (int k = 1; k < 4; ++k)
        if (v[k] > v[best]) best = k;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int yy = y + (k >> 1), xx = x + (k & 1);
        if (yy < g.H && xx < g.W)
          dqkv[(((long long)w.b * g.H + yy) * g.W + xx) * row3 + col] =
              __float2bfloat16(k == best ? dq[j][half * 2 + e] : 0.f);
      }
    }
  }
}

// dq tile (bf16, [64 rows][LD] in shared memory, rows = queries q0..q0+63) -> q third of dqkv with 16-byte stores, two
// threads per row.  With q pooling the gradient of a pooled query goes to the arg-max of its 2x2 source tokens per
// channel (first maximum in scan order, like ATen's max_pool2d backward) and the other three receive zero.
template <int HDP>
__device__ __forceinline__ void store_dq_rows(bf16* __restrict__ dqkv, const bf16* __restrict__ qkv,
                                              const float* __restrict__ bias, const Geom& g, const Win& w, int head,
                                              int q0, const bf16* Ts) {
  constexpr int LD = HDP + 8, CH = HDP / 8;
  const int r = threadIdx.x >> 1, par = threadIdx.x & 1;
  const int idx = q0 + r;
  if (idx >= w.nq) return;
  const int C = g.nh * g.hd;
  const long long row3 = 3LL * C;
  const int py = idx / w.qrw, px = idx - py * w.qrw;
  if (!g.pool) {
    bf16* dst = dqkv + (((long long)w.b * g.H + (w.wy * g.wh + py)) * g.W + (w.wx * g.ww + px)) * row3 + head * g.hd;
#pragma unroll
    for (int c = par; c < CH; c += 2)
      if (c * 8 < g.hd) *reinterpret_cast<uint4*>(dst + c * 8) = *reinterpret_cast<const uint4*>(Ts + r * LD + c * 8);
    return;
  }
  const int y = w.wy * g.wh + 2 * py, x = w.wx * g.ww + 2 * px;
#pragma unroll
  for (int c = par; c < CH; c += 2) {
    if (c * 8 >= g.hd) continue;
    const int col = head * g.hd + c * 8;
    uint4 src[4], o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) src[k] = tok8(qkv, bias, g, w.b, y + (k >> 1), x + (k & 1), col);
    const uint4 dv = *reinterpret_cast<const uint4*>(Ts + r * LD + c * 8);
    const bf16* de = reinterpret_cast<const bf16*>(&dv);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      int best = 0;
      float vb = __bfloat162float(reinterpret_cast<const bf16*>(&src[0])[e]);
#pragma unroll
      for (int k = 1; k < 4; ++k) {
        const float vk = __bfloat162float(reinterpret_cast<const bf16*>(&src[k])[e]);
        if (vk > vb) { vb = vk; best = k; }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) reinterpret_cast<bf16*>(&o[k])[e] = k == best ? de[e] : __float2bfloat16(0.f);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int yy = y + (k >> 1), xx = x + (k & 1);
      if (yy < g.H && xx < g.W)
        *reinterpret_cast<uint4*>(dqkv + (((long long)w.b * g.H + yy) * g.W + xx) * row3 + col) = o[k];
    }
  }
}

// acc[j][0..3] (j = 8-column tile) = A(16 x HDP, this warp's rows of `As`) . Bt(64 x HDP, rows of `Bs`)^T
template <int HDP>
__device__ __forceinline__ void mm_ab_t(float (*acc)[4], const bf16* As, const bf16* Bs, int warp, int lane) {
  constexpr int LD = HDP + 8;
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int t = 0; t < 4; ++t) acc[j][t] = 0.f;
#pragma unroll
  for (int kk = 0; kk < HDP / 16; ++kk) {
    uint32_t a[4];
    ldsm4(smem_u32(As + (warp * 16 + (lane & 15)) * LD + kk * 16 + (lane >> 4) * 8), a[0], a[1], a[2], a[3]);
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {
      uint32_t b0, b1, b2, b3;
      ldsm4(smem_u32(Bs + (jp * 16 + (lane & 7) + (lane >> 4) * 8) * LD + kk * 16 + ((lane >> 3) & 1) * 8), b0, b1,
            b2, b3);
      mma16816(acc[2 * jp], a, b0, b1);
      mma16816(acc[2 * jp + 1], a, b2, b3);
    }
  }
}

// out[j][0..3] (j over HDP/8 column tiles) += P(16 x 64, fragments re-packed from `p`) . Bs(64 x HDP)
template <int HDP>
__device__ __forceinline__ void mm_p_b(float (*out)[4], const float (*p)[4], const bf16* Bs, int lane) {
  constexpr int LD = HDP + 8;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t a[4];
    a[0] = pack_bf16(p[2 * kk][0], p[2 * kk][1]);
    a[1] = pack_bf16(p[2 * kk][2], p[2 * kk][3]);
    a[2] = pack_bf16(p[2 * kk + 1][0], p[2 * kk + 1][1]);
    a[3] = pack_bf16(p[2 * kk + 1][2], p[2 * kk + 1][3]);
#pragma unroll
    for (int dp = 0; dp < HDP / 16; ++dp) {
      uint32_t b0, b1, b2, b3;
      ldsm4t(smem_u32(Bs + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + dp * 16 + (lane >> 4) * 8), b0, b1,
             b2, b3);
      mma16816(out[2 * dp], a, b0, b1);
      mma16816(out[2 * dp + 1], a, b2, b3);
    }
  }
}

// ----------------------------------------------------------------------------------------------- forward
template <int HDP>
__global__ void __launch_bounds__(NT) fwd_kernel(const bf16* __restrict__ qkv, const float* __restrict__ bias,
                                                bf16* __restrict__ out, float* __restrict__ lse, Geom g) {
  pdl_sync();
  constexpr int LD = HDP + 8;
  extern __shared__ __align__(16) uint8_t smraw[];
  bf16* Qs = reinterpret_cast<bf16*>(smraw);
  bf16* KV = Qs + BM * LD;                                      // [2 buffers][K | V]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.z;
  const Win w = make_win(g, blockIdx.y);
  const int q0 = blockIdx.x * BM;
  if (q0 >= w.nq) return;                                       // tile of padded / cropped queries only
  const int nk = w.nk;
  const float sl2 = g.scale * 1.4426950408889634f;

  load_q<HDP>(Qs, qkv, bias, g, w, head, q0);
  load_kv<HDP>(KV, KV + BN * LD, qkv, bias, g, w, head, 0);
  cp_async_commit();
  float o[HDP / 8][4];
#pragma unroll
  for (int j = 0; j < HDP / 8; ++j)
#pragma unroll
    for (int t = 0; t < 4; ++t) o[j][t] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;     // rows lane/4 and lane/4 + 8 of this warp's 16
  for (int k0 = 0, it = 0; k0 < nk; k0 += BN, ++it) {
    const bf16* Ks = KV + (it & 1) * 2 * BN * LD;
    const bf16* Vs = Ks + BN * LD;
    if (k0 + BN < nk) {                                         // prefetch the next key tile into the other buffer
      bf16* nxt = KV + ((it + 1) & 1) * 2 * BN * LD;
      load_kv<HDP>(nxt, nxt + BN * LD, qkv, bias, g, w, head, k0 + BN);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    float s[8][4];
    mm_ab_t<HDP>(s, Qs, Ks, warp, lane);
    float t0 = -INFINITY, t1 = -INFINITY;
    const bool ragged = k0 + BN > w.n_real;                     // the tile holds the pad key and / or the window's end
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (ragged) {
        const int c = k0 + j * 8 + 2 * (lane & 3);
        if (c == w.n_real) { s[j][0] += w.bonus; s[j][2] += w.bonus; }
        if (c + 1 == w.n_real) { s[j][1] += w.bonus; s[j][3] += w.bonus; }
        if (c >= nk) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
        if (c + 1 >= nk) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
      }
      t0 = fmaxf(t0, fmaxf(s[j][0], s[j][1]));
      t1 = fmaxf(t1, fmaxf(s[j][2], s[j][3]));
    }
    t0 = fmaxf(t0, __shfl_xor_sync(0xffffffffu, t0, 1));
    t0 = fmaxf(t0, __shfl_xor_sync(0xffffffffu, t0, 2));
    t1 = fmaxf(t1, __shfl_xor_sync(0xffffffffu, t1, 1));
    t1 = fmaxf(t1, __shfl_xor_sync(0xffffffffu, t1, 2));
    const float n0 = fmaxf(m0, t0), n1 = fmaxf(m1, t1);
    const float a0 = ex2((m0 - n0) * sl2), a1 = ex2((m1 - n1) * sl2);
    float p0 = 0.f, p1 = 0.f;
    const float b0 = -n0 * sl2, b1 = -n1 * sl2;                 // one FMA per score: exp2(s * sl2 - n * sl2)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = ex2(fmaf(s[j][0], sl2, b0));
      s[j][1] = ex2(fmaf(s[j][1], sl2, b0));
      s[j][2] = ex2(fmaf(s[j][2], sl2, b1));
      s[j][3] = ex2(fmaf(s[j][3], sl2, b1));
      p0 += s[j][0] + s[j][1];
      p1 += s[j][2] + s[j][3];
    }
    p0 += __shfl_xor_sync(0xffffffffu, p0, 1);
    p0 += __shfl_xor_sync(0xffffffffu, p0, 2);
    p1 += __shfl_xor_sync(0xffffffffu, p1, 1);
    p1 += __shfl_xor_sync(0xffffffffu, p1, 2);
    l0 = l0 * a0 + p0;
    l1 = l1 * a1 + p1;
    m0 = n0;
    m1 = n1;
#pragma unroll
    for (int j = 0; j < HDP / 8; ++j) {
      o[j][0] *= a0; o[j][1] *= a0; o[j][2] *= a1; o[j][3] *= a1;
    }
    mm_p_b<HDP>(o, s, Vs, lane);
    __syncthreads();                                            // buffer (it & 1) is refilled two iterations later
  }
  const int C = g.nh * g.hd;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int row = warp * 16 + (lane >> 2) + half * 8;
    const long long tok = out_token(g, w, q0 + row);
    if (tok < 0) continue;
    const float inv = 1.f / (half ? l1 : l0);
    bf16* orow = out + tok * C + head * g.hd;
#pragma unroll
    for (int j = 0; j < HDP / 8; ++j) {
      const int d = j * 8 + 2 * (lane & 3);
      if (d < g.hd)
        *reinterpret_cast<__nv_bfloat162*>(orow + d) =
            __floats2bfloat162_rn(o[j][half * 2] * inv, o[j][half * 2 + 1] * inv);
    }
    if ((lane & 3) == 0) lse[tok * g.nh + head] = (half ? m1 : m0) * g.scale + __logf(half ? l1 : l0);
  }
}

// ------------------------------------------------------------------------------------------- backward: dQ
template <int HDP>
__global__ void __launch_bounds__(NT) bwd_dq_kernel(const bf16* __restrict__ qkv, const float* __restrict__ bias,
                                                   const bf16* __restrict__ out, const float* __restrict__ lse,
                                                   float* __restrict__ Dv, const bf16* __restrict__ dout,
                                                   bf16* __restrict__ dqkv, Geom g) {
  pdl_sync();
  constexpr int LD = HDP + 8;
  extern __shared__ __align__(16) uint8_t smraw[];
  bf16* Qs = reinterpret_cast<bf16*>(smraw);
  bf16* dOs = Qs + BM * LD;
  bf16* KV = dOs + BM * LD;                                     // [2 buffers][K | V]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.z;
  const Win w = make_win(g, blockIdx.y);
  const int b = w.b, wy = w.wy, wx = w.wx;
  const int q0 = blockIdx.x * BM;
  if (q0 >= w.nq) return;
  const int nk = w.nk;
  const float sl2 = g.scale * 1.4426950408889634f;

  __shared__ float Dsm[BM];
  load_q<HDP>(Qs, qkv, bias, g, w, head, q0);
  load_o<HDP>(dOs, dout, g, w, head, q0);
  bf16* Os = KV + 2 * BN * LD;                                  // second K/V buffer: free until the first prefetch
  load_o<HDP>(Os, out, g, w, head, q0);
  load_kv<HDP>(KV, KV + BN * LD, qkv, bias, g, w, head, 0);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  {
    // D[row] = sum_d dO * O (the softmax-backward row term), two threads per row; written out for the dK/dV kernel,
    // which runs after this one on the same stream
    const int r = threadIdx.x >> 1, par = threadIdx.x & 1;
    float acc = 0.f;
#pragma unroll
    for (int c = par; c < HDP / 8; c += 2) {
      const uint4 ua = *reinterpret_cast<const uint4*>(dOs + r * LD + c * 8);
      const uint4 ub = *reinterpret_cast<const uint4*>(Os + r * LD + c * 8);
      const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&ua);
      const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&ub);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 fa = __bfloat1622float2(ha[e]), fb = __bfloat1622float2(hb[e]);
        acc = fmaf(fa.x, fb.x, fmaf(fa.y, fb.y, acc));
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (par == 0) {
      Dsm[r] = acc;
      const long long tok = out_token(g, w, q0 + r);
      if (tok >= 0) Dv[tok * g.nh + head] = acc;
    }
  }
  __syncthreads();
  const int r0 = warp * 16 + (lane >> 2);
  const long long tok0 = out_token(g, w, q0 + r0), tok1 = out_token(g, w, q0 + r0 + 8);
  const float L0 = tok0 >= 0 ? lse[tok0 * g.nh + head] * 1.4426950408889634f : INFINITY;
  const float L1 = tok1 >= 0 ? lse[tok1 * g.nh + head] * 1.4426950408889634f : INFINITY;
  const float D0 = Dsm[r0], D1 = Dsm[r0 + 8];
  float dq[HDP / 8][4];
#pragma unroll
  for (int j = 0; j < HDP / 8; ++j)
#pragma unroll
    for (int t = 0; t < 4; ++t) dq[j][t] = 0.f;
  for (int k0 = 0, it = 0; k0 < nk; k0 += BN, ++it) {
    const bf16* Ks = KV + (it & 1) * 2 * BN * LD;
    const bf16* Vs = Ks + BN * LD;
    if (k0 + BN < nk) {
      bf16* nxt = KV + ((it + 1) & 1) * 2 * BN * LD;
      load_kv<HDP>(nxt, nxt + BN * LD, qkv, bias, g, w, head, k0 + BN);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    float s[8][4], dp[8][4];
    mm_ab_t<HDP>(s, Qs, Ks, warp, lane);
    mm_ab_t<HDP>(dp, dOs, Vs, warp, lane);
    const bool ragged = k0 + BN > w.n_real;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      bool v0 = true, v1 = true;
      if (ragged) {
        const int c = k0 + j * 8 + 2 * (lane & 3);
        if (c == w.n_real) { s[j][0] += w.bonus; s[j][2] += w.bonus; }
        if (c + 1 == w.n_real) { s[j][1] += w.bonus; s[j][3] += w.bonus; }
        v0 = c < nk;
        v1 = c + 1 < nk;
      }
      const float p00 = v0 ? ex2(s[j][0] * sl2 - L0) : 0.f, p01 = v1 ? ex2(s[j][1] * sl2 - L0) : 0.f;
      const float p10 = v0 ? ex2(s[j][2] * sl2 - L1) : 0.f, p11 = v1 ? ex2(s[j][3] * sl2 - L1) : 0.f;
      s[j][0] = p00 * (dp[j][0] - D0) * g.scale;
      s[j][1] = p01 * (dp[j][1] - D0) * g.scale;
      s[j][2] = p10 * (dp[j][2] - D1) * g.scale;
      s[j][3] = p11 * (dp[j][3] - D1) * g.scale;
    }
    mm_p_b<HDP>(dq, s, Ks, lane);
    __syncthreads();
  }
  // dq fragments -> shared tile (the Q tile is dead after the loop's last barrier) -> routed 16-byte stores
  bf16* Ts = Qs;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
#pragma unroll
    for (int j = 0; j < HDP / 8; ++j)
      *reinterpret_cast<__nv_bfloat162*>(Ts + (r0 + half * 8) * LD + j * 8 + 2 * (lane & 3)) =
          __floats2bfloat162_rn(dq[j][half * 2], dq[j][half * 2 + 1]);
  }
  __syncthreads();
  store_dq_rows<HDP>(dqkv, qkv, bias, g, w, head, q0, Ts);
}

// --------------------------------------------------------------------------------------- backward: dK, dV
template <int HDP>
__global__ void __launch_bounds__(NT, 3) bwd_dkv_kernel(const bf16* __restrict__ qkv, const float* __restrict__ bias,
                                                    const float* __restrict__ lse, const float* __restrict__ Dv,
                                                    const bf16* __restrict__ dout, bf16* __restrict__ dqkv, Geom g) {
  pdl_sync();
  constexpr int LD = HDP + 8;
  extern __shared__ __align__(16) uint8_t smraw[];
  bf16* Ks = reinterpret_cast<bf16*>(smraw);
  bf16* Vs = Ks + BN * LD;
  bf16* QO = Vs + BN * LD;                                 // [2 buffers][Q | dO]
  float* LD2 = reinterpret_cast<float*>(QO + 4 * BM * LD); // [2 buffers][lse*log2e (64) | D (64)]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.z;
  const Win w = make_win(g, blockIdx.y);
  const int b = w.b, wy = w.wy, wx = w.wx;
  const int k0 = blockIdx.x * BN;
  if (k0 >= w.n_real) return;                                  // only real keys receive gradients
  const int nk = w.n_real, nq = w.nq;
  const float sl2 = g.scale * 1.4426950408889634f;

  auto stage_q = [&](int buf, int q0) {                     // Q, dO, lse, D of one query tile -> buffer `buf`
    bf16* qs = QO + buf * 2 * BM * LD;
    load_q<HDP>(qs, qkv, bias, g, w, head, q0);
    load_o<HDP>(qs + BM * LD, dout, g, w, head, q0);
    if (threadIdx.x < BM) {
      const long long tok = out_token(g, w, q0 + threadIdx.x);
      LD2[buf * 2 * BM + threadIdx.x] = tok >= 0 ? lse[tok * g.nh + head] * 1.4426950408889634f : INFINITY;
      LD2[buf * 2 * BM + BM + threadIdx.x] = tok >= 0 ? Dv[tok * g.nh + head] : 0.f;
    }
  };
  load_kv<HDP>(Ks, Vs, qkv, bias, g, w, head, k0);
  stage_q(0, 0);
  cp_async_commit();
  float dk[HDP / 8][4], dv[HDP / 8][4];
#pragma unroll
  for (int j = 0; j < HDP / 8; ++j)
#pragma unroll
    for (int t = 0; t < 4; ++t) { dk[j][t] = 0.f; dv[j][t] = 0.f; }
  const int kr0 = k0 + warp * 16 + (lane >> 2);            // this thread's key rows: kr0, kr0 + 8
  for (int q0 = 0, it = 0; q0 < nq; q0 += BM, ++it) {
    const bf16* Qs = QO + (it & 1) * 2 * BM * LD;
    const bf16* dOs = Qs + BM * LD;
    const float* Ls = LD2 + (it & 1) * 2 * BM;
    const float* Ds = Ls + BM;
    if (q0 + BM < nq) stage_q((it + 1) & 1, q0 + BM);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    float st[8][4], dpt[8][4];                              // [16 keys x 64 queries]
    mm_ab_t<HDP>(st, Ks, Qs, warp, lane);
    mm_ab_t<HDP>(dpt, Vs, dOs, warp, lane);
#pragma unroll
    for (int j = 0; j < 8; ++j) {                            // st <- P^T
      const int c = j * 8 + 2 * (lane & 3);                  // query column inside the tile
      const float La = Ls[c], Lb = Ls[c + 1];
      const bool ka = kr0 < nk, kb = kr0 + 8 < nk;
      st[j][0] = ka ? ex2(st[j][0] * sl2 - La) : 0.f;
      st[j][1] = ka ? ex2(st[j][1] * sl2 - Lb) : 0.f;
      st[j][2] = kb ? ex2(st[j][2] * sl2 - La) : 0.f;
      st[j][3] = kb ? ex2(st[j][3] * sl2 - Lb) : 0.f;
    }
    mm_p_b<HDP>(dv, st, dOs, lane);
#pragma unroll
    for (int j = 0; j < 8; ++j) {                            // st <- dS^T = P^T o (dP^T - D) * scale
      const int c = j * 8 + 2 * (lane & 3);
      const float Da = Ds[c], Db = Ds[c + 1];
      st[j][0] *= (dpt[j][0] - Da) * g.scale;
      st[j][1] *= (dpt[j][1] - Db) * g.scale;
      st[j][2] *= (dpt[j][2] - Da) * g.scale;
      st[j][3] *= (dpt[j][3] - Db) * g.scale;
    }
    mm_p_b<HDP>(dk, st, Qs, lane);
    __syncthreads();
  }
  const int C = g.nh * g.hd;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int idx = kr0 + half * 8;
    if (idx >= nk) continue;                                 // (the virtual pad key is the frozen bias: no gradient)
    const int ty = idx / w.rw, tx = idx - ty * w.rw;
    const int y = wy * g.wh + ty, x = wx * g.ww + tx;
    bf16* dst = dqkv + (((long long)b * g.H + y) * g.W + x) * (3LL * C) + head * g.hd;
#pragma unroll
    for (int j = 0; j < HDP / 8; ++j) {
      const int d = j * 8 + 2 * (lane & 3);
      if (d < g.hd) {
        *reinterpret_cast<__nv_bfloat162*>(dst + C + d) = __floats2bfloat162_rn(dk[j][half * 2], dk[j][half * 2 + 1]);
        *reinterpret_cast<__nv_bfloat162*>(dst + 2 * C + d) = __floats2bfloat162_rn(dv[j][half * 2], dv[j][half * 2 + 1]);
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------------------
// Whole-window variant for windows of <= 64 tokens (Hiera stage 1: 8x8 windows, and the pooled transition into
// stage 2): the window is ONE 64-row tile, so one CTA = (window, head) does the entire forward or the entire
// backward with every operand loaded once by cp.async:
//   * q max-pool: the window's raw q rows are staged in shared memory and pooled there (the pooled loader of the
//     general kernel issues four dependent global loads per 16 bytes); the backward routes dq to the arg-max from the
//     same staged rows instead of re-reading global memory;
//   * backward: dQ part (warp = 16 queries: S, P, dP, D = rowsum(P o dP), dS, dQ = dS K) and dK/dV part (warp = 16
//     keys: S^T = K Q^T needs no transpose) share the tiles; log-sum-exp and D cross between them through shared
//     memory instead of the lse / D workspaces, and dK / dV / dQ leave through the dead K / V tiles and a dQ tile with
//     16-byte row stores.
// Traffic per call: qkv read once + dqkv written once (the two-kernel path reads qkv twice and O once).
// ------------------------------------------------------------------------------------------------------------

// raw q rows of the window's real tokens (row-major in the real extent) -> Qraw by cp.async; rows beyond: zero
template <int HDP>
__device__ __forceinline__ void load_qraw(bf16* dst, const bf16* qkv, const Geom& g, const Win& w, int head) {
  constexpr int LD = HDP + 8, CH = HDP / 8;
  const int r = threadIdx.x >> 1, par = threadIdx.x & 1;
  bf16* d = dst + r * LD;
  if (r < w.n_real) {
    const int ty = r / w.rw, tx = r - ty * w.rw;
    const bf16* src = qkv + (((long long)w.b * g.H + (w.wy * g.wh + ty)) * g.W + (w.wx * g.ww + tx)) *
                                (3LL * g.nh * g.hd) + head * g.hd;
#pragma unroll
    for (int c = par; c < CH; c += 2) {
      if (c * 8 < g.hd) cp_async16(d + c * 8, src + c * 8);
      else *reinterpret_cast<uint4*>(d + c * 8) = make_uint4(0, 0, 0, 0);
    }
  } else {
#pragma unroll
    for (int c = par; c < CH; c += 2) *reinterpret_cast<uint4*>(d + c * 8) = make_uint4(0, 0, 0, 0);
  }
}
// Qs[pooled query] = max over its 2x2 raw rows (real extents are even when pooling); rows beyond nq: zero
template <int HDP>
__device__ __forceinline__ void pool_q_smem(bf16* Qs, const bf16* Qraw, const Win& w) {
  constexpr int LD = HDP + 8, CH = HDP / 8;
  const int r = threadIdx.x >> 1, par = threadIdx.x & 1;
  if (r < w.nq) {
    const int py = r / w.qrw, px = r - py * w.qrw;
    const bf16* s0 = Qraw + ((2 * py) * w.rw + 2 * px) * LD;
    const bf16* s1 = s0 + w.rw * LD;
#pragma unroll
    for (int c = par; c < CH; c += 2)
      *reinterpret_cast<uint4*>(Qs + r * LD + c * 8) =
          max8(max8(*reinterpret_cast<const uint4*>(s0 + c * 8), *reinterpret_cast<const uint4*>(s0 + LD + c * 8)),
               max8(*reinterpret_cast<const uint4*>(s1 + c * 8), *reinterpret_cast<const uint4*>(s1 + LD + c * 8)));
  } else {
#pragma unroll
    for (int c = par; c < CH; c += 2) *reinterpret_cast<uint4*>(Qs + r * LD + c * 8) = make_uint4(0, 0, 0, 0);
  }
}

template <int HDP>
__global__ void __launch_bounds__(NT, 3) fwd_w64_kernel(const bf16* __restrict__ qkv, const float* __restrict__ bias,
                                                       bf16* __restrict__ out, float* __restrict__ lse, Geom g) {
  pdl_sync();
  constexpr int LD = HDP + 8;
  extern __shared__ __align__(16) uint8_t smraw[];
  bf16* Qs = reinterpret_cast<bf16*>(smraw);
  bf16* Ks = Qs + BM * LD;
  bf16* Vs = Ks + BN * LD;
  bf16* Qraw = Vs + BN * LD;                                    // pooling only
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.z;
  const Win w = make_win(g, blockIdx.y);
  const int nk = w.nk;
  const float sl2 = g.scale * 1.4426950408889634f;
  if (g.pool) load_qraw<HDP>(Qraw, qkv, g, w, head);
  else load_q<HDP>(Qs, qkv, bias, g, w, head, 0);
  load_kv<HDP>(Ks, Vs, qkv, bias, g, w, head, 0);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  if (g.pool) {
    pool_q_smem<HDP>(Qs, Qraw, w);
    __syncthreads();
  }
  if (warp * 16 >= w.nq) return;                                // no barrier below
  float s[8][4];
  mm_ab_t<HDP>(s, Qs, Ks, warp, lane);
  float t0 = -INFINITY, t1 = -INFINITY;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = j * 8 + 2 * (lane & 3);
    if (c == w.n_real) { s[j][0] += w.bonus; s[j][2] += w.bonus; }
    if (c + 1 == w.n_real) { s[j][1] += w.bonus; s[j][3] += w.bonus; }
    if (c >= nk) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
    if (c + 1 >= nk) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
    t0 = fmaxf(t0, fmaxf(s[j][0], s[j][1]));
    t1 = fmaxf(t1, fmaxf(s[j][2], s[j][3]));
  }
  t0 = fmaxf(t0, __shfl_xor_sync(0xffffffffu, t0, 1));
  t0 = fmaxf(t0, __shfl_xor_sync(0xffffffffu, t0, 2));
  t1 = fmaxf(t1, __shfl_xor_sync(0xffffffffu, t1, 1));
  t1 = fmaxf(t1, __shfl_xor_sync(0xffffffffu, t1, 2));
  float p0 = 0.f, p1 = 0.f;
  const float b0 = -t0 * sl2, b1 = -t1 * sl2;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s[j][0] = ex2(fmaf(s[j][0], sl2, b0));
    s[j][1] = ex2(fmaf(s[j][1], sl2, b0));
    s[j][2] = ex2(fmaf(s[j][2], sl2, b1));
    s[j][3] = ex2(fmaf(s[j][3], sl2, b1));
    p0 += s[j][0] + s[j][1];
    p1 += s[j][2] + s[j][3];
  }
  p0 += __shfl_xor_sync(0xffffffffu, p0, 1);
  p0 += __shfl_xor_sync(0xffffffffu, p0, 2);
  p1 += __shfl_xor_sync(0xffffffffu, p1, 1);
  p1 += __shfl_xor_sync(0xffffffffu, p1, 2);
  float o[HDP / 8][4];
#pragma unroll
  for (int j = 0; j < HDP / 8; ++j)
#pragma unroll
    for (int t = 0; t < 4; ++t) o[j][t] = 0.f;
  mm_p_b<HDP>(o, s, Vs, lane);
  const int C = g.nh * g.hd;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int row = warp * 16 + (lane >> 2) + half * 8;
    const long long tok = out_token(g, w, row);
    if (tok < 0) continue;
    const float inv = 1.f / (half ? p1 : p0);
    bf16* orow = out + tok * C + head * g.hd;
#pragma unroll
    for (int j = 0; j < HDP / 8; ++j) {
      const int d = j * 8 + 2 * (lane & 3);
      if (d < g.hd)
        *reinterpret_cast<__nv_bfloat162*>(orow + d) = __floats2bfloat162_rn(o[j][half * 2] * inv, o[j][half * 2 + 1] * inv);
    }
    if ((lane & 3) == 0) lse[tok * g.nh + head] = (half ? t1 : t0) * g.scale + __logf(half ? p1 : p0);
  }
}

template <int HDP>
__global__ void __launch_bounds__(NT, 3) bwd_w64_kernel(const bf16* __restrict__ qkv, const float* __restrict__ bias,
                                                       const bf16* __restrict__ dout, bf16* __restrict__ dqkv, Geom g) {
  pdl_sync();
  constexpr int LD = HDP + 8, CH = HDP / 8;
  extern __shared__ __align__(16) uint8_t smraw[];
  bf16* Qs = reinterpret_cast<bf16*>(smraw);
  bf16* Ks = Qs + BM * LD;
  bf16* Vs = Ks + BN * LD;
  bf16* dOs = Vs + BN * LD;
  bf16* Ts = dOs + BM * LD;                                     // dQ tile
  bf16* Qraw = Ts + BM * LD;                                    // pooling only
  __shared__ float Lsm[BM], Dsm[BM];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.z;
  const Win w = make_win(g, blockIdx.y);
  const int nk = w.nk;
  const float sl2 = g.scale * 1.4426950408889634f;
  if (g.pool) load_qraw<HDP>(Qraw, qkv, g, w, head);
  else load_q<HDP>(Qs, qkv, bias, g, w, head, 0);
  load_kv<HDP>(Ks, Vs, qkv, bias, g, w, head, 0);
  load_o<HDP>(dOs, dout, g, w, head, 0);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  if (g.pool) {
    pool_q_smem<HDP>(Qs, Qraw, w);
    __syncthreads();
  }
  const int r0 = warp * 16 + (lane >> 2);
  // ---- part 1: this warp's 16 queries
  {
    float s[8][4], dp[8][4];
    mm_ab_t<HDP>(s, Qs, Ks, warp, lane);
    mm_ab_t<HDP>(dp, dOs, Vs, warp, lane);
    float t0 = -INFINITY, t1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = j * 8 + 2 * (lane & 3);
      if (c == w.n_real) { s[j][0] += w.bonus; s[j][2] += w.bonus; }
      if (c + 1 == w.n_real) { s[j][1] += w.bonus; s[j][3] += w.bonus; }
      if (c >= nk) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
      if (c + 1 >= nk) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
      t0 = fmaxf(t0, fmaxf(s[j][0], s[j][1]));
      t1 = fmaxf(t1, fmaxf(s[j][2], s[j][3]));
    }
    t0 = fmaxf(t0, __shfl_xor_sync(0xffffffffu, t0, 1));
    t0 = fmaxf(t0, __shfl_xor_sync(0xffffffffu, t0, 2));
    t1 = fmaxf(t1, __shfl_xor_sync(0xffffffffu, t1, 1));
    t1 = fmaxf(t1, __shfl_xor_sync(0xffffffffu, t1, 2));
    float p0 = 0.f, p1 = 0.f;
    const float b0 = -t0 * sl2, b1 = -t1 * sl2;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = ex2(fmaf(s[j][0], sl2, b0));
      s[j][1] = ex2(fmaf(s[j][1], sl2, b0));
      s[j][2] = ex2(fmaf(s[j][2], sl2, b1));
      s[j][3] = ex2(fmaf(s[j][3], sl2, b1));
      p0 += s[j][0] + s[j][1];
      p1 += s[j][2] + s[j][3];
    }
    p0 += __shfl_xor_sync(0xffffffffu, p0, 1);
    p0 += __shfl_xor_sync(0xffffffffu, p0, 2);
    p1 += __shfl_xor_sync(0xffffffffu, p1, 1);
    p1 += __shfl_xor_sync(0xffffffffu, p1, 2);
    const float i0 = 1.f / p0, i1 = 1.f / p1;
    float D0 = 0.f, D1 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] *= i0; s[j][1] *= i0; s[j][2] *= i1; s[j][3] *= i1;
      D0 += s[j][0] * dp[j][0] + s[j][1] * dp[j][1];
      D1 += s[j][2] * dp[j][2] + s[j][3] * dp[j][3];
    }
    D0 += __shfl_xor_sync(0xffffffffu, D0, 1);
    D0 += __shfl_xor_sync(0xffffffffu, D0, 2);
    D1 += __shfl_xor_sync(0xffffffffu, D1, 1);
    D1 += __shfl_xor_sync(0xffffffffu, D1, 2);
    if ((lane & 3) == 0) {
      // log2-domain log-sum-exp of the row (infinite for rows beyond nq: their P^T column is zero in part 2)
      Lsm[r0] = r0 < w.nq ? t0 * sl2 + __log2f(p0) : INFINITY;
      Lsm[r0 + 8] = r0 + 8 < w.nq ? t1 * sl2 + __log2f(p1) : INFINITY;
      Dsm[r0] = D0;
      Dsm[r0 + 8] = D1;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] *= (dp[j][0] - D0) * g.scale;
      s[j][1] *= (dp[j][1] - D0) * g.scale;
      s[j][2] *= (dp[j][2] - D1) * g.scale;
      s[j][3] *= (dp[j][3] - D1) * g.scale;
    }
    float dq[CH][4];
#pragma unroll
    for (int j = 0; j < CH; ++j)
#pragma unroll
      for (int t = 0; t < 4; ++t) dq[j][t] = 0.f;
    mm_p_b<HDP>(dq, s, Ks, lane);
#pragma unroll
    for (int half = 0; half < 2; ++half)
#pragma unroll
      for (int j = 0; j < CH; ++j)
        *reinterpret_cast<__nv_bfloat162*>(Ts + (r0 + half * 8) * LD + j * 8 + 2 * (lane & 3)) =
            __floats2bfloat162_rn(dq[j][half * 2], dq[j][half * 2 + 1]);
  }
  __syncthreads();
  // ---- part 2: this warp's 16 keys
  {
    float st[8][4], dpt[8][4];
    mm_ab_t<HDP>(st, Ks, Qs, warp, lane);
    mm_ab_t<HDP>(dpt, Vs, dOs, warp, lane);
    float dk[CH][4], dv[CH][4];
#pragma unroll
    for (int j = 0; j < CH; ++j)
#pragma unroll
      for (int t = 0; t < 4; ++t) { dk[j][t] = 0.f; dv[j][t] = 0.f; }
    const bool ka = r0 < w.n_real, kb = r0 + 8 < w.n_real;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = j * 8 + 2 * (lane & 3);
      const float La = Lsm[c], Lb = Lsm[c + 1];
      st[j][0] = ka ? ex2(st[j][0] * sl2 - La) : 0.f;
      st[j][1] = ka ? ex2(st[j][1] * sl2 - Lb) : 0.f;
      st[j][2] = kb ? ex2(st[j][2] * sl2 - La) : 0.f;
      st[j][3] = kb ? ex2(st[j][3] * sl2 - Lb) : 0.f;
    }
    mm_p_b<HDP>(dv, st, dOs, lane);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = j * 8 + 2 * (lane & 3);
      const float Da = Dsm[c], Db = Dsm[c + 1];
      st[j][0] *= (dpt[j][0] - Da) * g.scale;
      st[j][1] *= (dpt[j][1] - Db) * g.scale;
      st[j][2] *= (dpt[j][2] - Da) * g.scale;
      st[j][3] *= (dpt[j][3] - Db) * g.scale;
    }
    mm_p_b<HDP>(dk, st, Qs, lane);
    // this warp's rows of the K / V tiles are read by this warp only: they become the dK / dV staging rows
    __syncwarp();
#pragma unroll
    for (int half = 0; half < 2; ++half)
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        const int o = (r0 + half * 8) * LD + j * 8 + 2 * (lane & 3);
        *reinterpret_cast<__nv_bfloat162*>(Ks + o) = __floats2bfloat162_rn(dk[j][half * 2], dk[j][half * 2 + 1]);
        *reinterpret_cast<__nv_bfloat162*>(Vs + o) = __floats2bfloat162_rn(dv[j][half * 2], dv[j][half * 2 + 1]);
      }
  }
  __syncthreads();
  // ---- rows out: two threads per row, 16-byte stores
  const int C = g.nh * g.hd;
  const long long row3 = 3LL * C;
  const int r = threadIdx.x >> 1, par = threadIdx.x & 1;
  if (r < w.n_real) {                                           // (the virtual pad key is the frozen bias: no gradient)
    const int ty = r / w.rw, tx = r - ty * w.rw;
    bf16* dst = dqkv + (((long long)w.b * g.H + (w.wy * g.wh + ty)) * g.W + (w.wx * g.ww + tx)) * row3 + head * g.hd;
#pragma unroll
    for (int c = par; c < CH; c += 2) {
      if (c * 8 >= g.hd) continue;
      *reinterpret_cast<uint4*>(dst + C + c * 8) = *reinterpret_cast<const uint4*>(Ks + r * LD + c * 8);
      *reinterpret_cast<uint4*>(dst + 2 * C + c * 8) = *reinterpret_cast<const uint4*>(Vs + r * LD + c * 8);
    }
    if (!g.pool) {
#pragma unroll
      for (int c = par; c < CH; c += 2)
        if (c * 8 < g.hd) *reinterpret_cast<uint4*>(dst + c * 8) = *reinterpret_cast<const uint4*>(Ts + r * LD + c * 8);
    } else {
      // dq of raw token r: the pooled query's gradient where this token is the arg-max of its 2x2 group (first
      // maximum in scan order, like ATen's max_pool2d backward), zero elsewhere
      const int py = ty >> 1, px = tx >> 1, me = (ty & 1) * 2 + (tx & 1);
      const bf16* g0 = Qraw + ((2 * py) * w.rw + 2 * px) * LD;
      const bf16* src[4] = {g0, g0 + LD, g0 + w.rw * LD, g0 + w.rw * LD + LD};
      const bf16* dqr = Ts + (py * w.qrw + px) * LD;
#pragma unroll
      for (int c = par; c < CH; c += 2) {
        if (c * 8 >= g.hd) continue;
        uint4 v[4], o;
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = *reinterpret_cast<const uint4*>(src[k] + c * 8);
        const uint4 dv4 = *reinterpret_cast<const uint4*>(dqr + c * 8);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          int best = 0;
          float vb = __bfloat162float(reinterpret_cast<const bf16*>(&v[0])[e]);
#pragma unroll
          for (int k = 1; k < 4; ++k) {
            const float vk = __bfloat162float(reinterpret_cast<const bf16*>(&v[k])[e]);
            if (vk > vb) { vb = vk; best = k; }
          }
          reinterpret_cast<bf16*>(&o)[e] = best == me ? reinterpret_cast<const bf16*>(&dv4)[e] : __float2bfloat16(0.f);
        }
        *reinterpret_cast<uint4*>(dst + c * 8) = o;
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------------------
// Packed variant for tiny windows (window area <= 16 tokens: Hiera stage 2 and its transition block): one CTA serves
// FOUR windows of one head, warp w <-> window w (a warp's 16 MMA rows are exactly one window's query slot), so the
// score tile of a window is a single 16x16 MMA block and nothing is wasted on cross-window pairs.  The backward is ONE
// kernel: with a single key tile the softmax is recomputed from scratch (no lse / D workspace; D_r = sum_j P_rj dP_rj)
// and P^T / dS^T reach the tensor core through a 16x16 shared-memory transpose per warp.
// ------------------------------------------------------------------------------------------------------------
constexpr int SLOT = 16, PACK = BM / SLOT;

__device__ __forceinline__ void packed_wins(Win* wins, const Geom& g) {
  if (threadIdx.x < PACK) {
    const int id = blockIdx.y * PACK + threadIdx.x;
    if (id < g.B * g.nwy * g.nwx) {
      wins[threadIdx.x] = make_win(g, id);
    } else {
      Win z;
      z.b = z.wy = z.wx = 0; z.rh = z.rw = 1; z.n_real = z.n_pad = z.nk = 0; z.qrh = z.qrw = 1; z.nq = 0; z.bonus = 0.f;
      wins[threadIdx.x] = z;
    }
  }
}
// kind 0: q, 1: k, 2: v, 3: tensor `src` at the output tokens (dO)
template <int HDP>
__device__ __forceinline__ void load_packed(bf16* dst, int kind, const bf16* qkv, const float* bias, const bf16* src,
                                            const Geom& g, const Win* wins, int head) {
  constexpr int LD = HDP + 8, CH = HDP / 8;
  for (int e = threadIdx.x; e < BM * CH; e += NT) {
    const int r = e / CH, c = e - r * CH;
    const Win& w = wins[r / SLOT];
    const int idx = r % SLOT;
    bf16* d = dst + r * LD + c * 8;
    if (kind == 0) put_q(d, qkv, bias, g, w, head, idx, c);
    else if (kind == 3) put_o(d, src, g, w, head, idx, c);
    else put_kv(d, qkv, bias, g, w, head, kind, idx, c);
  }
}
// acc[2][4] = A(rows 16w..16w+15 of As) . B(rows 16w..16w+15 of Bs)^T       (one 16x16 block)
template <int HDP>
__device__ __forceinline__ void mm_16x16(float (*acc)[4], const bf16* As, const bf16* Bs, int warp, int lane) {
  constexpr int LD = HDP + 8;
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int t = 0; t < 4; ++t) acc[j][t] = 0.f;
#pragma unroll
  for (int kk = 0; kk < HDP / 16; ++kk) {
    uint32_t a[4], b0, b1, b2, b3;
    ldsm4(smem_u32(As + (warp * 16 + (lane & 15)) * LD + kk * 16 + (lane >> 4) * 8), a[0], a[1], a[2], a[3]);
    ldsm4(smem_u32(Bs + (warp * 16 + (lane & 7) + (lane >> 4) * 8) * LD + kk * 16 + ((lane >> 3) & 1) * 8), b0, b1, b2,
          b3);
    mma16816(acc[0], a, b0, b1);
    mma16816(acc[1], a, b2, b3);
  }
}
// out[j][4] (j over HDP/8) += A(16x16 fragment a[4]) . B(rows 16w..16w+15 of Bs, [16 x HDP])
template <int HDP>
__device__ __forceinline__ void mm_16xd(float (*out)[4], const uint32_t* a, const bf16* Bs, int warp, int lane) {
  constexpr int LD = HDP + 8;
#pragma unroll
  for (int dp = 0; dp < HDP / 16; ++dp) {
    uint32_t b0, b1, b2, b3;
    ldsm4t(smem_u32(Bs + (warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + dp * 16 + (lane >> 4) * 8), b0, b1, b2,
           b3);
    mma16816(out[2 * dp], a, b0, b1);
    mma16816(out[2 * dp + 1], a, b2, b3);
  }
}
// masked softmax of a warp's 16x16 score block (raw dot products in s); rows lane/4 and lane/4+8; returns P in s and
// the per-row (max * scale, sum) needed for the lse
__device__ __forceinline__ void softmax16(float (*s)[4], const Win& w, float sl2, int lane, float* mx, float* sum) {
  float t0 = -INFINITY, t1 = -INFINITY;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int c = j * 8 + 2 * (lane & 3) + e;
      if (c == w.n_real) { s[j][e] += w.bonus; s[j][2 + e] += w.bonus; }
      if (c >= w.nk) { s[j][e] = -INFINITY; s[j][2 + e] = -INFINITY; }
    }
    t0 = fmaxf(t0, fmaxf(s[j][0], s[j][1]));
    t1 = fmaxf(t1, fmaxf(s[j][2], s[j][3]));
  }
  t0 = fmaxf(t0, __shfl_xor_sync(0xffffffffu, t0, 1));
  t0 = fmaxf(t0, __shfl_xor_sync(0xffffffffu, t0, 2));
  t1 = fmaxf(t1, __shfl_xor_sync(0xffffffffu, t1, 1));
  t1 = fmaxf(t1, __shfl_xor_sync(0xffffffffu, t1, 2));
  if (w.nk == 0) { t0 = 0.f; t1 = 0.f; }                     // dead window slot: keep everything finite (p = 0)
  float p0 = 0.f, p1 = 0.f;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    s[j][0] = ex2((s[j][0] - t0) * sl2);
    s[j][1] = ex2((s[j][1] - t0) * sl2);
    s[j][2] = ex2((s[j][2] - t1) * sl2);
    s[j][3] = ex2((s[j][3] - t1) * sl2);
    p0 += s[j][0] + s[j][1];
    p1 += s[j][2] + s[j][3];
  }
  p0 += __shfl_xor_sync(0xffffffffu, p0, 1);
  p0 += __shfl_xor_sync(0xffffffffu, p0, 2);
  p1 += __shfl_xor_sync(0xffffffffu, p1, 1);
  p1 += __shfl_xor_sync(0xffffffffu, p1, 2);
  mx[0] = t0; mx[1] = t1; sum[0] = p0; sum[1] = p1;
}

template <int HDP>
__global__ void __launch_bounds__(NT) fwd_packed_kernel(const bf16* __restrict__ qkv, const float* __restrict__ bias,
                                                       bf16* __restrict__ out, float* __restrict__ lse, Geom g) {
  pdl_sync();
  constexpr int LD = HDP + 8;
  extern __shared__ __align__(16) uint8_t smraw[];
  __shared__ Win wins[PACK];
  bf16* Qs = reinterpret_cast<bf16*>(smraw);
  bf16* Ks = Qs + BM * LD;
  bf16* Vs = Ks + BM * LD;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, head = blockIdx.z;
  packed_wins(wins, g);
  __syncthreads();
  load_packed<HDP>(Qs, 0, qkv, bias, nullptr, g, wins, head);
  load_packed<HDP>(Ks, 1, qkv, bias, nullptr, g, wins, head);
  load_packed<HDP>(Vs, 2, qkv, bias, nullptr, g, wins, head);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  const Win w = wins[warp];
  const float sl2 = g.scale * 1.4426950408889634f;
  float s[2][4], mx[2], sum[2];
  mm_16x16<HDP>(s, Qs, Ks, warp, lane);
  softmax16(s, w, sl2, lane, mx, sum);
  uint32_t a[4] = {pack_bf16(s[0][0], s[0][1]), pack_bf16(s[0][2], s[0][3]), pack_bf16(s[1][0], s[1][1]),
                   pack_bf16(s[1][2], s[1][3])};
  float o[HDP / 8][4];
#pragma unroll
  for (int j = 0; j < HDP / 8; ++j)
#pragma unroll
    for (int t = 0; t < 4; ++t) o[j][t] = 0.f;
  mm_16xd<HDP>(o, a, Vs, warp, lane);
  const int C = g.nh * g.hd;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const long long tok = out_token(g, w, (lane >> 2) + half * 8);
    if (tok < 0) continue;
    const float inv = 1.f / sum[half];
    bf16* orow = out + tok * C + head * g.hd;
#pragma unroll
    for (int j = 0; j < HDP / 8; ++j) {
      const int d = j * 8 + 2 * (lane & 3);
      if (d < g.hd)
        *reinterpret_cast<__nv_bfloat162*>(orow + d) =
            __floats2bfloat162_rn(o[j][half * 2] * inv, o[j][half * 2 + 1] * inv);
    }
    if ((lane & 3) == 0) lse[tok * g.nh + head] = mx[half] * g.scale + __logf(sum[half]);
  }
}

template <int HDP>
__global__ void __launch_bounds__(NT) bwd_packed_kernel(const bf16* __restrict__ qkv, const float* __restrict__ bias,
                                                       const bf16* __restrict__ dout, bf16* __restrict__ dqkv,
                                                       Geom g) {
  pdl_sync();
  constexpr int LD = HDP + 8, TP = 24;                         // TP: pitch of the 16x16 transpose tiles (48 B rows)
  extern __shared__ __align__(16) uint8_t smraw[];
  __shared__ Win wins[PACK];
  bf16* Qs = reinterpret_cast<bf16*>(smraw);
  bf16* Ks = Qs + BM * LD;
  bf16* Vs = Ks + BM * LD;
  bf16* dOs = Vs + BM * LD;
  bf16* Ts = dOs + BM * LD;                                    // [4 warps][2 tiles (P, dS)][16][TP]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, head = blockIdx.z;
  packed_wins(wins, g);
  __syncthreads();
  load_packed<HDP>(Qs, 0, qkv, bias, nullptr, g, wins, head);
  load_packed<HDP>(Ks, 1, qkv, bias, nullptr, g, wins, head);
  load_packed<HDP>(Vs, 2, qkv, bias, nullptr, g, wins, head);
  load_packed<HDP>(dOs, 3, qkv, bias, dout, g, wins, head);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  const Win w = wins[warp];
  const float sl2 = g.scale * 1.4426950408889634f;
  float s[2][4], dp[2][4], mx[2], sum[2];
  mm_16x16<HDP>(s, Qs, Ks, warp, lane);
  mm_16x16<HDP>(dp, dOs, Vs, warp, lane);
  softmax16(s, w, sl2, lane, mx, sum);
  const bool q0ok = (lane >> 2) < w.nq, q1ok = (lane >> 2) + 8 < w.nq;
  const float i0 = q0ok ? 1.f / sum[0] : 0.f, i1 = q1ok ? 1.f / sum[1] : 0.f;   // cropped queries: P = 0
  float d0 = 0.f, d1 = 0.f;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    s[j][0] *= i0; s[j][1] *= i0; s[j][2] *= i1; s[j][3] *= i1;
    d0 += s[j][0] * dp[j][0] + s[j][1] * dp[j][1];
    d1 += s[j][2] * dp[j][2] + s[j][3] * dp[j][3];
  }
  d0 += __shfl_xor_sync(0xffffffffu, d0, 1);
  d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
  d1 += __shfl_xor_sync(0xffffffffu, d1, 1);
  d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
  // P and dS = P o (dP - D) * scale -> bf16 into the per-warp transpose tiles, rows = query, columns = key
  bf16* Pt = Ts + warp * 2 * 16 * TP;
  bf16* St = Pt + 16 * TP;
  float ds[2][4];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    ds[j][0] = s[j][0] * (dp[j][0] - d0) * g.scale;
    ds[j][1] = s[j][1] * (dp[j][1] - d0) * g.scale;
    ds[j][2] = s[j][2] * (dp[j][2] - d1) * g.scale;
    ds[j][3] = s[j][3] * (dp[j][3] - d1) * g.scale;
    const int c = j * 8 + 2 * (lane & 3), r = lane >> 2;
    *reinterpret_cast<__nv_bfloat162*>(Pt + r * TP + c) = __floats2bfloat162_rn(s[j][0], s[j][1]);
    *reinterpret_cast<__nv_bfloat162*>(Pt + (r + 8) * TP + c) = __floats2bfloat162_rn(s[j][2], s[j][3]);
    *reinterpret_cast<__nv_bfloat162*>(St + r * TP + c) = __floats2bfloat162_rn(ds[j][0], ds[j][1]);
    *reinterpret_cast<__nv_bfloat162*>(St + (r + 8) * TP + c) = __floats2bfloat162_rn(ds[j][2], ds[j][3]);
  }
  __syncwarp();
  // dQ = dS . K
  {
    float dq[HDP / 8][4];
#pragma unroll
    for (int j = 0; j < HDP / 8; ++j)
#pragma unroll
      for (int t = 0; t < 4; ++t) dq[j][t] = 0.f;
    uint32_t a[4] = {pack_bf16(ds[0][0], ds[0][1]), pack_bf16(ds[0][2], ds[0][3]), pack_bf16(ds[1][0], ds[1][1]),
                     pack_bf16(ds[1][2], ds[1][3])};
    mm_16xd<HDP>(dq, a, Ks, warp, lane);
    scatter_dq<HDP / 8>(dqkv, qkv, bias, g, w, head, lane >> 2, dq, 0, lane);
    scatter_dq<HDP / 8>(dqkv, qkv, bias, g, w, head, (lane >> 2) + 8, dq, 1, lane);
  }
  // dV = P^T . dO, dK = dS^T . Q : A fragments = transposed 16x16 tiles (ldmatrix.trans)
  float dk[HDP / 8][4], dv[HDP / 8][4];
#pragma unroll
  for (int j = 0; j < HDP / 8; ++j)
#pragma unroll
    for (int t = 0; t < 4; ++t) { dk[j][t] = 0.f; dv[j][t] = 0.f; }
  {
    uint32_t a[4];
    const int trow = (lane & 7) + (lane >> 4) * 8, tcol = ((lane >> 3) & 1) * 8;
    ldsm4t(smem_u32(Pt + trow * TP + tcol), a[0], a[1], a[2], a[3]);
    mm_16xd<HDP>(dv, a, dOs, warp, lane);
    ldsm4t(smem_u32(St + trow * TP + tcol), a[0], a[1], a[2], a[3]);
    mm_16xd<HDP>(dk, a, Qs, warp, lane);
  }
  const int C = g.nh * g.hd;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int idx = (lane >> 2) + half * 8;
    if (idx >= w.n_real) continue;                             // only real keys receive gradients
    const int ty = idx / w.rw, tx = idx - ty * w.rw;
    bf16* dst = dqkv + (((long long)w.b * g.H + w.wy * g.wh + ty) * g.W + w.wx * g.ww + tx) * (3LL * C) + head * g.hd;
#pragma unroll
    for (int j = 0; j < HDP / 8; ++j) {
      const int d = j * 8 + 2 * (lane & 3);
      if (d < g.hd) {
        *reinterpret_cast<__nv_bfloat162*>(dst + C + d) = __floats2bfloat162_rn(dk[j][half * 2], dk[j][half * 2 + 1]);
        *reinterpret_cast<__nv_bfloat162*>(dst + 2 * C + d) = __floats2bfloat162_rn(dv[j][half * 2], dv[j][half * 2 + 1]);
      }
    }
  }
}

static int make_geom(Geom& g, int B, int H, int W, int nh, int hd, int window, int pool) {
  if (B <= 0 || H <= 0 || W <= 0 || nh <= 0 || hd <= 0) return S2U_EINVAL;
  if ((hd & 7) || hd > 96) return S2U_EUNSUPPORTED;
  g.B = B; g.H = H; g.W = W; g.nh = nh; g.hd = hd;
  g.wh = window > 0 ? window : H;
  g.ww = window > 0 ? window : W;
  g.nwy = (H + g.wh - 1) / g.wh;
  g.nwx = (W + g.ww - 1) / g.ww;
  g.pool = pool ? 1 : 0;
  if (pool && ((g.wh & 1) || (g.ww & 1) || (H & 1) || (W & 1))) return S2U_EUNSUPPORTED;
  g.Ho = pool ? H / 2 : H;
  g.Wo = pool ? W / 2 : W;
  g.qh = pool ? g.wh / 2 : g.wh;
  g.qw = pool ? g.ww / 2 : g.ww;
  g.scale = 1.0f / sqrtf((float)hd);
  return 0;
}

template <int HDP>
static int launch_fwd(const bf16* qkv, const float* bias, bf16* out, float* lse, const Geom& g, cudaStream_t st) {
  if (g.wh * g.ww <= SLOT) {                                   // tiny windows: 4 per CTA
    dim3 pgrid(1, ceil_div(g.B * g.nwy * g.nwx, PACK), g.nh);
    const size_t psmem = (size_t)3 * BM * (HDP + 8) * sizeof(bf16);
    S2U_ALLOW_SMEM(fwd_packed_kernel<HDP>);
    S2U_LAUNCH((fwd_packed_kernel<HDP>), pgrid, NT, psmem, st, qkv, bias, out, lse, g);
    S2U_LAUNCH_CHECK();
    return 0;
  }
  if (g.wh * g.ww <= BN) {                                     // one tile per window (pooling: raw q staged and pooled in smem)
    dim3 wgrid(1, g.B * g.nwy * g.nwx, g.nh);
    const size_t wsmem = (size_t)(BM + (g.pool ? 3 : 2) * BN) * (HDP + 8) * sizeof(bf16);
    S2U_ALLOW_SMEM(fwd_w64_kernel<HDP>);
    S2U_LAUNCH((fwd_w64_kernel<HDP>), wgrid, NT, wsmem, st, qkv, bias, out, lse, g);
    S2U_LAUNCH_CHECK();
    return 0;
  }
  dim3 grid(ceil_div(g.qh * g.qw, BM), g.B * g.nwy * g.nwx, g.nh);
  const size_t smem = (size_t)(BM + 4 * BN) * (HDP + 8) * sizeof(bf16);
  S2U_ALLOW_SMEM(fwd_kernel<HDP>);
  S2U_LAUNCH((fwd_kernel<HDP>), grid, NT, smem, st, qkv, bias, out, lse, g);
  S2U_LAUNCH_CHECK();
  return 0;
}

template <int HDP>
static int launch_bwd(const bf16* qkv, const float* bias, const bf16* out, const float* lse, const bf16* dout,
                      bf16* dqkv, float* Dws, const Geom& g, cudaStream_t st) {
  if (g.wh * g.ww <= SLOT) {
    dim3 pgrid(1, ceil_div(g.B * g.nwy * g.nwx, PACK), g.nh);
    const size_t psmem = (size_t)(4 * BM * (HDP + 8) + 4 * 2 * 16 * 24) * sizeof(bf16);
    S2U_ALLOW_SMEM(bwd_packed_kernel<HDP>);
    S2U_LAUNCH((bwd_packed_kernel<HDP>), pgrid, NT, psmem, st, qkv, bias, dout, dqkv, g);
    S2U_LAUNCH_CHECK();
    return 0;
  }
  if (g.wh * g.ww <= BN) {                                     // whole window in one tile: fused single kernel
    dim3 wgrid(1, g.B * g.nwy * g.nwx, g.nh);
    const size_t wsmem = (size_t)((g.pool ? 3 : 2) * BM + 3 * BN) * (HDP + 8) * sizeof(bf16);
    S2U_ALLOW_SMEM(bwd_w64_kernel<HDP>);
    S2U_LAUNCH((bwd_w64_kernel<HDP>), wgrid, NT, wsmem, st, qkv, bias, dout, dqkv, g);
    S2U_LAUNCH_CHECK();
    return 0;
  }
  {
    dim3 grid(ceil_div(g.qh * g.qw, BM), g.B * g.nwy * g.nwx, g.nh);
    const size_t smem = (size_t)(2 * BM + 4 * BN) * (HDP + 8) * sizeof(bf16);
    S2U_ALLOW_SMEM(bwd_dq_kernel<HDP>);
    S2U_LAUNCH((bwd_dq_kernel<HDP>), grid, NT, smem, st, qkv, bias, out, lse, Dws, dout, dqkv, g);
    S2U_LAUNCH_CHECK();
  }
  {
    dim3 grid(ceil_div(g.wh * g.ww, BN), g.B * g.nwy * g.nwx, g.nh);
    const size_t smem = (size_t)(4 * BM + 2 * BN) * (HDP + 8) * sizeof(bf16) + 4 * BM * sizeof(float);
    S2U_ALLOW_SMEM(bwd_dkv_kernel<HDP>);
    S2U_LAUNCH((bwd_dkv_kernel<HDP>), grid, NT, smem, st, qkv, bias, lse, Dws, dout, dqkv, g);
    S2U_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace amma

// entry points used by attention.cu's dispatch (bf16 only)
int s2u_attn_mma_fwd(const void* qkv, const float* bias, void* out, float* lse, int B, int H, int W, int nh, int hd,
                     int window, int pool, cudaStream_t st) {
  amma::Geom g;
  int rc = amma::make_geom(g, B, H, W, nh, hd, window, pool);
  if (rc) return rc;
  const bf16* q = (const bf16*)qkv;
  bf16* o = (bf16*)out;
  if (hd <= 32) return amma::launch_fwd<32>(q, bias, o, lse, g, st);
  if (hd <= 64) return amma::launch_fwd<64>(q, bias, o, lse, g, st);
  if (hd <= 80) return amma::launch_fwd<80>(q, bias, o, lse, g, st);
  return amma::launch_fwd<96>(q, bias, o, lse, g, st);
}

int s2u_attn_mma_bwd(const void* qkv, const float* bias, const void* out, const float* lse, const void* dout,
                     void* dqkv, float* Dws, int B, int H, int W, int nh, int hd, int window, int pool,
                     cudaStream_t st) {
  amma::Geom g;
  int rc = amma::make_geom(g, B, H, W, nh, hd, window, pool);
  if (rc) return rc;
  const bf16 *q = (const bf16*)qkv, *o = (const bf16*)out, *d = (const bf16*)dout;
  bf16* dq = (bf16*)dqkv;
  if (hd <= 32) return amma::launch_bwd<32>(q, bias, o, lse, d, dq, Dws, g, st);
  if (hd <= 64) return amma::launch_bwd<64>(q, bias, o, lse, d, dq, Dws, g, st);
  if (hd <= 80) return amma::launch_bwd<80>(q, bias, o, lse, d, dq, Dws, g, st);
  return amma::launch_bwd<96>(q, bias, o, lse, d, dq, Dws, g, st);
}
