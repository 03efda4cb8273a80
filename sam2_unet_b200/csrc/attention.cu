// Windowed / global multi-head attention of the Hiera trunk, forward and backward, with the window
// partition, the zero padding, the q max-pool and the un-partition + crop folded into the addressing.
//
// Reference semantics (/root/reference/sam2/modeling/backbones/hieradet.py:56-81,141-162 and
// backbones/utils.py:16-55): tokens are zero-padded AFTER norm1 and BEFORE the qkv projection, so a padded
// token carries q = k = v = bias and takes part in the softmax as a key; queries at padded positions are
// computed by the reference and then cropped away.  Here the qkv GEMM runs on real tokens only; this kernel
// substitutes the bias for padded positions, and never materialises a window-major copy.
//
// Layouts: qkv [B,H,W,3*nh*hd] (q | k | v, head-major inside each third, hieradet.py:59-61);
//          out [B,Ho,Wo,nh*hd] in un-windowed NHWC token order; lse [B,Ho,Wo,nh] fp32.
// This file is the exact-arithmetic SIMT implementation (fp32 math, any head dim that is a multiple of 4,
// any window size); it is the reference point for the tensor-core variant.
#include "common.cuh"

struct AttnGeom {
  int B, H, W;        // token grid
  int nh, hd;         // heads, head dim
  int wh, ww;         // window extent in tokens (== H, W for global attention)
  int nwy, nwx;       // windows per image
  int pool;           // 1: q is 2x2 max-pooled inside each window (hieradet.py:64-67)
  int Ho, Wo;         // output token grid (H, W or H/2, W/2)
  int qh, qw;         // query extent of a window (wh, ww or wh/2, ww/2)
  float scale;        // 1/sqrt(hd)
};

constexpr int QT = 32, KT = 32, NTH = 128, DMAX = 24;   // DMAX*4 = 96 = largest head dim (Hiera-T/S)

template <typename T>
__device__ __forceinline__ float tok_val(const T* __restrict__ qkv, const float* __restrict__ bias,
                                         const AttnGeom& g, int b, int y, int x, int col) {
  // value of column `col` of the qkv row at token (y, x); padded tokens hold the bias (rounded like a GEMM output)
  if (y < g.H && x < g.W) return ldf(qkv + (((long long)b * g.H + y) * g.W + x) * (3LL * g.nh * g.hd) + col);
  return rnd(bias[col], (const T*)nullptr);
}

// K or V tile: rows = window positions k0..k0+KT-1 (row-major inside the window)
template <typename T>
__device__ __forceinline__ void load_kv_tile(float* dst, int ldd, const T* qkv, const float* bias, const AttnGeom& g,
                                             int b, int wy, int wx, int head, int which, int k0) {
  const int C = g.nh * g.hd;
  for (int e = threadIdx.x; e < KT * g.hd; e += NTH) {
    const int kk = e / g.hd, d = e - kk * g.hd;
    const int idx = k0 + kk;
    float v = 0.f;
    if (idx < g.wh * g.ww) {
      const int ty = idx / g.ww, tx = idx - ty * g.ww;
      v = tok_val(qkv, bias, g, b, wy * g.wh + ty, wx * g.ww + tx, which * C + head * g.hd + d);
    }
    dst[kk * ldd + d] = v;
  }
}

// Q tile: rows = query positions q0..q0+QT-1 of the window (pooled grid when g.pool)
template <typename T>
__device__ __forceinline__ void load_q_tile(float* dst, int ldd, const T* qkv, const float* bias, const AttnGeom& g,
                                            int b, int wy, int wx, int head, int q0) {
  for (int e = threadIdx.x; e < QT * g.hd; e += NTH) {
    const int qq = e / g.hd, d = e - qq * g.hd;
    const int idx = q0 + qq;
    float v = 0.f;
    if (idx < g.qh * g.qw) {
      const int py = idx / g.qw, px = idx - py * g.qw;
      const int col = head * g.hd + d;
      if (!g.pool) {
        v = tok_val(qkv, bias, g, b, wy * g.wh + py, wx * g.ww + px, col);
      } else {
        const int y = wy * g.wh + 2 * py, x = wx * g.ww + 2 * px;
        v = fmaxf(fmaxf(tok_val(qkv, bias, g, b, y, x, col), tok_val(qkv, bias, g, b, y, x + 1, col)),
                  fmaxf(tok_val(qkv, bias, g, b, y + 1, x, col), tok_val(qkv, bias, g, b, y + 1, x + 1, col)));
      }
    }
    dst[qq * ldd + d] = v;
  }
}

// output token of query `idx` of window (wy, wx); returns false when it is cropped away
__device__ __forceinline__ bool out_pos(const AttnGeom& g, int wy, int wx, int idx, int& oy, int& ox) {
  if (idx >= g.qh * g.qw) return false;
  const int py = idx / g.qw, px = idx - py * g.qw;
  oy = wy * g.qh + py;
  ox = wx * g.qw + px;
  return oy < g.Ho && ox < g.Wo;
}

template <typename T>
__global__ void __launch_bounds__(NTH) attn_fwd_kernel(const T* __restrict__ qkv, const float* __restrict__ bias,
                                                      T* __restrict__ out, float* __restrict__ lse, AttnGeom g) {
  pdl_sync();
  extern __shared__ float sm[];
  const int ldd = g.hd + 1;
  float* Qs = sm;
  float* Ks = Qs + QT * ldd;
  float* Vs = Ks + KT * ldd;
  float* Ps = Vs + KT * ldd;              // [QT][KT+1]
  const int head = blockIdx.z;
  const int win = blockIdx.y;
  const int b = win / (g.nwy * g.nwx);
  const int wy = (win / g.nwx) % g.nwy, wx = win % g.nwx;
  const int q0 = blockIdx.x * QT;
  const int qi = threadIdx.x >> 2, sub = threadIdx.x & 3;
  const int nd = (g.hd + 3) >> 2;

  load_q_tile(Qs, ldd, qkv, bias, g, b, wy, wx, head, q0);
  float o[DMAX];
#pragma unroll
  for (int i = 0; i < DMAX; ++i) o[i] = 0.f;
  float m = -INFINITY, l = 0.f;
  const int nk = g.wh * g.ww;
  for (int k0 = 0; k0 < nk; k0 += KT) {
    __syncthreads();
    load_kv_tile(Ks, ldd, qkv, bias, g, b, wy, wx, head, 1, k0);
    load_kv_tile(Vs, ldd, qkv, bias, g, b, wy, wx, head, 2, k0);
    __syncthreads();
    float s[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) s[jj] = 0.f;
    for (int d = 0; d < g.hd; ++d) {
      const float qv = Qs[qi * ldd + d];
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) s[jj] = fmaf(qv, Ks[(sub + 4 * jj) * ldd + d], s[jj]);
    }
    float tmax = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      s[jj] = (k0 + sub + 4 * jj < nk) ? s[jj] * g.scale : -INFINITY;
      tmax = fmaxf(tmax, s[jj]);
    }
    tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 1));
    tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 2));
    const float mnew = fmaxf(m, tmax);
    const float alpha = __expf(m - mnew);       // m = -inf on the first tile -> 0
    float psum = 0.f;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const float p = __expf(s[jj] - mnew);
      psum += p;
      Ps[qi * (KT + 1) + sub + 4 * jj] = p;
    }
    psum += __shfl_xor_sync(0xffffffffu, psum, 1);
    psum += __shfl_xor_sync(0xffffffffu, psum, 2);
    l = l * alpha + psum;
    m = mnew;
    __syncwarp();
#pragma unroll
    for (int i = 0; i < DMAX; ++i) {
      if (i < nd) {
        const int d = sub + 4 * i;
        float acc = o[i] * alpha;
        if (d < g.hd) {
#pragma unroll 8
          for (int j = 0; j < KT; ++j) acc = fmaf(Ps[qi * (KT + 1) + j], Vs[j * ldd + d], acc);
        }
        o[i] = acc;
      }
    }
  }
  int oy, ox;
  if (!out_pos(g, wy, wx, q0 + qi, oy, ox)) return;
  const long long tok = ((long long)b * g.Ho + oy) * g.Wo + ox;
  const float inv = 1.f / l;
  T* orow = out + tok * (g.nh * g.hd) + head * g.hd;
#pragma unroll
  for (int i = 0; i < DMAX; ++i) {
    const int d = sub + 4 * i;
    if (i < nd && d < g.hd) stf(orow + d, o[i] * inv);
  }
  if (sub == 0) lse[tok * g.nh + head] = m + __logf(l);
}

// ---- backward, part 1: dQ (and the q max-pool routing).  One CTA = 32 queries of one (window, head).
template <typename T>
__global__ void __launch_bounds__(NTH) attn_bwd_dq_kernel(const T* __restrict__ qkv, const float* __restrict__ bias,
                                                         const T* __restrict__ out, const float* __restrict__ lse,
                                                         const T* __restrict__ dout, T* __restrict__ dqkv,
                                                         AttnGeom g) {
  pdl_sync();
  extern __shared__ float sm[];
  const int ldd = g.hd + 1;
  float* Qs = sm;
  float* dOs = Qs + QT * ldd;
  float* Ks = dOs + QT * ldd;
  float* Vs = Ks + KT * ldd;
  float* Ss = Vs + KT * ldd;              // dS tile [QT][KT+1]
  const int head = blockIdx.z;
  const int win = blockIdx.y;
  const int b = win / (g.nwy * g.nwx);
  const int wy = (win / g.nwx) % g.nwy, wx = win % g.nwx;
  const int q0 = blockIdx.x * QT;
  const int qi = threadIdx.x >> 2, sub = threadIdx.x & 3;
  const int nd = (g.hd + 3) >> 2;
  const int C = g.nh * g.hd;

  load_q_tile(Qs, ldd, qkv, bias, g, b, wy, wx, head, q0);
  int oy = 0, ox = 0;
  const bool live = out_pos(g, wy, wx, q0 + qi, oy, ox);
  const long long tok = live ? ((long long)b * g.Ho + oy) * g.Wo + ox : 0;
  // dO tile + D_i = sum_d dO_i[d] * O_i[d]
  float dpart = 0.f;
#pragma unroll
  for (int i = 0; i < DMAX; ++i) {
    const int d = sub + 4 * i;
    if (i < nd && d < g.hd) {
      const float dv = live ? ldf(dout + tok * C + head * g.hd + d) : 0.f;
      const float ov = live ? ldf(out + tok * C + head * g.hd + d) : 0.f;
      dOs[qi * ldd + d] = dv;
      dpart += dv * ov;
    }
  }
  dpart += __shfl_xor_sync(0xffffffffu, dpart, 1);
  dpart += __shfl_xor_sync(0xffffffffu, dpart, 2);
  const float Di = dpart;
  const float li = live ? lse[tok * g.nh + head] : 0.f;
  float dq[DMAX];
#pragma unroll
  for (int i = 0; i < DMAX; ++i) dq[i] = 0.f;
  const int nk = g.wh * g.ww;
  for (int k0 = 0; k0 < nk; k0 += KT) {
    __syncthreads();
    load_kv_tile(Ks, ldd, qkv, bias, g, b, wy, wx, head, 1, k0);
    load_kv_tile(Vs, ldd, qkv, bias, g, b, wy, wx, head, 2, k0);
    __syncthreads();
    float s[8], dp[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) { s[jj] = 0.f; dp[jj] = 0.f; }
    for (int d = 0; d < g.hd; ++d) {
      const float qv = Qs[qi * ldd + d], dv = dOs[qi * ldd + d];
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        s[jj] = fmaf(qv, Ks[(sub + 4 * jj) * ldd + d], s[jj]);
        dp[jj] = fmaf(dv, Vs[(sub + 4 * jj) * ldd + d], dp[jj]);
      }
    }
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const bool ok = live && (k0 + sub + 4 * jj < nk);
      const float p = ok ? __expf(s[jj] * g.scale - li) : 0.f;
      Ss[qi * (KT + 1) + sub + 4 * jj] = p * (dp[jj] - Di) * g.scale;
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < DMAX; ++i) {
      const int d = sub + 4 * i;
      if (i < nd && d < g.hd) {
        float acc = dq[i];
#pragma unroll 8
        for (int j = 0; j < KT; ++j) acc = fmaf(Ss[qi * (KT + 1) + j], Ks[j * ldd + d], acc);
        dq[i] = acc;
      }
    }
  }
  // scatter dq into the q third of dqkv
  const int idx = q0 + qi;
  if (idx >= g.qh * g.qw) return;
  const int py = idx / g.qw, px = idx - py * g.qw;
  const long long row3 = 3LL * C;
  if (!g.pool) {
    const int y = wy * g.wh + py, x = wx * g.ww + px;
    if (y >= g.H || x >= g.W) return;
    T* dst = dqkv + (((long long)b * g.H + y) * g.W + x) * row3 + head * g.hd;
#pragma unroll
    for (int i = 0; i < DMAX; ++i) {
      const int d = sub + 4 * i;
      if (i < nd && d < g.hd) stf(dst + d, dq[i]);
    }
  } else {
    const int y = wy * g.wh + 2 * py, x = wx * g.ww + 2 * px;
#pragma unroll
    for (int i = 0; i < DMAX; ++i) {
      const int d = sub + 4 * i;
      if (i < nd && d < g.hd) {
        const int col = head * g.hd + d;
        float v[4];
        v[0] = tok_val(qkv, bias, g, b, y, x, col);
        v[1] = tok_val(qkv, bias, g, b, y, x + 1, col);
        v[2] = tok_val(qkv, bias, g, b, y + 1, x, col);
        v[3] = tok_val(qkv, bias, g, b, y + 1, x + 1, col);
        int best = 0;
#pragma unroll
        for (int k = 1; k < 4; ++k)
          if (v[k] > v[best]) best = k;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int yy = y + (k >> 1), xx = x + (k & 1);
          if (yy < g.H && xx < g.W)
            stf(dqkv + (((long long)b * g.H + yy) * g.W + xx) * row3 + col, k == best ? dq[i] : 0.f);
        }
      }
    }
  }
}

// ---- backward, part 2: dK, dV.  One CTA = 32 keys of one (window, head); loops over the window's queries.
template <typename T>
__global__ void __launch_bounds__(NTH) attn_bwd_dkv_kernel(const T* __restrict__ qkv, const float* __restrict__ bias,
                                                          const T* __restrict__ out, const float* __restrict__ lse,
                                                          const T* __restrict__ dout, T* __restrict__ dqkv,
                                                          AttnGeom g) {
  pdl_sync();
  extern __shared__ float sm[];
  const int ldd = g.hd + 1;
  float* Ks = sm;
  float* Vs = Ks + KT * ldd;
  float* Qs = Vs + KT * ldd;
  float* dOs = Qs + QT * ldd;
  float* Ps = dOs + QT * ldd;             // [KT][QT+1]
  float* Ss = Ps + KT * (QT + 1);         // [KT][QT+1]
  float* Ls = Ss + KT * (QT + 1);         // [QT] lse
  float* Ds = Ls + QT;                    // [QT] D
  const int head = blockIdx.z;
  const int win = blockIdx.y;
  const int b = win / (g.nwy * g.nwx);
  const int wy = (win / g.nwx) % g.nwy, wx = win % g.nwx;
  const int k0 = blockIdx.x * KT;
  const int kj = threadIdx.x >> 2, sub = threadIdx.x & 3;
  const int nd = (g.hd + 3) >> 2;
  const int C = g.nh * g.hd;
  const int nk = g.wh * g.ww, nq = g.qh * g.qw;

  load_kv_tile(Ks, ldd, qkv, bias, g, b, wy, wx, head, 1, k0);
  load_kv_tile(Vs, ldd, qkv, bias, g, b, wy, wx, head, 2, k0);
  float dk[DMAX], dv[DMAX];
#pragma unroll
  for (int i = 0; i < DMAX; ++i) { dk[i] = 0.f; dv[i] = 0.f; }
  for (int q0 = 0; q0 < nq; q0 += QT) {
    __syncthreads();
    load_q_tile(Qs, ldd, qkv, bias, g, b, wy, wx, head, q0);
    {
      // dO tile, lse and D for the 32 queries (thread layout: query = tid>>2, dims strided by 4)
      const int qi = threadIdx.x >> 2;
      int oy = 0, ox = 0;
      const bool live = out_pos(g, wy, wx, q0 + qi, oy, ox);
      const long long tok = live ? ((long long)b * g.Ho + oy) * g.Wo + ox : 0;
      float dpart = 0.f;
      for (int i = 0; i < nd; ++i) {
        const int d = sub + 4 * i;
        if (d < g.hd) {
          const float dvv = live ? ldf(dout + tok * C + head * g.hd + d) : 0.f;
          const float ov = live ? ldf(out + tok * C + head * g.hd + d) : 0.f;
          dOs[qi * ldd + d] = dvv;
          dpart += dvv * ov;
        }
      }
      dpart += __shfl_xor_sync(0xffffffffu, dpart, 1);
      dpart += __shfl_xor_sync(0xffffffffu, dpart, 2);
      if (sub == 0) {
        Ds[qi] = dpart;
        Ls[qi] = live ? lse[tok * g.nh + head] : INFINITY;    // +inf -> p = 0 for cropped queries
      }
    }
    __syncthreads();
    float s[8], dp[8];
#pragma unroll
    for (int ii = 0; ii < 8; ++ii) { s[ii] = 0.f; dp[ii] = 0.f; }
    for (int d = 0; d < g.hd; ++d) {
      const float kv = Ks[kj * ldd + d], vv = Vs[kj * ldd + d];
#pragma unroll
      for (int ii = 0; ii < 8; ++ii) {
        s[ii] = fmaf(kv, Qs[(sub + 4 * ii) * ldd + d], s[ii]);
        dp[ii] = fmaf(vv, dOs[(sub + 4 * ii) * ldd + d], dp[ii]);
      }
    }
#pragma unroll
    for (int ii = 0; ii < 8; ++ii) {
      const int qi = sub + 4 * ii;
      const bool ok = (q0 + qi < nq) && (k0 + kj < nk);
      const float p = ok ? __expf(s[ii] * g.scale - Ls[qi]) : 0.f;
      Ps[kj * (QT + 1) + qi] = p;
      Ss[kj * (QT + 1) + qi] = p * (dp[ii] - Ds[qi]) * g.scale;
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < DMAX; ++i) {
      const int d = sub + 4 * i;
      if (i < nd && d < g.hd) {
        float av = dv[i], ak = dk[i];
#pragma unroll 8
        for (int q = 0; q < QT; ++q) {
          av = fmaf(Ps[kj * (QT + 1) + q], dOs[q * ldd + d], av);
          ak = fmaf(Ss[kj * (QT + 1) + q], Qs[q * ldd + d], ak);
        }
        dv[i] = av;
        dk[i] = ak;
      }
    }
  }
  const int idx = k0 + kj;
  if (idx >= nk) return;
  const int ty = idx / g.ww, tx = idx - ty * g.ww;
  const int y = wy * g.wh + ty, x = wx * g.ww + tx;
  if (y >= g.H || x >= g.W) return;        // padded key: its k/v are the frozen bias, no gradient
  T* dst = dqkv + (((long long)b * g.H + y) * g.W + x) * (3LL * C) + head * g.hd;
#pragma unroll
  for (int i = 0; i < DMAX; ++i) {
    const int d = sub + 4 * i;
    if (i < nd && d < g.hd) {
      stf(dst + C + d, dk[i]);
      stf(dst + 2 * C + d, dv[i]);
    }
  }
}

static int make_geom(AttnGeom& g, int B, int H, int W, int nh, int hd, int window, int pool) {
  if (B <= 0 || H <= 0 || W <= 0 || nh <= 0 || hd <= 0 || hd > 4 * DMAX) return S2U_EINVAL;
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

// tensor-core variants (attention_mma.cu), bf16 only; return S2U_EUNSUPPORTED for shapes they do not cover
int s2u_attn_mma_fwd(const void* qkv, const float* bias, void* out, float* lse, int B, int H, int W, int nh, int hd,
                     int window, int pool, cudaStream_t st);
int s2u_attn_mma_bwd(const void* qkv, const float* bias, const void* out, const float* lse, const void* dout,
                     void* dqkv, float* Dws, int B, int H, int W, int nh, int hd, int window, int pool,
                     cudaStream_t st);

// tcgen05 / TMEM / TMA variants (attention_tc.cu), bf16 only, same convention
int s2u_attn_tc_fwd(const void* qkv, const float* bias, void* out, float* lse, int B, int H, int W, int nh, int hd,
                    int window, int pool, cudaStream_t st);

int s2u_attn_tc_bwd(const void* qkv, const float* bias, const void* out, const float* lse, const void* dout,
                    void* dqkv, float* dws, int B, int H, int W, int nh, int hd, int window, int pool,
                    cudaStream_t st);

// 0 = auto (tcgen05 kernels where they apply, else mma.sync, else fp32 SIMT), 1 = never tcgen05, 2 = tcgen05 or fail
static int g_attn_backend = -1;
static int attn_backend() {
  if (g_attn_backend < 0) {
    const char* e = getenv("S2U_ATTN_BACKEND");
    g_attn_backend = e ? atoi(e) : 0;
  }
  return g_attn_backend;
}

extern "C" {

int s2u_set_attn_backend(int backend) {
  if (backend < 0 || backend > 2) return S2U_EINVAL;
  g_attn_backend = backend;
  return 0;
}

int s2u_win_attn_fwd(const void* qkv, const float* bias, void* out, float* lse, int B, int H, int W, int nh, int hd,
                     int window, int pool, int dtype, void* stream) {
  if (dtype == S2U_BF16 && attn_backend() != 1) {
    const int rc3 = s2u_attn_tc_fwd(qkv, bias, out, lse, B, H, W, nh, hd, window, pool, (cudaStream_t)stream);
    if (rc3 != S2U_EUNSUPPORTED) return rc3;
    if (attn_backend() == 2) return rc3;
  }
  if (dtype == S2U_BF16) {
    const int rc2 = s2u_attn_mma_fwd(qkv, bias, out, lse, B, H, W, nh, hd, window, pool, (cudaStream_t)stream);
    if (rc2 != S2U_EUNSUPPORTED) return rc2;
  }
  AttnGeom g;
  int rc = make_geom(g, B, H, W, nh, hd, window, pool);
  if (rc) return rc;
  dim3 grid(ceil_div(g.qh * g.qw, QT), B * g.nwy * g.nwx, nh);
  const size_t smem = (size_t)((QT + 2 * KT) * (hd + 1) + QT * (KT + 1)) * sizeof(float);
  S2U_DISPATCH_T(dtype, {
    S2U_ALLOW_SMEM(attn_fwd_kernel<T>);
    S2U_LAUNCH((attn_fwd_kernel<T>), grid, NTH, smem, (cudaStream_t)stream, (const T*)qkv, bias, (T*)out, lse, g);
  })
  S2U_LAUNCH_CHECK();
  return 0;
}

// dqkv is fully overwritten: every real token receives dq (part 1) and dk, dv (part 2).
// dws: fp32 workspace of B*Ho*Wo*nh entries (rowsum(dO o O) of the tensor-core path).
int s2u_win_attn_bwd(const void* qkv, const float* bias, const void* out, const float* lse, const void* dout,
                     void* dqkv, float* dws, int B, int H, int W, int nh, int hd, int window, int pool, int dtype,
                     void* stream) {
  if (dtype == S2U_BF16 && dws && attn_backend() != 1) {
    const int rc3 = s2u_attn_tc_bwd(qkv, bias, out, lse, dout, dqkv, dws, B, H, W, nh, hd, window, pool,
                                    (cudaStream_t)stream);
    if (rc3 != S2U_EUNSUPPORTED) return rc3;
    if (attn_backend() == 2) return rc3;
  }
  if (dtype == S2U_BF16 && dws) {
    const int rc2 = s2u_attn_mma_bwd(qkv, bias, out, lse, dout, dqkv, dws, B, H, W, nh, hd, window, pool,
                                     (cudaStream_t)stream);
    if (rc2 != S2U_EUNSUPPORTED) return rc2;
  }
  AttnGeom g;
  int rc = make_geom(g, B, H, W, nh, hd, window, pool);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  {
    dim3 grid(ceil_div(g.qh * g.qw, QT), B * g.nwy * g.nwx, nh);
    const size_t smem = (size_t)((2 * QT + 2 * KT) * (hd + 1) + QT * (KT + 1)) * sizeof(float);
    S2U_DISPATCH_T(dtype, {
      S2U_ALLOW_SMEM(attn_bwd_dq_kernel<T>);
      S2U_LAUNCH((attn_bwd_dq_kernel<T>), grid, NTH, smem, st, (const T*)qkv, bias, (const T*)out, lse, (const T*)dout,
                                                     (T*)dqkv, g);
    })
    S2U_LAUNCH_CHECK();
  }
  {
    dim3 grid(ceil_div(g.wh * g.ww, KT), B * g.nwy * g.nwx, nh);
    const size_t smem = (size_t)((2 * QT + 2 * KT) * (hd + 1) + 2 * KT * (QT + 1) + 2 * QT) * sizeof(float);
    S2U_DISPATCH_T(dtype, {
      S2U_ALLOW_SMEM(attn_bwd_dkv_kernel<T>);
      S2U_LAUNCH((attn_bwd_dkv_kernel<T>), grid, NTH, smem, st, (const T*)qkv, bias, (const T*)out, lse, (const T*)dout,
                                                      (T*)dqkv, g);
    })
    S2U_LAUNCH_CHECK();
  }
  return 0;
}

}  // extern "C"
