// LayerNorm (eps 1e-6, /root/reference/sam2/modeling/backbones/hieradet.py:99-100,104,120) forward and
// input-gradient.  The trunk's affine parameters are frozen (SAM2UNet.py:146-147) so there is no weight gradient.
// Algorithmic traffic: forward reads x and writes y (2 elements/elem), backward reads dy, x (+ residual gradient,
// + the adapter's saved activation) and writes dx (+ dx2).
//
// Layout: a row is owned by a group of LPR = 8 / 16 / 32 lanes (chosen on the host so that narrow rows still use
// every lane: C = 144 -> 8 lanes x 3 chunks, 4 rows per warp); lane l of the group holds the 8-element chunks
// l, l + LPR, ... (NCH <= 5 of them) in registers.  Every global load of a row is issued before the first
// reduction, so one HBM round trip covers the whole row, and nothing is read twice.
#include "common.cuh"

constexpr int LN_MAX_CHUNKS = 5;             // C <= 5 * 32 * 8 = 1280 (Hiera-L stage 4: 1152)
constexpr int LN_NREP = 32;                  // replicated column-sum accumulators (see ln_bwd_kernel)

template <typename T> struct Raw8;
template <> struct Raw8<float> { float4 a, b; };
template <> struct Raw8<bf16> { uint4 u; };
__device__ __forceinline__ Raw8<float> ldraw(const float* p) {
  Raw8<float> r;
  r.a = *reinterpret_cast<const float4*>(p);
  r.b = *reinterpret_cast<const float4*>(p + 4);
  return r;
}
__device__ __forceinline__ Raw8<bf16> ldraw(const bf16* p) {
  Raw8<bf16> r;
  r.u = *reinterpret_cast<const uint4*>(p);
  return r;
}
__device__ __forceinline__ F8 cvt(const Raw8<float>& r) {
  F8 o;
  o.v[0] = r.a.x; o.v[1] = r.a.y; o.v[2] = r.a.z; o.v[3] = r.a.w; o.v[4] = r.b.x; o.v[5] = r.b.y; o.v[6] = r.b.z; o.v[7] = r.b.w;
  return o;
}
__device__ __forceinline__ F8 cvt(const Raw8<bf16>& r) {
  F8 o;
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r.u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    o.v[2 * i] = f.x; o.v[2 * i + 1] = f.y;
  }
  return o;
}
__device__ __forceinline__ float group_sum(float v, int lpr) {
  for (int o = lpr >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// TX = storage type of the normalised tensor (the residual stream: fp32 in both modes), T = compute dtype of y
template <typename T, typename TX, int NCH>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const TX* __restrict__ x, const float* __restrict__ gamma,
                                                    const float* __restrict__ beta, T* __restrict__ y,
                                                    float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                    long long R, int C, float eps, int lpr) {
  pdl_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rpw = 32 / lpr, sub = lane & (lpr - 1);
  const long long row = ((long long)blockIdx.x * 8 + warp) * rpw + lane / lpr;
  const bool valid = row < R;
  const int nv = C >> 3;
  const TX* xr = x + row * C;
  F8 v[NCH];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const int c = sub + k * lpr;
    if (valid && c < nv) {
      v[k] = ld8(xr + c * 8);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[k].v[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[k].v[j];
  }
  const float mean = group_sum(s, lpr) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    if (sub + k * lpr < nv) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[k].v[j] - mean;
        q += d * d;
      }
    }
  }
  const float rstd = rsqrtf(group_sum(q, lpr) / (float)C + eps);
  if (!valid) return;
  T* yr = y + row * C;
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const int c = sub + k * lpr;
    if (c < nv) {
      const F8 g = ld8(gamma + c * 8), b = ld8(beta + c * 8);
      F8 o;
#pragma unroll
      for (int j = 0; j < 8; ++j) o.v[j] = (v[k].v[j] - mean) * rstd * g.v[j] + b.v[j];
      st8(yr + c * 8, o);
    }
  }
  if (sub == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)) + dres,  g = dy * gamma,  xhat = (x - mean) * rstd.
// Optional fused tail of the adapter backward (SAM2UNet.py:57-63): dx2 = dx * gelu'(h) and colsum += sum_rows dx2
// (the gradient w.r.t. the adapter's second bias), which saves one elementwise pass and one reduction pass.  `pre`
// holds h (pre_is_grad = 0) or gelu'(h) as saved by the forward GEMM epilogue (pre_is_grad = 1).
// Column sums: per block through a shared [rows][C] tile, then ONE atomic per column into replica (block % LN_NREP)
// of the workspace `ws` - many hundred blocks adding into the same C addresses serialise in L2 (measured: +15 us on a
// 12 us kernel); ln_colsum_fold_kernel then folds the replicas into `colsum`.
template <typename T, typename TX, int NCH>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const T* __restrict__ dy, const TX* __restrict__ x,
                                                    const float* __restrict__ gamma, const float* __restrict__ mean,
                                                    const float* __restrict__ rstd, const T* __restrict__ dres,
                                                    T* __restrict__ dx, const T* __restrict__ pre,
                                                    T* __restrict__ dx2, float* __restrict__ colsum,
                                                    float* __restrict__ ws, int pre_is_grad, long long R, int C,
                                                    int lpr) {
  pdl_sync();
  extern __shared__ float tile[];            // [8 * rpw rows][C] dx2 values of this block (only when colsum != null)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rpw = 32 / lpr, sub = lane & (lpr - 1);
  const int slot = warp * rpw + lane / lpr;                  // row slot inside the block
  const long long row = (long long)blockIdx.x * (8 * rpw) + slot;
  const bool valid = row < R;
  const int nv = C >> 3;
  Raw8<TX> xr[NCH];
  Raw8<T> dr[NCH], rr[NCH], pr[NCH];
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const int c = sub + k * lpr;
    if (valid && c < nv) {
      const long long off = row * C + c * 8;
      xr[k] = ldraw(x + off);
      dr[k] = ldraw(dy + off);
      if (dres) rr[k] = ldraw(dres + off);
      if (dx2) pr[k] = ldraw(pre + off);
    }
  }
  float mu = 0.f, rs = 0.f;
  if (valid) {
    mu = mean[row];
    rs = rstd[row];
  }
  F8 xh[NCH], g[NCH];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const int c = sub + k * lpr;
    if (valid && c < nv) {
      const F8 xv = cvt(xr[k]), dv = cvt(dr[k]), gm = ld8(gamma + c * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        xh[k].v[j] = (xv.v[j] - mu) * rs;
        g[k].v[j] = dv.v[j] * gm.v[j];
        s1 += g[k].v[j];
        s2 += g[k].v[j] * xh[k].v[j];
      }
    }
  }
  s1 = group_sum(s1, lpr) / (float)C;
  s2 = group_sum(s2, lpr) / (float)C;
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const int c = sub + k * lpr;
    if (c >= nv) continue;
    F8 o2;
#pragma unroll
    for (int j = 0; j < 8; ++j) o2.v[j] = 0.f;
    if (valid) {
      const long long off = row * C + c * 8;
      F8 o;
#pragma unroll
      for (int j = 0; j < 8; ++j) o.v[j] = rs * (g[k].v[j] - s1 - xh[k].v[j] * s2);
      if (dres) {
        const F8 r = cvt(rr[k]);
#pragma unroll
        for (int j = 0; j < 8; ++j) o.v[j] += r.v[j];
      }
      st8(dx + off, o);
      if (dx2) {
        const F8 p = cvt(pr[k]);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          o2.v[j] = rnd(rnd(o.v[j], (const T*)nullptr) * (pre_is_grad ? p.v[j] : dgelu_f(p.v[j])), (const T*)nullptr);
        st8(dx2 + off, o2);
      }
    }
    if (colsum) {
      float* t = tile + slot * C + c * 8;
      *reinterpret_cast<float4*>(t) = make_float4(o2.v[0], o2.v[1], o2.v[2], o2.v[3]);
      *reinterpret_cast<float4*>(t + 4) = make_float4(o2.v[4], o2.v[5], o2.v[6], o2.v[7]);
    }
  }
  if (!colsum) return;
  __syncthreads();
  const int rows = 8 * rpw;
  float* rep = ws + (size_t)(blockIdx.x % LN_NREP) * C;
  for (int j = threadIdx.x; j < C; j += 256) {
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += tile[r * C + j];
    atomicAdd(rep + j, s);
  }
}

// folds the replicated column sums into the gradient and leaves the workspace zeroed for the next call.  A separate
// (tiny) launch on purpose: a "last block done" ticket inside ln_bwd_kernel needs a __threadfence and an atomic round
// trip at the end of every block, which doubled the lifetime of these short blocks (measured 35 us instead of 17).
__global__ void __launch_bounds__(256) ln_colsum_fold_kernel(float* __restrict__ ws, float* __restrict__ colsum, int C) {
  pdl_sync();
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= C) return;
  float v[LN_NREP];
#pragma unroll
  for (int r = 0; r < LN_NREP; ++r) v[r] = ws[(size_t)r * C + j];
  float s = 0.f;
#pragma unroll
  for (int r = 0; r < LN_NREP; ++r) {
    s += v[r];
    ws[(size_t)r * C + j] = 0.f;
  }
  colsum[j] += s;
}

// lanes per row: the narrowest group that keeps a row within 3 chunks per lane, else within LN_MAX_CHUNKS
static int ln_lanes(int nv, int* nch) {
  for (int lim : {3, LN_MAX_CHUNKS})
    for (int lpr : {8, 16, 32})
      if (nv <= lpr * lim) {
        *nch = (nv + lpr - 1) / lpr;
        return lpr;
      }
  *nch = 0;
  return 0;
}

#define LN_DISPATCH_NCH(nch, ...)                    \
  switch (nch) {                                     \
    case 1: { constexpr int NCH = 1; __VA_ARGS__ } break; \
    case 2: { constexpr int NCH = 2; __VA_ARGS__ } break; \
    case 3: { constexpr int NCH = 3; __VA_ARGS__ } break; \
    case 4: { constexpr int NCH = 4; __VA_ARGS__ } break; \
    case 5: { constexpr int NCH = 5; __VA_ARGS__ } break; \
    default: return S2U_EUNSUPPORTED;                \
  }

extern "C" {

int s2u_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                      long long R, int C, float eps, int x_f32, int dtype, void* stream) {
  if (R <= 0 || C <= 0 || (C & 7)) return S2U_EINVAL;
  int nch = 0;
  const int lpr = ln_lanes(C >> 3, &nch);
  if (!lpr) return S2U_EUNSUPPORTED;
  const int grid = ceil_div(R, 8 * (32 / lpr));
  cudaStream_t st = (cudaStream_t)stream;
  if (x_f32 || dtype == S2U_F32) {
    S2U_DISPATCH_T(dtype, {
      LN_DISPATCH_NCH(nch, {
        S2U_LAUNCH((ln_fwd_kernel<T, float, NCH>), grid, 256, 0, st, (const float*)x, gamma, beta, (T*)y, mean, rstd, R, C, eps,
                                                          lpr);
      })
    })
  } else {
    LN_DISPATCH_NCH(nch, {
      S2U_LAUNCH((ln_fwd_kernel<bf16, bf16, NCH>), grid, 256, 0, st, (const bf16*)x, gamma, beta, (bf16*)y, mean, rstd, R, C, eps,
                                                          lpr);
    })
  }
  S2U_LAUNCH_CHECK();
  return 0;
}

// ws: fp32 workspace of s2u_layernorm_ws_floats(C) elements, zero before the first use (the kernel leaves it zeroed);
// required when colsum is given, one per stream that may run this concurrently
int s2u_layernorm_ws_floats(int C) { return LN_NREP * C; }

int s2u_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                      const void* dres, void* dx, const void* pre, void* dx2, float* colsum, float* ws,
                      int pre_is_grad, long long R, int C, int x_f32, int dtype, void* stream) {
  if (R <= 0 || C <= 0 || (C & 7) || (dx2 && !pre) || (colsum && (!dx2 || !ws))) return S2U_EINVAL;
  int nch = 0;
  const int lpr = ln_lanes(C >> 3, &nch);
  if (!lpr) return S2U_EUNSUPPORTED;
  const int rows = 8 * (32 / lpr);
  const size_t smem = colsum ? (size_t)rows * C * sizeof(float) : 0;
  if (smem > 48 * 1024) return S2U_EUNSUPPORTED;
  const int grid = ceil_div(R, rows);
  cudaStream_t st = (cudaStream_t)stream;
  if (x_f32 || dtype == S2U_F32) {
    S2U_DISPATCH_T(dtype, {
      LN_DISPATCH_NCH(nch, {
        S2U_LAUNCH((ln_bwd_kernel<T, float, NCH>), grid, 256, smem, st, (const T*)dy, (const float*)x, gamma, mean, rstd,
                                                             (const T*)dres, (T*)dx, (const T*)pre, (T*)dx2, colsum,
                                                             ws, pre_is_grad, R, C, lpr);
      })
    })
  } else {
    LN_DISPATCH_NCH(nch, {
      S2U_LAUNCH((ln_bwd_kernel<bf16, bf16, NCH>), grid, 256, smem, st, (const bf16*)dy, (const bf16*)x, gamma, mean, rstd,
                                                             (const bf16*)dres, (bf16*)dx, (const bf16*)pre,
                                                             (bf16*)dx2, colsum, ws, pre_is_grad, R, C, lpr);
    })
  }
  S2U_LAUNCH_CHECK();
  if (colsum) {
    S2U_LAUNCH((ln_colsum_fold_kernel), ceil_div(C, 256), 256, 0, st, ws, colsum, C);
    S2U_LAUNCH_CHECK();
  }
  return 0;
}

}  // extern "C"
