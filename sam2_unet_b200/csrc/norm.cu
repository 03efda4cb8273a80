// LayerNorm (eps 1e-6, /root/reference/sam2/modeling/backbones/hieradet.py:99-100,104,120) forward and
// input-gradient, one warp per token row, 128-bit accesses.  The trunk's affine parameters are frozen
// (SAM2UNet.py:146-147) so there is no weight gradient.  Algorithmic traffic: forward reads x and writes y
// (2 elements/elem), backward reads dy, x (+ residual gradient) and writes dx.
#include "common.cuh"

// TX = storage type of the normalised tensor (the residual stream: fp32 in both modes), T = compute dtype of y
template <typename T, typename TX>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const TX* __restrict__ x, const float* __restrict__ gamma,
                                                    const float* __restrict__ beta, T* __restrict__ y,
                                                    float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                    long long R, int C, float eps) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  const TX* xr = x + row * C;
  const int nv = C >> 3;
  float s = 0.f;
  for (int c = lane; c < nv; c += 32) {
    const F8 v = ld8(xr + c * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v.v[j];
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
  for (int c = lane; c < nv; c += 32) {
    const F8 v = ld8(xr + c * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = v.v[j] - mean;
      q += d * d;
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
  T* yr = y + row * C;
  for (int c = lane; c < nv; c += 32) {
    F8 v = ld8(xr + c * 8);
    const F8 g = ld8(gamma + c * 8), b = ld8(beta + c * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) v.v[j] = (v.v[j] - mean) * rstd * g.v[j] + b.v[j];
    st8(yr + c * 8, v);
  }
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)) + dres,  g = dy * gamma,  xhat = (x - mean) * rstd.
// Optional fused tail of the adapter backward (SAM2UNet.py:57-63): dx2 = dx * gelu'(pre) and colsum += sum_rows dx2
// (the gradient w.r.t. the adapter's second bias), which saves one elementwise pass and one reduction pass.
// One block = 8 warps, one row each.
constexpr int LN_BWD_ROWS = 8;               // one row per warp: enough blocks to fill the GPU at 5,808 rows
constexpr int LN_MAX_CHUNKS = 5;             // C <= 5 * 32 * 8 = 1280 (Hiera-L stage 4: 1152)
template <typename T, typename TX>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const T* __restrict__ dy, const TX* __restrict__ x,
                                                    const float* __restrict__ gamma, const float* __restrict__ mean,
                                                    const float* __restrict__ rstd, const T* __restrict__ dres,
                                                    T* __restrict__ dx, const T* __restrict__ pre,
                                                    T* __restrict__ dx2, float* __restrict__ colsum, long long R,
                                                    int C) {
  extern __shared__ float csum[];            // [C] block-level column sums of dx2 (only when colsum != null)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = C >> 3;
  if (colsum) {
    for (int i = threadIdx.x; i < C; i += 256) csum[i] = 0.f;
    __syncthreads();
  }
  float cs[LN_MAX_CHUNKS][8];                // this lane's columns (chunks lane, lane+32, ...) summed over its rows
#pragma unroll
  for (int k = 0; k < LN_MAX_CHUNKS; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) cs[k][j] = 0.f;
  for (int it = 0; it < LN_BWD_ROWS / 8; ++it) {
    const long long row = (long long)blockIdx.x * LN_BWD_ROWS + it * 8 + warp;
    if (row >= R) continue;
    const TX* xr = x + row * C;
    const T* dyr = dy + row * C;
    const float mu = mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < nv; c += 32) {
      const F8 v = ld8(xr + c * 8), d = ld8(dyr + c * 8), g = ld8(gamma + c * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float gg = d.v[j] * g.v[j];
        s1 += gg;
        s2 += gg * (v.v[j] - mu) * rs;
      }
    }
    s1 = warp_sum(s1) / (float)C;
    s2 = warp_sum(s2) / (float)C;
    T* dxr = dx + row * C;
#pragma unroll
    for (int k = 0; k < LN_MAX_CHUNKS; ++k) {
      const int c = lane + 32 * k;
      if (c >= nv) break;
      const F8 v = ld8(xr + c * 8), d = ld8(dyr + c * 8), g = ld8(gamma + c * 8);
      F8 o;
#pragma unroll
      for (int j = 0; j < 8; ++j) o.v[j] = rs * (d.v[j] * g.v[j] - s1 - (v.v[j] - mu) * rs * s2);
      if (dres) {
        const F8 r = ld8(dres + row * C + c * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) o.v[j] += r.v[j];
      }
      st8(dxr + c * 8, o);
      if (dx2) {
        const F8 pr = ld8(pre + row * C + c * 8);
        F8 o2;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          o2.v[j] = rnd(o.v[j], (const T*)nullptr) * dgelu_f(pr.v[j]);
          cs[k][j] += rnd(o2.v[j], (const T*)nullptr);
        }
        st8(dx2 + row * C + c * 8, o2);
      }
    }
  }
  if (colsum) {
#pragma unroll
    for (int k = 0; k < LN_MAX_CHUNKS; ++k) {
      const int c = lane + 32 * k;
      if (c < nv) {
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(&csum[c * 8 + j], cs[k][j]);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += 256) atomicAdd(colsum + i, csum[i]);
  }
}

extern "C" {

int s2u_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                      long long R, int C, float eps, int x_f32, int dtype, void* stream) {
  if (R <= 0 || C <= 0 || (C & 7)) return S2U_EINVAL;
  if (x_f32 || dtype == S2U_F32) {
    S2U_DISPATCH_T(dtype, {
      ln_fwd_kernel<T, float><<<ceil_div(R, 8), 256, 0, (cudaStream_t)stream>>>((const float*)x, gamma, beta, (T*)y,
                                                                              mean, rstd, R, C, eps);
    })
  } else {
    ln_fwd_kernel<bf16, bf16><<<ceil_div(R, 8), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, gamma, beta, (bf16*)y,
                                                                              mean, rstd, R, C, eps);
  }
  S2U_LAUNCH_CHECK();
  return 0;
}

int s2u_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                      const void* dres, void* dx, const void* pre, void* dx2, float* colsum, long long R, int C,
                      int x_f32, int dtype, void* stream) {
  if (R <= 0 || C <= 0 || (C & 7) || (dx2 && !pre) || (colsum && !dx2)) return S2U_EINVAL;
  if (C > LN_MAX_CHUNKS * 256) return S2U_EUNSUPPORTED;
  const size_t smem = colsum ? (size_t)C * sizeof(float) : 0;
  const int grid = ceil_div(R, LN_BWD_ROWS);
  cudaStream_t st = (cudaStream_t)stream;
  if (x_f32 || dtype == S2U_F32) {
    S2U_DISPATCH_T(dtype, {
      ln_bwd_kernel<T, float><<<grid, 256, smem, st>>>((const T*)dy, (const float*)x, gamma, mean, rstd, (const T*)dres,
                                                      (T*)dx, (const T*)pre, (T*)dx2, colsum, R, C);
    })
  } else {
    ln_bwd_kernel<bf16, bf16><<<grid, 256, smem, st>>>((const bf16*)dy, (const bf16*)x, gamma, mean, rstd,
                                                      (const bf16*)dres, (bf16*)dx, (const bf16*)pre, (bf16*)dx2,
                                                      colsum, R, C);
  }
  S2U_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
