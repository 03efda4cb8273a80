// LayerNorm (eps 1e-6, /root/reference/sam2/modeling/backbones/hieradet.py:99-100,104,120) forward and
// input-gradient, one warp per token row, 128-bit accesses.  The trunk's affine parameters are frozen
// (SAM2UNet.py:146-147) so there is no weight gradient.  Algorithmic traffic: forward reads x and writes y
// (2 elements/elem), backward reads dy, x (+ residual gradient) and writes dx.
#include "common.cuh"

template <typename T>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                                                    const float* __restrict__ beta, T* __restrict__ y,
                                                    float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                    long long R, int C, float eps) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  const T* xr = x + row * C;
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

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)) + dres,  g = dy * gamma,  xhat = (x - mean) * rstd
template <typename T>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                    const float* __restrict__ gamma, const float* __restrict__ mean,
                                                    const float* __restrict__ rstd, const T* __restrict__ dres,
                                                    T* __restrict__ dx, long long R, int C) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  const T* xr = x + row * C;
  const T* dyr = dy + row * C;
  const float mu = mean[row], rs = rstd[row];
  const int nv = C >> 3;
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
  for (int c = lane; c < nv; c += 32) {
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
  }
}

extern "C" {

int s2u_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                      long long R, int C, float eps, int dtype, void* stream) {
  if (R <= 0 || C <= 0 || (C & 7)) return S2U_EINVAL;
  S2U_DISPATCH_T(dtype, {
    ln_fwd_kernel<T><<<ceil_div(R, 8), 256, 0, (cudaStream_t)stream>>>((const T*)x, gamma, beta, (T*)y, mean, rstd, R,
                                                                     C, eps);
  })
  S2U_LAUNCH_CHECK();
  return 0;
}

int s2u_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                      const void* dres, void* dx, long long R, int C, int dtype, void* stream) {
  if (R <= 0 || C <= 0 || (C & 7)) return S2U_EINVAL;
  S2U_DISPATCH_T(dtype, {
    ln_bwd_kernel<T><<<ceil_div(R, 8), 256, 0, (cudaStream_t)stream>>>((const T*)dy, (const T*)x, gamma, mean, rstd,
                                                                     (const T*)dres, (T*)dx, R, C);
  })
  S2U_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
