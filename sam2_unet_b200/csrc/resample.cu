// Bilinear resampling (forward and its transpose) and the 64->1 prediction heads.
//
//  * Up: nn.Upsample(scale_factor=2, bilinear, align_corners=True) feeding cat([skip, up])
//    (/root/reference/SAM2UNet.py:32-49): the forward writes straight into the channel slice of the concat
//    buffer (ld_out), so the concat never exists as a separate copy.
//  * heads: 1x1 conv 64->1 with bias, then F.interpolate(scale_factor=16/8/4, bilinear, align_corners=False)
//    (SAM2UNet.py:160-172), fp32 logits out.
// Both interpolations are separable 2-tap filters; the host passes per-axis tables (index pair + weights for the
// forward, the transposed adjacency list for the backward) built by sam2_unet_b200/resample_tables.py, so the
// backward is a deterministic gather, not an atomic scatter.
#include "common.cuh"

struct AxisFwd { const int* i0; const int* i1; const float* w0; const float* w1; };
struct AxisBwd { const int* idx; const float* w; int taps; };   // [n_in, taps], weight 0 marks an unused slot

template <typename T>
__global__ void resample_fwd_kernel(const T* __restrict__ x, int ldx, T* __restrict__ out, int ld_out, int B, int Hi,
                                    int Wi, int Ho, int Wo, int C, AxisFwd ay, AxisFwd ax) {
  pdl_sync();
  const int cg = C >> 3;
  const long long total = (long long)B * Ho * Wo * cg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    long long t = i / cg;
    const int ox = (int)(t % Wo); t /= Wo;
    const int oy = (int)(t % Ho);
    const int b = (int)(t / Ho);
    const int y0 = ay.i0[oy], y1 = ay.i1[oy], x0 = ax.i0[ox], x1 = ax.i1[ox];
    const float wy0 = ay.w0[oy], wy1 = ay.w1[oy], wx0 = ax.w0[ox], wx1 = ax.w1[ox];
    const T* base = x + (long long)b * Hi * Wi * ldx + g * 8;
    const F8 a = ld8(base + ((long long)y0 * Wi + x0) * ldx), bb = ld8(base + ((long long)y0 * Wi + x1) * ldx),
             c = ld8(base + ((long long)y1 * Wi + x0) * ldx), d = ld8(base + ((long long)y1 * Wi + x1) * ldx);
    F8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      o.v[j] = wy0 * (wx0 * a.v[j] + wx1 * bb.v[j]) + wy1 * (wx0 * c.v[j] + wx1 * d.v[j]);
    st8(out + (((long long)b * Ho + oy) * Wo + ox) * ld_out + g * 8, o);
  }
}

template <typename T>
__global__ void resample_bwd_kernel(const T* __restrict__ dout, int ld_do, T* __restrict__ dx, int ld_dx, int B,
                                    int Hi, int Wi, int Ho, int Wo, int C, AxisBwd ay, AxisBwd ax) {
  pdl_sync();
  const int cg = C >> 3;
  const long long total = (long long)B * Hi * Wi * cg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    long long t = i / cg;
    const int ix = (int)(t % Wi); t /= Wi;
    const int iy = (int)(t % Hi);
    const int b = (int)(t / Hi);
    F8 acc;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc.v[j] = 0.f;
    for (int a = 0; a < ay.taps; ++a) {
      const float wy = ay.w[iy * ay.taps + a];
      if (wy == 0.f) continue;
      const int oy = ay.idx[iy * ay.taps + a];
      for (int c = 0; c < ax.taps; ++c) {
        const float wx = ax.w[ix * ax.taps + c];
        if (wx == 0.f) continue;
        const int ox = ax.idx[ix * ax.taps + c];
        const F8 v = ld8(dout + (((long long)b * Ho + oy) * Wo + ox) * ld_do + g * 8);
        const float w = wy * wx;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc.v[j] = fmaf(w, v.v[j], acc.v[j]);
      }
    }
    st8(dx + (((long long)b * Hi + iy) * Wi + ix) * ld_dx + g * 8, acc);
  }
}

// single-channel fp32 variants (the logits maps)
__global__ void resample1_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int Hi, int Wi,
                                     int Ho, int Wo, AxisFwd ay, AxisFwd ax) {
  pdl_sync();
  const long long total = (long long)B * Ho * Wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % Wo);
    const int oy = (int)((i / Wo) % Ho);
    const int b = (int)(i / ((long long)Wo * Ho));
    const float* p = x + (long long)b * Hi * Wi;
    const int y0 = ay.i0[oy], y1 = ay.i1[oy], x0 = ax.i0[ox], x1 = ax.i1[ox];
    out[i] = ay.w0[oy] * (ax.w0[ox] * p[y0 * Wi + x0] + ax.w1[ox] * p[y0 * Wi + x1]) +
             ay.w1[oy] * (ax.w0[ox] * p[y1 * Wi + x0] + ax.w1[ox] * p[y1 * Wi + x1]);
  }
}

__global__ void resample1_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dx, int B, int Hi, int Wi,
                                     int Ho, int Wo, AxisBwd ay, AxisBwd ax) {
  pdl_sync();
  // one warp per input pixel: lanes split the (up to taps^2) contributing output pixels
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = (long long)B * Hi * Wi;
  if (warp >= total) return;
  const int ix = (int)(warp % Wi);
  const int iy = (int)((warp / Wi) % Hi);
  const int b = (int)(warp / ((long long)Wi * Hi));
  const float* p = dout + (long long)b * Ho * Wo;
  float acc = 0.f;
  const int n = ay.taps * ax.taps;
  for (int e = lane; e < n; e += 32) {
    const int a = e / ax.taps, c = e - a * ax.taps;
    const float w = ay.w[iy * ay.taps + a] * ax.w[ix * ax.taps + c];
    if (w != 0.f) acc = fmaf(w, p[(long long)ay.idx[iy * ay.taps + a] * Wo + ax.idx[ix * ax.taps + c]], acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) dx[warp] = acc;
}

// logit[m] = bias + sum_c feat[m, c] * w[c]   (C = 64: 8 lanes x 8 channels per pixel)
template <typename T>
__global__ void head_fwd_kernel(const T* __restrict__ feat, int ldf_, const float* __restrict__ w,
                                const float* __restrict__ bias, float* __restrict__ out, long long M, int C) {
  pdl_sync();
  const int cg = C >> 3;                    // lanes per pixel (8 for C = 64)
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long m = gid / cg;
  const int g = (int)(gid % cg);
  float s = 0.f;
  if (m < M) {
    const F8 v = ld8(feat + m * ldf_ + g * 8), ww = ld8(w + g * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) s = fmaf(v.v[j], ww.v[j], s);
  }
  for (int o = cg >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (m < M && g == 0) out[m] = s + bias[0];
}

// dfeat[m, c] (+)= dlogit[m] * w[c];  dw[c] += sum_m dlogit[m] * feat[m, c];  db += sum_m dlogit[m]
template <typename T>
__global__ void __launch_bounds__(256) head_bwd_kernel(const T* __restrict__ feat, int ldf_,
                                                      const float* __restrict__ w, const float* __restrict__ dlogit,
                                                      T* __restrict__ dfeat, int ld_df, int accumulate,
                                                      float* __restrict__ dw, float* __restrict__ db, long long M,
                                                      int C, int rows_per_block) {
  pdl_sync();
  __shared__ float red[32][65];
  const int cg = C >> 3;                    // 8
  const int g = threadIdx.x % cg, sub = threadIdx.x / cg;     // sub: 0..31
  const F8 ww = ld8(w + g * 8);
  float aw[8], ab = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) aw[j] = 0.f;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  for (long long m = r0 + sub; m < min(M, r0 + rows_per_block); m += 256 / cg) {
    const float d = dlogit[m];
    const F8 v = ld8(feat + m * ldf_ + g * 8);
    F8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) { aw[j] = fmaf(d, v.v[j], aw[j]); o.v[j] = d * ww.v[j]; }
    if (accumulate) {
      const F8 old = ld8(dfeat + m * ld_df + g * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) o.v[j] += old.v[j];
    }
    st8(dfeat + m * ld_df + g * 8, o);
    if (g == 0) ab += d;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[sub][g * 8 + j] = aw[j];
  if (g == 0) red[sub][64] = ab;
  __syncthreads();
  if (threadIdx.x < 65) {
    float t = 0.f;
    for (int k = 0; k < 32; ++k) t += red[k][threadIdx.x];
    if (threadIdx.x < 64) atomicAdd(dw + threadIdx.x, t);
    else atomicAdd(db, t);
  }
}

static inline int grid_for(long long n, int threads) {
  long long g = (n + threads - 1) / threads;
  if (g > 148LL * 16) g = 148LL * 16;
  if (g < 1) g = 1;
  return (int)g;
}

extern "C" {

int s2u_resample_fwd(const void* x, int ldx, void* out, int ld_out, int B, int Hi, int Wi, int Ho, int Wo, int C,
                     const int* y_i0, const int* y_i1, const float* y_w0, const float* y_w1, const int* x_i0,
                     const int* x_i1, const float* x_w0, const float* x_w1, int dtype, void* stream) {
  if (B <= 0 || (C & 7) || (ldx & 7) || (ld_out & 7)) return S2U_EINVAL;
  AxisFwd ay{y_i0, y_i1, y_w0, y_w1}, ax{x_i0, x_i1, x_w0, x_w1};
  S2U_DISPATCH_T(dtype, {
    S2U_LAUNCH((resample_fwd_kernel<T>), grid_for((long long)B * Ho * Wo * (C >> 3), 256), 256, 0, (cudaStream_t)stream, 
        (const T*)x, ldx, (T*)out, ld_out, B, Hi, Wi, Ho, Wo, C, ay, ax);
  })
  S2U_LAUNCH_CHECK();
  return 0;
}

int s2u_resample_bwd(const void* dout, int ld_do, void* dx, int ld_dx, int B, int Hi, int Wi, int Ho, int Wo, int C,
                     const int* y_idx, const float* y_w, int y_taps, const int* x_idx, const float* x_w, int x_taps,
                     int dtype, void* stream) {
  if (B <= 0 || (C & 7) || (ld_do & 7) || (ld_dx & 7)) return S2U_EINVAL;
  AxisBwd ay{y_idx, y_w, y_taps}, ax{x_idx, x_w, x_taps};
  S2U_DISPATCH_T(dtype, {
    S2U_LAUNCH((resample_bwd_kernel<T>), grid_for((long long)B * Hi * Wi * (C >> 3), 256), 256, 0, (cudaStream_t)stream, 
        (const T*)dout, ld_do, (T*)dx, ld_dx, B, Hi, Wi, Ho, Wo, C, ay, ax);
  })
  S2U_LAUNCH_CHECK();
  return 0;
}

int s2u_resample1_fwd(const float* x, float* out, int B, int Hi, int Wi, int Ho, int Wo, const int* y_i0,
                      const int* y_i1, const float* y_w0, const float* y_w1, const int* x_i0, const int* x_i1,
                      const float* x_w0, const float* x_w1, void* stream) {
  if (B <= 0) return S2U_EINVAL;
  AxisFwd ay{y_i0, y_i1, y_w0, y_w1}, ax{x_i0, x_i1, x_w0, x_w1};
  S2U_LAUNCH((resample1_fwd_kernel), grid_for((long long)B * Ho * Wo, 256), 256, 0, (cudaStream_t)stream, x, out, B, Hi, Wi, Ho,
                                                                                             Wo, ay, ax);
  S2U_LAUNCH_CHECK();
  return 0;
}

int s2u_resample1_bwd(const float* dout, float* dx, int B, int Hi, int Wi, int Ho, int Wo, const int* y_idx,
                      const float* y_w, int y_taps, const int* x_idx, const float* x_w, int x_taps, void* stream) {
  if (B <= 0) return S2U_EINVAL;
  AxisBwd ay{y_idx, y_w, y_taps}, ax{x_idx, x_w, x_taps};
  const long long warps = (long long)B * Hi * Wi;
  S2U_LAUNCH((resample1_bwd_kernel), ceil_div(warps * 32, 256), 256, 0, (cudaStream_t)stream, dout, dx, B, Hi, Wi, Ho, Wo, ay,
                                                                                  ax);
  S2U_LAUNCH_CHECK();
  return 0;
}

int s2u_head_fwd(const void* feat, int ldf_, const float* w, const float* bias, float* out, long long M, int C,
                 int dtype, void* stream) {
  if (M <= 0 || C != 64 || (ldf_ & 7)) return S2U_EINVAL;
  S2U_DISPATCH_T(dtype, {
    S2U_LAUNCH((head_fwd_kernel<T>), ceil_div(M * (C >> 3), 256), 256, 0, (cudaStream_t)stream, (const T*)feat, ldf_, w, bias,
                                                                                    out, M, C);
  })
  S2U_LAUNCH_CHECK();
  return 0;
}

int s2u_head_bwd(const void* feat, int ldf_, const float* w, const float* dlogit, void* dfeat, int ld_df,
                 int accumulate, float* dw, float* db, long long M, int C, int dtype, void* stream) {
  if (M <= 0 || C != 64 || (ldf_ & 7) || (ld_df & 7)) return S2U_EINVAL;
  const int rows = 1024;
  S2U_DISPATCH_T(dtype, {
    S2U_LAUNCH((head_bwd_kernel<T>), ceil_div(M, rows), 256, 0, (cudaStream_t)stream, (const T*)feat, ldf_, w, dlogit, (T*)dfeat,
                                                                           ld_df, accumulate, dw, db, M, C, rows);
  })
  S2U_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
