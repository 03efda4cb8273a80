// Convolution support for the stem and the RFB / decoder (NHWC everywhere inside the model).
//
//  * patch_embed: 7x7 stride-4 pad-3 conv 3->E on the NCHW fp32 input image, bias and the precomputed
//    position-embedding table added in the epilogue, NHWC tokens out
//    (/root/reference/sam2/modeling/backbones/utils.py:80-88, hieradet.py:268-283).  No backward: the stem is
//    frozen and the image needs no gradient.
//  * im2col: gathers the taps of a stride-1 "same" convolution (any kh x kw, dilation) of an NHWC map into a
//    row-major [pixels, taps*Cin] matrix, so that conv forward, input-gradient and weight-gradient all run on
//    the GEMM kernels of gemm.cu (SAM2UNet.py:68-125 RFB, :9-49 decoder).
//  * conv_weight_pack: nn.Conv2d weight [Cout,Cin,kh,kw] (the state-dict layout, kept as the fp32 master) ->
//    forward operand [Cout][(ky,kx,ci)] and flipped/transposed input-gradient operand [Cin][(ky,kx,co)].
#include "common.cuh"

// out: fp32 tokens (the residual stream) when out_f32, else T; out_copy (optional): the same values in T
template <typename T>
__global__ void __launch_bounds__(256) patch_embed_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias,
                                                         const float* __restrict__ pos, void* __restrict__ out,
                                                         int out_f32, T* __restrict__ out_copy, int B, int S, int E) {
  pdl_sync();
  extern __shared__ float sm[];
  float* wT = sm;                 // [147][E]
  float* patch = sm + 147 * E;    // [3][35][36]
  const int Hp = S / 4;
  const int tiles_x = (Hp + 7) / 8;
  const int b = blockIdx.y;
  const int oy0 = (blockIdx.x / tiles_x) * 8, ox0 = (blockIdx.x % tiles_x) * 8;
  for (int i = threadIdx.x; i < 147 * E; i += 256) {
    const int c = i / 147, k = i - c * 147;
    wT[k * E + c] = w[i];
  }
  const int iy0 = oy0 * 4 - 3, ix0 = ox0 * 4 - 3;
  for (int i = threadIdx.x; i < 3 * 35 * 35; i += 256) {
    const int ci = i / (35 * 35);
    const int r = i - ci * 35 * 35;
    const int py = r / 35, px = r - py * 35;
    const int iy = iy0 + py, ix = ix0 + px;
    float v = 0.f;
    if (iy >= 0 && iy < S && ix >= 0 && ix < S) v = x[(((long long)b * 3 + ci) * S + iy) * S + ix];
    patch[(ci * 35 + py) * 36 + px] = v;
  }
  __syncthreads();
  for (int item = threadIdx.x; item < 8 * E; item += 256) {
    const int ty = item / E, c = item - ty * E;
    float acc[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t] = 0.f;
    for (int ci = 0; ci < 3; ++ci)
      for (int ky = 0; ky < 7; ++ky) {
        const float* prow = patch + (ci * 35 + ty * 4 + ky) * 36;
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
          const float wv = wT[(ci * 49 + ky * 7 + kx) * E + c];
#pragma unroll
          for (int t = 0; t < 8; ++t) acc[t] = fmaf(prow[kx + 4 * t], wv, acc[t]);
        }
      }
    const int oy = oy0 + ty;
    if (oy >= Hp) continue;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int ox = ox0 + t;
      if (ox >= Hp) continue;
      const long long o = ((long long)oy * Hp + ox) * E + c;
      const long long go = (long long)b * Hp * Hp * E + o;
      const float val = acc[t] + bias[c] + pos[o];
      if (out_f32) reinterpret_cast<float*>(out)[go] = val;
      else stf(reinterpret_cast<T*>(out) + go, val);
      if (out_copy) stf(out_copy + go, val);
    }
  }
}

// Tensor-core form of the stem (bf16 mode): rows of the 7x7/s4/p3 patch matrix, one row per token,
//   out[m, k]       = bf16(x)              k = ci*49 + ky*7 + kx < 147   (nn.Conv2d weight order, utils.py:80-88)
//   out[m, 160 + k] = bf16(x - bf16(x))    the rounding residual, so that the image enters the GEMM with ~16
//                                          mantissa bits (the weight operand repeats W in both halves)
// and zeros elsewhere (row pitch PE_K = 320).  One thread = one 16-byte chunk of a row.
constexpr int PE_K = 320, PE_HALF = 160;
__global__ void __launch_bounds__(256) patch_im2col_kernel(const float* __restrict__ x, bf16* __restrict__ out, int B,
                                                          int S) {
  pdl_sync();
  const int Hp = S / 4;
  const long long total = (long long)B * Hp * Hp * (PE_K / 8);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % (PE_K / 8));
    const long long m = i / (PE_K / 8);
    const int ox = (int)(m % Hp), oy = (int)((m / Hp) % Hp);
    const long long b = m / ((long long)Hp * Hp);
    const bool lo = c >= PE_HALF / 8;
    const int kbase = (lo ? c - PE_HALF / 8 : c) * 8;
    F8 v;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = kbase + j;
      float val = 0.f;
      if (k < 147) {
        const int ci = k / 49, r = k - ci * 49, ky = r / 7, kx = r - ky * 7;
        const int iy = oy * 4 - 3 + ky, ix = ox * 4 - 3 + kx;
        if (iy >= 0 && iy < S && ix >= 0 && ix < S) {
          val = x[((b * 3 + ci) * S + iy) * S + ix];
          if (lo) val -= __bfloat162float(__float2bfloat16(val));
        }
      }
      v.v[j] = val;
    }
    st8(out + m * PE_K + c * 8, v);
  }
}

// out[m, (ky*KW+kx)*Cin + ci] = x[b, y + ky*dh - ph, x + kx*dw - pw, ci]  (0 outside the map)
// One thread = one (pixel, 8-channel group); it walks the taps, so the pixel decomposition is done once and a warp
// writes full 16-byte-per-lane rows of the output matrix for every tap.
template <typename T>
__global__ void __launch_bounds__(256) im2col_kernel(const T* __restrict__ x, int ldx, T* __restrict__ out, int B,
                                                    int H, int W, int Cin, int KH, int KW, int dh, int dw, int ph,
                                                    int pw) {
  pdl_sync();
  const int C8 = Cin >> 3;
  const int taps = KH * KW;
  const long long total = (long long)B * H * W * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    const long long m = i / C8;
    const int xx = (int)(m % W);
    const int yy = (int)((m / W) % H);
    const long long img = (m / ((long long)W * H)) * H;
    T* orow = out + m * taps * Cin + c * 8;
    for (int ky = 0; ky < KH; ++ky) {
      const int sy = yy + ky * dh - ph;
      const bool yok = sy >= 0 && sy < H;
      for (int kx = 0; kx < KW; ++kx) {
        const int sx = xx + kx * dw - pw;
        F8 v;
        if (yok && sx >= 0 && sx < W) {
          v = ld8(x + ((img + sy) * W + sx) * ldx + c * 8);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v.v[j] = 0.f;
        }
        st8(orow + (ky * KW + kx) * Cin, v);
      }
    }
  }
}

template <typename T>
__global__ void conv_weight_pack_kernel(const float* __restrict__ w, T* __restrict__ wf, T* __restrict__ wd, int Cout,
                                        int Cin, int KH, int KW) {
  pdl_sync();
  const int taps = KH * KW;
  const long long total = (long long)Cout * Cin * taps;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % taps);
    const int ci = (int)((i / taps) % Cin);
    const int co = (int)(i / ((long long)taps * Cin));
    const float v = w[i];
    if (wf) stf(wf + ((long long)co * taps + tap) * Cin + ci, v);
    if (wd) stf(wd + ((long long)ci * taps + (taps - 1 - tap)) * Cout + co, v);
  }
}

static inline int grid_for(long long n, int threads) {
  long long g = (n + threads - 1) / threads;
  if (g > 148LL * 32) g = 148LL * 32;
  if (g < 1) g = 1;
  return (int)g;
}

extern "C" {

int s2u_patch_embed(const float* x, const float* w, const float* bias, const float* pos, void* out, int out_f32,
                    void* out_copy, int B, int S, int E, int dtype, void* stream) {
  if (B <= 0 || S <= 0 || (S % 4) || E <= 0) return S2U_EINVAL;
  const int Hp = S / 4;
  const size_t smem = (size_t)(147 * E + 3 * 35 * 36) * sizeof(float);
  if (smem > 200 * 1024) return S2U_EUNSUPPORTED;
  dim3 grid(((Hp + 7) / 8) * ((Hp + 7) / 8), B);
  S2U_DISPATCH_T(dtype, {
    S2U_ALLOW_SMEM(patch_embed_kernel<T>);
    S2U_LAUNCH((patch_embed_kernel<T>), grid, 256, smem, (cudaStream_t)stream, x, w, bias, pos, out, out_f32, (T*)out_copy, B, S,
                                                                     E);
  })
  S2U_LAUNCH_CHECK();
  return 0;
}

// bf16 patch matrix [B*(S/4)^2, 320] of the stem for s2u_gemm (see patch_im2col_kernel)
int s2u_patch_im2col(const float* x, void* out, int B, int S, void* stream) {
  if (B <= 0 || S <= 0 || (S % 4)) return S2U_EINVAL;
  const long long total = (long long)B * (S / 4) * (S / 4) * (PE_K / 8);
  S2U_LAUNCH(patch_im2col_kernel, grid_for(total, 256), 256, 0, (cudaStream_t)stream, x, (bf16*)out, B, S);
  S2U_LAUNCH_CHECK();
  return 0;
}

int s2u_im2col(const void* x, int ldx, void* out, int B, int H, int W, int Cin, int KH, int KW, int dil_h, int dil_w,
               int pad_h, int pad_w, int dtype, void* stream) {
  if (B <= 0 || H <= 0 || W <= 0 || Cin <= 0 || (Cin & 7) || (ldx & 7)) return S2U_EINVAL;
  const long long total = (long long)B * H * W * (Cin / 8);
  S2U_DISPATCH_T(dtype, {
    S2U_LAUNCH((im2col_kernel<T>), grid_for(total, 256), 256, 0, (cudaStream_t)stream, (const T*)x, ldx, (T*)out, B, H, W, Cin,
                                                                          KH, KW, dil_h, dil_w, pad_h, pad_w);
  })
  S2U_LAUNCH_CHECK();
  return 0;
}

int s2u_conv_weight_pack(const float* w, void* wf, void* wd, int Cout, int Cin, int KH, int KW, int dtype,
                         void* stream) {
  if (Cout <= 0 || Cin <= 0 || KH <= 0 || KW <= 0) return S2U_EINVAL;
  const long long total = (long long)Cout * Cin * KH * KW;
  S2U_DISPATCH_T(dtype, {
    S2U_LAUNCH((conv_weight_pack_kernel<T>), grid_for(total, 256), 256, 0, (cudaStream_t)stream, w, (T*)wf, (T*)wd, Cout, Cin,
                                                                                    KH, KW);
  })
  S2U_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
