// Small HBM-bound helpers of the trunk: GELU' product, 2x2 max-pool (hieradet.py:21-32,108-110) forward and
// backward on NHWC tokens, dtype casts / transposes that refresh the low-precision weight shadows.
#include "common.cuh"

template <typename T>
__global__ void dgelu_mul_kernel(const T* __restrict__ dy, const T* __restrict__ pre, T* __restrict__ out,
                                 long long n8) {
  pdl_sync();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const F8 d = ld8(dy + i * 8), p = ld8(pre + i * 8);
    F8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) o.v[j] = d.v[j] * dgelu_f(p.v[j]);
    st8(out + i * 8, o);
  }
}

template <typename T>
__global__ void add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, long long n8) {
  pdl_sync();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const F8 x = ld8(a + i * 8), y = ld8(b + i * 8);
    F8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) o.v[j] = x.v[j] + y.v[j];
    st8(out + i * 8, o);
  }
}

// x [B,H,W,C] -> out [B,H/2,W/2,C]  (floor mode)
template <typename T>
__global__ void maxpool2_fwd_kernel(const T* __restrict__ x, T* __restrict__ out, int B, int H, int W, int C) {
  pdl_sync();
  const int Ho = H / 2, Wo = W / 2, C8 = C >> 3;
  const long long total = (long long)B * Ho * Wo * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    long long t = i / C8;
    const int ox = (int)(t % Wo); t /= Wo;
    const int oy = (int)(t % Ho);
    const int b = (int)(t / Ho);
    const T* p = x + (((long long)b * H + 2 * oy) * W + 2 * ox) * C + c * 8;
    F8 m = ld8(p);
    const F8 v1 = ld8(p + C), v2 = ld8(p + (long long)W * C), v3 = ld8(p + (long long)W * C + C);
#pragma unroll
    for (int j = 0; j < 8; ++j) m.v[j] = fmaxf(fmaxf(m.v[j], v1.v[j]), fmaxf(v2.v[j], v3.v[j]));
    st8(out + (((long long)b * Ho + oy) * Wo + ox) * C + c * 8, m);
  }
}

// dx[argmax of each 2x2 cell] = dout, other positions 0; first maximum in (y,x) scan order wins, as ATen does.
// Rows/columns of x beyond 2*floor(H/2) receive no gradient.
template <typename T>
__global__ void maxpool2_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dout, T* __restrict__ dx, int B,
                                    int H, int W, int C) {
  pdl_sync();
  const int Ho = H / 2, Wo = W / 2, C8 = C >> 3;
  const int Hc = (H + 1) / 2, Wc = (W + 1) / 2;
  const long long total = (long long)B * Hc * Wc * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    long long t = i / C8;
    const int ox = (int)(t % Wc); t /= Wc;
    const int oy = (int)(t % Hc);
    const int b = (int)(t / Hc);
    F8 zero;
#pragma unroll
    for (int j = 0; j < 8; ++j) zero.v[j] = 0.f;
    if (oy >= Ho || ox >= Wo) {            // ragged border: no pooling cell covers it
      for (int dy = 0; dy < 2; ++dy)
        for (int dxx = 0; dxx < 2; ++dxx) {
          const int yy = 2 * oy + dy, xx = 2 * ox + dxx;
          if (yy < H && xx < W) st8(dx + (((long long)b * H + yy) * W + xx) * C + c * 8, zero);
        }
      continue;
    }
    const long long base = (((long long)b * H + 2 * oy) * W + 2 * ox) * C + c * 8;
    const long long offs[4] = {0, C, (long long)W * C, (long long)W * C + C};
    F8 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = ld8(x + base + offs[k]);
    const F8 g = ld8(dout + (((long long)b * Ho + oy) * Wo + ox) * C + c * 8);
    F8 o[4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int best = 0;
      float bv = v[0].v[j];
#pragma unroll
      for (int k = 1; k < 4; ++k)
        if (v[k].v[j] > bv) { bv = v[k].v[j]; best = k; }
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k].v[j] = (k == best) ? g.v[j] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) st8(dx + base + offs[k], o[k]);
  }
}

// dst[c, r] = (T) src[r, c]  (or plain cast when transpose == 0)
template <typename T>
__global__ void cast_kernel(const float* __restrict__ src, T* __restrict__ dst, int R, int C, int transpose) {
  pdl_sync();
  const long long total = (long long)R * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    if (!transpose) {
      stf(dst + i, src[i]);
    } else {
      const int r = (int)(i / C), c = (int)(i % C);
      stf(dst + (long long)c * R + r, src[i]);
    }
  }
}

// One launch refreshes every low-precision operand of the trainable weights from the fp32 masters (192 adapter
// casts/transposes + 66 conv re-layouts per step otherwise).  `blocks[b]` = (entry, chunk): block b converts elements
// [chunk*1024, chunk*1024+1024) of entry `entry`.
struct S2uRefreshEntry {
  unsigned long long src, dst0, dst1;   // fp32 master; compute-dtype outputs (0 = skip)
  int kind;                             // 0: [d0, d1] matrix -> dst0 (same layout) and dst1 (transposed)
  int d0, d1, d2, d3;                   // 1: conv weight [Cout=d0][Cin=d1][KH=d2][KW=d3] -> dst0 [Cout][tap][Cin],
  int pad;                              //    dst1 [Cin][flipped tap][Cout]; pad != 0: row pitch of dst1 (the five 1x1
                                        //    convs of an RFB write column slices of one [Cin][320] operand)
};

template <typename T>
__global__ void __launch_bounds__(256) refresh_kernel(const S2uRefreshEntry* __restrict__ entries,
                                                     const int2* __restrict__ blocks) {
  pdl_sync();
  const int2 bk = blocks[blockIdx.x];
  const S2uRefreshEntry e = entries[bk.x];
  const float* src = reinterpret_cast<const float*>(e.src);
  T* dst0 = reinterpret_cast<T*>(e.dst0);
  T* dst1 = reinterpret_cast<T*>(e.dst1);
  const long long total = e.kind == 0 ? (long long)e.d0 * e.d1 : (long long)e.d0 * e.d1 * e.d2 * e.d3;
  const long long base = (long long)bk.y * 1024;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long i = base + k * 256 + threadIdx.x;
    if (i >= total) break;
    const float v = src[i];
    if (e.kind == 0) {
      if (dst0) stf(dst0 + i, v);
      if (dst1) {
        const int r = (int)(i / e.d1), c = (int)(i % e.d1);
        stf(dst1 + (long long)c * e.d0 + r, v);
      }
    } else {
      const int taps = e.d2 * e.d3;
      const int tap = (int)(i % taps);
      const int ci = (int)((i / taps) % e.d1);
      const int co = (int)(i / ((long long)taps * e.d1));
      if (dst0) stf(dst0 + ((long long)co * taps + tap) * e.d1 + ci, v);
      if (dst1) {
        if (e.pad) stf(dst1 + (long long)ci * e.pad + (taps - 1 - tap) * e.d0 + co, v);
        else stf(dst1 + ((long long)ci * taps + (taps - 1 - tap)) * e.d0 + co, v);
      }
    }
  }
}

static inline int grid_for(long long n, int threads) {
  long long g = (n + threads - 1) / threads;
  if (g > 148LL * 16) g = 148LL * 16;
  if (g < 1) g = 1;
  return (int)g;
}

extern "C" {

int s2u_dgelu_mul(const void* dy, const void* pre, void* out, long long n, int dtype, void* stream) {
  if (n <= 0 || (n & 7)) return S2U_EINVAL;
  S2U_DISPATCH_T(dtype, {
    S2U_LAUNCH((dgelu_mul_kernel<T>), grid_for(n / 8, 256), 256, 0, (cudaStream_t)stream, (const T*)dy, (const T*)pre, (T*)out,
                                                                             n / 8);
  })
  S2U_LAUNCH_CHECK();
  return 0;
}

int s2u_add(const void* a, const void* b, void* out, long long n, int dtype, void* stream) {
  if (n <= 0 || (n & 7)) return S2U_EINVAL;
  S2U_DISPATCH_T(dtype, {
    S2U_LAUNCH((add_kernel<T>), grid_for(n / 8, 256), 256, 0, (cudaStream_t)stream, (const T*)a, (const T*)b, (T*)out, n / 8);
  })
  S2U_LAUNCH_CHECK();
  return 0;
}

int s2u_maxpool2_fwd(const void* x, void* out, int B, int H, int W, int C, int dtype, void* stream) {
  if (B <= 0 || H < 2 || W < 2 || (C & 7)) return S2U_EINVAL;
  const long long total = (long long)B * (H / 2) * (W / 2) * (C / 8);
  S2U_DISPATCH_T(dtype, {
    S2U_LAUNCH((maxpool2_fwd_kernel<T>), grid_for(total, 256), 256, 0, (cudaStream_t)stream, (const T*)x, (T*)out, B, H, W, C);
  })
  S2U_LAUNCH_CHECK();
  return 0;
}

int s2u_maxpool2_bwd(const void* x, const void* dout, void* dx, int B, int H, int W, int C, int dtype, void* stream) {
  if (B <= 0 || H < 2 || W < 2 || (C & 7)) return S2U_EINVAL;
  const long long total = (long long)B * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8);
  S2U_DISPATCH_T(dtype, {
    S2U_LAUNCH((maxpool2_bwd_kernel<T>), grid_for(total, 256), 256, 0, (cudaStream_t)stream, (const T*)x, (const T*)dout,
                                                                                (T*)dx, B, H, W, C);
  })
  S2U_LAUNCH_CHECK();
  return 0;
}

int s2u_refresh_shadows(const void* entries, const void* blocks, int nblocks, int dtype, void* stream) {
  if (nblocks <= 0 || !entries || !blocks) return S2U_EINVAL;
  S2U_DISPATCH_T(dtype, {
    S2U_LAUNCH((refresh_kernel<T>), nblocks, 256, 0, (cudaStream_t)stream, (const S2uRefreshEntry*)entries, (const int2*)blocks);
  })
  S2U_LAUNCH_CHECK();
  return 0;
}

int s2u_cast(const float* src, void* dst, int R, int C, int transpose, int dtype, void* stream) {
  if (R <= 0 || C <= 0) return S2U_EINVAL;
  S2U_DISPATCH_T(dtype, {
    S2U_LAUNCH((cast_kernel<T>), grid_for((long long)R * C, 256), 256, 0, (cudaStream_t)stream, src, (T*)dst, R, C, transpose);
  })
  S2U_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
