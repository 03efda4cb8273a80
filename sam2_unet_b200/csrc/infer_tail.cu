// Post-forward inference tail on the device (/root/reference/test.py:66-76, train.py:103-112): crop the letterbox
// padding off the [S,S] logit map, bilinear-resize it to the ground-truth size (align_corners = False, ATen's
// upsample_bilinear2d arithmetic), sigmoid, min-max normalise over the image and quantise to 8 bits.  The reference
// copies the fp32 map to the host and does the last three steps in numpy; here one byte per pixel leaves the device.
//
// Two launches: (1) min / max of the resized logits (sigmoid is monotonic, so the extrema of sigmoid(resized) are the
// sigmoid of these), one ordered-integer atomic pair per block; (2) resize again, sigmoid, normalise, truncate to
// uint8 (numpy's astype) - the map is ~0.5 MB, recomputing the four-tap resize is cheaper than storing it.
#include "common.cuh"

struct TailGeom {
  int S, x0, y0, in_w, in_h, out_w, out_h;
  float scale_x, scale_y;
};

// ATen area_pixel_compute_source_index, align_corners = false, not cubic
__device__ __forceinline__ float src_index(float scale, int dst) {
  const float s = scale * ((float)dst + 0.5f) - 0.5f;
  return s < 0.f ? 0.f : s;
}
__device__ __forceinline__ float resized(const float* __restrict__ logits, const TailGeom& g, int oy, int ox) {
  const float sy = src_index(g.scale_y, oy), sx = src_index(g.scale_x, ox);
  const int y1 = (int)sy, x1 = (int)sx;
  const int yp = y1 < g.in_h - 1 ? 1 : 0, xp = x1 < g.in_w - 1 ? 1 : 0;
  const float ly1 = sy - (float)y1, ly0 = 1.f - ly1, lx1 = sx - (float)x1, lx0 = 1.f - lx1;
  const float* p = logits + (long long)(g.y0 + y1) * g.S + g.x0 + x1;
  const float v00 = p[0], v01 = p[xp], v10 = p[(long long)yp * g.S], v11 = p[(long long)yp * g.S + xp];
  // same association as ATen: h0 * (w0 * v00 + w1 * v01) + h1 * (w0 * v10 + w1 * v11), no contraction
  const float top = __fadd_rn(__fmul_rn(lx0, v00), __fmul_rn(lx1, v01));
  const float bot = __fadd_rn(__fmul_rn(lx0, v10), __fmul_rn(lx1, v11));
  return __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
}
// order-preserving map float -> int (for atomicMin / atomicMax on floats)
__device__ __forceinline__ int f2ord(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// mm[0] = min, mm[1] = max (ordered ints); must hold {INT_MAX, INT_MIN} on entry
__global__ void __launch_bounds__(256) tail_minmax_kernel(const float* __restrict__ logits, TailGeom g,
                                                         int* __restrict__ mm) {
  pdl_sync();
  const long long total = (long long)g.out_h * g.out_w;
  float lo = INFINITY, hi = -INFINITY;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const float v = resized(logits, g, (int)(i / g.out_w), (int)(i % g.out_w));
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
  lo = -warp_max(-lo);
  hi = warp_max(hi);
  __shared__ float slo[8], shi[8];
  if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { lo = fminf(lo, slo[w]); hi = fmaxf(hi, shi[w]); }
    if (lo <= hi) {
      atomicMin(mm, f2ord(lo));
      atomicMax(mm + 1, f2ord(hi));
    }
  }
}

__device__ __forceinline__ float sigmoid_ref(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(256) tail_quantise_kernel(const float* __restrict__ logits, TailGeom g,
                                                           int* __restrict__ mm, unsigned char* __restrict__ out) {
  pdl_sync();
  const float smin = sigmoid_ref(ord2f(mm[0])), smax = sigmoid_ref(ord2f(mm[1]));
  const float denom = __fadd_rn(__fsub_rn(smax, smin), 1e-8f);
  const long long total = (long long)g.out_h * g.out_w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const float s = sigmoid_ref(resized(logits, g, (int)(i / g.out_w), (int)(i % g.out_w)));
    const float n = __fdiv_rn(__fsub_rn(s, smin), denom);
    out[i] = (unsigned char)(int)__fmul_rn(n, 255.f);        // numpy astype(uint8): truncation
  }
}

// re-arms the min / max words for the next image (stream-ordered after the quantise pass)
__global__ void tail_reset_kernel(int* __restrict__ mm) {
  pdl_sync();
  mm[0] = 0x7fffffff;
  mm[1] = (int)0x80000000;
}

extern "C" {

// logits: [S,S] fp32 map of one image (the padded network output); pad_*: letterbox padding to remove; out: [out_h,
// out_w] uint8.  ws: two ints of device scratch holding {INT_MAX, INT_MIN} on entry (s2u_infer_tail_init) - the call
// leaves them re-armed, so one workspace serves a whole stream of images.
int s2u_infer_tail_init(int* ws, void* stream) {
  S2U_LAUNCH(tail_reset_kernel, 1, 1, 0, (cudaStream_t)stream, ws);
  S2U_LAUNCH_CHECK();
  return 0;
}

int s2u_infer_tail(const float* logits, int S, int pad_left, int pad_top, int pad_right, int pad_bottom, int out_h,
                   int out_w, int* ws, unsigned char* out, void* stream) {
  TailGeom g;
  g.S = S;
  g.x0 = pad_left;
  g.y0 = pad_top;
  g.in_w = S - pad_left - pad_right;
  g.in_h = S - pad_top - pad_bottom;
  g.out_w = out_w;
  g.out_h = out_h;
  if (S <= 0 || pad_left < 0 || pad_top < 0 || pad_right < 0 || pad_bottom < 0 || g.in_w <= 0 || g.in_h <= 0 ||
      out_h <= 0 || out_w <= 0)
    return S2U_EINVAL;
  g.scale_x = (float)g.in_w / (float)out_w;
  g.scale_y = (float)g.in_h / (float)out_h;
  const long long total = (long long)out_h * out_w;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  cudaStream_t st = (cudaStream_t)stream;
  S2U_LAUNCH(tail_minmax_kernel, (int)blocks, 256, 0, st, logits, g, ws);
  S2U_LAUNCH_CHECK();
  S2U_LAUNCH(tail_quantise_kernel, (int)blocks, 256, 0, st, logits, g, ws, out);
  S2U_LAUNCH_CHECK();
  S2U_LAUNCH(tail_reset_kernel, 1, 1, 0, st, ws);
  S2U_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
