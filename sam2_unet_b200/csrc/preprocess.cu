// Test-time input path of the reference on the device (/root/reference/dataset.py:336-407: ImageToTensor ->
// LongestMaxSizeAndPad -> NormalizeImage, used by test.py:41 and train.py:60): uint8 HWC image -> fp32 /255 ->
// resize so that the longest side is S (torchvision F.resize on a tensor = bilinear WITH antialiasing, ATen's
// _upsample_bilinear2d_aa: separable triangle filter whose support grows with the down-scale factor) -> zero padding
// to S x S, centred -> (x - mean) / std (the padding is normalised too, as in the reference: pad comes first).
// Two passes like ATen: horizontal into an fp32 intermediate [3, H, new_w], then vertical + pad + normalise.
#include "common.cuh"

struct AaAxis {
  float scale, support, invscale;
  int in_size, out_size;
};
__host__ __device__ inline AaAxis aa_axis(int in_size, int out_size) {
  AaAxis a;
  a.in_size = in_size;
  a.out_size = out_size;
  a.scale = (float)in_size / (float)out_size;             // area_pixel_compute_scale, align_corners = false
  a.support = a.scale >= 1.f ? a.scale : 1.f;             // (interp_size / 2) * scale, interp_size = 2
  a.invscale = a.scale >= 1.f ? 1.f / a.scale : 1.f;
  return a;
}
// ATen _compute_indices_min_size_weights_aa for output index i: first tap, tap count, and the weight of tap j
__device__ __forceinline__ void aa_span(const AaAxis& a, int i, int& xmin, int& xsize, float& center) {
  // ATen evaluates these with fp32 operands and a double 0.5 literal: the same promotions here, so that the integer
  // truncations land on the same side
  center = __fmul_rn(a.scale, (float)((double)i + 0.5));   // rounded product: no FMA contraction into the uses below
  xmin = max((int)((double)(center - a.support) + 0.5), 0);
  xsize = min((int)((double)(center + a.support) + 0.5), a.in_size) - xmin;
}
__device__ __forceinline__ float aa_weight(const AaAxis& a, int j, int xmin, float center) {
  const float x = (float)(((double)((float)(j + xmin) - center) + 0.5) * (double)a.invscale);
  const float ax = fabsf(x);
  return ax < 1.f ? 1.f - ax : 0.f;
}

// tmp[c, y, ox] = sum_j w_j * img[y, xmin + j, c] / 255
__global__ void __launch_bounds__(256) prep_h_kernel(const unsigned char* __restrict__ img, int H, int W, AaAxis ax,
                                                    float* __restrict__ tmp) {
  pdl_sync();
  const long long total = (long long)H * ax.out_size;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % ax.out_size), y = (int)(i / ax.out_size);
    int xmin, xsize;
    float center;
    aa_span(ax, ox, xmin, xsize, center);
    float tw = 0.f;
    for (int j = 0; j < xsize; ++j) tw += aa_weight(ax, j, xmin, center);
    float acc[3] = {0.f, 0.f, 0.f};
    const unsigned char* row = img + ((long long)y * W + xmin) * 3;
    for (int j = 0; j < xsize; ++j) {
      const float w = tw != 0.f ? aa_weight(ax, j, xmin, center) / tw : aa_weight(ax, j, xmin, center);
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[c] += w * ((float)row[j * 3 + c] / 255.f);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) tmp[((long long)c * H + y) * ax.out_size + ox] = acc[c];
  }
}

// out[c, Y, X] = ((inside ? sum_j w_j * tmp[c, ymin + j, X - pad_left] : 0) - mean[c]) / std[c]
__global__ void __launch_bounds__(256) prep_v_kernel(const float* __restrict__ tmp, int H, AaAxis ay, int new_w, int S,
                                                    int pad_left, int pad_top, float m0, float m1, float m2, float s0,
                                                    float s1, float s2, float* __restrict__ out) {
  pdl_sync();
  const long long total = (long long)S * S;
  const float mean[3] = {m0, m1, m2}, stdv[3] = {s0, s1, s2};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(i % S), Y = (int)(i / S);
    const int ox = X - pad_left, oy = Y - pad_top;
    float acc[3] = {0.f, 0.f, 0.f};
    if (ox >= 0 && ox < new_w && oy >= 0 && oy < ay.out_size) {
      int ymin, ysize;
      float center;
      aa_span(ay, oy, ymin, ysize, center);
      float tw = 0.f;
      for (int j = 0; j < ysize; ++j) tw += aa_weight(ay, j, ymin, center);
      for (int j = 0; j < ysize; ++j) {
        const float w = tw != 0.f ? aa_weight(ay, j, ymin, center) / tw : aa_weight(ay, j, ymin, center);
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] += w * tmp[((long long)c * H + ymin + j) * new_w + ox];
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) out[((long long)c * S + Y) * S + X] = (acc[c] - mean[c]) / stdv[c];
  }
}

extern "C" {

// img: uint8 [H, W, 3] (RGB, as PIL delivers it); out: fp32 [3, S, S]; tmp: fp32 scratch of 3 * H * new_w elements;
// new_h / new_w / pad_left / pad_top as computed by LongestMaxSizeAndPad (dataset.py:361-384, host integers)
int s2u_preprocess(const unsigned char* img, int H, int W, int S, int new_h, int new_w, int pad_left, int pad_top,
                   const float* mean3_host, const float* std3_host, float* tmp, float* out, void* stream) {
  if (H <= 0 || W <= 0 || S <= 0 || new_h <= 0 || new_w <= 0 || new_h > S || new_w > S || pad_left < 0 || pad_top < 0 ||
      pad_left + new_w > S || pad_top + new_h > S)
    return S2U_EINVAL;
  const AaAxis ax = aa_axis(W, new_w), ay = aa_axis(H, new_h);
  cudaStream_t st = (cudaStream_t)stream;
  long long n1 = (long long)H * new_w, n2 = (long long)S * S;
  int g1 = (int)((n1 + 255) / 256 > 148 * 8 ? 148 * 8 : (n1 + 255) / 256);
  int g2 = (int)((n2 + 255) / 256 > 148 * 8 ? 148 * 8 : (n2 + 255) / 256);
  S2U_LAUNCH(prep_h_kernel, g1, 256, 0, st, img, H, W, ax, tmp);
  S2U_LAUNCH_CHECK();
  S2U_LAUNCH(prep_v_kernel, g2, 256, 0, st, (const float*)tmp, H, ay, new_w, S, pad_left, pad_top, mean3_host[0],
             mean3_host[1], mean3_host[2], std3_host[0], std3_host[1], std3_host[2], out);
  S2U_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
