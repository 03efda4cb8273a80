// Training-time input pipeline of the reference on the device (/root/reference/dataset.py:34-333: FullDataset's
// transform = ToTensor -> ResizeLongestSideAndPad -> RandomRotate -> ToGray -> ColorAugmentations -> GaussianBlur ->
// Normalize).  The random DECISIONS stay on the host (Python's `random`, drawn in the reference's order, see
// sam2_unet_b200/augment.py); these kernels are the pixel work, one launch per transform, on a [3, S, S] fp32 image and
// a [1, S, S] fp32 label.  Arithmetic follows torchvision's tensor kernels operation by operation (_functional_tensor.py:
// rgb_to_grayscale, _blend, adjust_contrast / saturation / hue / gamma, gaussian_blur) and ATen's antialiased bilinear /
// nearest resize, so results agree with the reference to float rounding.
#include "common.cuh"

namespace aug {

struct AaAxis {
  float scale, support, invscale;
  int in_size, out_size;
};
static AaAxis aa_axis(int in_size, int out_size) {
  AaAxis a;
  a.in_size = in_size;
  a.out_size = out_size;
  a.scale = (float)in_size / (float)out_size;
  a.support = a.scale >= 1.f ? a.scale : 1.f;
  a.invscale = a.scale >= 1.f ? 1.f / a.scale : 1.f;
  return a;
}
// ATen _compute_indices_min_size_weights_aa (same promotions as csrc/preprocess.cu)
__device__ __forceinline__ void aa_span(const AaAxis& a, int i, int& xmin, int& xsize, float& center) {
  center = __fmul_rn(a.scale, (float)((double)i + 0.5));
  xmin = max((int)((double)(center - a.support) + 0.5), 0);
  xsize = min((int)((double)(center + a.support) + 0.5), a.in_size) - xmin;
}
__device__ __forceinline__ float aa_weight(const AaAxis& a, int j, int xmin, float center) {
  const float x = (float)(((double)((float)(j + xmin) - center) + 0.5) * (double)a.invscale);
  const float ax = fabsf(x);
  return ax < 1.f ? 1.f - ax : 0.f;
}

// The "processed" image of ResizeLongestSideAndPad (dataset.py:57-103) as a virtual [ph, pw] view of the uint8 source:
// mode 0 = the source padded by (off_y, off_x) on the top / left with fill 1.0 (label 0) everywhere outside,
// mode 1 = the crop starting at (off_y, off_x).
struct Src {
  const unsigned char* img;   // [H, W, 3]
  const unsigned char* lab;   // [H, W]
  int H, W, mode, off_y, off_x;
};
__device__ __forceinline__ bool src_pos(const Src& s, int y, int x, int& sy, int& sx) {
  sy = s.mode ? y + s.off_y : y - s.off_y;
  sx = s.mode ? x + s.off_x : x - s.off_x;
  return sy >= 0 && sy < s.H && sx >= 0 && sx < s.W;
}

// tmp[c, y, ox] = sum_j w_j * V[y, xmin + j, c]   (V = the virtual processed image, values in [0, 1])
__global__ void __launch_bounds__(256) resize_h_kernel(Src s, int ph, AaAxis ax, float* __restrict__ tmp) {
  pdl_sync();
  const long long total = (long long)ph * ax.out_size;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % ax.out_size), y = (int)(i / ax.out_size);
    int xmin, xsize;
    float center;
    aa_span(ax, ox, xmin, xsize, center);
    float tw = 0.f;
    for (int j = 0; j < xsize; ++j) tw += aa_weight(ax, j, xmin, center);
    float acc[3] = {0.f, 0.f, 0.f};
    for (int j = 0; j < xsize; ++j) {
      const float w = tw != 0.f ? aa_weight(ax, j, xmin, center) / tw : aa_weight(ax, j, xmin, center);
      int sy, sx;
      if (src_pos(s, y, xmin + j, sy, sx)) {
        const unsigned char* p = s.img + ((long long)sy * s.W + sx) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] += w * ((float)p[c] / 255.f);
      } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] += w;                       // white padding (fill = 1.0)
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) tmp[((long long)c * ph + y) * ax.out_size + ox] = acc[c];
  }
}

// vertical pass + centred zero padding to S x S; the label by nearest neighbour (ATen: floor(dst * in / out))
__global__ void __launch_bounds__(256) resize_v_kernel(const float* __restrict__ tmp, Src s, int ph, int pw, AaAxis ay,
                                                      int new_w, int S, int pad_left, int pad_top, float lab_sy,
                                                      float lab_sx, float* __restrict__ out, float* __restrict__ lab) {
  pdl_sync();
  const long long total = (long long)S * S;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(i % S), Y = (int)(i / S);
    const int ox = X - pad_left, oy = Y - pad_top;
    float acc[3] = {0.f, 0.f, 0.f};
    float lv = 0.f;
    if (ox >= 0 && ox < new_w && oy >= 0 && oy < ay.out_size) {
      int ymin, ysize;
      float center;
      aa_span(ay, oy, ymin, ysize, center);
      float tw = 0.f;
      for (int j = 0; j < ysize; ++j) tw += aa_weight(ay, j, ymin, center);
      for (int j = 0; j < ysize; ++j) {
        const float w = tw != 0.f ? aa_weight(ay, j, ymin, center) / tw : aa_weight(ay, j, ymin, center);
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] += w * tmp[((long long)c * ph + ymin + j) * new_w + ox];
      }
      const int ly = min((int)floorf((float)oy * lab_sy), ph - 1), lx = min((int)floorf((float)ox * lab_sx), pw - 1);
      int sy, sx;
      if (src_pos(s, ly, lx, sy, sx)) lv = (float)s.lab[(long long)sy * s.W + sx] / 255.f;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) out[((long long)c * S + Y) * S + X] = acc[c];
    lab[i] = lv;
  }
}

// torch.rot90(x, k, (1, 2)) of [C, S, S] (counter-clockwise, what F.rotate(angle = 90 k) does to a square map)
__global__ void __launch_bounds__(256) rot90_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int S,
                                                   int k) {
  pdl_sync();
  const long long total = (long long)C * S * S;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % S), y = (int)((i / S) % S);
    const long long c = i / ((long long)S * S);
    int sy, sx;
    if (k == 1) { sy = x; sx = S - 1 - y; }
    else if (k == 2) { sy = S - 1 - y; sx = S - 1 - x; }
    else { sy = S - 1 - x; sx = y; }
    out[i] = in[(c * S + sy) * S + sx];
  }
}

__device__ __forceinline__ float gray_of(float r, float g, float b) {
  return __fadd_rn(__fadd_rn(__fmul_rn(0.2989f, r), __fmul_rn(0.587f, g)), __fmul_rn(0.114f, b));
}
// torchvision _blend, bound = 1: ratio and (1.0 - ratio) are Python doubles rounded to fp32 separately
__device__ __forceinline__ float blend(float a, float b, float ratio, float comp) {
  return fminf(fmaxf(__fadd_rn(__fmul_rn(ratio, a), __fmul_rn(comp, b)), 0.f), 1.f);
}

// sum of the grayscale image (adjust_contrast's mean), fp64 atomics
__global__ void __launch_bounds__(256) gray_sum_kernel(const float* __restrict__ img, long long n, double* __restrict__ out) {
  pdl_sync();
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    s += (double)gray_of(img[i], img[n + i], img[2 * n + i]);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ double red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(out, t);
  }
}

enum Op { OP_GRAY = 0, OP_BRIGHTNESS = 1, OP_CONTRAST = 2, OP_SATURATION = 3, OP_HUE = 4, OP_GAMMA = 5, OP_NORMALIZE = 6 };
struct OpArgs {
  float f, fc;                // factor / gamma / hue shift; fc = fp32(1.0 - factor) for the blends
  const double* gray_sum;     // contrast: sum of the grayscale image
  float m[3], s[3];           // normalise
};

__global__ void __launch_bounds__(256) color_kernel(float* __restrict__ img, long long n, int op, OpArgs a) {
  pdl_sync();
  const float mean = op == OP_CONTRAST ? (float)(*a.gray_sum / (double)n) : 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float r = img[i], g = img[n + i], b = img[2 * n + i];
    if (op == OP_GRAY) {
      r = g = b = gray_of(r, g, b);
    } else if (op == OP_BRIGHTNESS) {
      r = blend(r, 0.f, a.f, a.fc); g = blend(g, 0.f, a.f, a.fc); b = blend(b, 0.f, a.f, a.fc);
    } else if (op == OP_CONTRAST) {
      r = blend(r, mean, a.f, a.fc); g = blend(g, mean, a.f, a.fc); b = blend(b, mean, a.f, a.fc);
    } else if (op == OP_SATURATION) {
      const float l = gray_of(r, g, b);
      r = blend(r, l, a.f, a.fc); g = blend(g, l, a.f, a.fc); b = blend(b, l, a.f, a.fc);
    } else if (op == OP_GAMMA) {
      r = fminf(fmaxf(powf(r, a.f), 0.f), 1.f);
      g = fminf(fmaxf(powf(g, a.f), 0.f), 1.f);
      b = fminf(fmaxf(powf(b, a.f), 0.f), 1.f);
    } else if (op == OP_NORMALIZE) {
      r = __fdiv_rn(r - a.m[0], a.s[0]); g = __fdiv_rn(g - a.m[1], a.s[1]); b = __fdiv_rn(b - a.m[2], a.s[2]);
    } else if (op == OP_HUE) {
      // _rgb2hsv
      const float maxc = fmaxf(r, fmaxf(g, b)), minc = fminf(r, fminf(g, b));
      const bool eqc = maxc == minc;
      const float cr = maxc - minc;
      const float s = __fdiv_rn(cr, eqc ? 1.f : maxc);
      const float div = eqc ? 1.f : cr;
      const float rc = __fdiv_rn(maxc - r, div), gc = __fdiv_rn(maxc - g, div), bc = __fdiv_rn(maxc - b, div);
      const float hr = maxc == r ? bc - gc : 0.f;
      const float hg = (maxc == g && maxc != r) ? __fadd_rn(2.0f + rc, -bc) : 0.f;
      const float hb = (maxc != g && maxc != r) ? __fadd_rn(4.0f + gc, -rc) : 0.f;
      float h = __fadd_rn(__fadd_rn(hr, hg), hb);
      h = fmodf(__fadd_rn(__fdiv_rn(h, 6.0f), 1.0f), 1.0f);
      // h = (h + f) % 1.0  (Python semantics: result in [0, 1))
      h = __fadd_rn(h, a.f);
      h = h - floorf(h);
      // _hsv2rgb
      const float v = maxc;
      const float h6 = __fmul_rn(h, 6.0f);
      const float fi = floorf(h6);
      const float f = h6 - fi;
      int ii = (int)fi % 6;
      if (ii < 0) ii += 6;
      const float p = fminf(fmaxf(__fmul_rn(v, 1.0f - s), 0.f), 1.f);
      const float q = fminf(fmaxf(__fmul_rn(v, 1.0f - __fmul_rn(s, f)), 0.f), 1.f);
      const float t = fminf(fmaxf(__fmul_rn(v, 1.0f - __fmul_rn(s, 1.0f - f)), 0.f), 1.f);
      switch (ii) {
        case 0: r = v; g = t; b = p; break;
        case 1: r = q; g = v; b = p; break;
        case 2: r = p; g = v; b = t; break;
        case 3: r = p; g = q; b = v; break;
        case 4: r = t; g = p; b = v; break;
        default: r = v; g = p; b = q; break;
      }
    }
    img[i] = r; img[n + i] = g; img[2 * n + i] = b;
  }
}

// depthwise k x k gaussian (k = 3 or 5, weights w1[k], kernel2d = outer product) with reflect padding
struct Blur { float w[5]; int k; };
__global__ void __launch_bounds__(256) blur_kernel(const float* __restrict__ in, float* __restrict__ out, int S, Blur bl) {
  pdl_sync();
  const long long total = 3LL * S * S;
  const int r = bl.k / 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % S), y = (int)((i / S) % S);
    const float* plane = in + (i / ((long long)S * S)) * S * S;
    float acc = 0.f;
    for (int dy = -r; dy <= r; ++dy) {
      int yy = y + dy;
      yy = yy < 0 ? -yy : (yy >= S ? 2 * S - 2 - yy : yy);
      for (int dx = -r; dx <= r; ++dx) {
        int xx = x + dx;
        xx = xx < 0 ? -xx : (xx >= S ? 2 * S - 2 - xx : xx);
        acc = fmaf(__fmul_rn(bl.w[dy + r], bl.w[dx + r]), plane[(long long)yy * S + xx], acc);
      }
    }
    out[i] = acc;
  }
}

static int grid_for(long long n) {
  long long g = (n + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace aug

extern "C" {

// ResizeLongestSideAndPad (dataset.py:34-143) from the uint8 sources: mode 0 = pad the [H, W] source to [ph, pw] with
// (off_y, off_x) = (pad_top, pad_left) and white fill (label 0), mode 1 = crop [ph, pw] at (off_y, off_x); then resize to
// [new_h, new_w] (image: antialiased bilinear, label: nearest) and centre in S x S with zeros.  tmp: 3 * ph * new_w floats.
int s2u_aug_resize_pad(const unsigned char* img, const unsigned char* lab, int H, int W, int mode, int off_y, int off_x,
                       int ph, int pw, int S, int new_h, int new_w, int pad_left, int pad_top, float* tmp, float* out_img,
                       float* out_lab, void* stream) {
  if (H <= 0 || W <= 0 || ph <= 0 || pw <= 0 || S <= 0 || new_h <= 0 || new_w <= 0 || new_h > S || new_w > S ||
      pad_left < 0 || pad_top < 0 || pad_left + new_w > S || pad_top + new_h > S || off_y < 0 || off_x < 0)
    return S2U_EINVAL;
  if (mode == 1 && (off_y + ph > H || off_x + pw > W)) return S2U_EINVAL;
  const aug::AaAxis ax = aug::aa_axis(pw, new_w), ay = aug::aa_axis(ph, new_h);
  const aug::Src s{img, lab, H, W, mode, off_y, off_x};
  cudaStream_t st = (cudaStream_t)stream;
  S2U_LAUNCH(aug::resize_h_kernel, aug::grid_for((long long)ph * new_w), 256, 0, st, s, ph, ax, tmp);
  S2U_LAUNCH_CHECK();
  S2U_LAUNCH(aug::resize_v_kernel, aug::grid_for((long long)S * S), 256, 0, st, (const float*)tmp, s, ph, pw, ay, new_w, S,
             pad_left, pad_top, (float)ph / (float)new_h, (float)pw / (float)new_w, out_img, out_lab);
  S2U_LAUNCH_CHECK();
  return 0;
}

// out = rot90(in, k) for a [C, S, S] map, k in 1..3 (RandomRotate, dataset.py:156-173)
int s2u_aug_rot90(const float* in, float* out, int C, int S, int k, void* stream) {
  if (C <= 0 || S <= 0 || k < 1 || k > 3) return S2U_EINVAL;
  S2U_LAUNCH(aug::rot90_kernel, aug::grid_for((long long)C * S * S), 256, 0, (cudaStream_t)stream, in, out, C, S, k);
  S2U_LAUNCH_CHECK();
  return 0;
}

// In-place colour transform of a [3, S, S] image (dataset.py:176-262, 146-153): op 0 grayscale (3 channels), 1 brightness,
// 2 contrast (ws: one double of device scratch), 3 saturation, 4 hue, 5 gamma, 6 normalise (m3 / s3: host arrays of 3
// floats).  factor_c = fp32(1.0 - factor) for the blends (brightness / contrast / saturation).
int s2u_aug_color(float* img, int S, int op, float factor, float factor_c, double* ws, const float* m3_host,
                  const float* s3_host, void* stream) {
  if (S <= 0 || op < 0 || op > 6 || (op == 2 && !ws) || (op == 6 && (!m3_host || !s3_host))) return S2U_EINVAL;
  const long long n = (long long)S * S;
  cudaStream_t st = (cudaStream_t)stream;
  aug::OpArgs a{};
  a.f = factor;
  a.fc = factor_c;
  a.gray_sum = ws;
  if (op == 6)
    for (int c = 0; c < 3; ++c) { a.m[c] = m3_host[c]; a.s[c] = s3_host[c]; }
  if (op == 2) {
    cudaError_t ce = cudaMemsetAsync(ws, 0, sizeof(double), st);
    if (ce != cudaSuccess) return (int)ce;
    S2U_LAUNCH(aug::gray_sum_kernel, aug::grid_for(n), 256, 0, st, (const float*)img, n, ws);
    S2U_LAUNCH_CHECK();
  }
  S2U_LAUNCH(aug::color_kernel, aug::grid_for(n), 256, 0, st, img, n, op, a);
  S2U_LAUNCH_CHECK();
  return 0;
}

// out = gaussian_blur(in, [k, k]) of a [3, S, S] image, k = 3 or 5, w1_host = the k normalised 1-D weights
int s2u_aug_blur(const float* in, float* out, int S, int k, const float* w1_host, void* stream) {
  if (S <= 2 || (k != 3 && k != 5) || !w1_host) return S2U_EINVAL;
  aug::Blur bl{};
  bl.k = k;
  for (int i = 0; i < k; ++i) bl.w[i] = w1_host[i];
  S2U_LAUNCH(aug::blur_kernel, aug::grid_for(3LL * S * S), 256, 0, (cudaStream_t)stream, in, out, S, bl);
  S2U_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
