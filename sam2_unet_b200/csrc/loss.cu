// structure_loss (/root/reference/train.py:21-29) forward and backward for up to three prediction maps that
// share one mask (train.py:76-79), and the fused AdamW step (train.py:48-52,83).
//
//   weit = 1 + 5 |avgpool31x31(mask) - mask|      (stride 1, zero pad 15, divisor always 961)
//   bce  = mean over the whole batch of BCE-with-logits  (the reference's reduce="none" is the legacy boolean
//          kwarg, i.e. reduction="mean"; the weighted ratio that follows returns the same scalar)
//   I_b  = sum sigma(p) m weit,  U_b = sum (sigma(p) + m) weit,  loss = bce + mean_b (1 - (I_b+1)/(U_b-I_b+1))
//
// Pass 1 (one CTA per 32x64 pixel tile): separable 31-tap box sum of the mask in shared memory, weit written
// once, per-image partial sums for every head by warp shuffles + one fp64 atomic per CTA and quantity.
// Pass 2: closed-form gradient.  Algorithmic traffic: 12 B/pixel/head (read logit,
// read mask, write grad) + 8 B/pixel for weit.
#include "common.cuh"

constexpr int LT_H = 32, LT_W = 64, LR = 15;
constexpr int MAXH = 3;

struct LossPtrs { const float* pred[MAXH]; float* grad[MAXH]; };

// sums layout (fp64): [head][b][0] = I, [head][b][1] = U, then bce[head] at offset nheads*B*2 + head
__global__ void __launch_bounds__(256) loss_fwd_kernel(LossPtrs p, const float* __restrict__ mask,
                                                      float* __restrict__ weit, double* __restrict__ sums, int B,
                                                      int H, int W, int nheads) {
  pdl_sync();
  __shared__ float tile[LT_H + 2 * LR][LT_W + 2 * LR + 2];   // mask halo tile
  __shared__ float hsum[LT_H + 2 * LR][LT_W + 1];            // horizontal 31-tap sums
  __shared__ float red[8][MAXH * 3];
  const int b = blockIdx.z;
  const int y0 = blockIdx.y * LT_H, x0 = blockIdx.x * LT_W;
  const float* mb = mask + (long long)b * H * W;
  for (int i = threadIdx.x; i < (LT_H + 2 * LR) * (LT_W + 2 * LR); i += 256) {
    const int ty = i / (LT_W + 2 * LR), tx = i - ty * (LT_W + 2 * LR);
    const int y = y0 + ty - LR, x = x0 + tx - LR;
    tile[ty][tx] = (y >= 0 && y < H && x >= 0 && x < W) ? mb[(long long)y * W + x] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (LT_H + 2 * LR) * LT_W; i += 256) {
    const int ty = i / LT_W, tx = i - ty * LT_W;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k <= 2 * LR; ++k) s += tile[ty][tx + k];
    hsum[ty][tx] = s;
  }
  __syncthreads();
  float acc[MAXH * 3];
#pragma unroll
  for (int k = 0; k < MAXH * 3; ++k) acc[k] = 0.f;
  for (int i = threadIdx.x; i < LT_H * LT_W; i += 256) {
    const int ty = i / LT_W, tx = i - ty * LT_W;
    const int y = y0 + ty, x = x0 + tx;
    if (y >= H || x >= W) continue;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k <= 2 * LR; ++k) s += hsum[ty + k][tx];
    const float m = tile[ty + LR][tx + LR];
    const float wv = 1.f + 5.f * fabsf(s * (1.f / 961.f) - m);
    const long long o = ((long long)b * H + y) * W + x;
    weit[o] = wv;
    for (int h = 0; h < nheads; ++h) {
      const float pv = p.pred[h][o];
      const float sg = 1.f / (1.f + __expf(-pv));
      acc[h * 3 + 0] += sg * m * wv;
      acc[h * 3 + 1] += (sg + m) * wv;
      acc[h * 3 + 2] += fmaxf(pv, 0.f) - pv * m + log1pf(__expf(-fabsf(pv)));
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < MAXH * 3; ++k) {
    const float v = warp_sum(acc[k]);
    if (lane == 0) red[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < nheads * 3) {
    float t = 0.f;
    for (int w8 = 0; w8 < 8; ++w8) t += red[w8][threadIdx.x];
    const int h = threadIdx.x / 3, q = threadIdx.x % 3;
    if (q < 2) atomicAdd(sums + ((long long)h * B + b) * 2 + q, (double)t);
    else atomicAdd(sums + (long long)nheads * B * 2 + h, (double)t);
  }
}


// ---- vectorised path (W % 4 == 0: every row starts 16-byte aligned) ------------------------------------------------
// Forward: one CTA = 16 rows x 128 columns.  The 31x31 box sum is separable AND sliding: a thread starts a 31-tap sum
// once and then slides it (one add, one subtract per pixel) - first vertically (lanes = consecutive columns), then
// horizontally over 16-pixel runs with lanes = rows of an array whose pitch is 1 (mod 32), so both passes are free of
// bank conflicts: ~5 shared-memory adds per pixel instead of 62.  The last phase is a pure row-major stream: float4
// loads of the (up to) three logit maps, float4 store of weit, one exp / rcp / log per head and pixel (sigmoid and
// softplus share exp(-|p|)), per-image sums by warp shuffles and one fp64 atomic per CTA and quantity.
constexpr int VT_H = 16, VT_W = 128, VHALO_H = VT_H + 2 * LR, VHALO_W = VT_W + 2 * LR;
constexpr int VPITCH = VHALO_W + 2;                                  // 160 floats: halo rows stay 16-byte aligned
constexpr int VS_PITCH = 161, BX_PITCH = 129;                        // 1 (mod 32): lanes = rows hit distinct banks
constexpr size_t VEC_SMEM = sizeof(float) * (VHALO_H * VPITCH + VT_H * VS_PITCH + VT_H * BX_PITCH);

__device__ __forceinline__ void sig_softplus(float pv, float& sg, float& sp) {
  // e = exp(-|p|): sigmoid(p) = p >= 0 ? 1/(1+e) : e/(1+e); softplus(-|p|) = log(1+e)
  const float e = __expf(-fabsf(pv));
  const float r = __frcp_rn(1.f + e);
  sg = pv >= 0.f ? r : e * r;
  sp = __logf(1.f + e);
}

__global__ void __launch_bounds__(256) loss_fwd_vec_kernel(LossPtrs p, const float* __restrict__ mask,
                                                          float* __restrict__ weit, double* __restrict__ sums, int B,
                                                          int H, int W, int nheads) {
  pdl_sync();
  extern __shared__ __align__(16) uint8_t vsm[];            // 48 KB: opted in by the launcher
  float (*tile)[VPITCH] = reinterpret_cast<float (*)[VPITCH]>(vsm);                        // mask + 15-pixel halo
  float* vsum = reinterpret_cast<float*>(vsm) + VHALO_H * VPITCH;                          // [16][VS_PITCH] column sums
  float* box = vsum + VT_H * VS_PITCH;                                                     // [16][BX_PITCH] 31x31 sums
  __shared__ float red[8][MAXH * 3];
  const int b = blockIdx.z;
  const int y0 = blockIdx.y * VT_H, x0 = blockIdx.x * VT_W;
  const float* mb = mask + (long long)b * H * W;
  // this thread's two 4-pixel groups of the last phase: issue their logit loads now, consume them after the filter
  constexpr int NG = VT_H * (VT_W / 4) / 256;
  float4 pre[NG][MAXH];
#pragma unroll
  for (int it = 0; it < NG; ++it) {
    const int i = threadIdx.x + it * 256;
    const int ty = i / (VT_W / 4), tx = (i - ty * (VT_W / 4)) * 4;
    const bool ok = y0 + ty < H && x0 + tx < W;
    const long long o = ((long long)b * H + y0 + ty) * W + x0 + tx;
#pragma unroll
    for (int h = 0; h < MAXH; ++h)
      pre[it][h] = (ok && h < nheads) ? __ldg(reinterpret_cast<const float4*>(p.pred[h] + o)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // halo tile: tile column c = image column x0 - 16 + c, in 4-pixel groups (16-byte aligned since x0 % 4 == 0).
  // All of a thread's loads are issued before the first store: a load -> store loop pays the DRAM latency per trip
  {
    constexpr int NV = (VHALO_H * (VPITCH / 4) + 255) / 256;
    float4 v[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int i = threadIdx.x + k * 256;
      const int ty = i / (VPITCH / 4), g = i - ty * (VPITCH / 4);
      const int y = y0 + ty - LR, x = x0 - 16 + 4 * g;
      v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      // W % 4 == 0 and x % 4 == 0: a group is all in or all out
      if (i < VHALO_H * (VPITCH / 4) && y >= 0 && y < H && x >= 0 && x + 3 < W)
        v[k] = __ldg(reinterpret_cast<const float4*>(mb + (long long)y * W + x));
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int i = threadIdx.x + k * 256;
      const int ty = i / (VPITCH / 4), g = i - ty * (VPITCH / 4);
      if (i < VHALO_H * (VPITCH / 4)) *reinterpret_cast<float4*>(&tile[ty][4 * g]) = v[k];
    }
  }
  __syncthreads();
  // vertical 31-tap sums of every halo column for the 16 output rows: (column, 8-row run), lanes = consecutive columns
  for (int i = threadIdx.x; i < VPITCH * 2; i += 256) {
    const int c = i % VPITCH, r0 = (i / VPITCH) * 8;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k <= 2 * LR; ++k) s += tile[r0 + k][c];
    vsum[r0 * VS_PITCH + c] = s;
#pragma unroll
    for (int j = 1; j < 8; ++j) {
      s += tile[r0 + j + 2 * LR][c] - tile[r0 + j - 1][c];
      vsum[(r0 + j) * VS_PITCH + c] = s;
    }
  }
  __syncthreads();
  // horizontal 31-tap sums: (row, 16-column run); a warp = 16 rows x 2 adjacent runs -> banks r + 16 * run + k, distinct
  if (threadIdx.x < VT_H * (VT_W / 16)) {
    const int r = threadIdx.x & 15, tx0 = (threadIdx.x >> 4) * 16;
    const float* v = vsum + r * VS_PITCH + tx0 + 1;         // window of output column tx: tile columns tx + 1 .. tx + 31
    float s = 0.f;
#pragma unroll
    for (int k = 0; k <= 2 * LR; ++k) s += v[k];
    box[r * BX_PITCH + tx0] = s;
#pragma unroll
    for (int j = 1; j < 16; ++j) {
      s += v[j + 2 * LR] - v[j - 1];
      box[r * BX_PITCH + tx0 + j] = s;
    }
  }
  __syncthreads();
  float acc[MAXH * 3];
#pragma unroll
  for (int k = 0; k < MAXH * 3; ++k) acc[k] = 0.f;
#pragma unroll
  for (int it = 0; it < NG; ++it) {                                     // 2 float4 groups per thread, row-major
    const int i = threadIdx.x + it * 256;
    const int ty = i / (VT_W / 4), tx = (i - ty * (VT_W / 4)) * 4;
    const int y = y0 + ty, x = x0 + tx;
    if (y >= H || x >= W) continue;                                    // W % 4 == 0: a group is all in or all out
    const float4 m4 = *reinterpret_cast<const float4*>(&tile[ty + LR][tx + 16]);
    const float mv[4] = {m4.x, m4.y, m4.z, m4.w};
    float wv[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) wv[e] = 1.f + 5.f * fabsf(box[ty * BX_PITCH + tx + e] * (1.f / 961.f) - mv[e]);
    const long long o = ((long long)b * H + y) * W + x;
    *reinterpret_cast<float4*>(weit + o) = make_float4(wv[0], wv[1], wv[2], wv[3]);
#pragma unroll
    for (int h = 0; h < MAXH; ++h) {
      if (h >= nheads) break;
      const float4 p4 = pre[it][h];
      const float pv[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float sg, sp;
        sig_softplus(pv[e], sg, sp);
        acc[h * 3 + 0] = fmaf(sg * mv[e], wv[e], acc[h * 3 + 0]);
        acc[h * 3 + 1] = fmaf(sg + mv[e], wv[e], acc[h * 3 + 1]);
        acc[h * 3 + 2] += fmaxf(pv[e], 0.f) - pv[e] * mv[e] + sp;
      }
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < MAXH * 3; ++k) {
    const float v = warp_sum(acc[k]);
    if (lane == 0) red[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < nheads * 3) {
    float t = 0.f;
    for (int w8 = 0; w8 < 8; ++w8) t += red[w8][threadIdx.x];
    const int h = threadIdx.x / 3, q = threadIdx.x % 3;
    if (q < 2) atomicAdd(sums + ((long long)h * B + b) * 2 + q, (double)t);
    else atomicAdd(sums + (long long)nheads * B * 2 + h, (double)t);
  }
}

// Backward, vectorised: 4 pixels per thread and iteration, everything float4 (H * W % 4 == 0 keeps a group inside one
// image); the per-image N, D of every head are hoisted out of the pixel loop when the whole block stays in one image
__global__ void __launch_bounds__(256) loss_bwd_vec_kernel(LossPtrs p, const float* __restrict__ mask,
                                                          const float* __restrict__ weit,
                                                          const double* __restrict__ sums,
                                                          const float* __restrict__ gscale, int B, int H, int W,
                                                          int nheads) {
  pdl_sync();
  const long long hw = (long long)H * W;
  const long long groups = (long long)B * hw / 4;
  const float inv_n = 1.f / (float)((long long)B * hw), inv_b = 1.f / (float)B;
  for (long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x; gi < groups;
       gi += (long long)gridDim.x * blockDim.x) {
    const long long i = gi * 4;
    const int b = (int)(i / hw);
    const float4 m4 = __ldg(reinterpret_cast<const float4*>(mask + i));
    const float4 w4 = __ldg(reinterpret_cast<const float4*>(weit + i));
    const float mv[4] = {m4.x, m4.y, m4.z, m4.w}, wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
    for (int h = 0; h < MAXH; ++h) {
      if (h >= nheads) break;
      const double I = sums[((long long)h * B + b) * 2], U = sums[((long long)h * B + b) * 2 + 1];
      const float N = (float)(I + 1.0), D = (float)(U - I + 1.0);
      const float gs = gscale ? gscale[h] : 1.f;
      const float k1 = gs * inv_n, k2 = gs * inv_b / (D * D);
      const float4 p4 = __ldg(reinterpret_cast<const float4*>(p.pred[h] + i));
      const float pv[4] = {p4.x, p4.y, p4.z, p4.w};
      float g[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float ex = __expf(-fabsf(pv[e]));
        const float r = __frcp_rn(1.f + ex);
        const float sg = pv[e] >= 0.f ? r : ex * r;
        // (sg - m)/(B H W) - (1/B) sg (1 - sg) w (m D - (1 - m) N) / D^2
        g[e] = (sg - mv[e]) * k1 - k2 * sg * (1.f - sg) * wv[e] * (mv[e] * D - (1.f - mv[e]) * N);
      }
      *reinterpret_cast<float4*>(p.grad[h] + i) = make_float4(g[0], g[1], g[2], g[3]);
    }
  }
}

// loss[h] = bce_h / (B H W) + (1/B) sum_b (1 - N_b / D_b)
__global__ void loss_finalize_kernel(const double* __restrict__ sums, float* __restrict__ loss, int B, int H, int W,
                                     int nheads) {
  pdl_sync();
  const int h = threadIdx.x;
  if (h >= nheads) return;
  double l = sums[(long long)nheads * B * 2 + h] / ((double)B * H * W);
  for (int b = 0; b < B; ++b) {
    const double I = sums[((long long)h * B + b) * 2], U = sums[((long long)h * B + b) * 2 + 1];
    l += (1.0 - (I + 1.0) / (U - I + 1.0)) / (double)B;
  }
  loss[h] = (float)l;
}

// grad = gscale[h] * [ (s - m)/(B H W) - (1/B) s (1-s) w (m D - (1-m) N) / D^2 ]
__global__ void loss_bwd_kernel(LossPtrs p, const float* __restrict__ mask, const float* __restrict__ weit,
                                const double* __restrict__ sums, const float* __restrict__ gscale, int B, int H,
                                int W, int nheads) {
  pdl_sync();
  const long long hw = (long long)H * W;
  const long long total = (long long)B * hw;
  const float inv_n = 1.f / (float)total, inv_b = 1.f / (float)B;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / hw);
    const float m = mask[i], wv = weit[i];
    for (int h = 0; h < nheads; ++h) {
      const double I = sums[((long long)h * B + b) * 2], U = sums[((long long)h * B + b) * 2 + 1];
      const float N = (float)(I + 1.0), D = (float)(U - I + 1.0);
      const float pv = p.pred[h][i];
      const float sg = 1.f / (1.f + __expf(-pv));
      const float g = (sg - m) * inv_n - inv_b * sg * (1.f - sg) * wv * (m * D - (1.f - m) * N) / (D * D);
      p.grad[h][i] = (gscale ? gscale[h] : 1.f) * g;
    }
  }
}

// AdamW on one flat fp32 buffer.  hyper (device, fp32): [0] lr, [1] gradient scale, [2] beta1^t, [3] beta2^t.
// The step counter lives on the device ([2], [3] are advanced by adamw_tick_kernel right before the update), so a
// captured CUDA graph replays the step without any host-side hyper-parameter traffic.
__global__ void adamw_tick_kernel(float* __restrict__ hyper, float beta1, float beta2) {
  pdl_sync();
  hyper[2] *= beta1;
  hyper[3] *= beta2;
}

__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, long long n, const float* __restrict__ hyper, float beta1,
                             float beta2, float eps, float wd) {
  pdl_sync();
  const float lr = hyper[0], gs = hyper[1], bc1 = 1.f - hyper[2], bc2 = 1.f - hyper[3];
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gv = g[i] * gs;
    float pv = p[i] * (1.f - lr * wd);
    const float mv = beta1 * m[i] + (1.f - beta1) * gv;
    const float vv = beta2 * v[i] + (1.f - beta2) * gv * gv;
    const float denom = sqrtf(vv) * inv_sqrt_bc2 + eps;
    pv -= step_size * mv / denom;
    p[i] = pv;
    m[i] = mv;
    v[i] = vv;
  }
}

extern "C" {

int s2u_structure_loss_fwd(const float* pred0, const float* pred1, const float* pred2, const float* mask, float* weit,
                           double* sums, float* loss, int B, int H, int W, int nheads, void* stream) {
  if (B <= 0 || H <= 0 || W <= 0 || nheads < 1 || nheads > MAXH) return S2U_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  LossPtrs p{{pred0, pred1, pred2}, {nullptr, nullptr, nullptr}};
  cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * ((size_t)nheads * B * 2 + nheads), st);
  if (e != cudaSuccess) return (int)e;
  const bool vec = (W % 4 == 0) && (((uintptr_t)pred0 | (uintptr_t)pred1 | (uintptr_t)pred2 | (uintptr_t)mask |
                                    (uintptr_t)weit) % 16 == 0);
  if (vec) {
    dim3 grid(ceil_div(W, VT_W), ceil_div(H, VT_H), B);
    constexpr size_t smem = VEC_SMEM;
    S2U_ALLOW_SMEM(loss_fwd_vec_kernel);
    S2U_LAUNCH((loss_fwd_vec_kernel), grid, 256, smem, st, p, mask, weit, sums, B, H, W, nheads);
  } else {
    dim3 grid(ceil_div(W, LT_W), ceil_div(H, LT_H), B);
    S2U_LAUNCH((loss_fwd_kernel), grid, 256, 0, st, p, mask, weit, sums, B, H, W, nheads);
  }
  S2U_LAUNCH_CHECK();
  S2U_LAUNCH((loss_finalize_kernel), 1, 32, 0, st, sums, loss, B, H, W, nheads);
  S2U_LAUNCH_CHECK();
  return 0;
}

int s2u_structure_loss_bwd(const float* pred0, const float* pred1, const float* pred2, const float* mask,
                           const float* weit, const double* sums, const float* gscale, float* grad0, float* grad1,
                           float* grad2, int B, int H, int W, int nheads, void* stream) {
  if (B <= 0 || H <= 0 || W <= 0 || nheads < 1 || nheads > MAXH) return S2U_EINVAL;
  LossPtrs p{{pred0, pred1, pred2}, {grad0, grad1, grad2}};
  long long total = (long long)B * H * W;
  const bool vec = (W % 4 == 0) && (((uintptr_t)pred0 | (uintptr_t)pred1 | (uintptr_t)pred2 | (uintptr_t)mask |
                                    (uintptr_t)weit | (uintptr_t)grad0 | (uintptr_t)grad1 | (uintptr_t)grad2) % 16 == 0);
  if (vec) {
    long long g = (total / 4 + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    S2U_LAUNCH((loss_bwd_vec_kernel), (int)g, 256, 0, (cudaStream_t)stream, p, mask, weit, sums, gscale, B, H, W, nheads);
  } else {
    long long g = (total + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    S2U_LAUNCH((loss_bwd_kernel), (int)g, 256, 0, (cudaStream_t)stream, p, mask, weit, sums, gscale, B, H, W, nheads);
  }
  S2U_LAUNCH_CHECK();
  return 0;
}

int s2u_adamw(float* p, const float* g, float* m, float* v, long long n, float* hyper, float beta1, float beta2,
              float eps, float wd, void* stream) {
  if (n <= 0) return S2U_EINVAL;
  long long gsz = (n + 255) / 256;
  if (gsz > 148 * 16) gsz = 148 * 16;
  S2U_LAUNCH((adamw_tick_kernel), 1, 1, 0, (cudaStream_t)stream, hyper, beta1, beta2);
  S2U_LAUNCH_CHECK();
  S2U_LAUNCH((adamw_kernel), (int)gsz, 256, 0, (cudaStream_t)stream, p, g, m, v, n, hyper, beta1, beta2, eps, wd);
  S2U_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
