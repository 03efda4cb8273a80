// BatchNorm2d (eps 1e-5, momentum 0.1) around the RFB / decoder convolutions, on NHWC rows [M = B*H*W, C].
// Reference: BasicConv2d = bn(conv(x)) with NO activation (/root/reference/SAM2UNet.py:68-86), DoubleConv =
// conv-bn-relu twice (SAM2UNet.py:9-26), RFB tail relu(bn(conv_cat) + bn(conv_res)) (SAM2UNet.py:117-125).
// Training mode normalises with the biased batch variance and updates the running statistics with the
// unbiased one; eval mode uses the running statistics.  Per-channel sums are accumulated in fp64 so the
// E[x^2]-E[x]^2 form loses nothing against ATen's Welford pass.
//
// Launch count matters here (66 layers, most of them tiny): the per-channel finalisation runs inside the reduction
// kernel, in whichever block finishes last ("threadfence reduction"), so training forward is 2 launches per layer
// (statistics+finalize, apply) and backward is 2 (reduce+finalize, apply).
//
// The fp64 workspace `sums` holds BN_NREP replicas of the 2C accumulators followed by one ticket word
// (s2u_bn_ws_doubles(C) doubles); it is left zeroed for the next use.  Block b adds into replica b % BN_NREP: with a
// single set, the ~1,500 blocks of a 88x88x12 map all add into the same 128 addresses and L2 serialises them
// (measured 29 us for a 12 MB pass); the reduction grids are also persistent (<= 4 blocks per SM, grid-stride over
// rows with four loads in flight per thread) so that each block contributes once.
#include "common.cuh"
#include "bn_common.cuh"

__global__ void bn_finalize_kernel(double* __restrict__ sums, BnFin f, long long M, int C, int training) {
  pdl_sync();
  bn_finalize_block(sums, f, M, C, training);
}

// dgamma/dbeta accumulate into the parameter gradients; per-channel coefficients for the backward apply pass
__device__ __forceinline__ void bn_bwd_finalize_block(volatile double* sums, float* dgamma, float* dbeta, float* c1,
                                                      float* c2, long long M, int C) {
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double sb = rep_sum_clear(sums, c, 2 * C), sg = rep_sum_clear(sums, C + c, 2 * C);
    dbeta[c] += (float)sb;
    dgamma[c] += (float)sg;
    c1[c] = (float)(sb / (double)M);
    c2[c] = (float)(sg / (double)M);
  }
}

// block-level column reduction of per-thread partials (s, q) into the fp64 accumulators
__device__ __forceinline__ void reduce_to_sums(float* red, const float* s, const float* q, double* sums, int C,
                                               int g, int sub, int nsub) {
  if (sub < nsub) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      red[sub * 2 * C + g * 8 + j] = s[j];
      red[sub * 2 * C + C + g * 8 + j] = q[j];
    }
  }
  __syncthreads();
  double* rep = sums + (size_t)(blockIdx.x % BN_NREP) * 2 * C;
  for (int i = threadIdx.x; i < 2 * C; i += 256) {
    float t = 0.f;
    for (int k = 0; k < nsub; ++k) t += red[k * 2 * C + i];
    atomicAdd(rep + i, (double)t);
  }
}

// sums[0..C) += sum x, sums[C..2C) += sum x^2 ; optionally finalises in the last block
template <typename T>
__global__ void __launch_bounds__(256) bn_stats_kernel(const T* __restrict__ x, int ldx, double* __restrict__ sums,
                                                      long long M, int C, int fuse_finalize, BnFin fin) {
  pdl_sync();
  extern __shared__ float red[];              // [row lanes][2*C]
  const int cg = C >> 3;                      // 8-channel groups
  const int nsub = 256 / cg;                  // row lanes per block
  const int g = threadIdx.x % cg, sub = threadIdx.x / cg;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
  if (sub < nsub) {
    const long long step = (long long)gridDim.x * nsub;
    long long r = (long long)blockIdx.x * nsub + sub;
    for (; r + (BN_UNROLL - 1) * step < M; r += BN_UNROLL * step) {
      F8 v[BN_UNROLL];
#pragma unroll
      for (int u = 0; u < BN_UNROLL; ++u) v[u] = ld8(x + (r + u * step) * ldx + g * 8);
#pragma unroll
      for (int u = 0; u < BN_UNROLL; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j] += v[u].v[j]; q[j] = fmaf(v[u].v[j], v[u].v[j], q[j]); }
    }
    for (; r < M; r += step) {
      const F8 v = ld8(x + r * ldx + g * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] += v.v[j]; q[j] = fmaf(v.v[j], v.v[j], q[j]); }
    }
  }
  reduce_to_sums(red, s, q, sums, C, g, sub, nsub);
  if (fuse_finalize && last_block(reinterpret_cast<unsigned int*>(sums + BN_NREP * 2 * C))) bn_finalize_block(sums, fin, M, C, 1);
}

// out = act(x * scale + shift (+ resid)); out may be a channel slice of a wider (concat) buffer via ld_out
template <typename T>
__global__ void bn_apply_kernel(const T* __restrict__ x, int ldx, const float* __restrict__ scale,
                                const float* __restrict__ shift, const T* __restrict__ resid, int ld_res,
                                T* __restrict__ out, int ld_out, long long M, int C, int relu) {
  pdl_sync();
  const int cg = C >> 3;
  const long long total = M * cg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    const long long r = i / cg;
    F8 v = ld8(x + r * ldx + g * 8);
    const F8 sc = ld8(scale + g * 8), sh = ld8(shift + g * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) v.v[j] = fmaf(v.v[j], sc.v[j], sh.v[j]);
    if (resid) {
      const F8 rr = ld8(resid + r * ld_res + g * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) v.v[j] += rr.v[j];
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v.v[j] = fmaxf(v.v[j], 0.f);
    }
    st8(out + r * ld_out + g * 8, v);
  }
}

// g = dy * (y > 0): gradient through a ReLU given its OUTPUT y
template <typename T>
__global__ void relu_bwd_kernel(const T* __restrict__ dy, int ld_dy, const T* __restrict__ y, int ld_y,
                                T* __restrict__ g, int ld_g, long long M, int C) {
  pdl_sync();
  const int cg = C >> 3;
  const long long total = M * cg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cg);
    const long long r = i / cg;
    const F8 d = ld8(dy + r * ld_dy + c * 8), yy = ld8(y + r * ld_y + c * 8);
    F8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) o.v[j] = yy.v[j] > 0.f ? d.v[j] : 0.f;
    st8(g + r * ld_g + c * 8, o);
  }
}

// sums[0..C) += sum g, sums[C..2C) += sum g * xhat, with g = dy * (y > 0 if y given); finalises in the last block
template <typename T>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const T* __restrict__ dy, int ld_dy,
                                                           const T* __restrict__ y, int ld_y,
                                                           const T* __restrict__ x, int ldx,
                                                           const float* __restrict__ mean,
                                                           const float* __restrict__ rstd,
                                                           double* __restrict__ sums, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta, float* __restrict__ c1,
                                                           float* __restrict__ c2, long long M, int C) {
  pdl_sync();
  extern __shared__ float red[];
  const int cg = C >> 3;
  const int nsub = 256 / cg;
  const int g = threadIdx.x % cg, sub = threadIdx.x / cg;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
  if (sub < nsub) {
    const F8 mu = ld8(mean + g * 8), rs = ld8(rstd + g * 8);
    const long long step = (long long)gridDim.x * nsub;
    long long r = (long long)blockIdx.x * nsub + sub;
    auto acc = [&](F8 d, const F8& v, const F8& yy) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float dj = yy.v[j] > 0.f ? d.v[j] : 0.f;
        s[j] += dj;
        q[j] = fmaf(dj, (v.v[j] - mu.v[j]) * rs.v[j], q[j]);
      }
    };
    F8 one;
#pragma unroll
    for (int j = 0; j < 8; ++j) one.v[j] = 1.f;
    for (; r + (BN_UNROLL - 1) * step < M; r += BN_UNROLL * step) {
      F8 d[BN_UNROLL], v[BN_UNROLL], yy[BN_UNROLL];
#pragma unroll
      for (int u = 0; u < BN_UNROLL; ++u) {
        const long long ru = r + u * step;
        d[u] = ld8(dy + ru * ld_dy + g * 8);
        v[u] = ld8(x + ru * ldx + g * 8);
        yy[u] = y ? ld8(y + ru * ld_y + g * 8) : one;
      }
#pragma unroll
      for (int u = 0; u < BN_UNROLL; ++u) acc(d[u], v[u], yy[u]);
    }
    for (; r < M; r += step) acc(ld8(dy + r * ld_dy + g * 8), ld8(x + r * ldx + g * 8), y ? ld8(y + r * ld_y + g * 8) : one);
  }
  reduce_to_sums(red, s, q, sums, C, g, sub, nsub);
  if (last_block(reinterpret_cast<unsigned int*>(sums + BN_NREP * 2 * C))) bn_bwd_finalize_block(sums, dgamma, dbeta, c1, c2, M, C);
}

// dx = gamma * rstd * (g - c1 - xhat * c2)
template <typename T>
__global__ void bn_bwd_apply_kernel(const T* __restrict__ dy, int ld_dy, const T* __restrict__ y, int ld_y,
                                    const T* __restrict__ x, int ldx, const float* __restrict__ mean,
                                    const float* __restrict__ rstd, const float* __restrict__ gamma,
                                    const float* __restrict__ c1, const float* __restrict__ c2, T* __restrict__ dx,
                                    int ld_dx, long long M, int C) {
  pdl_sync();
  const int cg = C >> 3;
  const long long total = M * cg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    const long long r = i / cg;
    F8 d = ld8(dy + r * ld_dy + g * 8);
    const F8 v = ld8(x + r * ldx + g * 8);
    if (y) {
      const F8 yy = ld8(y + r * ld_y + g * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) d.v[j] = yy.v[j] > 0.f ? d.v[j] : 0.f;
    }
    const F8 mu = ld8(mean + g * 8), rs = ld8(rstd + g * 8), ga = ld8(gamma + g * 8), a = ld8(c1 + g * 8),
             b = ld8(c2 + g * 8);
    F8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (v.v[j] - mu.v[j]) * rs.v[j];
      o.v[j] = ga.v[j] * rs.v[j] * (d.v[j] - a.v[j] - xh * b.v[j]);
    }
    st8(dx + r * ld_dx + g * 8, o);
  }
}

// BatchNorm backward in ONE pass over the inputs (C = 64, bf16): a persistent grid of at most one CTA per SM keeps
// its rows' g = dy * (y > 0) and x in shared memory between the reduction and the apply phase, which are separated
// by a grid-wide barrier (cooperative launch: the grid - never more than one CTA per SM - starts only when all of it
// is resident, also when several of these kernels are in flight on different streams).
// Traffic: dy, y, x read once + dx written (the two-kernel path reads them twice).
// `sums` layout as above; the two words after the accumulators are the barrier counter and the exit ticket.
constexpr int BNF_THREADS = 512;
__global__ void __launch_bounds__(BNF_THREADS, 1)
bn_bwd_fused_kernel(const bf16* __restrict__ dy, int ld_dy, const bf16* __restrict__ y, int ld_y,
                    const bf16* __restrict__ x, int ldx, const float* __restrict__ mean, const float* __restrict__ rstd,
                    const float* __restrict__ gamma, double* __restrict__ sums, float* __restrict__ dgamma,
                    float* __restrict__ dbeta, bf16* __restrict__ dx, int ld_dx, long long M, int rows_per_cta) {
  constexpr int C = 64, CG = 8, NSUB = BNF_THREADS / CG;
  pdl_sync();
  extern __shared__ __align__(16) uint8_t bnf_smem[];
  uint4* gs = reinterpret_cast<uint4*>(bnf_smem);                      // [rows][8] masked gradient (8 bf16 each)
  uint4* xs = gs + (size_t)rows_per_cta * CG;                           // [rows][8] conv output
  float* red = reinterpret_cast<float*>(xs + (size_t)rows_per_cta * CG);   // [NSUB][2C], later coef[2C]
  const int g = threadIdx.x % CG, sub = threadIdx.x / CG;
  const long long r_beg = (long long)blockIdx.x * rows_per_cta;
  const long long r_end = min(M, r_beg + rows_per_cta);
  const F8 mu = ld8(mean + g * 8), rs = ld8(rstd + g * 8);
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
  for (long long r0 = r_beg + sub; r0 < r_end; r0 += 4 * NSUB) {
    uint4 d[4], v[4], yy[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long r = r0 + u * NSUB;
      if (r < r_end) {
        d[u] = *reinterpret_cast<const uint4*>(dy + r * ld_dy + g * 8);
        v[u] = *reinterpret_cast<const uint4*>(x + r * ldx + g * 8);
        if (y) yy[u] = *reinterpret_cast<const uint4*>(y + r * ld_y + g * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long r = r0 + u * NSUB;
      if (r >= r_end) continue;
      __nv_bfloat162* dh = reinterpret_cast<__nv_bfloat162*>(&d[u]);
      const __nv_bfloat162* vh = reinterpret_cast<const __nv_bfloat162*>(&v[u]);
      const __nv_bfloat162* yh = reinterpret_cast<const __nv_bfloat162*>(&yy[u]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float2 df = __bfloat1622float2(dh[e]);
        const float2 vf = __bfloat1622float2(vh[e]);
        if (y) {
          const float2 yf = __bfloat1622float2(yh[e]);
          if (!(yf.x > 0.f)) df.x = 0.f;
          if (!(yf.y > 0.f)) df.y = 0.f;
          dh[e] = __floats2bfloat162_rn(df.x, df.y);
        }
        s[2 * e] += df.x;
        s[2 * e + 1] += df.y;
        q[2 * e] = fmaf(df.x, (vf.x - mu.v[2 * e]) * rs.v[2 * e], q[2 * e]);
        q[2 * e + 1] = fmaf(df.y, (vf.y - mu.v[2 * e + 1]) * rs.v[2 * e + 1], q[2 * e + 1]);
      }
      gs[(r - r_beg) * CG + g] = d[u];
      xs[(r - r_beg) * CG + g] = v[u];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[sub * 2 * C + g * 8 + j] = s[j];
    red[sub * 2 * C + C + g * 8 + j] = q[j];
  }
  __syncthreads();
  double* rep = sums + (size_t)(blockIdx.x % BN_NREP) * 2 * C;
  if (threadIdx.x < 2 * C) {
    float t = 0.f;
    for (int k = 0; k < NSUB; ++k) t += red[k * 2 * C + threadIdx.x];
    atomicAdd(rep + threadIdx.x, (double)t);
  }
  // ---- grid barrier
  unsigned int* bar = reinterpret_cast<unsigned int*>(sums + BN_NREP * 2 * C);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(bar, 1u);
    const long long t0 = clock64();
    while (*reinterpret_cast<volatile unsigned int*>(bar) < gridDim.x) {
      if (clock64() - t0 > 4000000000LL) __trap();            // a missing CTA must fail loudly, not hang
    }
    __threadfence();
  }
  __syncthreads();
  // ---- totals -> per-channel coefficients (every CTA), parameter gradients (the last CTA to get here)
  float* coef = red;                                           // [0,C): sum g / M, [C,2C): sum g xhat / M
  if (threadIdx.x < 2 * C) {
    double t = 0.0;
    const volatile double* vs = sums;
#pragma unroll
    for (int r = 0; r < BN_NREP; ++r) t += vs[r * 2 * C + threadIdx.x];
    coef[threadIdx.x] = (float)(t / (double)M);
    red[2 * C + threadIdx.x] = (float)t;
  }
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicAdd(bar + 1, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {                                                  // everyone has read the totals: clear for the next use
    if (threadIdx.x < C) {
      dbeta[threadIdx.x] += red[2 * C + threadIdx.x];
      dgamma[threadIdx.x] += red[2 * C + C + threadIdx.x];
    }
    for (int i = threadIdx.x; i < BN_NREP * 2 * C; i += BNF_THREADS) sums[i] = 0.0;
    if (threadIdx.x == 0) { bar[0] = 0u; bar[1] = 0u; }
  }
  const F8 ga = ld8(gamma + g * 8);
  float k0[8], k1[8], k2[8];                                   // dx = k0 * g - k1 - k2 * (x - mean)
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float gr = ga.v[j] * rs.v[j];
    k0[j] = gr;
    k1[j] = gr * coef[g * 8 + j];
    k2[j] = gr * coef[C + g * 8 + j] * rs.v[j];
  }
  for (long long r = r_beg + sub; r < r_end; r += NSUB) {
    const uint4 d = gs[(r - r_beg) * CG + g], v = xs[(r - r_beg) * CG + g];
    const __nv_bfloat162* dh = reinterpret_cast<const __nv_bfloat162*>(&d);
    const __nv_bfloat162* vh = reinterpret_cast<const __nv_bfloat162*>(&v);
    uint4 o;
    __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 df = __bfloat1622float2(dh[e]), vf = __bfloat1622float2(vh[e]);
      oh[e] = __floats2bfloat162_rn(fmaf(k0[2 * e], df.x, -fmaf(k2[2 * e], vf.x - mu.v[2 * e], k1[2 * e])),
                                    fmaf(k0[2 * e + 1], df.y, -fmaf(k2[2 * e + 1], vf.y - mu.v[2 * e + 1], k1[2 * e + 1])));
    }
    *reinterpret_cast<uint4*>(dx + r * ld_dx + g * 8) = o;
  }
}

static inline int grid_for(long long n, int threads) {
  long long g = (n + threads - 1) / threads;
  if (g > 148LL * 16) g = 148LL * 16;
  if (g < 1) g = 1;
  return (int)g;
}
static inline bool bn_fused_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("S2U_BN_FUSED");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}
static inline bool bn_c_ok(int C) { return C >= 8 && (C & 7) == 0 && (C >> 3) <= 256; }
// reduction grid: every block gets at least BN_UNROLL rows per row lane, at most 4 blocks per SM
static inline int bn_reduce_grid(long long M, int C) {
  const int nsub = 256 / (C >> 3);
  long long g = (M + (long long)nsub * BN_UNROLL - 1) / ((long long)nsub * BN_UNROLL);
  if (g > 148 * 4) g = 148 * 4;
  if (g < 1) g = 1;
  return (int)g;
}

extern "C" {

// size of the fp64 workspace `sums` of the functions below (replicated accumulators + ticket)
int s2u_bn_ws_doubles(int C) { return BN_NREP * 2 * C + 2; }

// statistics only (the caller finalises with s2u_bn_finalize)
int s2u_bn_stats(const void* x, int ldx, double* sums, long long M, int C, int dtype, void* stream) {
  if (M <= 0 || !bn_c_ok(C) || (ldx & 7)) return S2U_EINVAL;
  const int nsub = 256 / (C >> 3);
  const size_t smem = (size_t)nsub * 2 * C * sizeof(float);
  BnFin none{};
  S2U_DISPATCH_T(dtype, {
    S2U_ALLOW_SMEM(bn_stats_kernel<T>);
    S2U_LAUNCH((bn_stats_kernel<T>), bn_reduce_grid(M, C), 256, smem, (cudaStream_t)stream, (const T*)x, ldx, sums, M, C, 0, none);
  })
  S2U_LAUNCH_CHECK();
  return 0;
}

// training forward in ONE launch: statistics + (last block) scale/shift, saved mean/rstd, running-stat update.
// `sums` must hold s2u_bn_ws_doubles(C) doubles, zero on entry; it is zero again on exit.
int s2u_bn_stats_finalize(const void* x, int ldx, double* sums, const float* gamma, const float* beta,
                          float* running_mean, float* running_var, long long* num_batches, float* scale, float* shift,
                          float* save_mean, float* save_rstd, long long M, int C, float eps, float momentum, int dtype,
                          void* stream) {
  if (M <= 0 || !bn_c_ok(C) || (ldx & 7)) return S2U_EINVAL;
  const int nsub = 256 / (C >> 3);
  const size_t smem = (size_t)nsub * 2 * C * sizeof(float);
  BnFin f{gamma, beta, running_mean, running_var, num_batches, scale, shift, save_mean, save_rstd, eps, momentum};
  S2U_DISPATCH_T(dtype, {
    S2U_ALLOW_SMEM(bn_stats_kernel<T>);
    S2U_LAUNCH((bn_stats_kernel<T>), bn_reduce_grid(M, C), 256, smem, (cudaStream_t)stream, (const T*)x, ldx, sums, M, C, 1, f);
  })
  S2U_LAUNCH_CHECK();
  return 0;
}

int s2u_bn_finalize(double* sums, const float* gamma, const float* beta, float* running_mean, float* running_var,
                    long long* num_batches, float* scale, float* shift, float* save_mean, float* save_rstd,
                    long long M, int C, float eps, float momentum, int training, void* stream) {
  if (C <= 0) return S2U_EINVAL;
  BnFin f{gamma, beta, running_mean, running_var, num_batches, scale, shift, save_mean, save_rstd, eps, momentum};
  S2U_LAUNCH((bn_finalize_kernel), 1, 256, 0, (cudaStream_t)stream, sums, f, M, C, training);
  S2U_LAUNCH_CHECK();
  return 0;
}

int s2u_bn_apply(const void* x, int ldx, const float* scale, const float* shift, const void* resid, int ld_res,
                 void* out, int ld_out, long long M, int C, int relu, int dtype, void* stream) {
  if (M <= 0 || !bn_c_ok(C) || (ldx & 7) || (ld_out & 7) || (resid && (ld_res & 7))) return S2U_EINVAL;
  S2U_DISPATCH_T(dtype, {
    S2U_LAUNCH((bn_apply_kernel<T>), grid_for(M * (C >> 3), 256), 256, 0, (cudaStream_t)stream, 
        (const T*)x, ldx, scale, shift, (const T*)resid, ld_res, (T*)out, ld_out, M, C, relu);
  })
  S2U_LAUNCH_CHECK();
  return 0;
}

int s2u_relu_bwd(const void* dy, int ld_dy, const void* y, int ld_y, void* g, int ld_g, long long M, int C, int dtype,
                 void* stream) {
  if (M <= 0 || (C & 7) || (ld_dy & 7) || (ld_y & 7) || (ld_g & 7)) return S2U_EINVAL;
  S2U_DISPATCH_T(dtype, {
    S2U_LAUNCH((relu_bwd_kernel<T>), grid_for(M * (C >> 3), 256), 256, 0, (cudaStream_t)stream, (const T*)dy, ld_dy, (const T*)y,
                                                                                   ld_y, (T*)g, ld_g, M, C);
  })
  S2U_LAUNCH_CHECK();
  return 0;
}

// y == null: plain BN backward; y given: backward through relu(bn(x)) using the saved output for the mask.
// `sums`: s2u_bn_ws_doubles(C) doubles, zero on entry and on exit.
int s2u_bn_bwd(const void* dy, int ld_dy, const void* y, int ld_y, const void* x, int ldx, const float* mean,
               const float* rstd, const float* gamma, double* sums, float* dgamma, float* dbeta, float* c1, float* c2,
               void* dx, int ld_dx, long long M, int C, int dtype, void* stream) {
  if (M <= 0 || !bn_c_ok(C) || (ldx & 7) || (ld_dy & 7) || (ld_dx & 7) || (y && (ld_y & 7))) return S2U_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == S2U_BF16 && C == 64 && bn_fused_enabled()) {
    // one pass: the rows stay in shared memory across a grid barrier (fits when M / #SMs rows x 256 B <= ~190 KB)
    static int sms = 0;
    if (sms == 0) {
      int dev = 0;
      cudaGetDevice(&dev);
      if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    int rows = (int)((M + sms - 1) / sms);
    if (rows < 64) rows = 64;
    const int grid = (int)((M + rows - 1) / rows);
    const size_t fsmem = (size_t)rows * 256 + (size_t)(BNF_THREADS / 8) * 2 * 64 * sizeof(float) + 3 * 64 * sizeof(float);
    if (fsmem <= 200 * 1024) {
      S2U_ALLOW_SMEM(bn_bwd_fused_kernel);
      // COOPERATIVE launch: the grid barrier needs every CTA resident at once.  Several of these kernels can be in
      // flight on different streams (the four RFB backwards); with plain launches two of them could each hold part
      // of the SMs and spin on CTAs that cannot be scheduled.  The cooperative attribute makes the hardware start
      // the grid only when all of it fits.  (No programmatic-dependent-launch attribute on this one.)
      cudaLaunchConfig_t lc = {};
      lc.gridDim = dim3(grid);
      lc.blockDim = dim3(BNF_THREADS);
      lc.dynamicSmemBytes = fsmem;
      lc.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeCooperative;
      at[0].val.cooperative = 1;
      lc.attrs = at;
      lc.numAttrs = 1;
      cudaError_t ce = cudaLaunchKernelEx(&lc, bn_bwd_fused_kernel, (const bf16*)dy, ld_dy, (const bf16*)y, ld_y,
                                          (const bf16*)x, ldx, mean, rstd, gamma, sums, dgamma, dbeta, (bf16*)dx, ld_dx,
                                          M, rows);
      if (ce != cudaSuccess) return (int)ce;
      S2U_LAUNCH_CHECK();
      return 0;
    }
  }
  const int nsub = 256 / (C >> 3);
  const size_t smem = (size_t)nsub * 2 * C * sizeof(float);
  S2U_DISPATCH_T(dtype, {
    S2U_ALLOW_SMEM(bn_bwd_reduce_kernel<T>);
    S2U_LAUNCH((bn_bwd_reduce_kernel<T>), bn_reduce_grid(M, C), 256, smem, st, (const T*)dy, ld_dy, (const T*)y, ld_y,
                                                                    (const T*)x, ldx, mean, rstd, sums, dgamma, dbeta,
                                                                    c1, c2, M, C);
  })
  S2U_LAUNCH_CHECK();
  S2U_DISPATCH_T(dtype, {
    S2U_LAUNCH((bn_bwd_apply_kernel<T>), grid_for(M * (C >> 3), 256), 256, 0, st, (const T*)dy, ld_dy, (const T*)y, ld_y,
                                                                       (const T*)x, ldx, mean, rstd, gamma, c1, c2,
                                                                       (T*)dx, ld_dx, M, C);
  })
  S2U_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
