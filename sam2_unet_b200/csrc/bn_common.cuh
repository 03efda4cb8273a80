// BatchNorm pieces shared by bn.cu and the convolution kernel whose epilogue produces the batch statistics
// (conv_igemm.cu): the replicated fp64 accumulator workspace, the per-channel finalisation and the "last block" ticket.
#pragma once
#include "common.cuh"

constexpr int BN_NREP = 16;   // replicated accumulator sets
constexpr int BN_UNROLL = 4;  // independent row loads in flight per thread

// sum of one accumulator over the replicas; clears them
__device__ __forceinline__ double rep_sum_clear(volatile double* sums, int i, int C2) {
  double t = 0.0;
#pragma unroll
  for (int r = 0; r < BN_NREP; ++r) {
    t += sums[r * C2 + i];
    sums[r * C2 + i] = 0.0;
  }
  return t;
}

struct BnFin {
  const float* gamma; const float* beta; float* running_mean; float* running_var; long long* num_batches;
  float* scale; float* shift; float* save_mean; float* save_rstd; float eps, momentum;
};

// batch statistics -> (scale, shift) for the apply pass, saved (mean, rstd) for backward, running-stat update;
// clears the accumulators.  Executed by ONE block (the standalone kernel or the last block of the statistics pass).
__device__ __forceinline__ void bn_finalize_block(volatile double* sums, const BnFin& f, long long M, int C,
                                                  int training) {
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mean, rstd;
    if (training) {
      const double m = rep_sum_clear(sums, c, 2 * C) / (double)M;
      double var = rep_sum_clear(sums, C + c, 2 * C) / (double)M - m * m;
      if (var < 0) var = 0;
      mean = (float)m;
      rstd = (float)(1.0 / sqrt(var + (double)f.eps));
      const double unbiased = M > 1 ? var * (double)M / (double)(M - 1) : var;
      f.running_mean[c] = (1.f - f.momentum) * f.running_mean[c] + f.momentum * mean;
      f.running_var[c] = (1.f - f.momentum) * f.running_var[c] + f.momentum * (float)unbiased;
    } else {
      mean = f.running_mean[c];
      rstd = rsqrtf(f.running_var[c] + f.eps);
    }
    const float sc = f.gamma[c] * rstd;
    f.scale[c] = sc;
    f.shift[c] = f.beta[c] - mean * sc;
    if (f.save_mean) f.save_mean[c] = mean;
    if (f.save_rstd) f.save_rstd[c] = rstd;
  }
  if (training && f.num_batches && threadIdx.x == 0) *f.num_batches += 1;
}

// "last block done": true in exactly one block of the grid, after every block's atomics are visible to it.
__device__ __forceinline__ bool last_block(unsigned int* ticket) {
  __shared__ bool last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ticket, 1u);
    last = (t == gridDim.x - 1);
    if (last) *ticket = 0u;
  }
  __syncthreads();
  if (last) __threadfence();
  return last;
}

