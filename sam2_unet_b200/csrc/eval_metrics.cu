// Segmentation metrics of the reference's eval.py on the device (/root/reference/eval.py:55-171): pixel counts for the
// semantic IoU / Dice, 8-connected component labelling of the thresholded prediction and ground truth
// (skimage.measure.label with its default full connectivity, eval.py:105-106), component areas and the
// (prediction component, ground-truth component) intersection counts that the instance matching needs.  All integer
// work: results are exact.  The greedy matching itself (eval.py:120-160, a few dozen components) stays on the host.
#include "common.cuh"

// counts[0..3] += |P & G|, |P | G|, |P|, |G|   with P = pred > thr, G = gt > thr  (eval.py:85-101)
__global__ void __launch_bounds__(256) seg_counts_kernel(const unsigned char* __restrict__ pred,
                                                        const unsigned char* __restrict__ gt, long long n, float thr,
                                                        unsigned long long* __restrict__ counts) {
  pdl_sync();
  unsigned int c[4] = {0, 0, 0, 0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const bool p = (float)pred[i] > thr, g = (float)gt[i] > thr;
    c[0] += p && g;
    c[1] += p || g;
    c[2] += p;
    c[3] += g;
  }
  __shared__ unsigned int sh[4];
  if (threadIdx.x < 4) sh[threadIdx.x] = 0;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    unsigned int v = c[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&sh[k], v);
  }
  __syncthreads();
  if (threadIdx.x < 4 && sh[threadIdx.x]) atomicAdd(counts + threadIdx.x, (unsigned long long)sh[threadIdx.x]);
}

// ---- connected components: lock-free union-find, the root of a component is its smallest raster index (so sorting
// the roots gives skimage's label order: components numbered by their first pixel in raster order)
__device__ __forceinline__ int cc_find(const int* parent, int x) {
  while (true) {
    const int p = reinterpret_cast<const volatile int*>(parent)[x];   // other threads re-parent concurrently
    if (p == x) return x;
    x = p;
  }
}
__device__ __forceinline__ void cc_union(int* parent, int a, int b) {
  while (true) {
    a = cc_find(parent, a);
    b = cc_find(parent, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }        // a > b: hang the larger root under the smaller one
    const int old = atomicMin(parent + a, b);
    if (old == a) return;
    a = old;                                              // someone re-parented a meanwhile: merge that chain too
  }
}
__global__ void cc_init_kernel(const unsigned char* __restrict__ mask, float thr, int* __restrict__ labels,
                               long long n) {
  pdl_sync();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    labels[i] = (float)mask[i] > thr ? (int)i : -1;
}
// every foreground pixel joins its W, NW, N, NE foreground neighbours (the other four directions are covered from the
// neighbour's side)
__global__ void cc_merge_kernel(int* __restrict__ labels, int H, int W) {
  pdl_sync();
  const long long n = (long long)H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (labels[i] < 0) continue;
    const int y = (int)(i / W), x = (int)(i % W);
    if (x > 0 && labels[i - 1] >= 0) cc_union(labels, (int)i, (int)i - 1);
    if (y > 0) {
      const long long up = i - W;
      if (labels[up] >= 0) cc_union(labels, (int)i, (int)up);
      if (x > 0 && labels[up - 1] >= 0) cc_union(labels, (int)i, (int)up - 1);
      if (x + 1 < W && labels[up + 1] >= 0) cc_union(labels, (int)i, (int)up + 1);
    }
  }
}
// path compression to the root; the roots are appended (unordered) to `roots`, their number to *nroots
__global__ void cc_compress_kernel(int* __restrict__ labels, long long n, int* __restrict__ roots,
                                   int* __restrict__ nroots, int cap) {
  pdl_sync();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (labels[i] < 0) continue;
    const int r = cc_find(labels, (int)i);
    if (r == (int)i) {
      const int k = atomicAdd(nroots, 1);
      if (k < cap) roots[k] = r;
    }
  }
}
__global__ void cc_flatten_kernel(int* __restrict__ labels, long long n) {
  pdl_sync();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (labels[i] >= 0) labels[i] = cc_find(labels, (int)i);   // roots are fixed points, so concurrent rewrites are safe
}

// areas and pairwise intersections.  parea / garea: [n] ints indexed by root (zero elsewhere).  Intersections go to an
// open-addressing hash table keys[cap] (0xFFFF... = empty) / vals[cap], cap a power of two.
__global__ void cc_stats_kernel(const int* __restrict__ pl, const int* __restrict__ gl, long long n,
                                int* __restrict__ parea, int* __restrict__ garea,
                                unsigned long long* __restrict__ keys, int* __restrict__ vals, int cap,
                                int* __restrict__ overflow) {
  pdl_sync();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int p = pl[i], g = gl[i];
    if (p >= 0) atomicAdd(parea + p, 1);
    if (g >= 0) atomicAdd(garea + g, 1);
    if (p < 0 || g < 0) continue;
    const unsigned long long key = ((unsigned long long)(unsigned)p << 32) | (unsigned)g;
    unsigned long long h = key * 0x9E3779B97F4A7C15ull;
    int slot = (int)(h >> 40) & (cap - 1);
    bool done = false;
    for (int probe = 0; probe < cap; ++probe) {
      const unsigned long long cur = atomicCAS(keys + slot, 0xFFFFFFFFFFFFFFFFull, key);
      if (cur == 0xFFFFFFFFFFFFFFFFull || cur == key) {
        atomicAdd(vals + slot, 1);
        done = true;
        break;
      }
      slot = (slot + 1) & (cap - 1);
    }
    if (!done) *overflow = 1;
  }
}

static inline int eval_grid(long long n) {
  long long g = (n + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  if (g < 1) g = 1;
  return (int)g;
}

extern "C" {

int s2u_seg_counts(const unsigned char* pred, const unsigned char* gt, long long n, float threshold,
                   unsigned long long* counts, void* stream) {
  if (n <= 0) return n == 0 ? 0 : S2U_EINVAL;
  S2U_LAUNCH(seg_counts_kernel, eval_grid(n), 256, 0, (cudaStream_t)stream, pred, gt, n, threshold, counts);
  S2U_LAUNCH_CHECK();
  return 0;
}

// labels[H*W]: root (smallest raster index of the 8-connected component) or -1; roots[cap] / nroots: the component
// roots in no particular order (*nroots must be 0 on entry; more than cap components -> *nroots > cap, list truncated)
int s2u_cc_label(const unsigned char* mask, float threshold, int H, int W, int* labels, int* roots, int* nroots,
                 int cap, void* stream) {
  if (H <= 0 || W <= 0 || (long long)H * W > 0x7fffffffLL || cap <= 0) return S2U_EINVAL;
  const long long n = (long long)H * W;
  cudaStream_t st = (cudaStream_t)stream;
  S2U_LAUNCH(cc_init_kernel, eval_grid(n), 256, 0, st, mask, threshold, labels, n);
  S2U_LAUNCH_CHECK();
  S2U_LAUNCH(cc_merge_kernel, eval_grid(n), 256, 0, st, labels, H, W);
  S2U_LAUNCH_CHECK();
  S2U_LAUNCH(cc_compress_kernel, eval_grid(n), 256, 0, st, labels, n, roots, nroots, cap);
  S2U_LAUNCH_CHECK();
  S2U_LAUNCH(cc_flatten_kernel, eval_grid(n), 256, 0, st, labels, n);
  S2U_LAUNCH_CHECK();
  return 0;
}

// parea / garea [n] and vals [cap] zero on entry, keys [cap] all-ones on entry, cap a power of two
int s2u_cc_stats(const int* plabels, const int* glabels, long long n, int* parea, int* garea, unsigned long long* keys,
                 int* vals, int cap, int* overflow, void* stream) {
  if (n <= 0 || cap <= 0 || (cap & (cap - 1))) return S2U_EINVAL;
  S2U_LAUNCH(cc_stats_kernel, eval_grid(n), 256, 0, (cudaStream_t)stream, plabels, glabels, n, parea, garea, keys, vals,
             cap, overflow);
  S2U_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
