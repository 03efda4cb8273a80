// Blackwell-native windowed / global attention for the Hiera trunk (bf16): tcgen05.mma with the score and output
// accumulators in tensor memory, operands staged by TMA, forward and backward.
// Reference semantics: /root/reference/sam2/modeling/backbones/hieradet.py:56-81 (MultiScaleAttention.forward) with
// window_partition / window_unpartition (/root/reference/sam2/modeling/backbones/utils.py:16-55) folded into the TMA
// box coordinates; same tensors and layouts as attention_mma.cu (qkv [B,H,W,3,nh,hd], out [B,H,W,nh,hd], lse
// [B,H,W,nh]).
//
// Window partition as TMA addressing: qkv is described to the TMA unit as a 5-D tensor (d, 3*nh, x, y, b).  A window's
// K (or V, or a tile of its Q rows) for one head is ONE box (64 | C1, 1, rw, rows, 1) at (d0, part*nh + head, wx*ww,
// wy*wh, b): the box lands in shared memory as [tokens of the window, row-major][64 | C1 head-dim columns] in exactly
// the 128-byte (32/64-byte for the head-dim remainder 64..hd) swizzled K-major layout tcgen05.mma reads.  Windows on
// the right / bottom border of the map are ragged; each of the (up to) four window shapes of a map has its own tensor
// maps whose box is the REAL extent of the window, so the tiles hold real tokens only.  The zero-padded tokens of the
// reference (q = k = v = bias, utils.py:30 + hieradet.py:59) all carry the same key and value: they are ONE extra key
// row (k = bias, v = bias) whose score gets + ln(n_pad) - the same softmax.  Head dim 72 is padded to 80 in shared
// memory only: the tensor map's d extent is hd, so columns 72..79 of the remainder box are out-of-bounds zero fill.
//
// Forward, one CTA = (window, head, tile of <= 128 queries), 128 threads, thread = query row = TMEM lane:
//   S = Q K^T          tcgen05.mma  (A, B from shared memory, M = 128, N = keys rounded up to 16 (<= 256), K = hd')
//   softmax            tcgen05.ld of the thread's row, no cross-thread reduction; P (bf16) written back OVER S with
//                      tcgen05.st, so it is the tensor-memory A operand of the next product
//   O (+)= P V         tcgen05.mma  (A from tensor memory, B = V tile read MN-major), O in tensor memory
// Windows (<= 256 keys) need one pass and no rescaling; global attention streams 128-key blocks (double-buffered TMA)
// with the online-softmax correction applied to O in tensor memory.
// 100 KB of shared memory and 256 TMEM columns per CTA -> two CTAs per SM overlap each other's TMA / MMA / softmax.
#include <unordered_map>

#include "common.cuh"
#include "umma.cuh"

namespace atc {

using namespace umma;

// -DS2U_ATC_TIMING: thread 0 of every forward CTA stamps its phases (globaltimer ns) into dws[8 * cta] (tuning only)
#ifdef S2U_ATC_TIMING
__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define ATC_STAMP(i) do { if (threadIdx.x == 0 && p.dws) reinterpret_cast<long long*>(p.dws)[8LL * (blockIdx.y * gridDim.x + blockIdx.x) + (i)] = gtime(); } while (0)
#define ATC_STAMP_T(i, thr) do { if (threadIdx.x == (thr) && p.dws) reinterpret_cast<long long*>(p.dws)[8LL * (blockIdx.y * gridDim.x + blockIdx.x) + (i)] = gtime(); } while (0)
#else
#define ATC_STAMP(i) do {} while (0)
#define ATC_STAMP_T(i, thr) do {} while (0)
#endif

constexpr int NTHR = 128;
constexpr int MAXCLS = 4;

struct TcClass {
  int rw, rh;              // real extent of the windows of this class (stream mode: rw = H*W, rh = 1)
  int wx0, wy0, cnx, cny;  // first window and number of windows (per image) of the class in x / y
  int n_real, n_pad;       // real and zero-padded tokens of a window
  int qbh, n_qt;           // window rows per query tile, query tiles per window
  int item0, n_items;      // work items (window, query tile) of the class: [item0, item0 + n_items)
  int n_kb, kitem0;        // backward: 128-key blocks per window, first (window, key block) work item of the class
  int witem0;              // fused backward: first window index of the class
};

struct TcParams {
  CUtensorMap q0[MAXCLS], q1[MAXCLS];      // query-tile boxes of qkv: head-dim chunk 0 (64 wide) and remainder
  CUtensorMap k0[MAXCLS], k1[MAXCLS];      // key / value boxes of qkv
  CUtensorMap o0[MAXCLS], o1[MAXCLS];      // query-tile boxes of dout (backward only)
  TcClass cls[MAXCLS];
  int n_cls;
  int B, H, W, nh, hd, wh, ww;
  float scale;
  const float* bias;       // [3*nh*hd] qkv bias (pad tokens)
  const bf16* qkv;
  bf16* out;               // forward: output; backward: unused
  float* lse;
  const bf16* o_in;        // backward: forward output
  const bf16* dout;
  bf16* dqkv;
  float* dws;              // backward: rowsum(dO o O) per (token, head)
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void sts16(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// byte address of 16-byte chunk c16 of row `row` in a K-major tile: chunk 0 = 64 columns, 128-byte swizzle
// (Swizzle<3,4,3>); remainder = C1 columns, 32-byte (Swizzle<1,4,3>) or 64-byte (Swizzle<2,4,3>) swizzle
__device__ __forceinline__ uint32_t tile0_addr(uint32_t base, int row, int c16) {
  return base + (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((c16 ^ (row & 7)) << 4));
}
template <int C1>
__device__ __forceinline__ uint32_t tile1_addr(uint32_t base, int row, int c16) {
  if (C1 == 16) return base + (uint32_t)(row * 32 + ((c16 ^ ((row >> 2) & 1)) << 4));
  return base + (uint32_t)(row * 64 + ((c16 ^ ((row >> 1) & 3)) << 4));
}

// shared-memory carve-up (bytes, every tile 1024-aligned).  Q / dO tiles: 128 rows; K / V: 256 rows (one window) or two
// 128-row blocks (streaming)
template <int C1>
struct Smem {
  static constexpr int C1B = C1 * 2;                       // bytes per row of the remainder chunk
  static constexpr int Q0 = 0, Q0_BYTES = 128 * 128;
  static constexpr int Q1 = Q0 + Q0_BYTES, Q1_BYTES = ((128 * C1B) + 1023) / 1024 * 1024;
  static constexpr int K0 = Q1 + Q1_BYTES, K0_BYTES = 256 * 128;
  static constexpr int K1 = K0 + K0_BYTES, K1_BYTES = 256 * C1B;
  static constexpr int V0 = K1 + K1_BYTES;
  static constexpr int V1 = V0 + K0_BYTES;
  static constexpr int FWD_BYTES = V1 + K1_BYTES + 1024;   // + alignment slack
  static constexpr uint32_t LAYOUT1 = C1 == 16 ? 6u : 4u;  // SWIZZLE_32B : SWIZZLE_64B
  static constexpr uint32_t SBO1 = 8 * C1B;                // 8 rows of the remainder chunk
};

constexpr uint32_t IDESC_BASE = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24);   // f32 acc, bf16 x bf16, M = 128
__device__ __forceinline__ uint32_t idesc_n(int n, bool b_mn_major) {
  return IDESC_BASE | ((uint32_t)(n >> 3) << 17) | (b_mn_major ? (1u << 16) : 0u);
}

struct Item {
  int c, b, wx, wy, t;
  int n_real, n_pad, n_keys, rw, rh;
  int q_rows;              // valid query rows of this tile
  int x0, yq0, yk0;        // box origins
};
template <bool STREAM>
__device__ __forceinline__ Item decode_item(const TcParams& p, int item) {
  Item it;
  int c = 0;
#pragma unroll
  for (int i = 1; i < MAXCLS; ++i)
    if (i < p.n_cls && item >= p.cls[i].item0) c = i;
  const TcClass& k = p.cls[c];
  const int local = item - k.item0;
  it.c = c;
  it.t = local % k.n_qt;
  const int widx = local / k.n_qt;
  it.wx = k.wx0 + widx % k.cnx;
  it.wy = k.wy0 + (widx / k.cnx) % k.cny;
  it.b = widx / (k.cnx * k.cny);
  it.n_real = k.n_real; it.n_pad = k.n_pad; it.n_keys = k.n_real + (k.n_pad > 0 ? 1 : 0);
  it.rw = k.rw; it.rh = k.rh;
  if (STREAM) {
    it.q_rows = min(128, k.rw - it.t * 128);
    it.x0 = it.t * 128; it.yq0 = 0; it.yk0 = 0;
  } else {
    it.q_rows = min(k.qbh, k.rh - it.t * k.qbh) * k.rw;
    it.x0 = it.wx * p.ww; it.yk0 = it.wy * p.wh; it.yq0 = it.yk0 + it.t * k.qbh;
  }
  return it;
}
// global token index of query row r of the item's tile (r < q_rows)
template <bool STREAM>
__device__ __forceinline__ long long query_token(const TcParams& p, const Item& it, int r) {
  if (STREAM) return (long long)it.b * p.H * p.W + it.x0 + r;
  const int ry = r / it.rw, rx = r - ry * it.rw;
  return ((long long)it.b * p.H + it.yq0 + ry) * p.W + it.x0 + rx;
}

// ------------------------------------------------------------------------------------------------ backward
// Shared by the two backward kernels: keys are visited in blocks of <= 128 (window mode: the same row boxes as the
// query tiles, i.e. qbh window rows; streaming: 128 tokens).
struct KeyBlock {
  int key0;                // first key index of the block
  int real_here;           // real keys in the block
  int keys_here;           // + the virtual pad key when it lives in this block
  int n_mma;               // rounded up to 16
  int x, y;                // box origin
};
template <bool STREAM>
__device__ __forceinline__ KeyBlock key_block(const TcParams& p, const TcClass& cl, const Item& it, int j) {
  KeyBlock kb;
  const int per = STREAM ? 128 : cl.qbh * cl.rw;
  kb.key0 = j * per;
  kb.real_here = min(per, it.n_real - kb.key0);
  kb.keys_here = kb.real_here + ((it.n_pad > 0 && j == cl.n_kb - 1) ? 1 : 0);
  kb.n_mma = max(16, (kb.keys_here + 15) & ~15);
  kb.x = STREAM ? j * 128 : it.x0;
  kb.y = STREAM ? 0 : it.yk0 + j * cl.qbh;
  return kb;
}

// rows [first, n_mma) of one K/V block tile: the virtual pad key (if it is row `pad_row`) or zeros
template <int C1>
__device__ __forceinline__ void write_tail_rows(uint32_t k0, uint32_t k1, uint32_t v0, uint32_t v1, const TcParams& p,
                                                int head, int first, int n_mma, int pad_row) {
  constexpr int NCH = 8 + C1 / 8;
  const int C = p.nh * p.hd;
  const int rows = n_mma - first;
  for (int e = threadIdx.x; e < rows * NCH * 2; e += blockDim.x) {
    const int which = e & 1;
    const int rc = e >> 1;
    const int row = first + rc / NCH, c16 = rc % NCH;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (row == pad_row) {
      const float* bsrc = p.bias + (which + 1) * C + head * p.hd;
      uint32_t w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int d = c16 * 8 + 2 * i;
        w[i] = pack2(d < p.hd ? bsrc[d] : 0.f, d + 1 < p.hd ? bsrc[d + 1] : 0.f);
      }
      val = make_uint4(w[0], w[1], w[2], w[3]);
    }
    const uint32_t b0 = which ? v0 : k0, b1 = which ? v1 : k1;
    sts16(c16 < 8 ? tile0_addr(b0, row, c16) : tile1_addr<C1>(b1, row, c16 - 8), val);
  }
}

// ------------------------------------------------------------------------------------------------- forward
// 256 threads: the two warps of a TMEM lane quarter (warp w and w + 4) share 32 query rows and split the key columns
// of each row in two contiguous ranges; row maximum and row sum are exchanged through shared memory.  Each thread
// writes its bf16 P INSIDE the column range it has already read (the partner may still be reading its own), so the
// packed A operand sits at columns [0, 16*h0) and [32*h0, ...): the MMA issuer steps its tensor-memory address
// accordingly.  O: columns [64,128) + [192,208) (windows: S is dead by then), [128,208) when streaming.
constexpr int NTHR_F = 256;
template <int C1, bool STREAM>
__global__ void __launch_bounds__(NTHR_F, C1 == 16 ? 2 : 1) fwd_kernel(const __grid_constant__ TcParams p) {
  using sm = Smem<C1>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[6];
  __shared__ uint32_t tmem_holder;
  __shared__ float red[2][2][128];                          // [max | sum][column half][row]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3, half = warp >> 2;
  const int row = quarter * 32 + lane;
  const int head = blockIdx.x;                               // fast dimension: the heavy (full-window) items of ALL heads first
  const Item it = decode_item<STREAM>(p, blockIdx.y);
  const TcClass& cl = p.cls[it.c];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t bar_k[2] = {bar0, bar0 + 8}, bar_v[2] = {bar0 + 16, bar0 + 24}, bar_s = bar0 + 32, bar_o = bar0 + 40;
  constexpr int KBCAP = STREAM ? 128 : 256;                  // keys per block
  const int nblk = STREAM ? (it.n_keys + 127) / 128 : 1;
  const int C = p.nh * p.hd;
  const int q_box_rows = STREAM ? 128 : cl.qbh * cl.rw;
  const int k_box_rows = STREAM ? 128 : cl.rh * cl.rw;
  ATC_STAMP(0);

  auto load_kv = [&](int j) {                               // key / value block j -> buffer j & 1
    const int buf = j & 1;
    const uint32_t kb0 = base + sm::K0 + buf * (128 * 128), kb1 = base + sm::K1 + buf * (128 * sm::C1B);
    const uint32_t vb0 = base + sm::V0 + buf * (128 * 128), vb1 = base + sm::V1 + buf * (128 * sm::C1B);
    const int x = STREAM ? j * 128 : it.x0;
    const uint32_t bytes = (uint32_t)k_box_rows * (128 + sm::C1B);
    mbar_expect_tx(bar_k[buf], bytes + (j == 0 ? (uint32_t)q_box_rows * (128 + sm::C1B) : 0u));
    tma_load_5d(kb0, &p.k0[it.c], bar_k[buf], 0, p.nh + head, x, it.yk0, it.b);
    if (C1) tma_load_5d(kb1, &p.k1[it.c], bar_k[buf], 64, p.nh + head, x, it.yk0, it.b);
    mbar_expect_tx(bar_v[buf], bytes);
    tma_load_5d(vb0, &p.k0[it.c], bar_v[buf], 0, 2 * p.nh + head, x, it.yk0, it.b);
    if (C1) tma_load_5d(vb1, &p.k1[it.c], bar_v[buf], 64, 2 * p.nh + head, x, it.yk0, it.b);
  };

  pdl_launch_dependents();
  if (warp == 0) {
    if (lane == 0) {
      // barriers, then straight to the TMA loads: their latency overlaps the TMEM allocation and the pad rows
      for (int i = 0; i < 6; ++i) mbar_init(bar0 + 8u * i, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      pdl_wait();                                           // qkv of the producer GEMM is complete
      load_kv(0);
      tma_load_5d(base + sm::Q0, &p.q0[it.c], bar_k[0], 0, head, it.x0, it.yq0, it.b);
      if (C1) tma_load_5d(base + sm::Q1, &p.q1[it.c], bar_k[0], 64, head, it.x0, it.yq0, it.b);
      if (nblk > 1) load_kv(1);
    }
  } else if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (!STREAM) {                                            // pad key / zero tail rows (rows the TMA boxes never touch)
    const int n_mma = max(16, (it.n_keys + 15) & ~15);
    write_tail_rows<C1>(base + sm::K0, base + sm::K1, base + sm::V0, base + sm::V1, p, head, it.n_real, n_mma,
                        it.n_pad > 0 ? it.n_real : -1);
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_holder;
  const uint32_t t_row = tmem + ((uint32_t)(quarter * 32) << 16);   // this warp's lane quarter
  ATC_STAMP(1);
  pdl_wait();
  ATC_STAMP(2);

  const float sl2 = p.scale * 1.4426950408889634f;
  const float bonus = it.n_pad > 0 ? __logf((float)it.n_pad) / p.scale : 0.f;
  float m_run = -INFINITY, l_run = 0.f;                     // l_run: this thread's column half only
  const uint32_t tS = tmem;
  // O: 64 + C1 columns that neither the packed P ranges nor (when streaming) the next S touch
  const bool o_high = STREAM || it.n_keys <= 128;
  const uint32_t o0col = o_high ? 128u : 64u;
  const uint32_t tO0 = tmem + o0col, tO1 = tmem + 192u;

  for (int j = 0; j < nblk; ++j) {
    const int buf = j & 1;
    const int keys_here = min(KBCAP, it.n_keys - j * KBCAP);
    const int n_mma = max(16, (keys_here + 15) & ~15);
    if (tid == 0) {
      if (j > 0) mbar_wait(bar_o, (uint32_t)(j - 1) & 1u);  // P of block j-1 (aliases S) has been consumed
      mbar_wait(bar_k[buf], (uint32_t)(j >> 1) & 1u);
      tc_fence_after();
      const uint32_t kb0 = base + sm::K0 + buf * (128 * 128), kb1 = base + sm::K1 + buf * (128 * sm::C1B);
      const uint64_t dq0 = smem_desc_sw128(base + sm::Q0), dk0 = smem_desc_sw128(kb0);
      const uint32_t id = idesc_n(n_mma, false);
#pragma unroll
      for (int k = 0; k < 4; ++k) tc_mma_bf16(tS, dq0 + 2 * k, dk0 + 2 * k, id, k > 0 ? 1u : 0u);
      if (C1) {
        const uint64_t dq1 = smem_desc(base + sm::Q1, 16, sm::SBO1, sm::LAYOUT1);
        const uint64_t dk1 = smem_desc(kb1, 16, sm::SBO1, sm::LAYOUT1);
#pragma unroll
        for (int k = 0; k < C1 / 16; ++k) tc_mma_bf16(tS, dq1 + 2 * k, dk1 + 2 * k, id, 1u);
      }
      tc_commit(bar_s);
      ATC_STAMP(3);
    }
    __syncwarp();
    mbar_wait(bar_s, (uint32_t)j & 1u);
    tc_fence_after();
    ATC_STAMP(4);

    // this thread's 32-column chunks of the row: [c_beg, c_end)
    const int nch = (n_mma + 31) >> 5;
    const int h0 = nch > 4 ? 4 : (nch + 1) >> 1;            // > 128 keys: P ranges [0,64) and [128,..), O in between
    const int c_beg = half ? h0 : 0, c_end = half ? nch : h0;
    const int key0 = j * KBCAP;
    // ---- pass 1: row maximum
    float mx = -INFINITY;
    for (int c = c_beg; c < c_end; ++c) {
      uint32_t r[32];
      tc_ld32(t_row + (uint32_t)(c * 32), r);
      tc_wait_ld();
      const int col0 = key0 + c * 32;
      if (col0 + 32 <= it.n_real) {
#pragma unroll
        for (int e = 0; e < 32; ++e) mx = fmaxf(mx, __uint_as_float(r[e]));
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          float v = __uint_as_float(r[e]);
          if (col0 + e == it.n_real) v += bonus;
          if (col0 + e < it.n_keys) mx = fmaxf(mx, v);
        }
      }
    }
    red[0][half][row] = mx;
    if (STREAM && j > 0) {
      // O holds blocks 0..j-1 once their PV product is complete
      mbar_wait(bar_o, (uint32_t)(j - 1) & 1u);
      tc_fence_after();
      if (tid == 0 && j + 1 < nblk) load_kv(j + 1);         // both buffers (j+1)&1 are free now
    }
    __syncthreads();
    const float m_new = fmaxf(m_run, fmaxf(red[0][0][row], red[0][1][row]));
    const float alpha = ex2((m_run - m_new) * sl2);         // 0 for the first block (m_run = -inf)
    const float mb = -m_new * sl2;
    if (STREAM && j > 0) {
      // rescale O to the new maximum; the two threads of a row take 48 and 16 + C1 of its columns
      for (int c = half ? 3 : 0; c < (half ? (64 + C1) / 16 : 3); ++c) {
        uint32_t r[16];
        const uint32_t a = t_row + (c < 4 ? 128u + (uint32_t)(c * 16) : 192u + (uint32_t)((c - 4) * 16));
        tc_ld16(a, r);
        tc_wait_ld();
#pragma unroll
        for (int e = 0; e < 16; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * alpha);
        tc_st16(a, r);
      }
    }
    // ---- pass 2: P = exp(S * scale - max) as bf16 inside this thread's column range, partial row sum
    float sum = 0.f;
    for (int c = c_beg; c < c_end; ++c) {
      uint32_t r[32], pk[16];
      tc_ld32(t_row + (uint32_t)(c * 32), r);
      tc_wait_ld();
      const int col0 = key0 + c * 32;
      if (col0 + 32 <= it.n_real) {
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const float a = ex2(fmaf(__uint_as_float(r[e]), sl2, mb)), b = ex2(fmaf(__uint_as_float(r[e + 1]), sl2, mb));
          const uint32_t w = pack2(a, b);
          pk[e >> 1] = w;
          sum += __uint_as_float(w << 16) + __uint_as_float(w & 0xffff0000u);   // the ROUNDED P: O is a convex combination
        }
      } else {
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          float v0 = __uint_as_float(r[e]), v1 = __uint_as_float(r[e + 1]);
          if (col0 + e == it.n_real) v0 += bonus;
          if (col0 + e + 1 == it.n_real) v1 += bonus;
          const float a = col0 + e < it.n_keys ? ex2(fmaf(v0, sl2, mb)) : 0.f;
          const float b = col0 + e + 1 < it.n_keys ? ex2(fmaf(v1, sl2, mb)) : 0.f;
          const uint32_t w = pack2(a, b);
          pk[e >> 1] = w;
          sum += __uint_as_float(w << 16) + __uint_as_float(w & 0xffff0000u);
        }
      }
      tc_st16(t_row + (uint32_t)(half ? 32 * h0 + 16 * (c - h0) : 16 * c), pk);
    }
    l_run = l_run * alpha + sum;
    m_run = m_new;
    tc_wait_st();
    tc_fence_before();
    __syncthreads();
    ATC_STAMP(5);
    if (tid == 0) {
      tc_fence_after();
      mbar_wait(bar_v[buf], (uint32_t)(j >> 1) & 1u);
      tc_fence_after();
      const uint32_t vb0 = base + sm::V0 + buf * (128 * 128), vb1 = base + sm::V1 + buf * (128 * sm::C1B);
      const uint64_t dv0 = smem_desc(vb0, 1024, 1024, 2), dv1 = smem_desc(vb1, sm::SBO1, sm::SBO1, sm::LAYOUT1);
      const uint32_t id0 = idesc_n(64, true), id1 = idesc_n(C1 ? C1 : 16, true);
      for (int ks = 0; ks < n_mma / 16; ++ks) {
        const uint32_t acc = (j > 0 || ks > 0) ? 1u : 0u;
        const int c = ks >> 1;                              // 16 keys = 8 packed columns of chunk c
        const uint32_t acol = (uint32_t)((c < h0 ? 16 * c : 32 * h0 + 16 * (c - h0)) + 8 * (ks & 1));
        tc_mma_bf16_ts(tO0, tS + acol, dv0 + (uint64_t)(ks * (2048 >> 4)), id0, acc);
        if (C1) tc_mma_bf16_ts(tO1, tS + acol, dv1 + (uint64_t)(ks * ((16 * sm::C1B) >> 4)), id1, acc);
      }
      tc_commit(bar_o);
    }
    __syncwarp();
  }
  red[1][half][row] = l_run;
  mbar_wait(bar_o, (uint32_t)(nblk - 1) & 1u);
  tc_fence_after();
  __syncthreads();
  ATC_STAMP(6);

  // ---- epilogue: O / l -> bf16 rows of `out` (16-byte stores), lse.  Column half 0: d 0..47, half 1: d 48..hd-1
  const bool live = row < it.q_rows;
  const long long tok = live ? query_token<STREAM>(p, it, row) : 0;
  const float l_tot = red[1][0][row] + red[1][1][row];
  const float inv = 1.f / l_tot;
  bf16* orow = p.out + tok * C + head * p.hd;
  for (int c = half ? 3 : 0; c < (half ? (64 + C1) / 16 : 3); ++c) {
    uint32_t r[16];
    tc_ld16(t_row + (c < 4 ? o0col + (uint32_t)(c * 16) : 192u + (uint32_t)((c - 4) * 16)), r);
    tc_wait_ld();
    if (live) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int d = c * 16 + h * 8;
        if (d < p.hd) {
          uint4 u;
          u.x = pack2(__uint_as_float(r[h * 8 + 0]) * inv, __uint_as_float(r[h * 8 + 1]) * inv);
          u.y = pack2(__uint_as_float(r[h * 8 + 2]) * inv, __uint_as_float(r[h * 8 + 3]) * inv);
          u.z = pack2(__uint_as_float(r[h * 8 + 4]) * inv, __uint_as_float(r[h * 8 + 5]) * inv);
          u.w = pack2(__uint_as_float(r[h * 8 + 6]) * inv, __uint_as_float(r[h * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + d) = u;
        }
      }
    }
  }
  if (live && half == 0) p.lse[tok * p.nh + head] = m_run * p.scale + __logf(l_tot);
  tc_fence_before();
  __syncthreads();
  ATC_STAMP(7);
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}


// D = sum_d dO[d] * O[d] of one query row (hd <= 64 + C1 bf16 values each): every 16-byte load is issued before the
// first use - with a runtime trip count the loads serialise behind the FMA chain (~0.5 us of L2 latency each)
template <int C1>
__device__ __forceinline__ float row_dot(const bf16* __restrict__ a, const bf16* __restrict__ b, int hd) {
  constexpr int NCH = (64 + C1) / 8;
  uint4 ua[NCH], ub[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    ua[c] = make_uint4(0, 0, 0, 0);
    ub[c] = make_uint4(0, 0, 0, 0);
    if (c * 8 < hd) {
      ua[c] = __ldg(reinterpret_cast<const uint4*>(a) + c);
      ub[c] = __ldg(reinterpret_cast<const uint4*>(b) + c);
    }
  }
  float d0 = 0.f, d1 = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&ua[c]);
    const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&ub[c]);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 fa = __bfloat1622float2(ha[e]), fb = __bfloat1622float2(hb[e]);
      d0 = fmaf(fa.x, fb.x, d0);
      d1 = fmaf(fa.y, fb.y, d1);
    }
  }
  return d0 + d1;
}

// rows of <= 64 + C1 head-dim values, one thread per row in shared memory (pitch 16 * NCH bytes), written out with
// CONSECUTIVE lanes on consecutive 16-byte chunks: a thread-per-row global store touches 32 cache lines per
// instruction (~1.2 us per [128 x 72] tile), this one ~4.  off[r] = element offset of row r's destination, < 0: skip
template <int C1>
__device__ __forceinline__ void copy_rows_out(const uint8_t* stage, const long long* off, bf16* __restrict__ dst,
                                              int hd, int t, int nthreads) {
  constexpr int NCH = (64 + C1) / 8;
  const int live = hd / 8;
  for (int f = t; f < 128 * live; f += nthreads) {
    const int r = f / live, c = f - r * live;
    const long long o = off[r];
    if (o >= 0)
      *reinterpret_cast<uint4*>(dst + o + c * 8) = *reinterpret_cast<const uint4*>(stage + r * (NCH * 16) + c * 16);
  }
}

template <int C1>
struct SmemDq {
  static constexpr int C1B = C1 * 2;
  static constexpr int T0 = 128 * 128, T1 = ((128 * C1B) + 1023) / 1024 * 1024;   // one 128-row tile: chunk 0, remainder
  static constexpr int Q0 = 0, Q1 = Q0 + T0;
  static constexpr int D0 = Q1 + T1, D1 = D0 + T0;           // dO
  static constexpr int K0 = D1 + T1, K1 = K0 + 2 * T0;       // two K buffers
  static constexpr int V0 = K1 + 2 * T1, V1 = V0 + T0;
  static constexpr int BYTES = V1 + T1 + 1024;
};

// dQ: one CTA = (window, head, tile of <= 128 queries), thread = query row = TMEM lane.  Per 128-key block:
//   S = Q K^T and dP = dO V^T (two accumulators, 256 TMEM columns) -> P = exp2(S * scale * log2e - lse),
//   dS = P o (dP - D) * scale written as bf16 over S -> dQ_block = dS K (A from tensor memory, K tile read MN-major)
//   into the columns dP occupied, added to the thread's fp32 row in registers.
// D = rowsum(dO o O) is computed here and kept for the dK/dV kernel (dws).
template <int C1, bool STREAM>
__global__ void __launch_bounds__(NTHR, C1 == 16 ? 2 : 1) bwd_dq_kernel(const __grid_constant__ TcParams p) {
  using sm = SmemDq<C1>;
  using s1 = Smem<C1>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[6];
  __shared__ uint32_t tmem_holder;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int head = blockIdx.x;                               // fast dimension: the heavy (full-window) items of ALL heads first
  const Item it = decode_item<STREAM>(p, blockIdx.y);
  const TcClass& cl = p.cls[it.c];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t bar_k[2] = {bar0, bar0 + 8}, bar_v = bar0 + 16, bar_s = bar0 + 24, bar_q = bar0 + 32;
  const int nblk = cl.n_kb;
  const int C = p.nh * p.hd;

  const uint32_t tile_bytes = (uint32_t)(STREAM ? 128 : cl.qbh * cl.rw) * (128 + s1::C1B);
  auto load_k = [&](int j) {
    const KeyBlock kb = key_block<STREAM>(p, cl, it, j);
    const int buf = j & 1;
    mbar_expect_tx(bar_k[buf], tile_bytes * (j == 0 ? 3u : 1u));       // block 0 also carries Q and dO
    tma_load_5d(base + sm::K0 + buf * sm::T0, &p.q0[it.c], bar_k[buf], 0, p.nh + head, kb.x, kb.y, it.b);
    if (C1) tma_load_5d(base + sm::K1 + buf * sm::T1, &p.q1[it.c], bar_k[buf], 64, p.nh + head, kb.x, kb.y, it.b);
  };
  auto load_v = [&](int j) {
    const KeyBlock kb = key_block<STREAM>(p, cl, it, j);
    mbar_expect_tx(bar_v, tile_bytes);
    tma_load_5d(base + sm::V0, &p.q0[it.c], bar_v, 0, 2 * p.nh + head, kb.x, kb.y, it.b);
    if (C1) tma_load_5d(base + sm::V1, &p.q1[it.c], bar_v, 64, 2 * p.nh + head, kb.x, kb.y, it.b);
  };
  pdl_launch_dependents();
  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < 5; ++i) mbar_init(bar0 + 8u * i, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      pdl_wait();
      load_k(0);
      tma_load_5d(base + sm::Q0, &p.q0[it.c], bar_k[0], 0, head, it.x0, it.yq0, it.b);
      if (C1) tma_load_5d(base + sm::Q1, &p.q1[it.c], bar_k[0], 64, head, it.x0, it.yq0, it.b);
      tma_load_5d(base + sm::D0, &p.o0[it.c], bar_k[0], 0, head, it.x0, it.yq0, it.b);
      if (C1) tma_load_5d(base + sm::D1, &p.o1[it.c], bar_k[0], 64, head, it.x0, it.yq0, it.b);
      load_v(0);
      if (nblk > 1) load_k(1);
    }
  } else if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_holder;
  const uint32_t t_row = tmem + ((uint32_t)(warp * 32) << 16);
  pdl_wait();

  // this thread's query row: lse, D = sum_d dO * O (also written out for the dK/dV kernel)
  const bool live = tid < it.q_rows;
  const long long tok = live ? query_token<STREAM>(p, it, tid) : 0;
  float Lrow = INFINITY, Drow = 0.f;
  if (live) {
    Lrow = p.lse[tok * p.nh + head] * 1.4426950408889634f;
    Drow = row_dot<C1>(p.dout + tok * C + head * p.hd, p.o_in + tok * C + head * p.hd, p.hd);
    p.dws[tok * p.nh + head] = Drow;
  }
  const float sl2 = p.scale * 1.4426950408889634f;
  const float bonus = it.n_pad > 0 ? __logf((float)it.n_pad) / p.scale : 0.f;
  float acc[64 + C1];
#pragma unroll
  for (int e = 0; e < 64 + C1; ++e) acc[e] = 0.f;

  for (int j = 0; j < nblk; ++j) {
    const int buf = j & 1;
    const KeyBlock kb = key_block<STREAM>(p, cl, it, j);
    const uint32_t k0 = base + sm::K0 + buf * sm::T0, k1 = base + sm::K1 + buf * sm::T1;
    if (!STREAM && kb.real_here < kb.n_mma) {
      // rows beyond the real keys of this block: pad key / zeros, after the TMA boxes (which may cover them) landed
      mbar_wait(bar_k[buf], (uint32_t)(j >> 1) & 1u);
      mbar_wait(bar_v, (uint32_t)j & 1u);
      write_tail_rows<C1>(k0, k1, base + sm::V0, base + sm::V1, p, head, kb.real_here, kb.n_mma,
                          kb.keys_here > kb.real_here ? kb.real_here : -1);
      fence_proxy_async();
      __syncthreads();
    }
    if (tid == 0) {
      mbar_wait(bar_k[buf], (uint32_t)(j >> 1) & 1u);
      mbar_wait(bar_v, (uint32_t)j & 1u);
      tc_fence_after();
      const uint32_t id = idesc_n(kb.n_mma, false);
      const uint64_t dq0 = smem_desc_sw128(base + sm::Q0), dk0 = smem_desc_sw128(k0);
      const uint64_t dd0 = smem_desc_sw128(base + sm::D0), dv0 = smem_desc_sw128(base + sm::V0);
#pragma unroll
      for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem, dq0 + 2 * k, dk0 + 2 * k, id, k > 0 ? 1u : 0u);
      if (C1) {
        const uint64_t dq1 = smem_desc(base + sm::Q1, 16, s1::SBO1, s1::LAYOUT1), dk1 = smem_desc(k1, 16, s1::SBO1, s1::LAYOUT1);
#pragma unroll
        for (int k = 0; k < C1 / 16; ++k) tc_mma_bf16(tmem, dq1 + 2 * k, dk1 + 2 * k, id, 1u);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem + 128, dd0 + 2 * k, dv0 + 2 * k, id, k > 0 ? 1u : 0u);
      if (C1) {
        const uint64_t dd1 = smem_desc(base + sm::D1, 16, s1::SBO1, s1::LAYOUT1);
        const uint64_t dv1 = smem_desc(base + sm::V1, 16, s1::SBO1, s1::LAYOUT1);
#pragma unroll
        for (int k = 0; k < C1 / 16; ++k) tc_mma_bf16(tmem + 128, dd1 + 2 * k, dv1 + 2 * k, id, 1u);
      }
      tc_commit(bar_s);
    }
    __syncwarp();
    mbar_wait(bar_s, (uint32_t)j & 1u);
    tc_fence_after();
    if (tid == 0 && j + 1 < nblk) {
      load_v(j + 1);                                         // V_j has been consumed by dP
      if (j >= 1) load_k(j + 1);                             // its buffer held K_{j-1}: dQ_{j-1} is complete
    }
    __syncwarp();
    const int nch = (kb.n_mma + 31) >> 5;
    for (int c = 0; c < nch; ++c) {
      uint32_t s[32], d[32], pk[16];
      tc_ld32(t_row + (uint32_t)(c * 32), s);
      tc_ld32(t_row + 128u + (uint32_t)(c * 32), d);
      tc_wait_ld();
      const int col0 = c * 32;
      if (col0 + 32 <= kb.real_here) {
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const float p0 = ex2(fmaf(__uint_as_float(s[e]), sl2, -Lrow)), p1 = ex2(fmaf(__uint_as_float(s[e + 1]), sl2, -Lrow));
          pk[e >> 1] = pack2(p0 * (__uint_as_float(d[e]) - Drow) * p.scale, p1 * (__uint_as_float(d[e + 1]) - Drow) * p.scale);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          float v0 = __uint_as_float(s[e]), v1 = __uint_as_float(s[e + 1]);
          if (col0 + e == kb.real_here) v0 += bonus;
          if (col0 + e + 1 == kb.real_here) v1 += bonus;
          const float p0 = col0 + e < kb.keys_here ? ex2(fmaf(v0, sl2, -Lrow)) : 0.f;
          const float p1 = col0 + e + 1 < kb.keys_here ? ex2(fmaf(v1, sl2, -Lrow)) : 0.f;
          const float g0 = col0 + e < kb.keys_here ? p0 * (__uint_as_float(d[e]) - Drow) * p.scale : 0.f;
          const float g1 = col0 + e + 1 < kb.keys_here ? p1 * (__uint_as_float(d[e + 1]) - Drow) * p.scale : 0.f;
          pk[e >> 1] = pack2(g0, g1);
        }
      }
      tc_st16(t_row + (uint32_t)(c * 16), pk);
    }
    tc_wait_st();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint64_t dk0 = smem_desc(k0, 1024, 1024, 2), dk1 = smem_desc(k1, s1::SBO1, s1::SBO1, s1::LAYOUT1);
      const uint32_t id0 = idesc_n(64, true), id1 = idesc_n(C1 ? C1 : 16, true);
      for (int ks = 0; ks < kb.n_mma / 16; ++ks) {
        tc_mma_bf16_ts(tmem + 128, tmem + (uint32_t)(8 * ks), dk0 + (uint64_t)(ks * (2048 >> 4)), id0, ks > 0 ? 1u : 0u);
        if (C1) tc_mma_bf16_ts(tmem + 192, tmem + (uint32_t)(8 * ks), dk1 + (uint64_t)(ks * ((16 * s1::C1B) >> 4)), id1, ks > 0 ? 1u : 0u);
      }
      tc_commit(bar_q);
    }
    __syncwarp();
    mbar_wait(bar_q, (uint32_t)j & 1u);
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < (64 + C1) / 16; ++c) {
      uint32_t r[16];
      tc_ld16(t_row + 128u + (uint32_t)(c * 16), r);
      tc_wait_ld();
#pragma unroll
      for (int e = 0; e < 16; ++e) acc[c * 16 + e] += __uint_as_float(r[e]);
    }
    tc_fence_before();
    __syncthreads();                                         // the next block's S / dP overwrite these columns
    tc_fence_after();
  }
  if (live) {
    bf16* dst = p.dqkv + tok * (3LL * C) + head * p.hd;
#pragma unroll
    for (int c = 0; c < (64 + C1) / 8; ++c) {
      if (c * 8 < p.hd) {
        uint4 u;
        u.x = pack2(acc[c * 8 + 0], acc[c * 8 + 1]); u.y = pack2(acc[c * 8 + 2], acc[c * 8 + 3]);
        u.z = pack2(acc[c * 8 + 4], acc[c * 8 + 5]); u.w = pack2(acc[c * 8 + 6], acc[c * 8 + 7]);
        *reinterpret_cast<uint4*>(dst + c * 8) = u;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

// dK / dV: one CTA = (window, head, block of <= 128 keys), 256 threads, key row = TMEM lane, the two warps of a lane
// quarter split the query columns.  Per tile of <= 128 queries (double-buffered TMA):
//   S^T = K Q^T and dP^T = V dO^T -> P^T, dS^T (bf16, written over S^T / dP^T) ->
//   dV += P^T dO, dK += dS^T Q   (A from tensor memory, Q / dO tiles read MN-major), accumulated in tensor memory
// over the query tiles: 128 + 128 + 2 x (64 + C1) = 416 TMEM columns.
constexpr int NTHR_KV = 256;
template <int C1>
struct SmemKv {
  static constexpr int C1B = C1 * 2;
  static constexpr int T0 = 128 * 128, T1 = ((128 * C1B) + 1023) / 1024 * 1024;
  static constexpr int K0 = 0, K1 = K0 + T0;
  static constexpr int V0 = K1 + T1, V1 = V0 + T0;
  static constexpr int Q0 = V1 + T1, Q1 = Q0 + 2 * T0;       // two query-tile buffers
  static constexpr int D0 = Q1 + 2 * T1, D1 = D0 + 2 * T0;
  static constexpr int LD = D1 + 2 * T1;                     // [2 buffers][lse * log2e (128) | D (128)] floats
  static constexpr int BYTES = LD + 2 * 2 * 128 * 4 + 1024;
};

template <int C1, bool STREAM>
__global__ void __launch_bounds__(NTHR_KV, 1) bwd_dkv_kernel(const __grid_constant__ TcParams p) {
  using sm = SmemKv<C1>;
  using s1 = Smem<C1>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[6];
  __shared__ uint32_t tmem_holder;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int head = blockIdx.x;
  // work item -> (class, window, key block)
  int c = 0;
#pragma unroll
  for (int i = 1; i < MAXCLS; ++i)
    if (i < p.n_cls && (int)blockIdx.y >= p.cls[i].kitem0) c = i;
  const TcClass& cl = p.cls[c];
  const int local = blockIdx.y - cl.kitem0;
  const int jb = local % cl.n_kb;
  // decode_item on the window's first query tile gives the window coordinates
  const Item it = decode_item<STREAM>(p, cl.item0 + (local / cl.n_kb) * cl.n_qt);
  const KeyBlock kb = key_block<STREAM>(p, cl, it, jb);
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t bar_in[2] = {bar0, bar0 + 8}, bar_s = bar0 + 16, bar_d = bar0 + 24;
  float* LDs = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)) + sm::LD);
  const int n_qt = cl.n_qt;
  const int C = p.nh * p.hd;
  const int q_box_rows = STREAM ? 128 : cl.qbh * cl.rw;

  const uint32_t tile_bytes = (uint32_t)q_box_rows * (128 + s1::C1B);
  auto q_origin = [&](int i, int& x, int& y) {
    x = STREAM ? i * 128 : it.x0;
    y = STREAM ? 0 : it.yk0 + i * cl.qbh;
  };
  auto load_q = [&](int i) {
    const int buf = i & 1;
    int x, y;
    q_origin(i, x, y);
    mbar_expect_tx(bar_in[buf], tile_bytes * (i == 0 ? 4u : 2u));      // tile 0 also carries K and V
    tma_load_5d(base + sm::Q0 + buf * sm::T0, &p.q0[c], bar_in[buf], 0, head, x, y, it.b);
    if (C1) tma_load_5d(base + sm::Q1 + buf * sm::T1, &p.q1[c], bar_in[buf], 64, head, x, y, it.b);
    tma_load_5d(base + sm::D0 + buf * sm::T0, &p.o0[c], bar_in[buf], 0, head, x, y, it.b);
    if (C1) tma_load_5d(base + sm::D1 + buf * sm::T1, &p.o1[c], bar_in[buf], 64, head, x, y, it.b);
  };
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(bar0 + 8u * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    pdl_wait();                                              // dws from the dQ kernel, lse, dout
    load_q(0);
    tma_load_5d(base + sm::K0, &p.q0[c], bar_in[0], 0, p.nh + head, kb.x, kb.y, it.b);
    if (C1) tma_load_5d(base + sm::K1, &p.q1[c], bar_in[0], 64, p.nh + head, kb.x, kb.y, it.b);
    tma_load_5d(base + sm::V0, &p.q0[c], bar_in[0], 0, 2 * p.nh + head, kb.x, kb.y, it.b);
    if (C1) tma_load_5d(base + sm::V1, &p.q1[c], bar_in[0], 64, 2 * p.nh + head, kb.x, kb.y, it.b);
    if (n_qt > 1) load_q(1);
  }
  pdl_launch_dependents();
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // query rows the TMA boxes never write but the (rounded-up) MMAs read as the K dimension of dV / dK: zero
  {
    constexpr int NCH = 8 + C1 / 8;
    const int first = q_box_rows, rows = min(128, (q_box_rows + 15) & ~15) - first;
    for (int e = tid; e < rows * NCH * 4; e += NTHR_KV) {
      const int which = e & 3;                               // Q buf 0/1, dO buf 0/1
      const int rc = e >> 2;
      const int row = first + rc / NCH, c16 = rc % NCH;
      const uint32_t b0 = base + ((which & 2) ? sm::D0 : sm::Q0) + (which & 1) * sm::T0;
      const uint32_t b1 = base + ((which & 2) ? sm::D1 : sm::Q1) + (which & 1) * sm::T1;
      sts16(c16 < 8 ? tile0_addr(b0, row, c16) : tile1_addr<C1>(b1, row, c16 - 8), make_uint4(0, 0, 0, 0));
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_holder;
  const int quarter = warp & 3, half = warp >> 2;
  const uint32_t t_row = tmem + ((uint32_t)(quarter * 32) << 16);
  pdl_wait();                                                // dws / dqkv's q third from the dQ kernel, lse, dout

  // valid queries of tile i and their tokens
  auto q_rows_of = [&](int i) { return STREAM ? min(128, it.rw - i * 128) : min(cl.qbh, cl.rh - i * cl.qbh) * cl.rw; };
  auto fill_ld = [&](int i) {                                 // lse * log2e and D of tile i's queries -> LDs[i & 1]
    if (tid < 128) {
      const int rows = q_rows_of(i);
      float L = INFINITY, D = 0.f;
      if (tid < rows) {
        long long tk;
        if (STREAM) tk = (long long)it.b * p.H * p.W + i * 128 + tid;
        else {
          const int ry = tid / cl.rw, rx = tid - ry * cl.rw;
          tk = ((long long)it.b * p.H + it.yk0 + i * cl.qbh + ry) * p.W + it.x0 + rx;
        }
        L = p.lse[tk * p.nh + head] * 1.4426950408889634f;
        D = p.dws[tk * p.nh + head];
      }
      LDs[(i & 1) * 256 + tid] = L;
      LDs[(i & 1) * 256 + 128 + tid] = D;
    }
  };
  fill_ld(0);
  __syncthreads();
  const float sl2 = p.scale * 1.4426950408889634f;
  const uint32_t tK = tmem + 256, tV = tmem + 256 + 64 + C1;

  for (int i = 0; i < n_qt; ++i) {
    const int buf = i & 1;
    const int q_rows = q_rows_of(i);
    const int nq_mma = max(16, (q_rows + 15) & ~15);
    const uint32_t q0 = base + sm::Q0 + buf * sm::T0, q1 = base + sm::Q1 + buf * sm::T1;
    const uint32_t d0 = base + sm::D0 + buf * sm::T0, d1 = base + sm::D1 + buf * sm::T1;
    if (tid == 0) {
      if (i > 0) {
        mbar_wait(bar_d, (uint32_t)(i - 1) & 1u);            // dV / dK products of tile i-1 have consumed P^T, dS^T
        if (i + 1 < n_qt) load_q(i + 1);                     // ... and its Q / dO buffer
      }
      mbar_wait(bar_in[buf], (uint32_t)(i >> 1) & 1u);
      tc_fence_after();
      const uint32_t id = idesc_n(nq_mma, false);
      const uint64_t dk0 = smem_desc_sw128(base + sm::K0), dq0 = smem_desc_sw128(q0);
      const uint64_t dv0 = smem_desc_sw128(base + sm::V0), dd0 = smem_desc_sw128(d0);
#pragma unroll
      for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem, dk0 + 2 * k, dq0 + 2 * k, id, k > 0 ? 1u : 0u);
      if (C1) {
        const uint64_t dk1 = smem_desc(base + sm::K1, 16, s1::SBO1, s1::LAYOUT1), dq1 = smem_desc(q1, 16, s1::SBO1, s1::LAYOUT1);
#pragma unroll
        for (int k = 0; k < C1 / 16; ++k) tc_mma_bf16(tmem, dk1 + 2 * k, dq1 + 2 * k, id, 1u);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem + 128, dv0 + 2 * k, dd0 + 2 * k, id, k > 0 ? 1u : 0u);
      if (C1) {
        const uint64_t dv1 = smem_desc(base + sm::V1, 16, s1::SBO1, s1::LAYOUT1), dd1 = smem_desc(d1, 16, s1::SBO1, s1::LAYOUT1);
#pragma unroll
        for (int k = 0; k < C1 / 16; ++k) tc_mma_bf16(tmem + 128, dv1 + 2 * k, dd1 + 2 * k, id, 1u);
      }
      tc_commit(bar_s);
    }
    __syncwarp();
    if (i + 1 < n_qt) fill_ld(i + 1);                        // published by the barrier below
    mbar_wait(bar_s, (uint32_t)i & 1u);
    tc_fence_after();
    const float* Ls = LDs + buf * 256;
    const float* Ds = Ls + 128;
#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {
      const int col0 = half * 64 + cc * 32;
      if (col0 >= nq_mma) break;
      uint32_t s[32], d[32], pp[16], pd[16];
      tc_ld32(t_row + (uint32_t)col0, s);
      tc_ld32(t_row + 128u + (uint32_t)col0, d);
      tc_wait_ld();
#pragma unroll
      for (int e = 0; e < 32; e += 2) {
        const float2 L2 = *reinterpret_cast<const float2*>(Ls + col0 + e);
        const float2 D2 = *reinterpret_cast<const float2*>(Ds + col0 + e);
        float p0 = ex2(fmaf(__uint_as_float(s[e]), sl2, -L2.x)), p1 = ex2(fmaf(__uint_as_float(s[e + 1]), sl2, -L2.y));
        float g0 = p0 * (__uint_as_float(d[e]) - D2.x) * p.scale, g1 = p1 * (__uint_as_float(d[e + 1]) - D2.y) * p.scale;
        if (col0 + e >= q_rows) { p0 = 0.f; g0 = 0.f; }
        if (col0 + e + 1 >= q_rows) { p1 = 0.f; g1 = 0.f; }
        pp[e >> 1] = pack2(p0, p1);
        pd[e >> 1] = pack2(g0, g1);
      }
      // packed bf16 pairs go INSIDE this warp's own (already read) column range - the other warp of the lane quarter
      // may still be reading its 64 columns: queries [64h + 32cc, +32) -> columns [64h + 16cc, +16)
      tc_st16(t_row + (uint32_t)(half * 64 + cc * 16), pp);
      tc_st16(t_row + 128u + (uint32_t)(half * 64 + cc * 16), pd);
    }
    tc_wait_st();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint64_t bq0 = smem_desc(q0, 1024, 1024, 2), bq1 = smem_desc(q1, s1::SBO1, s1::SBO1, s1::LAYOUT1);
      const uint64_t bd0 = smem_desc(d0, 1024, 1024, 2), bd1 = smem_desc(d1, s1::SBO1, s1::SBO1, s1::LAYOUT1);
      const uint32_t id0 = idesc_n(64, true), id1 = idesc_n(C1 ? C1 : 16, true);
      for (int ks = 0; ks < nq_mma / 16; ++ks) {
        const uint32_t accf = (i > 0 || ks > 0) ? 1u : 0u;
        const uint64_t o0 = (uint64_t)(ks * (2048 >> 4)), o1 = (uint64_t)(ks * ((16 * s1::C1B) >> 4));
        const uint32_t acol = (uint32_t)((ks >> 2) * 64 + (ks & 3) * 8);              // 16 queries = 8 packed columns
        tc_mma_bf16_ts(tV, tmem + acol, bd0 + o0, id0, accf);                           // dV += P^T dO
        if (C1) tc_mma_bf16_ts(tV + 64, tmem + acol, bd1 + o1, id1, accf);
        tc_mma_bf16_ts(tK, tmem + 128 + acol, bq0 + o0, id0, accf);                     // dK += dS^T Q
        if (C1) tc_mma_bf16_ts(tK + 64, tmem + 128 + acol, bq1 + o1, id1, accf);
      }
      tc_commit(bar_d);
    }
    __syncwarp();
  }
  mbar_wait(bar_d, (uint32_t)(n_qt - 1) & 1u);
  tc_fence_after();
  // warps 0-3 store dK, warps 4-7 store dV: row = key, 16-byte stores
  const int kidx = kb.key0 + quarter * 32 + lane;
  const bool live = quarter * 32 + lane < kb.real_here;
  long long tk = 0;
  if (live) {
    if (STREAM) tk = (long long)it.b * p.H * p.W + kidx;
    else {
      const int ky = kidx / cl.rw, kx = kidx - ky * cl.rw;
      tk = ((long long)it.b * p.H + it.yk0 + ky) * p.W + it.x0 + kx;
    }
  }
  bf16* dst = p.dqkv + tk * (3LL * C) + (half ? 2 : 1) * C + head * p.hd;
  const uint32_t t_src = t_row + (half ? (256u + 64u + C1) : 256u);
#pragma unroll
  for (int cc = 0; cc < (64 + C1) / 16; ++cc) {
    uint32_t r[16];
    tc_ld16(t_src + (uint32_t)(cc * 16), r);
    tc_wait_ld();
    if (live) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int dd = cc * 16 + h * 8;
        if (dd < p.hd) {
          uint4 u;
          u.x = pack2(__uint_as_float(r[h * 8 + 0]), __uint_as_float(r[h * 8 + 1]));
          u.y = pack2(__uint_as_float(r[h * 8 + 2]), __uint_as_float(r[h * 8 + 3]));
          u.z = pack2(__uint_as_float(r[h * 8 + 4]), __uint_as_float(r[h * 8 + 5]));
          u.w = pack2(__uint_as_float(r[h * 8 + 6]), __uint_as_float(r[h * 8 + 7]));
          *reinterpret_cast<uint4*>(dst + dd) = u;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// ------------------------------------------------------------------------------------------------------------
// Fused backward for windows (<= 256 tokens, head dim <= 80): ONE CTA = (window, head) computes dQ, dK and dV, so the
// scores are recomputed once (the two-kernel path above recomputes them in each kernel) and no D workspace travels
// through HBM.  All eight operand tiles of the window (Q, dO per query tile; K, V per key block) are TMA-loaded once.
// Keys sit on the TMEM lanes.  Per (key block j, query tile i):
//   S^T = K_j Q_i^T, dP^T = V_j dO_i^T                     two accumulators, 256 TMEM columns
//   P^T, dS^T (bf16) over them (tcgen05.st), and dS^T also into a 128-byte-swizzled shared-memory tile [key][query]
//   dV_j += P^T dO_i,  dK_j += dS^T Q_i                    A from tensor memory, accumulated in TMEM over i
//   dQ_i part = dS K_j                                     A = the shared dS^T tile read MN-major (M = queries), B = K_j
//                                                          read MN-major; the [128 x hd'] result is added to fp32
//                                                          registers (queries on lanes), accumulated over j
// TMEM: 128 + 128 + 3 x (64 + C1) = 496 columns; shared memory: 8 tiles + dS^T tile = 195 KB; 256 threads.
template <int C1>
struct SmemWin {
  static constexpr int C1B = C1 * 2;
  static constexpr int T0 = 128 * 128, T1 = ((128 * C1B) + 1023) / 1024 * 1024, TILE = T0 + T1;
  static constexpr int Q = 0, DO = 2 * TILE, K = 4 * TILE, V = 6 * TILE;     // [2 tiles or blocks][chunk 0 | remainder]
  static constexpr int DS = 8 * TILE;                                        // dS^T: two [128 x 64] chunks
  static constexpr int LD = DS + 2 * T0;                                     // [2 tiles][lse*log2e (128) | D (128)]
  static constexpr int BYTES = LD + 2 * 256 * 4 + 1024;
};

template <int C1>
__global__ void __launch_bounds__(NTHR_KV, 1) bwd_win_kernel(const __grid_constant__ TcParams p) {
  using sm = SmemWin<C1>;
  using s1 = Smem<C1>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[6];
  __shared__ uint32_t tmem_holder;
  __shared__ long long offs[2][128];                         // destination rows of the staged stores
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3, half = warp >> 2;
  const int row = quarter * 32 + lane;
  const int head = blockIdx.x;
  int c = 0;
#pragma unroll
  for (int i = 1; i < MAXCLS; ++i)
    if (i < p.n_cls && (int)blockIdx.y >= p.cls[i].witem0) c = i;
  const TcClass& cl = p.cls[c];
  const Item it = decode_item<false>(p, cl.item0 + ((int)blockIdx.y - cl.witem0) * cl.n_qt);
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t bar_q[2] = {bar0, bar0 + 8}, bar_k[2] = {bar0 + 16, bar0 + 24}, bar_s = bar0 + 32, bar_d = bar0 + 40;
  float* LDs = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)) + sm::LD);
  const int n_qt = cl.n_qt;                                  // <= 2 query tiles = key blocks
  const int C = p.nh * p.hd;
  const int box_rows = cl.qbh * cl.rw;
  const uint32_t tile_bytes = (uint32_t)box_rows * (128 + s1::C1B);
  auto tile0 = [&](int what, int i) { return base + (uint32_t)(what + i * sm::TILE); };
  auto tile1 = [&](int what, int i) { return base + (uint32_t)(what + i * sm::TILE + sm::T0); };
  auto q_rows_of = [&](int i) { return min(cl.qbh, cl.rh - i * cl.qbh) * cl.rw; };
  ATC_STAMP(0);

  pdl_launch_dependents();
  if (warp == 0) {
    // the four tensor maps miss the descriptor cache on a cold SM (~1 us each when fetched one after the other by
    // the issuing thread): four lanes prefetch them at once, and four lanes issue the loads (one barrier group each)
    if (lane == 0) tma_prefetch_desc(&p.q0[c]);
    if (lane == 1) tma_prefetch_desc(&p.q1[c]);
    if (lane == 2) tma_prefetch_desc(&p.o0[c]);
    if (lane == 3) tma_prefetch_desc(&p.o1[c]);
    if (lane == 0) {
      for (int i = 0; i < 4; ++i) mbar_init(bar0 + 8u * i, 1);
      mbar_init(bar_s, 2);                                   // one commit per MMA-issuing thread, see below
      mbar_init(bar_d, 3);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (lane < 4) {
      pdl_wait();
      const int i = lane >> 1;                               // tile i of the queries == block i of the keys (same boxes)
      if (i < n_qt) {
        const int y = it.yk0 + i * cl.qbh;
        if (lane & 1) {
          mbar_expect_tx(bar_q[i], 2 * tile_bytes);
          tma_load_5d(tile0(sm::Q, i), &p.q0[c], bar_q[i], 0, head, it.x0, y, it.b);
          if (C1) tma_load_5d(tile1(sm::Q, i), &p.q1[c], bar_q[i], 64, head, it.x0, y, it.b);
          tma_load_5d(tile0(sm::DO, i), &p.o0[c], bar_q[i], 0, head, it.x0, y, it.b);
          if (C1) tma_load_5d(tile1(sm::DO, i), &p.o1[c], bar_q[i], 64, head, it.x0, y, it.b);
        } else {
          mbar_expect_tx(bar_k[i], 2 * tile_bytes);
          tma_load_5d(tile0(sm::K, i), &p.q0[c], bar_k[i], 0, p.nh + head, it.x0, y, it.b);
          if (C1) tma_load_5d(tile1(sm::K, i), &p.q1[c], bar_k[i], 64, p.nh + head, it.x0, y, it.b);
          tma_load_5d(tile0(sm::V, i), &p.q0[c], bar_k[i], 0, 2 * p.nh + head, it.x0, y, it.b);
          if (C1) tma_load_5d(tile1(sm::V, i), &p.q1[c], bar_k[i], 64, 2 * p.nh + head, it.x0, y, it.b);
        }
      }
    }
  } else if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // query rows the TMA boxes never write but the (rounded-up) dV / dK products read as their K dimension: zero
  {
    constexpr int NCH = 8 + C1 / 8;
    const int rows = min(128, (box_rows + 15) & ~15) - box_rows;
    for (int e = tid; e < rows * NCH * 4; e += NTHR_KV) {
      const int which = e & 3;                               // Q tile 0/1, dO tile 0/1
      const int rc = e >> 2;
      const int r = box_rows + rc / NCH, c16 = rc % NCH;
      const int what = (which & 2) ? sm::DO : sm::Q;
      sts16(c16 < 8 ? tile0_addr(tile0(what, which & 1), r, c16) : tile1_addr<C1>(tile1(what, which & 1), r, c16 - 8),
            make_uint4(0, 0, 0, 0));
    }
  }
  tc_fence_before();
  __syncthreads();                                           // barriers initialised, TMEM allocated
  tc_fence_after();
  pdl_wait();
  // lse and D = sum_d dO * O of the window's queries: thread t owns query (t / 128, t % 128).  The O row comes from
  // global memory (loads issued here, consumed below), the dO row from its TMA-loaded tile
  {
    constexpr int NCH = (64 + C1) / 8;
    const int i = tid >> 7, r = tid & 127;
    const bool liveq = i < n_qt && r < q_rows_of(i);
    float L = INFINITY, D = 0.f;
    uint4 uo[NCH];
#pragma unroll
    for (int cc = 0; cc < NCH; ++cc) uo[cc] = make_uint4(0, 0, 0, 0);
    if (liveq) {
      const int ry = r / cl.rw, rx = r - ry * cl.rw;
      const long long tk = ((long long)it.b * p.H + it.yk0 + i * cl.qbh + ry) * p.W + it.x0 + rx;
      L = p.lse[tk * p.nh + head] * 1.4426950408889634f;
      const uint4* src = reinterpret_cast<const uint4*>(p.o_in + tk * C + head * p.hd);
#pragma unroll
      for (int cc = 0; cc < NCH; ++cc)
        if (cc * 8 < p.hd) uo[cc] = __ldg(src + cc);
    }
    if (i < n_qt) {
      mbar_wait(bar_q[i], 0);                                // dO_i has landed
      float d0 = 0.f, d1 = 0.f;
#pragma unroll
      for (int cc = 0; cc < NCH; ++cc) {
        uint4 ud;
        const uint32_t ad = cc < 8 ? tile0_addr(tile0(sm::DO, i), r, cc) : tile1_addr<C1>(tile1(sm::DO, i), r, cc - 8);
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(ud.x), "=r"(ud.y), "=r"(ud.z), "=r"(ud.w) : "r"(ad));
        const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&ud);
        const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&uo[cc]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 fa = __bfloat1622float2(ha[e]), fb = __bfloat1622float2(hb[e]);
          d0 = fmaf(fa.x, fb.x, d0);
          d1 = fmaf(fa.y, fb.y, d1);
        }
      }
      D = liveq ? d0 + d1 : 0.f;                             // (rows beyond the window: uo = 0)
    }
    LDs[i * 256 + r] = L;
    LDs[i * 256 + 128 + r] = D;
  }
  fence_proxy_async();
  __syncthreads();
  const uint32_t tmem = tmem_holder;
  const uint32_t t_row = tmem + ((uint32_t)(quarter * 32) << 16);
  const uint32_t tK = tmem + 256, tV = tmem + 256 + 64 + C1, tQ = tmem + 256 + 2 * (64 + C1);
  const float sl2 = p.scale * 1.4426950408889634f;
  const float bonus = it.n_pad > 0 ? __logf((float)it.n_pad) / p.scale : 0.f;
  constexpr int NACC = 48;                                   // dQ row: column half 0 keeps d [0,48), half 1 [48, 64 + C1)
  const int a_beg = half ? 3 : 0, a_end = half ? (64 + C1) / 16 : 3;
  float acc[2][NACC];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int e = 0; e < NACC; ++e) acc[i][e] = 0.f;
  const uint32_t ds_base = base + sm::DS;
  uint32_t step = 0;

  for (int j = 0; j < n_qt; ++j) {
    const KeyBlock kb = key_block<false>(p, cl, it, j);
    if (kb.real_here < kb.n_mma) {
      // rows beyond the real keys of this block: pad key (k = v = bias) / zeros, after the TMA box (which may cover them)
      mbar_wait(bar_k[j], 0);
      write_tail_rows<C1>(tile0(sm::K, j), tile1(sm::K, j), tile0(sm::V, j), tile1(sm::V, j), p, head, kb.real_here, kb.n_mma,
                          kb.keys_here > kb.real_here ? kb.real_here : -1);
      fence_proxy_async();
      __syncthreads();
    }
    const bool key_live = row < kb.keys_here;                // this thread's key takes part in the softmax
    const float row_bonus = (kb.keys_here > kb.real_here && row == kb.real_here) ? bonus : 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (i >= n_qt) break;
      const int q_rows = q_rows_of(i);
      const int nq_mma = max(16, (q_rows + 15) & ~15);
      if (tid == 0 || tid == 32) {
        // the tensor core is driven by several threads (one per product): a single thread issues a tcgen05.mma
        // every ~35 ns, which made the 48 instructions of a step the longest part of it
        mbar_wait(bar_k[j], 0);
        mbar_wait(bar_q[i], 0);
        tc_fence_after();
        const uint32_t id = idesc_n(nq_mma, false);
        const int a_what = tid == 0 ? sm::K : sm::V, b_what = tid == 0 ? sm::Q : sm::DO;
        const uint32_t td = tmem + (tid == 0 ? 0u : 128u);
        const uint64_t da0 = smem_desc_sw128(tile0(a_what, j)), db0 = smem_desc_sw128(tile0(b_what, i));
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_bf16(td, da0 + 2 * k, db0 + 2 * k, id, k > 0 ? 1u : 0u);
        if (C1) {
          const uint64_t da1 = smem_desc(tile1(a_what, j), 16, s1::SBO1, s1::LAYOUT1), db1 = smem_desc(tile1(b_what, i), 16, s1::SBO1, s1::LAYOUT1);
#pragma unroll
          for (int k = 0; k < C1 / 16; ++k) tc_mma_bf16(td, da1 + 2 * k, db1 + 2 * k, id, 1u);
        }
        tc_commit(bar_s);
      }
      __syncwarp();
      mbar_wait(bar_s, step & 1u);
      tc_fence_after();
      const float* Ls = LDs + i * 256;
      const float* Ds = Ls + 128;
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int col0 = half * 64 + cc * 32;
        if (col0 >= nq_mma) break;
        uint32_t s[32], d[32], pp[16], pd[16];
        tc_ld32(t_row + (uint32_t)col0, s);
        tc_ld32(t_row + 128u + (uint32_t)col0, d);
        tc_wait_ld();
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const float2 L2 = *reinterpret_cast<const float2*>(Ls + col0 + e);
          const float2 D2 = *reinterpret_cast<const float2*>(Ds + col0 + e);
          float p0 = ex2(fmaf(__uint_as_float(s[e]) + row_bonus, sl2, -L2.x));
          float p1 = ex2(fmaf(__uint_as_float(s[e + 1]) + row_bonus, sl2, -L2.y));
          float g0 = p0 * (__uint_as_float(d[e]) - D2.x) * p.scale, g1 = p1 * (__uint_as_float(d[e + 1]) - D2.y) * p.scale;
          if (!key_live || col0 + e >= q_rows) { p0 = 0.f; g0 = 0.f; }
          if (!key_live || col0 + e + 1 >= q_rows) { p1 = 0.f; g1 = 0.f; }
          pp[e >> 1] = pack2(p0, p1);
          pd[e >> 1] = pack2(g0, g1);
        }
        // packed pairs inside this warp's own (already read) column range, see bwd_dkv_kernel
        tc_st16(t_row + (uint32_t)(half * 64 + cc * 16), pp);
        tc_st16(t_row + 128u + (uint32_t)(half * 64 + cc * 16), pd);
        // dS^T row of this key, queries col0 .. col0 + 31, into the [key][query] tile (A operand of the dQ product)
        const uint32_t chunk = ds_base + (uint32_t)((col0 >> 6) * sm::T0);
#pragma unroll
        for (int m = 0; m < 4; ++m)
          sts16(tile0_addr(chunk, row, ((col0 & 63) >> 3) + m), make_uint4(pd[4 * m], pd[4 * m + 1], pd[4 * m + 2], pd[4 * m + 3]));
      }
      tc_wait_st();
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      if (tid == 0 || tid == 32) {
        // dV_j += P^T dO_i (thread 0), dK_j += dS^T Q_i (thread 32): A from tensor memory
        tc_fence_after();
        const int b_what = tid == 0 ? sm::DO : sm::Q;
        const uint32_t td = tid == 0 ? tV : tK, ta = tmem + (tid == 0 ? 0u : 128u);
        const uint64_t b0 = smem_desc(tile0(b_what, i), 1024, 1024, 2), b1 = smem_desc(tile1(b_what, i), s1::SBO1, s1::SBO1, s1::LAYOUT1);
        const uint32_t id0 = idesc_n(64, true), id1 = idesc_n(C1 ? C1 : 16, true);
        for (int ks = 0; ks < nq_mma / 16; ++ks) {
          const uint32_t accf = (i > 0 || ks > 0) ? 1u : 0u;
          const uint32_t acol = (uint32_t)((ks >> 2) * 64 + (ks & 3) * 8);
          tc_mma_bf16_ts(td, ta + acol, b0 + (uint64_t)(ks * (2048 >> 4)), id0, accf);
          if (C1) tc_mma_bf16_ts(td + 64, ta + acol, b1 + (uint64_t)(ks * ((16 * s1::C1B) >> 4)), id1, accf);
        }
        tc_commit(bar_d);
      } else if (tid == 64) {
        // dQ_i part = dS K_j: both operands MN-major (A: queries contiguous in the dS^T tile, B: head dim contiguous)
        tc_fence_after();
        const uint64_t bk0 = smem_desc(tile0(sm::K, j), 1024, 1024, 2), bk1 = smem_desc(tile1(sm::K, j), s1::SBO1, s1::SBO1, s1::LAYOUT1);
        const uint32_t iq0 = idesc_n(64, true) | (1u << 15), iq1 = idesc_n(C1 ? C1 : 16, true) | (1u << 15);
        for (int ks = 0; ks < kb.n_mma / 16; ++ks) {
          const uint64_t a = smem_desc(ds_base + (uint32_t)(ks * 2048), sm::T0, 1024, 2);
          tc_mma_bf16(tQ, a, bk0 + (uint64_t)(ks * (2048 >> 4)), iq0, ks > 0 ? 1u : 0u);
          if (C1) tc_mma_bf16(tQ + 64, a, bk1 + (uint64_t)(ks * ((16 * s1::C1B) >> 4)), iq1, ks > 0 ? 1u : 0u);
        }
        tc_commit(bar_d);
      }
      __syncwarp();
      mbar_wait(bar_d, step & 1u);
      tc_fence_after();
      ++step;
      // dQ part: queries on the lanes now
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        if (a_beg + cc < a_end) {
          uint32_t r[16];
          tc_ld16(t_row + 256u + 2 * (64 + C1) + (uint32_t)((a_beg + cc) * 16), r);
          tc_wait_ld();
#pragma unroll
          for (int e = 0; e < 16; ++e) acc[i][cc * 16 + e] += __uint_as_float(r[e]);
        }
      }
      tc_fence_before();
      __syncthreads();                                       // the next step overwrites S^T / dP^T, the dS^T tile and dQ part
      tc_fence_after();
    }
    // dK_j (column half 0) and dV_j (half 1) are complete: rows = keys.  Their tiles K_j / V_j are dead now and stage
    // the bf16 rows for a coalesced copy-out
    {
      constexpr int RB = (64 + C1) / 8 * 16;                 // bytes per staged row
      uint8_t* stage = smem_raw + (base - smem_u32(smem_raw)) + (half ? sm::V : sm::K) + j * sm::TILE;
      const uint32_t t_src = t_row + (half ? (256u + 64u + C1) : 256u);
#pragma unroll
      for (int cc = 0; cc < (64 + C1) / 16; ++cc) {
        uint32_t r[16];
        tc_ld16(t_src + (uint32_t)(cc * 16), r);
        tc_wait_ld();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint4 u;
          u.x = pack2(__uint_as_float(r[h * 8 + 0]), __uint_as_float(r[h * 8 + 1]));
          u.y = pack2(__uint_as_float(r[h * 8 + 2]), __uint_as_float(r[h * 8 + 3]));
          u.z = pack2(__uint_as_float(r[h * 8 + 4]), __uint_as_float(r[h * 8 + 5]));
          u.w = pack2(__uint_as_float(r[h * 8 + 6]), __uint_as_float(r[h * 8 + 7]));
          *reinterpret_cast<uint4*>(stage + row * RB + cc * 32 + h * 16) = u;
        }
      }
      if (half == 0) {
        long long o = -1;
        if (row < kb.real_here) {
          const int kidx = kb.key0 + row;
          const int ky = kidx / cl.rw, kx = kidx - ky * cl.rw;
          o = (((long long)it.b * p.H + it.yk0 + ky) * p.W + it.x0 + kx) * (3LL * C);
        }
        offs[0][row] = o;
      }
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
      copy_rows_out<C1>(stage, offs[0], p.dqkv + (half ? 2 : 1) * C + head * p.hd, p.hd, tid & 127, 128);
      __syncthreads();                                       // offs / the next key block's accumulators
    }
  }
  // dQ: thread = (query row, column half) of both query tiles -> staged in the (dead) Q tiles, copied out coalesced
  {
    constexpr int RB = (64 + C1) / 8 * 16;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (i >= n_qt) break;
      uint8_t* stage = smem_raw + (base - smem_u32(smem_raw)) + sm::Q + i * sm::TILE;
#pragma unroll
      for (int cc = 0; cc < NACC / 8; ++cc) {
        if (a_beg * 16 + cc * 8 < a_end * 16) {
          uint4 u;
          u.x = pack2(acc[i][cc * 8 + 0], acc[i][cc * 8 + 1]); u.y = pack2(acc[i][cc * 8 + 2], acc[i][cc * 8 + 3]);
          u.z = pack2(acc[i][cc * 8 + 4], acc[i][cc * 8 + 5]); u.w = pack2(acc[i][cc * 8 + 6], acc[i][cc * 8 + 7]);
          *reinterpret_cast<uint4*>(stage + row * RB + (a_beg * 2 + cc) * 16) = u;
        }
      }
      if (half == 0) {
        long long o = -1;
        if (row < q_rows_of(i)) {
          const int ry = row / cl.rw, rx = row - ry * cl.rw;
          o = (((long long)it.b * p.H + it.yk0 + i * cl.qbh + ry) * p.W + it.x0 + rx) * (3LL * C);
        }
        offs[i][row] = o;
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (i >= n_qt) break;
      copy_rows_out<C1>(smem_raw + (base - smem_u32(smem_raw)) + sm::Q + i * sm::TILE, offs[i], p.dqkv + head * p.hd, p.hd, tid, NTHR_KV);
    }
  }
  tc_fence_before();
  __syncthreads();
  ATC_STAMP(7);
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// ----------------------------------------------------------------------------------------------- host side

// 5-D view (d, part*nh + head, x, y, b) of a [B, H, W, parts*nh*hd] bf16 tensor; stream mode: x = H*W, y = 1
static int make_map5(CUtensorMap* out, const void* ptr, int B, int H, int W, int parts_nh, int hd, bool stream,
                     int box_d, int box_x, int box_y, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return S2U_EUNSUPPORTED;
  const cuuint64_t row = (cuuint64_t)parts_nh * hd * 2;      // bytes per token
  cuuint64_t dims[5], strides[4];
  dims[0] = (cuuint64_t)hd; dims[1] = (cuuint64_t)parts_nh;
  strides[0] = (cuuint64_t)hd * 2;
  if (stream) {
    dims[2] = (cuuint64_t)H * W; dims[3] = 1;
    strides[1] = row; strides[2] = row * H * W;
  } else {
    dims[2] = (cuuint64_t)W; dims[3] = (cuuint64_t)H;
    strides[1] = row; strides[2] = row * W;
  }
  dims[4] = (cuuint64_t)B;
  strides[3] = row * H * W;
  cuuint32_t box[5] = {(cuuint32_t)box_d, 1, (cuuint32_t)box_x, (cuuint32_t)box_y, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -100 - (int)r;
}

struct PlanKey {
  const void *qkv, *dout;
  int B, H, W, nh, hd, window;
  bool operator==(const PlanKey& o) const {
    return qkv == o.qkv && dout == o.dout && B == o.B && H == o.H && W == o.W && nh == o.nh && hd == o.hd &&
           window == o.window;
  }
};
struct PlanKeyHash {
  size_t operator()(const PlanKey& k) const {
    size_t h = (size_t)k.qkv * 1000003u ^ (size_t)k.dout;
    for (int v : {k.B, k.H, k.W, k.nh, k.hd, k.window}) h = h * 1000003u ^ (size_t)v;
    return h;
  }
};
struct Plan {
  TcParams p;
  int total_items;          // query-tile work items
  int total_kitems;         // key-block work items (dK/dV kernel)
  int total_windows;        // window work items (fused backward)
  bool fused_ok;            // every window has <= 2 query tiles
  bool bwd_ok;              // every key block has room for the virtual pad key
  bool stream;
};

// can this path run the problem?  (else the caller falls back to the mma.sync kernels)
static bool supported(int H, int W, int hd, int window, int pool) {
  if (pool) return false;
  if (hd <= 64 || hd > 96 || (hd & 7)) return false;        // head dim = 64-column chunk + remainder of 8..32
  if (window == 0) return true;
  if (window * window > 256 || window * window <= 64) return false;   // small windows: packed / mma.sync kernels
  return true;
}

// window-shape classes, tensor maps, work-item ranges; cached per (tensor, geometry)
static int get_plan(Plan** out, const void* qkv, const void* dout, int B, int H, int W, int nh, int hd, int window) {
  static std::unordered_map<PlanKey, Plan, PlanKeyHash> cache;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  const PlanKey key{qkv, dout, B, H, W, nh, hd, window};
  auto f = cache.find(key);
  if (f != cache.end()) { *out = &f->second; return 0; }
  Plan pl;
  memset(&pl, 0, sizeof(pl));
  TcParams& p = pl.p;
  const int C1 = hd - 64;
  const CUtensorMapSwizzle swz1 = C1 == 16 || C1 == 8 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B;
  const int box1 = C1 <= 16 ? 16 : 32;
  p.B = B; p.H = H; p.W = W; p.nh = nh; p.hd = hd;
  p.scale = 1.0f / sqrtf((float)hd);
  pl.stream = window == 0;
  pl.bwd_ok = true;
  pl.fused_ok = !pl.stream;
  int items = 0, kitems = 0, witems = 0;
  auto add_class = [&](int rw, int rh, int wx0, int wy0, int cnx, int cny, int n_pad) -> int {
    const int c = p.n_cls++;
    TcClass& k = p.cls[c];
    k.rw = rw; k.rh = rh; k.wx0 = wx0; k.wy0 = wy0; k.cnx = cnx; k.cny = cny;
    k.n_real = rw * rh; k.n_pad = n_pad;
    int qbx, qby, kbx, kby;
    if (pl.stream) {
      k.qbh = 1; k.n_qt = (rw + 127) / 128;
      qbx = 128; qby = 1; kbx = 128; kby = 1;
      k.n_kb = k.n_qt;
    } else {
      k.qbh = std::min(rh, 128 / rw);
      k.n_qt = (rh + k.qbh - 1) / k.qbh;
      qbx = rw; qby = k.qbh; kbx = rw; kby = rh;
      k.n_kb = k.n_qt;                                       // key blocks = the query-tile row boxes
      const int last_real = k.n_real - (k.n_kb - 1) * k.qbh * rw;
      if (n_pad > 0 && last_real + 1 > 128) pl.bwd_ok = false;   // no room for the pad key: mma.sync backward
    }
    k.item0 = items;
    k.n_items = cnx * cny * B * k.n_qt;
    items += k.n_items;
    k.kitem0 = kitems;
    kitems += cnx * cny * B * k.n_kb;
    k.witem0 = witems;
    witems += cnx * cny * B;
    if (k.n_qt > 2) pl.fused_ok = false;
    int rc = make_map5(&p.q0[c], qkv, B, H, W, 3 * nh, hd, pl.stream, 64, qbx, qby, CU_TENSOR_MAP_SWIZZLE_128B);
    if (!rc) rc = make_map5(&p.q1[c], qkv, B, H, W, 3 * nh, hd, pl.stream, box1, qbx, qby, swz1);
    if (!rc) rc = make_map5(&p.k0[c], qkv, B, H, W, 3 * nh, hd, pl.stream, 64, kbx, kby, CU_TENSOR_MAP_SWIZZLE_128B);
    if (!rc) rc = make_map5(&p.k1[c], qkv, B, H, W, 3 * nh, hd, pl.stream, box1, kbx, kby, swz1);
    if (!rc && dout) {
      rc = make_map5(&p.o0[c], dout, B, H, W, nh, hd, pl.stream, 64, qbx, qby, CU_TENSOR_MAP_SWIZZLE_128B);
      if (!rc) rc = make_map5(&p.o1[c], dout, B, H, W, nh, hd, pl.stream, box1, qbx, qby, swz1);
    }
    return rc;
  };
  int rc = 0;
  if (pl.stream) {
    p.wh = H; p.ww = W;
    rc = add_class(H * W, 1, 0, 0, 1, 1, 0);
  } else {
    p.wh = p.ww = window;
    const int nfx = W / window, nfy = H / window, remx = W % window, remy = H % window;
    const int area = window * window;
    if (nfx && nfy) rc = add_class(window, window, 0, 0, nfx, nfy, 0);
    if (!rc && remx && nfy) rc = add_class(remx, window, nfx, 0, 1, nfy, area - remx * window);
    if (!rc && remy && nfx) rc = add_class(window, remy, 0, nfy, nfx, 1, area - remy * window);
    if (!rc && remx && remy) rc = add_class(remx, remy, nfx, nfy, 1, 1, area - remx * remy);
  }
  if (rc) return rc;
  pl.total_items = items;
  pl.total_kitems = kitems;
  pl.total_windows = witems;
  if (cache.size() > 4096) cache.clear();
  auto ins = cache.emplace(key, pl);
  *out = &ins.first->second;
  return 0;
}

template <int C1>
static int launch_fwd(const Plan& pl, const TcParams& p, cudaStream_t st) {
  using sm = Smem<C1>;
  dim3 grid(p.nh, pl.total_items);
  if (pl.stream) {
    S2U_ALLOW_SMEM((fwd_kernel<C1, true>));
    S2U_LAUNCH((fwd_kernel<C1, true>), grid, NTHR_F, sm::FWD_BYTES, st, p);
  } else {
    S2U_ALLOW_SMEM((fwd_kernel<C1, false>));
    S2U_LAUNCH((fwd_kernel<C1, false>), grid, NTHR_F, sm::FWD_BYTES, st, p);
  }
  S2U_LAUNCH_CHECK();
  return 0;
}

template <int C1>
static int launch_bwd(const Plan& pl, const TcParams& p, cudaStream_t st) {
  static int skip = -1;                                      // S2U_ATC_SKIP=1 / 2: leave out the dQ / dK-dV kernel (timing)
  if (skip < 0) { const char* e = getenv("S2U_ATC_SKIP"); skip = e ? atoi(e) : 0; }   // 3: never the fused kernel
  if (C1 == 16 && pl.fused_ok && skip != 3) {
    if constexpr (C1 == 16) {
      dim3 grid(p.nh, pl.total_windows);
      S2U_ALLOW_SMEM((bwd_win_kernel<16>));
      S2U_LAUNCH((bwd_win_kernel<16>), grid, NTHR_KV, SmemWin<16>::BYTES, st, p);
      S2U_LAUNCH_CHECK();
    }
    return 0;
  }
  if (skip != 1) {
    dim3 grid(p.nh, pl.total_items);
    if (pl.stream) {
      S2U_ALLOW_SMEM((bwd_dq_kernel<C1, true>));
      S2U_LAUNCH((bwd_dq_kernel<C1, true>), grid, NTHR, SmemDq<C1>::BYTES, st, p);
    } else {
      S2U_ALLOW_SMEM((bwd_dq_kernel<C1, false>));
      S2U_LAUNCH((bwd_dq_kernel<C1, false>), grid, NTHR, SmemDq<C1>::BYTES, st, p);
    }
    S2U_LAUNCH_CHECK();
  }
  if (skip != 2) {
    dim3 grid(p.nh, pl.total_kitems);
    if (pl.stream) {
      S2U_ALLOW_SMEM((bwd_dkv_kernel<C1, true>));
      S2U_LAUNCH((bwd_dkv_kernel<C1, true>), grid, NTHR_KV, SmemKv<C1>::BYTES, st, p);
    } else {
      S2U_ALLOW_SMEM((bwd_dkv_kernel<C1, false>));
      S2U_LAUNCH((bwd_dkv_kernel<C1, false>), grid, NTHR_KV, SmemKv<C1>::BYTES, st, p);
    }
    S2U_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace atc

// entry points used by attention.cu's dispatch (bf16 only); S2U_EUNSUPPORTED = not covered by this path
int s2u_attn_tc_fwd(const void* qkv, const float* bias, void* out, float* lse, int B, int H, int W, int nh, int hd,
                    int window, int pool, cudaStream_t st) {
  if (!atc::supported(H, W, hd, window, pool)) return S2U_EUNSUPPORTED;
  atc::Plan* pl;
  int rc = atc::get_plan(&pl, qkv, nullptr, B, H, W, nh, hd, window);
  if (rc) return rc;
  atc::TcParams p = pl->p;
  p.bias = bias; p.qkv = (const bf16*)qkv; p.out = (bf16*)out; p.lse = lse;
#ifdef S2U_ATC_TIMING
  { const char* e = getenv("S2U_ATC_TIMING_BUF"); p.dws = e ? (float*)strtoull(e, nullptr, 0) : nullptr; }
#endif
  if (hd - 64 <= 16) return atc::launch_fwd<16>(*pl, p, st);
  return atc::launch_fwd<32>(*pl, p, st);
}

int s2u_attn_tc_bwd(const void* qkv, const float* bias, const void* out, const float* lse, const void* dout,
                    void* dqkv, float* dws, int B, int H, int W, int nh, int hd, int window, int pool,
                    cudaStream_t st) {
  if (!atc::supported(H, W, hd, window, pool) || !dws) return S2U_EUNSUPPORTED;
  atc::Plan* pl;
  int rc = atc::get_plan(&pl, qkv, dout, B, H, W, nh, hd, window);
  if (rc) return rc;
  if (!pl->bwd_ok) return S2U_EUNSUPPORTED;
  atc::TcParams p = pl->p;
  p.bias = bias; p.qkv = (const bf16*)qkv; p.lse = const_cast<float*>(lse);
  p.o_in = (const bf16*)out; p.dout = (const bf16*)dout; p.dqkv = (bf16*)dqkv; p.dws = dws;
#ifdef S2U_ATC_TIMING
  { const char* e = getenv("S2U_ATC_TIMING_BUF"); if (e) p.dws = (float*)strtoull(e, nullptr, 0); }
#endif
  if (hd - 64 <= 16) return atc::launch_bwd<16>(*pl, p, st);
  return atc::launch_bwd<32>(*pl, p, st);
}
