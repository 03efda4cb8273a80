/* sam2unet_b200.h — C ABI of the B200-native (sm_100a) SAM2-UNet hot path.
 *
 * One shared object (sam2_unet_b200/libsam2unet_b200.so), plain pointers and sizes, no PyTorch types.  Every
 * entry point replaces the ATen call(s) the reference makes at the cited lines (paths relative to the
 * reference repository hanguyenh2/SAM2-UNet); the reference has no FFI of its own for this path — its Python
 * modules call torch.nn.functional directly — so these are the functions a reference-side binding (ctypes
 * stub) would bind in place of those calls.
 *
 * Conventions
 *   - return value: 0 = ok; > 0 = cudaError_t of the failed launch; -1 = invalid argument;
 *     -2 = shape/alignment this kernel does not support; <= -100 = -(100 + CUresult) from the driver.
 *     There is no CPU or library fallback behind any entry point.
 *   - `dtype`: 0 = float32, 1 = bfloat16; it selects the element type of every `void*` activation tensor.
 *     Parameters of norms / biases / BN and all statistics are float32; reductions accumulate in fp32/fp64.
 *   - all tensors are device pointers owned by the caller; kernels never allocate; everything is enqueued on
 *     `stream` (a cudaStream_t); activations are NHWC ("tokens x channels") row-major, `ld*` = row pitch in
 *     elements.  Vector paths need 16-byte aligned pointers and channel counts / pitches that are multiples of 8.
 */
#ifndef SAM2UNET_B200_H
#define SAM2UNET_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- GEMM family ------------------------------------------------------------------------------------------
 * C[M,N] = epi(A[M,K] . W[N,K]^T): nn.Linear forward (sam2/modeling/backbones/hieradet.py:59,79,
 * sam2/modeling/sam2_utils.py:127-132, SAM2UNet.py:57-63), its input gradient (W = pre-transposed weight), and
 * conv forward / input gradient on im2col rows (SAM2UNet.py:83-86).  epi(v): v += bias[n]; pre_out = v;
 * flags&1: v = gelu(v); flags&2: v *= gelu'(aux); flags&4: v += resid; flags&16: C is fp32; flags&32: resid is
 * fp32; flags&64: pre_out receives the FINAL value (compute-dtype copy of C) instead of the pre-activation one;
 * flags&128: pre_out receives gelu'(v) instead of v (so the backward pass only multiplies); flags&256: v *= aux;
 * flags&512: v = max(v, 0) last (eval-mode conv with folded BatchNorm, residual and ReLU in one call).
 * backend: 0 auto (tcgen05 + TMA for bf16: CTA-pair kernel when M > 128), 1 SIMT fp32-FMA, 2 tcgen05 required,
 * 16+bn one-CTA persistent tcgen05 kernel with N tile bn, 1024+bn CTA-pair (cta_group::2) kernel with N tile bn. */
int s2u_gemm(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K,
             const float* bias, void* pre_out, int ld_pre, const void* aux, int ld_aux, const void* resid,
             int ld_res, int flags, int dtype, int backend, void* stream);
/* G[P,Q] += A[M,P]^T . B[M,Q] (fp32 G): weight gradients of the adapters (SAM2UNet.py:57-59) and convs
 * (SAM2UNet.py:72-80); q_inner/q_taps map q = tap*Cin+ci to the [Cout,Cin,kh,kw] parameter layout (0,0: plain). */
/* Number of bf16 GEMM / weight-gradient calls so far that backend 0 routed to the fp32-FMA kernels because the
 * TMA / tcgen05 path cannot describe them (pointer or pitch alignment); reset != 0 clears the counter.  0 on every
 * shape of the Hiera trunks. */
int s2u_gemm_simt_fallbacks(int reset);
int s2u_gemm_wgrad(const void* A, int lda, const void* B, int ldb, float* G, int ldg, long long M, int P, int Q,
                   int q_inner, int q_taps, int dtype, void* stream);
/* Two weight gradients sharing the row count M in one launch (an adapter's dW2 and dW1, SAM2UNet.py:57-59): G0 +=
 * A0^T B0, G1 += A1^T B1, plain [P,Q] layouts. */
int s2u_gemm_wgrad_pair(const void* A0, int lda0, const void* B0, int ldb0, float* G0, int ldg0, int P0, int Q0,
                        const void* A1, int lda1, const void* B1, int ldb1, float* G1, int ldg1, int P1, int Q1,
                        long long M, int dtype, void* stream);
/* out[P] += column sums of A[M,P]: bias gradients. */
int s2u_colsum(const void* A, int lda, float* out, long long M, int P, int dtype, void* stream);

/* ---- LayerNorm (hieradet.py:99-100,104,120,134,166; eps 1e-6) ------------------------------------------- */
/* x_f32 != 0: x (the residual stream) is fp32 even when dtype selects bf16 for y / dy / dx. */
int s2u_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                      long long R, int C, float eps, int x_f32, int dtype, void* stream);
/* dx = LN'(dy) + dres (dres may be NULL); the affine parameters are frozen (SAM2UNet.py:146-147).
 * Optional fused adapter tail (SAM2UNet.py:57-63): dx2 = dx * gelu'(h) with pre = h (pre_is_grad 0) or pre = gelu'(h)
 * as saved by s2u_gemm's GEMM_SAVE_DGELU epilogue (pre_is_grad 1); colsum[C] += column sums of dx2 (NULL to skip).
 * ws: fp32 workspace of s2u_layernorm_ws_floats(C) elements, required with colsum; zero it once, the kernel leaves
 * it zeroed.  Calls that may overlap on different streams need separate workspaces. */
int s2u_layernorm_ws_floats(int C);
int s2u_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                      const void* dres, void* dx, const void* pre, void* dx2, float* colsum, float* ws,
                      int pre_is_grad, long long R, int C, int x_f32, int dtype, void* stream);

/* ---- fused adapter + LayerNorm (bf16 operands, fp32 residual stream) ---------------------------------------
 * Forward of SAM2UNet.py:57-63 followed by hieradet.py:134 in ONE launch:
 *   xa = x + gelu(gelu(x W1^T + b1) W2^T + b2)  (fp32),  n1 = LayerNorm(xa) (bf16), mean / rstd of the rows,
 *   and for the backward u = gelu(.) [R,32], g1 = gelu'(.) [R,32], g2 = gelu'(.) [R,C] (all three NULL at inference).
 * W1 [32,C], W2 [C,32] bf16.  Backward in ONE launch: dxa = LN'(dn1) + dres (dres may be NULL),
 *   dh2 = dxa * g2, dh1 = (dh2 W2) * g1, dx = dxa + dh1 W1, db1[32] += colsum(dh1), db2[C] += colsum(dh2);
 *   W2t [32,C], W1t [C,32] are the transposed weights; dh2 / dh1 are written for s2u_gemm_wgrad_pair.
 * s2u_adapter_supported(C) = 1 when C = 16 * {2,6,7,9} * {1,2,4,8} (every Hiera stage width), else the entry
 * points return -2 and the caller uses s2u_gemm + s2u_layernorm_*. */
int s2u_adapter_supported(int C);
int s2u_adapter_ln_fwd(const float* x, const void* W1, const float* b1, const void* W2, const float* b2,
                       const float* gamma, const float* beta, float eps, float* xa, void* n1, float* mean, float* rstd,
                       void* u, void* g1, void* g2, long long R, int C, void* stream);
int s2u_adapter_ln_bwd(const void* dn1, const float* xa, const float* mean, const float* rstd, const float* gamma,
                       const void* dres, const void* g2, const void* g1, const void* W2t, const void* W1t, void* dh2,
                       void* dh1, void* dx, float* db1, float* db2, long long R, int C, void* stream);

/* ---- element-wise helpers --------------------------------------------------------------------------------- */
int s2u_dgelu_mul(const void* dy, const void* pre, void* out, long long n, int dtype, void* stream);
int s2u_add(const void* a, const void* b, void* out, long long n, int dtype, void* stream);
/* 2x2/stride-2 max-pool of NHWC tokens and its argmax-routed gradient (hieradet.py:21-32,108-110,137-138). */
int s2u_maxpool2_fwd(const void* x, void* out, int B, int H, int W, int C, int dtype, void* stream);
int s2u_maxpool2_bwd(const void* x, const void* dout, void* dx, int B, int H, int W, int C, int dtype,
                     void* stream);
/* fp32 master -> compute-dtype shadow of a [R,C] parameter, optionally transposed. */
int s2u_cast(const float* src, void* dst, int R, int C, int transpose, int dtype, void* stream);
/* all shadows in one launch: `entries` = device array of {u64 src, u64 dst0, u64 dst1, i32 kind, i32 d0..d3, i32 pad}
 * (kind 0: [d0,d1] matrix -> dst0 as is + dst1 transposed; kind 1: conv weight [d0,d1,d2,d3] -> dst0 [Cout][tap][Cin],
 * dst1 [Cin][flipped tap][Cout]); `blocks` = device array of int2 (entry, 1024-element chunk), one per CUDA block. */
int s2u_refresh_shadows(const void* entries, const void* blocks, int nblocks, int dtype, void* stream);

/* ---- windowed / global attention (hieradet.py:56-81,141-162; backbones/utils.py:16-55) ------------------ *
 * qkv [B,H,W,3*nh*hd]; window = 0 means global; pool = 1 applies the 2x2 q max-pool inside each window;
 * bias = the qkv bias (value of zero-padded tokens); out [B,Ho,Wo,nh*hd], lse [B,Ho,Wo,nh];
 * dws = fp32 workspace [B,Ho,Wo,nh] of the backward.  bf16 runs on tensor cores, fp32 on an exact FFMA kernel. */
int s2u_win_attn_fwd(const void* qkv, const float* bias, void* out, float* lse, int B, int H, int W, int nh, int hd,
                     int window, int pool, int dtype, void* stream);
/* bf16 kernel family: 0 = auto (tcgen05 + TMEM + TMA kernels of attention_tc.cu where they apply - no q-pool, window of
 * 65..256 tokens or global, head dim 72..96 - else the mma.sync kernels), 1 = never tcgen05, 2 = tcgen05 or
 * S2U_EUNSUPPORTED (tests / A-B measurements).  Environment S2U_ATTN_BACKEND sets the initial value. */
int s2u_set_attn_backend(int backend);
int s2u_win_attn_bwd(const void* qkv, const float* bias, const void* out, const float* lse, const void* dout,
                     void* dqkv, float* dws, int B, int H, int W, int nh, int hd, int window, int pool, int dtype,
                     void* stream);

/* ---- stem and convolutions -------------------------------------------------------------------------------- *
 * 7x7/s4/p3 conv 3->E + bias + position-embedding table (backbones/utils.py:80-88, hieradet.py:268-283). */
/* out: fp32 when out_f32 else the compute dtype; out_copy (optional): compute-dtype copy of the same tokens. */
int s2u_patch_embed(const float* x, const float* w, const float* bias, const float* pos, void* out, int out_f32,
                    void* out_copy, int B, int S, int E, int dtype, void* stream);
/* Tensor-core form of the stem in bf16 mode: rows [B*(S/4)^2, 320] of the 7x7 / stride 4 / pad 3 patch matrix of the
 * NCHW fp32 image, columns [0,147) = bf16(x) in nn.Conv2d weight order (ci, ky, kx), columns [160,307) =
 * bf16(x - bf16(x)), zero elsewhere; multiplied by [W | W] with s2u_gemm (bias + position table as fp32 residual). */
int s2u_patch_im2col(const float* x, void* out, int B, int S, void* stream);
/* taps of a stride-1 conv gathered to [B*H*W, KH*KW*Cin] (SAM2UNet.py:68-125,9-26). */
int s2u_im2col(const void* x, int ldx, void* out, int B, int H, int W, int Cin, int KH, int KW, int dil_h, int dil_w,
               int pad_h, int pad_w, int dtype, void* stream);
/* Implicit-GEMM convolution (bf16, tcgen05 + 4-D TMA boxes, no im2col matrix): forward of SAM2UNet.py:83-86 and its
 * input gradient (x = d(raw), Wm = the flipped / transposed operand of s2u_conv_weight_pack).  Stride 1, "same" zero
 * padding, dilation dil; Cin % 64 == 0, N in {64, 128, 256} (s2u_conv_igemm_supported), else -2.
 * out = epi(conv): + bias[N] (NULL: none), + resid (NULL: none; resid == out accumulates), relu.  sums (N == 64, or
 * NULL): the fp64 workspace of s2u_bn_ws_doubles(64), receives the BatchNorm batch statistics of the rounded output
 * (finalise with s2u_bn_finalize) - no separate statistics pass. */
int s2u_conv_igemm_supported(int Cin, int N, int ldx, int ld_out);
int s2u_conv_igemm(const void* x, int ldx, int B, int H, int W, int Cin, const void* Wm, int N, int KH, int KW, int dil,
                   void* out, int ld_out, const float* bias, const void* resid, int ld_res, int relu, double* sums,
                   void* stream);
/* Training forward of conv + BatchNorm batch statistics in ONE launch (N = 64): the CTA that finishes last finalises
 * scale / shift (for s2u_bn_apply), the saved mean / rstd and the running statistics, like s2u_bn_stats_finalize. */
int s2u_conv_igemm_bn(const void* x, int ldx, int B, int H, int W, int Cin, const void* Wm, int KH, int KW, int dil,
                      void* out, int ld_out, double* sums, const float* gamma, const float* beta, float* running_mean,
                      float* running_var, long long* num_batches, float* scale, float* shift, float* save_mean,
                      float* save_rstd, float eps, float momentum, void* stream);
/* G [Cout][Cin][KH][KW] (fp32, accumulated) += weight gradient of the convolution from the un-expanded activations:
 * dy [B,H,W,Cout <= 64] pitch ld_dy, x [B,H,W,Cin % 64 == 0] pitch ldx (bf16). */
int s2u_conv_wgrad(const void* dy, int ld_dy, const void* x, int ldx, float* G, int B, int H, int W, int Cin, int Cout,
                   int KH, int KW, int dil, void* stream);
int s2u_conv_weight_pack(const float* w, void* wf, void* wd, int Cout, int Cin, int KH, int KW, int dtype,
                         void* stream);

/* ---- BatchNorm2d, eps 1e-5, momentum 0.1 (SAM2UNet.py:80,85,18,21) ---------------------------------------- */
/* `sums`: fp64 workspace of s2u_bn_ws_doubles(C) elements (replicated accumulators + 1 ticket word), zero on entry,
 * left zero by the finalising kernel.
 * s2u_bn_stats_finalize = training-mode statistics + finalisation in ONE launch (the last block finalises). */
int s2u_bn_ws_doubles(int C);
int s2u_bn_stats(const void* x, int ldx, double* sums, long long M, int C, int dtype, void* stream);
int s2u_bn_stats_finalize(const void* x, int ldx, double* sums, const float* gamma, const float* beta,
                          float* running_mean, float* running_var, long long* num_batches, float* scale, float* shift,
                          float* save_mean, float* save_rstd, long long M, int C, float eps, float momentum, int dtype,
                          void* stream);
int s2u_bn_finalize(double* sums, const float* gamma, const float* beta, float* running_mean, float* running_var,
                    long long* num_batches, float* scale, float* shift, float* save_mean, float* save_rstd,
                    long long M, int C, float eps, float momentum, int training, void* stream);
int s2u_bn_apply(const void* x, int ldx, const float* scale, const float* shift, const void* resid, int ld_res,
                 void* out, int ld_out, long long M, int C, int relu, int dtype, void* stream);
int s2u_relu_bwd(const void* dy, int ld_dy, const void* y, int ld_y, void* g, int ld_g, long long M, int C,
                 int dtype, void* stream);
int s2u_bn_bwd(const void* dy, int ld_dy, const void* y, int ld_y, const void* x, int ldx, const float* mean,
               const float* rstd, const float* gamma, double* sums, float* dgamma, float* dbeta, float* c1, float* c2,
               void* dx, int ld_dx, long long M, int C, int dtype, void* stream);

/* ---- bilinear resampling and heads (SAM2UNet.py:32-49,160-172) ------------------------------------------- */
int s2u_resample_fwd(const void* x, int ldx, void* out, int ld_out, int B, int Hi, int Wi, int Ho, int Wo, int C,
                     const int* y_i0, const int* y_i1, const float* y_w0, const float* y_w1, const int* x_i0,
                     const int* x_i1, const float* x_w0, const float* x_w1, int dtype, void* stream);
int s2u_resample_bwd(const void* dout, int ld_do, void* dx, int ld_dx, int B, int Hi, int Wi, int Ho, int Wo, int C,
                     const int* y_idx, const float* y_w, int y_taps, const int* x_idx, const float* x_w, int x_taps,
                     int dtype, void* stream);
int s2u_resample1_fwd(const float* x, float* out, int B, int Hi, int Wi, int Ho, int Wo, const int* y_i0,
                      const int* y_i1, const float* y_w0, const float* y_w1, const int* x_i0, const int* x_i1,
                      const float* x_w0, const float* x_w1, void* stream);
int s2u_resample1_bwd(const float* dout, float* dx, int B, int Hi, int Wi, int Ho, int Wo, const int* y_idx,
                      const float* y_w, int y_taps, const int* x_idx, const float* x_w, int x_taps, void* stream);
int s2u_head_fwd(const void* feat, int ldf, const float* w, const float* bias, float* out, long long M, int C,
                 int dtype, void* stream);
int s2u_head_bwd(const void* feat, int ldf, const float* w, const float* dlogit, void* dfeat, int ld_df,
                 int accumulate, float* dw, float* db, long long M, int C, int dtype, void* stream);

/* ---- structure_loss (train.py:21-29,76-79) and AdamW (train.py:48-52,83) --------------------------------- *
 * sums: fp64 workspace of nheads*B*2 + nheads entries; weit: fp32 [B,H,W] workspace shared by the heads. */
int s2u_structure_loss_fwd(const float* pred0, const float* pred1, const float* pred2, const float* mask,
                           float* weit, double* sums, float* loss, int B, int H, int W, int nheads, void* stream);
int s2u_structure_loss_bwd(const float* pred0, const float* pred1, const float* pred2, const float* mask,
                           const float* weit, const double* sums, const float* gscale, float* grad0, float* grad1,
                           float* grad2, int B, int H, int W, int nheads, void* stream);
/* hyper (device fp32[4]): lr, gradient scale (1/world_size under data parallelism), beta1^t, beta2^t; the last two
 * start at 1 and are advanced on the device by every call, so the step is replayable from a CUDA graph. */
int s2u_adamw(float* p, const float* g, float* m, float* v, long long n, float* hyper, float beta1,
              float beta2, float eps, float wd, void* stream);

/* ---- inference tail (test.py:66-76, train.py:103-112) ---------------------------------------------------------
 * One image: crop the letterbox padding off the [S,S] fp32 logit map, bilinear-resize to [out_h,out_w]
 * (align_corners = False), sigmoid, min-max normalise over the image, quantise to uint8 (truncation, like numpy's
 * astype).  ws: two ints of device scratch armed once with s2u_infer_tail_init; every call leaves them re-armed. */
int s2u_infer_tail_init(int* ws, void* stream);
int s2u_infer_tail(const float* logits, int S, int pad_left, int pad_top, int pad_right, int pad_bottom, int out_h,
                   int out_w, int* ws, unsigned char* out, void* stream);

/* ---- evaluation metrics (eval.py:55-171) --------------------------------------------------------------------------
 * counts[4] (uint64, accumulated) += |P & G|, |P | G|, |P|, |G| with P = pred > threshold, G = gt > threshold
 * (eval.py:81-101: semantic IoU and Dice). */
int s2u_seg_counts(const unsigned char* pred, const unsigned char* gt, long long n, float threshold,
                   unsigned long long* counts, void* stream);
/* 8-connected components of (mask > threshold) (skimage.measure.label, eval.py:105-106): labels[H*W] = smallest
 * raster index of the pixel's component, -1 for background; the component roots are appended to roots[cap] in no
 * particular order and counted in *nroots (zero on entry). */
int s2u_cc_label(const unsigned char* mask, float threshold, int H, int W, int* labels, int* roots, int* nroots,
                 int cap, void* stream);
/* Component areas (parea / garea [n], indexed by root, zero on entry) and the intersection counts of every
 * overlapping (prediction root, ground-truth root) pair in an open-addressing table: keys[cap] = (proot << 32) |
 * groot (all-ones = empty, on entry), vals[cap] (zero on entry), cap a power of two; *overflow = 1 if it filled up. */
int s2u_cc_stats(const int* plabels, const int* glabels, long long n, int* parea, int* garea, unsigned long long* keys,
                 int* vals, int cap, int* overflow, void* stream);

/* ---- test-time input path (dataset.py:336-407: ImageToTensor, LongestMaxSizeAndPad, NormalizeImage) ------------
 * img: uint8 [H,W,3] RGB on the device; out: fp32 [3,S,S] = normalise(pad(antialiased-bilinear-resize(img / 255)));
 * new_h, new_w, pad_left, pad_top: the host integers of LongestMaxSizeAndPad; mean / std: 3 floats each in HOST
 * memory; tmp: fp32 device scratch of 3 * H * new_w elements. */
int s2u_preprocess(const unsigned char* img, int H, int W, int S, int new_h, int new_w, int pad_left, int pad_top,
                   const float* mean3_host, const float* std3_host, float* tmp, float* out, void* stream);

/* ---- training-time input pipeline on the device (dataset.py:34-333, FullDataset's train transform) -----------
 * The random decisions stay on the host (sam2_unet_b200/augment.py draws them in the reference's order); one launch per
 * transform on a [3,S,S] fp32 image / [1,S,S] fp32 label.
 * s2u_aug_resize_pad: ResizeLongestSideAndPad from the uint8 sources (mode 0: pad to [ph,pw] at (off_y, off_x) with white
 *   fill / label 0; mode 1: crop [ph,pw] at (off_y, off_x)), antialiased bilinear (image) / nearest (label) resize to
 *   [new_h,new_w], centred in S x S with zeros; tmp = 3*ph*new_w floats.
 * s2u_aug_rot90: RandomRotate (k x 90 degrees counter-clockwise, k = 1..3).
 * s2u_aug_color: op 0 grayscale, 1 brightness, 2 contrast (ws = one double of scratch), 3 saturation, 4 hue, 5 gamma,
 *   6 normalise (host mean / std); factor_c = (float)(1.0 - factor) for the blends.
 * s2u_aug_blur: GaussianBlur k = 3 | 5 with the k normalised 1-D weights (host), reflect padding. */
int s2u_aug_resize_pad(const unsigned char* img, const unsigned char* lab, int H, int W, int mode, int off_y, int off_x,
                       int ph, int pw, int S, int new_h, int new_w, int pad_left, int pad_top, float* tmp, float* out_img,
                       float* out_lab, void* stream);
int s2u_aug_rot90(const float* in, float* out, int C, int S, int k, void* stream);
int s2u_aug_color(float* img, int S, int op, float factor, float factor_c, double* ws, const float* m3_host,
                  const float* s3_host, void* stream);
int s2u_aug_blur(const float* in, float* out, int S, int k, const float* w1_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SAM2UNET_B200_H */
