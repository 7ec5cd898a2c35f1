/* xcp.h -- C ABI of libxcp_sm100.so: the B200 (sm_100a) hot path of Tonmoy1321/Multimodal-DeepFake-Detection.
 *
 * The reference has no FFI layer: its boundary is the torch.nn.Module API of Xception.py / XceptionLSTMV.py /
 * XceptionLSTMA.py, and every FLOP is executed by ATen/cuDNN/cuBLAS behind nn.Conv2d / nn.BatchNorm2d /
 * nn.MaxPool2d / nn.LSTM / nn.Linear.  Each entry point below replaces the library call(s) cited beside it
 * (file:line in the reference) so that a maintainer can bind it with ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch types.  All pointers are DEVICE pointers unless noted.
 *   - activations: bf16, NHWC, i.e. a row-major [F*H*W, C] matrix, C % 8 == 0, 16-byte aligned.
 *   - every function takes the CUDA device ordinal and the cudaStream_t (as void*) to launch on, is re-entrant
 *     and thread-safe (autograd calls backward from another thread), never allocates or frees device memory and
 *     keeps no reference to its arguments.
 *   - return value: 0 = ok, < 0 = argument / shape / alignment error, > 0 = cudaError_t.
 *     xcp_last_error_string() (thread-local) explains the last failure.  Nothing throws, nothing exits.
 *   - there is no CPU path and no other-architecture path: xcp_check_device() fails on a non-sm_100 device.
 */
#ifndef XCP_H
#define XCP_H
#ifdef __cplusplus
extern "C" {
#endif

const char* xcp_last_error_string(void);
int xcp_version(void);
int xcp_check_device(int device);

/* ---- pointwise / skip 1x1 convolutions, LSTM input projection: tcgen05 + TMEM + TMA GEMMs ------------------
 * D[M,N] = A[M,K] * B[N,K]^T, bf16 operands, fp32 accumulate.  Replaces nn.Conv2d(k=1) in
 * SeparableConv2d.pointwise (Xception.py:42,46), Block.skip (Xception.py:55,93), and x_t W_ih^T of nn.LSTM
 * (XceptionLSTMV.py:18-23).  Also the data gradient (A = dY, B = W^T).
 * epi: 0 bf16 out | 1 bf16 out + per-channel (sum, sum-sq) partials stats[ceil(M/128)][2][N] for train-mode
 * BatchNorm (Xception.py:67,73,78) | 2 fp32 out (+ optional bias[N]).  lda/ldb/ldo are row pitches in elements. */
int xcp_gemm_tn(const void* A, long long lda, const void* B, long long ldb, void* out, long long ldo, int M, int N, int K,
                int epi, float* stats, const float* bias, int n_real, int k_real, int device, void* stream);
/* n_real / k_real (0 = N / K): logical channel counts when N / K are channel pitches whose tail is zero padding (728 in 768).
 * With XCP_GEMM_TRIM=1 in the environment the tensor-core work that would only multiply padding is skipped (the padded output
 * columns are written as zeros either way).  Off by default: measured slower in the training step (csrc/gemm.cu, set_trim). */
/* Inference plan (eval-mode BatchNorm folded into the weights; the reference's no_grad evaluation, test_visual.py:609-624):
 * out[M,N] = relu?( A[M,K] * B[N,K]^T + bias[N] + residual[M,N] ) as bf16.  B = pointwise weights pre-multiplied per output
 * channel by gamma * rsqrt(running_var + eps) (xcp_pack_weight_scaled), bias = beta - running_mean * that scale; residual
 * (optional, bf16, row pitch ld_res) is the identity-skip input of a Block (Xception.py:96-98).  N % 32 == 0. */
int xcp_gemm_tn_bias(const void* A, long long lda, const void* B, long long ldb, void* out, long long ldo, int M, int N, int K,
                     const float* bias, int relu, const void* residual, long long ld_res, int n_real, int k_real, int device,
                     void* stream);
/* rows of `stats` written by xcp_gemm_tn(epi=1) / xcp_conv3x3_gemm for an M x N problem (<= SM count when N fits one tile) */
int xcp_gemm_stats_parts(long long M, int N, int device);
/* dW[P,Q] += dY[R,P]^T * X[R,Q] (fp32 accumulate into dW): weight gradient of the layers above. */
int xcp_gemm_wgrad(const void* dY, long long ld_dy, const void* X, long long ld_x, float* dW, long long ld_dw, int R, int P,
                   int Q, int device, void* stream);
/* SIMT cross-check of the two GEMMs above (tests only; never on the product path). */
int xcp_gemm_ref(const void* A, long long lda, const void* B, long long ldb, float* out, long long ldo, int M, int N, int K,
                 int mn_major, int device, void* stream);
/* Dense 3x3 stem conv2 (Xception.py:122,172) and its data gradient as an implicit GEMM; see gemm.cu. */
int xcp_conv3x3_gemm(const void* a, const void* b, void* out, float* stats, int F, int Hg, int Wg, int Cin, int Cout, int Ho,
                     int Wo, int sign, int device, void* stream);
/* weight gradient of the same conv in one launch: gk[Cout][tap*Cin + i] (fp32, tap-major packing, accumulated) from
 * dy_grid [F,Hg,Wg,Cout] (zero outside the valid window) and the conv input x [F,Hg,Wg,Cin]  (conv2 backward of
 * Xception.py:122,172) */
int xcp_conv3x3_wgrad(const void* dy_grid, const void* x, float* gk, int F, int Hg, int Wg, int Cin, int Cout, int device,
                      void* stream);

/* ---- stem conv1 3->32 k3 s2 p0 (Xception.py:118,168): bf16 NHWC out + BN partials [xcp_stem_conv1_parts()][2][32].
 * Input x: x_u8_nhwc = 0 -> fp32 NCHW [F,3,H,W] in [0,1] (the tensor video_dataloader.py:35 builds); 1 -> uint8 NHWC [F,H,W,3]
 * (the on-disk frame format, video_dataloader.py:27-35), scaled by 1/255 inside the kernel. */
int xcp_stem_conv1_parts(int F, int H, int W, int device);
int xcp_stem_conv1_fwd(const void* x, int x_u8_nhwc, const float* w, void* y, float* partials, int F, int H, int W, int device,
                       void* stream);
/* inference form (row f-3): out = relu(scale*conv1(x) + shift), eval-mode bn1 folded (Xception.py:168-170 in one pass) */
int xcp_stem_conv1_fwd_affine(const void* x, int x_u8_nhwc, const float* w, const float* scale, const float* shift, void* out, int F,
                              int H, int W, int device, void* stream);
/* weight gradient: im2col (bf16 [M,32]) + MN-major tcgen05 split-K GEMM; `workspace` = xcp_stem_conv1_wgrad_ws_bytes() bytes */
long long xcp_stem_conv1_wgrad_ws_bytes(int F, int H, int W);
int xcp_stem_conv1_wgrad(const void* x, int x_u8_nhwc, const void* dy, float* dW, void* workspace, int F, int H, int W, int device,
                         void* stream);

/* ---- depthwise 3x3 s1 p1 (SeparableConv2d.conv1, Xception.py:41,45) fused with the preceding ReLU
 * (Xception.py:61-76) and the producer's BatchNorm affine.  w9 = tap-major weights [9][C] (xcp_pack_dw). */
int xcp_dw3x3_fwd(const void* x, const float* w9, const float* scale, const float* shift, int relu, void* out, int F, int H,
                  int W, int C, int device, void* stream);
/* backward: dz = mask*conv_transpose(dD) [+ add_full] [+ add_half at even pixels]; dw[C][9] (nn.Conv2d layout) and
 * bnsum[2][C] = (sum dz, sum dz*x) are accumulated with RED (bnsum only when scale/shift given; caller zero-fills) */
int xcp_dw3x3_bwd(const void* dD, const void* xin, const float* w9, const float* scale, const float* shift, int relu, void* dz,
                  const void* add_full, const void* add_half, float* dw, float* bnsum, int F, int H, int W, int C, int c_real,
                  int device, void* stream);

/* Channel padding: activations whose channel count is not a multiple of 64 (728 in the middle flow) are stored with a
 * physical pitch C rounded up to 64 (768) so every pixel row is whole 128-byte lines for TMA; the pad channels are kept
 * exactly zero (zero-padded packed weights, zero BN scale/shift).  Functions that touch per-channel PARAMETER arrays
 * take c_real = the logical channel count (the length of gamma / beta / running stats / dw rows); C is the pitch. */
/* ---- BatchNorm2d (Xception.py:56,67,73,78,119,123,143,147): statistics finalisation (train), affine folding (eval) */
int xcp_bn_finalize(const float* partials, int nparts, int C, int c_real, double count, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, float momentum, float eps, float* scale, float* shift,
                    float* mean_out, float* rstd_out, int device, void* stream);
int xcp_bn_eval_affine(const float* gamma, const float* beta, const float* rm, const float* rv, float eps, float* scale,
                       float* shift, float* mean_out, float* rstd_out, int C, int c_real, int device, void* stream);
/* out = relu?(scale*y + shift) */
int xcp_bn_act(const void* y, const float* scale, const float* shift, int relu, void* out, long long n, int C, int device,
               void* stream);
/* input sampling of the stride-2 skip conv (Xception.py:55,93): out[f,ho,wo,:] = act(x[f,2ho,2wo,:]) */
int xcp_gather_s2(const void* x, const float* scale, const float* shift, int relu, void* out, int F, int H, int W, int C,
                  int device, void* stream);
/* BN + MaxPool2d(3,2,1) (Xception.py:86) + skip BN + residual add (Xception.py:92-98); idx = arg-max taps (uint8);
 * ymax (optional, bf16 [F,Ho,Wo,C]) = the raw y at the arg-max, consumed by xcp_bn_bwd_sums in backward */
int xcp_pool_add_fwd(const void* y, const float* scale, const float* shift, const void* ys, const float* scale_s,
                     const float* shift_s, void* out, void* idx, void* ymax, int F, int H, int W, int C, int device,
                     void* stream);
/* BN + residual add (blocks 4-11, Xception.py:96-98); scale_s/shift_s != NULL applies the skip BN (stride-1 skip conv) */
int xcp_bn_add_fwd(const void* y, const float* scale, const float* shift, const void* skip, const float* scale_s,
                   const float* shift_s, void* out, long long n, int C, int device, void* stream);
/* bn4 + ReLU + adaptive_avg_pool2d (Xception.py:194-198): feat fp32 [F,C] */
int xcp_bn_relu_gap(const void* y, const float* scale, const float* shift, float* feat, int F, int HW, int C, int device,
                    void* stream);
int xcp_bnbwd_num_parts(void);
/* pass 1 of the BatchNorm backward alone: sums[2][C] = (sum G, sum G*y) over an [n_pix, C] bf16 pair; workspace = fp32
 * [xcp_bnbwd_num_parts()][2][C].  For the max-pool blocks it runs on (ymax, G) at pooled resolution (sum dz = sum G,
 * sum dz*y = sum G*y[arg-max]) and its result is handed to xcp_bn_bwd as `presums`. */
int xcp_bn_bwd_sums(const void* y, const void* G, float* workspace, float* sums, long long n_pix, int C, int device,
                    void* stream);
/* two-pass BatchNorm backward with the ReLU / MaxPool / GAP gradient routing folded into its loads; see elementwise.cu */
int xcp_bn_bwd(int mode, const void* y, const void* G, const void* idx, const float* dfeat, const float* scale,
               const float* shift, const float* gamma, const float* mean, const float* rstd, int training, const float* presums,
               float* workspace, float* coef, float* dgamma, float* dbeta, void* dy, int F, int H, int W, int C, int c_real,
               int grid_w, int grid_h, int device, void* stream);

/* ---- layout / packing */
int xcp_nchw_to_nhwc(const float* x, void* out, int F, int C, int Cp, int HW, int device, void* stream);   /* Cp = NHWC pitch */
int xcp_nhwc_to_nchw(const void* x, float* out, int F, int C, int Cp, int HW, int device, void* stream);
/* fp32 [R,Cc] -> bf16 [Rp,Cp] (zero padded) and optionally its transpose bf16 [Cp,Rp] */
int xcp_pack_weight(const float* w, void* out, void* out_t, int R, int Cc, int Rp, int Cp, int device, void* stream);
/* same, every row r multiplied by row_scale[r] before rounding (BatchNorm scale folded into the weights); no transpose */
int xcp_pack_weight_scaled(const float* w, const float* row_scale, void* out, int R, int Cc, int Rp, int Cp, int device,
                           void* stream);
int xcp_pack_dw(const float* w, float* w9, int C, int Cp, int device, void* stream);                      /* w9 = [9][Cp] */
int xcp_unpack_dw_grad(const float* g9, float* gw, int C, int accumulate, int device, void* stream);
/* multi-tensor xcp_pack_weight / xcp_pack_dw: `table` = n_tensors x {const float* src; void* out; void* out_t; int R, Cc, Rp,
 * Cp; int kind; int tile0} (48 bytes, device memory).  kind 0: [R,Cc] fp32 -> bf16 [Rp,Cp] (+ transpose [Cp,Rp] unless
 * out_t is NULL), ceil(Rp/32)*ceil(Cp/32) tiles; kind 1: depthwise [R=C,1,3,3] -> fp32 [9][Cp], ceil(9*Cp/1024) tiles;
 * tile0 = running sum of the tile counts, n_tiles = their total. */
int xcp_pack_multi(const void* table, int n_tensors, int n_tiles, int device, void* stream);
int xcp_pack_conv3x3(const float* w, void* wk, void* wk_t, int O, int I, int device, void* stream);
int xcp_unpack_conv3x3_grad(const float* gk, float* gw, int O, int I, int device, void* stream);
/* F.interpolate(size=(S,S), mode="bilinear", align_corners=False) of [planes, n, 1] (XceptionLSTMA.py:45-46) */
int xcp_bilinear_up(const float* x, float* out, long long planes, int n, int S, int device, void* stream);
int xcp_cast_f32_bf16(const float* x, void* out, long long n, int device, void* stream);

/* ---- nn.LSTM(2048,H,1,batch_first) recurrence + BPTT (XceptionLSTMV.py:18-23,67) */
int xcp_lstm_fwd(const float* xproj, const float* b_ih, const float* b_hh, const void* w_hh_t, float* h_out, float* gates,
                 float* cstate, float* hn, float* cn, int B, int T, int H, int device, void* stream);
int xcp_lstm_bwd(const float* dout, const float* dhn, const float* dcn, const float* gates, const float* cstate,
                 const float* hstate, const void* w_hh, void* dgates, void* hprev, float* dbias_ih, float* dbias_hh, int B, int T,
                 int H, int device, void* stream);

/* ---- classifier head + losses (XceptionLSTMV.py:25-44,68-70; train_audio.py:20; train_visual.py:455-474,532;
 *      train_au_face.py:423-458,659-674; train_au_patch.py:203-211) */
int xcp_linear_small_fwd(const float* a, const float* W, const float* bias, const void* mask, float drop_scale, int act,
                         float* out, int B, int N, int K, int device, void* stream);
int xcp_linear_small_bwd(const float* delta_raw, const float* out_act, float drop_scale, const float* a, const float* W,
                         float* dW, float* db, float* din, int B, int N, int K, int device, void* stream);
int xcp_sigmoid_fwd(const float* z, float* p, int n, int device, void* stream);
int xcp_sigmoid_bwd(const float* p, const float* dp, float* dz, int n, int device, void* stream);
int xcp_bce_fwd_bwd(const float* z, const float* y, float smoothing, float* probs, float* loss, float* dz, int B, int device,
                    void* stream);
/* The whole classifier head + loss in ONE launch per direction (BASELINE north_star (3); XceptionLSTMV.py:25-44,66-70,
 * train_audio.py:20,39, train_au_patch.py:203-211).  Input row b = x + b*row_stride + (row_index ? row_index[b]*H : 0): the
 * lstm_out[:, -1, :] select (or the per-clip last valid step) is an address, not a copy.  wb / dwb: HOST arrays of 10 device
 * pointers {W0,b0,...,W4,b4} (fc_layers[0,3,6,9], fc_out) and their gradient slots (accumulated into; entries may be NULL).
 * Dropout: `mask` = optional uint8 keep masks [4][B][Wd]; else, with p_drop > 0 and `rng` = device {u64 seed, u64 launch
 * counter}, the kernel draws the masks itself and advances the counter (CUDA-graph replays draw fresh masks); both NULL =
 * no dropout.  acts = [4][B][Wd] layer outputs saved for the backward.  loss_mode 0: none; 1: nn.BCELoss on the sigmoid
 * output (loss, dz = dL/dz); 2: BCE-with-logits on y(1-smoothing)+smoothing/2.  bar = 2 x u32, zero before the first
 * launch (every launch leaves it zero).  B <= 32, H % 4 == 0, Wd % 16 == 0, both <= 1024.
 * Backward: dz[b] = dsrc[b] * (prob ? p(1-p) : 1) * (gscale ? *gscale : 1); dacts = [4][B][Wd] scratch; dx (nullable) = gradient
 * wrt the LSTM output with the same row addressing as x; the kernel first zeroes dx_zero_n floats from dx_base (the whole
 * [B,T,H] gradient), then adds the selected rows. */
int xcp_head_mlp_fwd(const float* x, long long row_stride, const long long* row_index, const void* const* wb, const void* mask,
                     void* rng, float p_drop, float* acts, float* z, float* prob, int loss_mode, const float* y, float smoothing,
                     float* loss, float* dz, void* bar, int B, int H, int Wd, int device, void* stream);
int xcp_head_mlp_bwd(const float* dsrc, const float* prob, const float* gscale, const float* x, long long row_stride,
                     const long long* row_index, const float* acts, float drop_scale, const void* const* wb, void* const* dwb,
                     float* dacts, float* dx, float* dx_base, long long dx_zero_n, void* bar, int B, int H, int Wd, int device,
                     void* stream);
/* The fused region of train_au_face.py:659-674 in ONE launch per direction: mean-pool both token streams ([B,Tv,D], [B,Ta,D]), concat,
 * embed_head (Linear(2D,N0) -> ReLU -> Dropout -> Linear(N0,N3)), ArcFace margin logits on N3-wide embeddings (2 classes), loss_mode 0:
 * cross entropy / 1: class-balanced focal (class_w[2], gamma), + lambda_align * mse(v_pool, a_pool) + lambda_temp * 0.5 * (temporal
 * smoothness of both streams).  labels NULL: inference logits s*cos only.  Forward outputs: pooled [B,2D], h [B,N0], e [B,N3],
 * logits [B,2], loss, and the unit-loss gradients de [B,N3], darc_scratch (first 2*N3 floats = d loss / d arc_w; size B*2*N3);
 * rows = 2*B floats of scratch; mask / rng / bar as for xcp_head_mlp_fwd (mask = uint8 [B,N0]).
 * Backward: everything scaled by *gscale (NULL = 1): dW0, db0, dW3, db3, darc accumulated into; dv / da (nullable) written;
 * dh [B,N0], dpooled [B,2D] scratch. */
int xcp_fusion_head_fwd(const float* v, const float* a, int B, int Tv, int Ta, int D, const float* W0, const float* b0, const float* W3,
                        const float* b3, int N0, int N3, const float* arc_w, const long long* labels, float s, float m, int loss_mode,
                        const float* class_w, float gamma, float lambda_align, float lambda_temp, const void* mask, void* rng,
                        float p_drop, float* pooled, float* h, float* e, float* logits, float* loss, float* de, float* darc_scratch,
                        float* rows, void* bar, int device, void* stream);
int xcp_fusion_head_bwd(const float* gscale, const float* v, const float* a, int B, int Tv, int Ta, int D, const float* W0,
                        const float* W3, int N0, int N3, const float* pooled, const float* h, const float* de, const float* darc_unit,
                        float drop_scale, float lambda_align, float lambda_temp, float* dW0, float* db0, float* dW3, float* db3,
                        float* darc, float* dh, float* dpooled, float* dv, float* da, void* bar, int device, void* stream);
/* nn.BCELoss() (mean) on probabilities + its gradient wrt p in one launch (train_audio.py:20,39); dp may be NULL */
int xcp_bce_prob_fwd_bwd(const float* p, const float* y, float* loss, float* dp, int n, int device, void* stream);
int xcp_arcface_loss(const float* x, const float* w, const long long* labels, float s, float m, int loss_mode,
                     const float* class_w, float gamma, const float* dlogits_in, float* logits, float* loss, float* loss_rows,
                     float* dx, float* dw, int B, int D, float gscale, int device, void* stream);
int xcp_fusion_pool_reg(const float* v, const float* a, float* pooled, float* loss_reg, float* dv, float* da, int B, int T, int D,
                        float lambda_align, float lambda_temp, float gscale, int device, void* stream);
int xcp_fusion_pool_bwd(const float* dpooled, float* dv, float* da, int B, int T, int D, int device, void* stream);

/* ---- optimizer side: clip_grad_norm_ + Adam / AdamW over a flat fp32 arena */
int xcp_grad_sumsq(const float* g, long long n, float* out, int zero_first, int device, void* stream);
int xcp_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                  float weight_decay, int decoupled, int step, const float* sumsq, float max_norm, float grad_scale, int device,
                  void* stream);
/* multi-tensor form: `table` = n_tensors x {float* p; const float* g; float* m; float* v; long long n; int* step} in device
 * memory (each tensor's own device-side step counter is incremented by the call, so the launch is CUDA-graph capturable and
 * keeps torch.optim.Adam's per-parameter step semantics), `chunks` = n_chunks x {int tensor, int chunk} (8192-element chunks);
 * max_norm > 0 clips by the global norm first (sumsq_ws: device float scratch).  hyper: optional DEVICE pointer to
 * {lr, weight_decay} overriding the two host scalars, so that a captured CUDA graph follows the LR schedulers of
 * train_visual.py:534,627 / train_au_face.py:620-623 (the host refreshes the two floats before a replay).  Replaces the
 * per-tensor loop of torch.optim.Adam.step / clip_grad_norm_ (train_visual.py:533,574-577; train_au_face.py:616-619,678-693). */
int xcp_adam_multi(const void* table, int n_tensors, const void* chunks, int n_chunks, float lr, float beta1, float beta2,
                   float eps, float weight_decay, int decoupled, float* sumsq_ws, float max_norm, float grad_scale,
                   const float* hyper, int device, void* stream);

/* ---- fp32 validation path (forward only; north_star parity tolerance "fp32 logits within 1e-4 relative").  Plain fp32 FMA
 * kernels on NHWC fp32 activations that read the fp32 master parameters in torch's layouts; correctness instruments for the
 * plan's layout / indexing / BatchNorm bookkeeping, not the production path (host side: fp32_plan.py).
 * Xception.py:44-47 (separable conv), :89-99 (block), :167-199 (network). */
int xcp_f32_conv3x3(const float* x, int x_nchw, const float* w, float* out, int F, int H, int W, int Ci, int Co, int stride,
                    int device, void* stream);                        /* padding 0; w [Co][Ci][3][3]; out NHWC */
int xcp_f32_dw3x3(const float* x, const float* w, float* out, int F, int H, int W, int C, int device, void* stream);
int xcp_f32_gemm(const float* a, const float* w, const float* bias, float* out, long long M, int N, int K, int device,
                 void* stream);                                       /* out[M,N] = a[M,K] . w[N,K]^T (+ bias) */
int xcp_f32_bn_stats_parts(long long M);                              /* rows of the partials buffer for M pixels */
int xcp_f32_bn_stats(const float* y, float* partials, long long M, int C, int device, void* stream); /* -> xcp_bn_finalize */
int xcp_f32_affine(const float* y, const float* scale, const float* shift, int relu, float* out, long long n, int C,
                   int device, void* stream);                         /* scale == NULL: (optional) ReLU only */
int xcp_f32_pool_add(const float* y, const float* skip, float* out, int F, int H, int W, int C, int device, void* stream);
int xcp_f32_add(const float* a, const float* b, float* out, long long n, int device, void* stream);
int xcp_f32_gather(const float* x, float* out, int F, int H, int W, int C, int stride, int device, void* stream);
int xcp_f32_gap(const float* x, float* out, int F, int HW, int C, int device, void* stream);
/* nn.LSTM(I,H,1,batch_first) recurrence, zero initial state; xproj = x . W_ih^T [B*T,4H] (xcp_f32_gemm), w_hh fp32 [4H,H] */
int xcp_f32_lstm_fwd(const float* xproj, const float* b_ih, const float* b_hh, const float* w_hh, float* h_out, float* hn,
                     float* cn, int B, int T, int H, int device, void* stream);

/* ---- fp32 validation path, backward + fused-signature forward twins (csrc/f32_bwd.cu).  Each function shadows the production
 * entry point named beside it -- same arguments and fused semantics (producer-BN affine + ReLU prologue, pool / skip / add
 * routing, BN-backward sums out of the depthwise backward, channel padding with c_real), fp32 NHWC activations instead of
 * bf16 -- so executor.py drives both families through one code path and the chain rule that trains is checked against the
 * fp32 oracle at <= 1e-4 per tensor (Xception.py:89-99,167-201 backward). */
int xcp_f32_dw3x3_fused(const float* x, const float* w9, const float* scale, const float* shift, int relu, float* out, int F,
                        int H, int W, int C, int device, void* stream);                           /* ~ xcp_dw3x3_fwd */
int xcp_f32_pool_add_fused(const float* y, const float* scale, const float* shift, const float* ys, const float* scale_s,
                           const float* shift_s, float* out, void* idx, float* ymax, int F, int H, int W, int C, int device,
                           void* stream);                                                        /* ~ xcp_pool_add_fwd */
int xcp_f32_bn_bwd_sums(const float* y, const float* G, float* sums, long long n_pix, int C, int device,
                        void* stream);                                                           /* ~ xcp_bn_bwd_sums */
int xcp_f32_bn_add(const float* y, const float* scale, const float* shift, const float* skip, const float* scale_s,
                   const float* shift_s, float* out, long long n, int C, int device, void* stream); /* ~ xcp_bn_add_fwd */
int xcp_f32_bn_relu_gap(const float* y, const float* scale, const float* shift, float* feat, int F, int HW, int C, int device,
                        void* stream);                                                           /* ~ xcp_bn_relu_gap */
/* ~ xcp_bn_bwd; sums_ws = fp32 [2][C] scratch (unused when presums given), coef = fp32 [3][C] */
int xcp_f32_bn_bwd(int mode, const float* y, const float* G, const void* idx, const float* dfeat, const float* scale,
                   const float* shift, const float* gamma, const float* mean, const float* rstd, int training,
                   const float* presums, float* sums_ws, float* coef, float* dgamma, float* dbeta, float* dy, int F, int H, int W,
                   int C, int c_real, int grid_w, int grid_h, int device, void* stream);
int xcp_f32_dw3x3_bwd(const float* dD, const float* xin, const float* w9, const float* scale, const float* shift, int relu,
                      float* dz, const float* add_full, const float* add_half, float* dw, float* bnsum, int F, int H, int W, int C,
                      int c_real, int device, void* stream);                                     /* ~ xcp_dw3x3_bwd */
int xcp_f32_gemm_wgrad(const float* dY, long long ld_dy, const float* X, long long ld_x, float* dW, long long ld_dw, long long R,
                       int P, int Q, int device, void* stream);                                  /* ~ xcp_gemm_wgrad */
/* stem conv backward (Xception.py:118,122): data gradient of the stride-1 conv2 (dy [F,H-2,W-2,Co] -> dx [F,H,W,Ci]) and the
 * weight gradient dw[Co][Ci][3][3] += ... of conv1 (x fp32 NCHW, stride 2) / conv2 (x NHWC, stride 1) */
int xcp_f32_conv3x3_dgrad(const float* dy, const float* w, float* dx, int F, int H, int W, int Ci, int Co, int device,
                          void* stream);
int xcp_f32_conv3x3_wgrad(const float* x, int x_nchw, const float* dy, float* dw, int F, int H, int W, int Ci, int Co, int stride,
                          int device, void* stream);
/* 3-way bf16 split of an fp32 matrix, 6-fold concatenated (side 0: h,h,m,h,m,l; side 1: h,m,h,l,m,h) along the columns
 * (along_rows = 0: out bf16 [rows][6*cols]) or the rows (1: out [6*rows][cols]): feeding both split operands to xcp_gemm_tn
 * (epi 2) / xcp_gemm_wgrad makes the production tcgen05 kernels compute an fp32-grade product (tests). */
int xcp_split3_bf16(const float* x, void* out, long long rows, int cols, int side, int along_rows, int device, void* stream);

/* ---- audio front-end: waveform -> MFCC on the device, replacing the offline
 * librosa.feature.mfcc(y, sr, n_mfcc=13, n_fft=int(0.025 sr), hop_length=int(0.010 sr)).T of wavfake_audio_dataset.py:17-19,40-44
 * (centre-padded periodic-Hann STFT power, Slaney mel filters, power_to_db(top_db), orthonormal DCT-II).
 * wav [B][L] fp32; melfb_t [n_fft/2+1][n_mels] fp32 (host-built constant); logmel_ws [B][T][n_mels] fp32 and gmax_ws [B] int
 * are scratch; out [B][T][n_mfcc] fp32 with T = xcp_mfcc_frames(L, hop) = 1 + L/hop.  pad_reflect: librosa < 0.10 padding. */
int xcp_mfcc_frames(int L, int hop);
int xcp_mfcc(const float* wav, int B, int L, const float* melfb_t, int n_fft, int hop, int n_mels, int n_mfcc, int pad_reflect,
             float amin, float top_db, float* logmel_ws, int* gmax_ws, float* out, int device, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XCP_H */
