/*
 * refign_b200.h -- C ABI of librefign_b200.so (sm_100a kernels for the Refign
 * per-training-step hot path).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer into memory owned by the caller
 *     (the Python host allocates torch tensors and passes .data_ptr());
 *   - tensors are dense, contiguous, row-major in the layout stated per call;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - calls are asynchronous with respect to the host and re-entrant per stream;
 *   - return value: 0 on success, negative RF_E* on failure; the message of the
 *     last failure on the calling thread is returned by rf_last_error();
 *   - no torch types, no global state besides per-process cached attributes.
 *
 * The reference has no C ABI of its own: its only native boundary is the
 * pybind11 module `correlation` (forward/backward) in
 *   /root/reference/models/correlation_ops/correlation_sampler.cpp:62-132
 * reached through
 *   /root/reference/models/correlation_ops/correlation_function.py:14-94
 * and tried-first as `spatial_correlation_sampler.spatial_correlation_sample`
 *   /root/reference/models/modules.py:252-262.
 * Each entry point below cites the reference interface it replaces.
 */
#ifndef REFIGN_B200_H_
#define REFIGN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RF_OK 0
#define RF_EINVAL (-1)   /* bad argument / unsupported shape            */
#define RF_ECUDA (-2)    /* CUDA runtime / launch error                  */
#define RF_ENODEV (-3)   /* no sm_100 device                             */

/* ---- library ---------------------------------------------------------- */
const char* rf_last_error(void);
int rf_version(void);                 /* ABI version, currently 1         */
int rf_device_check(void);            /* RF_OK iff current device is sm_100 */
/* Bind `device` to the calling host thread for this library (it links the CUDA
 * runtime statically, so its per-thread current device is its own).  Call once
 * per host thread that uses the library -- the Python binding does it. */
int rf_set_device(int device);

/* ---- local (windowed) correlation ------------------------------------- */
/* Replaces correlation.forward (correlation_sampler.cpp:62-90 ->
 * correlation_cuda_kernel.cu:241-278; CPU twin correlation.cpp:80-129).
 *   in1,in2 : f32 [B,C,H,W]      out : f32 [B,pH,pW,oH,oW]
 *   oH = (H + 2*padH - ((kH-1)*dilH + 1)) / sH + 1   (likewise oW)
 * Terms whose shifted coordinate falls outside the image contribute 0.
 * Every element of `out` is written (no pre-zeroing needed).
 * fuse_relu_l2norm != 0 additionally applies
 * LocalFeatureCorrelationLayer's  F.normalize(F.relu(.), dim=1) over the
 * pH*pW displacement channels (models/modules.py:271-273); requires kH=kW=1,
 * stride 1, pad 0, dilation_patch 1.  norm_out (f32 [B,oH,oW], may be NULL)
 * then receives max(||relu(c)||_2, 1e-12) per pixel for the backward. */
int rf_local_corr_fwd(const float* in1, const float* in2, float* out, float* norm_out,
                      int B, int C, int H, int W,
                      int kH, int kW, int pH, int pW, int padH, int padW,
                      int dilH, int dilW, int dpH, int dpW, int sH, int sW,
                      int fuse_relu_l2norm, void* stream);

/* Replaces correlation.backward (correlation_sampler.cpp:92-127 ->
 * correlation_cuda_kernel.cu:280-332; CPU twin correlation.cpp:131-183).
 *   grad_out : f32 [B,pH,pW,oH,oW]   grad_in1, grad_in2 : f32 [B,C,H,W]
 * Both gradients are fully written.  scratch: caller-provided f32 buffer of
 * rf_local_corr_bwd_scratch_bytes() bytes (may be NULL if that is 0). */
int64_t rf_local_corr_bwd_scratch_bytes(int B, int C, int H, int W, int kH, int kW,
                                        int pH, int pW, int padH, int padW, int dilH,
                                        int dilW, int dpH, int dpW, int sH, int sW);
int rf_local_corr_bwd(const float* in1, const float* in2, const float* grad_out,
                      float* grad_in1, float* grad_in2, void* scratch,
                      int B, int C, int H, int W,
                      int kH, int kW, int pH, int pW, int padH, int padW,
                      int dilH, int dilW, int dpH, int dpW, int sH, int sW,
                      void* stream);

/* Backward of the fused ReLU + L2-norm epilogue: given y = normalize(relu(c))
 * (the forward output, [B,K,HW]) and dL/dy, writes dL/dc into grad_c.
 * Positions where the norm was clamped (all-zero vectors) get zero gradient
 * scaled by 1/eps exactly as autograd of F.normalize does for the reference
 * (models/modules.py:273). norm: f32 [B,HW] as saved by the forward. */
int rf_relu_l2norm_bwd(const float* y, const float* norm, const float* grad_y,
                       float* grad_c, int B, int K, int64_t HW, void* stream);

/* ---- global correlation ----------------------------------------------- */
/* Replaces GlobalFeatureCorrelationLayer.forward (models/modules.py:294-308:
 * torch.bmm :362-374, mutual_matching :310-333, relu + normalize :307).
 *   src : f32 [B,C,Ns]   trg : f32 [B,C,Nt]   out : f32 [B,Ns,Nt]
 *   out[b,s,t] = <src[b,:,s], trg[b,:,t]>, source index row-major (h_s,w_s).
 *   mode bit0: mutual matching (eps 1e-5); bit1: relu + L2-norm over s.
 *   workspace: f32 [B*(Ns+2*Nt)] scratch for the row/column maxima and norms.
 *   use_tensor_cores: 1 = tcgen05 TF32 path (needs C%32==0, Ns,Nt%128==0),
 *                     0 = fp32 FFMA path, -1 = pick automatically. */
int64_t rf_global_corr_workspace_bytes(int B, int64_t Ns, int64_t Nt);
int rf_global_corr_fwd(const float* src, const float* trg, float* out, void* workspace,
                       int B, int C, int64_t Ns, int64_t Nt, int mode,
                       int use_tensor_cores, void* stream);

/* ---- bilinear warp ---------------------------------------------------- */
/* Replaces helpers.matching_utils.warp (matching_utils.py:11-49) for
 * padding_mode='zeros': sample x at (col + flow_x, row + flow_y) with
 * grid_sample(bilinear, zeros, align_corners=True) arithmetic in fp32, and the
 * strict-inside validity mask of :45-47.  No host synchronisation: the
 * reference's `torch.all(flo == 0)` early exit (:19-22) is replaced by the
 * device-side flag `all_zero_flag` (int32, may be NULL): when non-NULL the
 * kernel reads *all_zero_flag (1 = flow identically zero, as computed by
 * rf_flow_is_zero) and then copies x and sets the mask to all-true.
 *   x : f32 [B,C,H,W]  flow : f32 [B,2,H,W]  out : f32 [B,C,H,W]
 *   mask : u8 [B,H,W] or NULL */
int rf_flow_is_zero(const float* flow, int64_t n, int32_t* flag, void* stream);
int rf_warp_bilinear_fwd(const float* x, const float* flow, float* out, uint8_t* mask,
                         const int32_t* all_zero_flag, int B, int C, int H, int W,
                         void* stream);
/* Gradients of the above wrt x (scatter-add; grad_x must be zero-filled by the
 * caller) and wrt flow (may be NULL). */
int rf_warp_bilinear_bwd(const float* x, const float* flow, const float* grad_out,
                         float* grad_x, float* grad_flow, const int32_t* all_zero_flag,
                         int B, int C, int H, int W, void* stream);

/* ---- confidence + label refinement ------------------------------------ */
/* Replaces estimate_probability_of_confidence_interval_of_mixture_density
 * (matching_utils.py:52-57), R = 1:  cert = 1 - exp(-1 / (2 exp(logvar))). */
int rf_cert_fwd(const float* logvar, float* cert, int64_t n, void* stream);

/* Replaces DomainAdaptationSegmentationModel.refine + eta
 * (models/segmentation_model.py:438-491) and the torch.max of
 * get_dacs_mix (:551).
 *   logits_trg, logits_ref : f32 [B,K,H*W]  (K <= 32; Refign: 19)
 *   certs  : f32 [B,H*W] confidence P_R, or NULL
 *   logvar : f32 [B,H*W] log-variance (P_R computed in-kernel), or NULL;
 *            certs == logvar == NULL  =>  P = 0.5  (:472-473)
 *   warp_mask : u8 [B,H*W] or NULL
 *   ent_fix : i64 [B] scratch (2^-40 fixed-point entropy sums), trust : f32 [B] out
 *   probs_out : f32 [B,K,H*W];  label_out : i64 [B,H*W] or NULL;
 *   maxprob_out : f32 [B,H*W] or NULL
 *   static_mask : bit k set iff class k is in the static-large set S (:452)
 *   flags bit0 = disable_M, bit1 = disable_P
 * Integer outputs are bit-exact w.r.t. oracle/refign_oracle.c by construction
 * (same IEEE operation sequence, order-independent fixed-point reduction). */
int rf_refine_fwd(const float* logits_trg, const float* logits_ref, const float* certs,
                  const float* logvar, const uint8_t* warp_mask, int64_t* ent_fix,
                  float* trust, float* probs_out, int64_t* label_out, float* maxprob_out,
                  int B, int K, int64_t HW, float gamma, uint64_t static_mask, int flags,
                  void* stream);

/* ---- depthwise 3x3 convolution, channels-last --------------------------- */
/* Replaces the two depthwise convolutions of the train step, both stride 1,
 * padding = dilation, on NHWC (== token [B,H*W,C]) tensors:
 *   - Mix-FFN DWConv + nn.GELU (models/backbones/mix_transformer.py:96-103,
 *     556-568): dilation 1, bias, gelu = 1 (exact erf form);
 *   - the depthwise stage of DAFormer's separable ASPP branches
 *     (models/heads/daformer.py:26-35, models/modules.py:29-36): dilation
 *     6/12/18, bias NULL, gelu = 0.
 *   x, y : dtype 0 = f32, 1 = bf16, [B,H,W,C], C % 8 == 0, 16-byte aligned
 *   weight : f32 [C,1,3,3] (native parameter layout), bias : f32 [C] or NULL
 * Accumulation is fp32. */
int rf_dwconv3x3_nhwc_fwd(const void* x, const float* weight, const float* bias, void* y,
                          int B, int H, int W, int C, int dilation, int gelu, int dtype,
                          void* stream);
/* grad_x = conv with the flipped kernel of grad_y (plain conv, no activation). */
int rf_dwconv3x3_nhwc_bwd_input(const void* grad_y, const float* weight, void* grad_x,
                                int B, int H, int W, int C, int dilation, int dtype,
                                void* stream);
/* GELU backward pre-pass: grad_pre = grad_out * gelu'(conv(x) + bias); the
 * pre-activation is recomputed instead of being saved by the forward. */
int rf_dwconv3x3_gelu_bwd_pre(const void* x, const float* weight, const float* bias,
                              const void* grad_out, void* grad_pre, int B, int H, int W,
                              int C, int dilation, int dtype, void* stream);
/* grad_weight f32 [C,1,3,3] and grad_bias f32 [C] (or NULL), accumulated with fp32 atomics.
 * accumulate == 0: both are zeroed by the call first (autograd's "fresh gradient" contract);
 * accumulate != 0: the sums are added to what the buffers hold -- the runtime passes views of
 * its flat gradient buffer, which replaces the reference's per-parameter AccumulateGrad adds. */
int rf_dwconv3x3_nhwc_bwd_weight(const void* x, const void* grad_pre, float* grad_weight,
                                 float* grad_bias, int B, int H, int W, int C, int dilation,
                                 int dtype, int accumulate, void* stream);

/* ---- MiT spatial-reduction attention core (tcgen05 / TMEM / TMA) --------- */
/* Replaces the attention core of Attention.forward
 * (models/backbones/mix_transformer.py:150-160):
 *   out[b,n,h,:] = softmax_m(scale * <q[b,n,h,:], k[b,m,h,:]>) v[b,m,h,:], head_dim 64
 *   q   : bf16 [B,N,heads*64]        (output of the q projection, read in place)
 *   kv  : bf16 [B,M,2*heads*64]      (output of the kv projection: k = [0,C), v = [C,2C))
 *   out : bf16 [B,N,heads*64];  lse : f32 [B,heads,N] log-sum-exp of the scaled scores, or NULL
 * The [B,heads,N,M] attention matrix is never written to memory. */
int rf_sr_attention_fwd(const void* q, const void* kv, void* out, float* lse, int B, int N,
                        int M, int heads, float scale, void* stream);

/* Backward of the above (autograd of mix_transformer.py:150-160) with the probabilities recomputed
 * from q, k and the forward's lse:
 *   out, grad_out, grad_q : bf16 [B,N,heads*64];  lse : f32 [B,heads,N] (from the forward)
 *   grad_kv_f32 : f32 [B,M,2*heads*64]  (dK | dV; zeroed by the call, query splits are combined with
 *                 fp32 atomics; the caller rounds it to the kv dtype)
 *   workspace   : rf_sr_attention_bwd_workspace_bytes() bytes of scratch */
int64_t rf_sr_attention_bwd_workspace_bytes(int B, int N, int M, int heads);
int rf_sr_attention_bwd(const void* q, const void* kv, const void* out, const void* grad_out,
                        const float* lse, void* grad_q, float* grad_kv_f32, void* workspace,
                        int B, int N, int M, int heads, float scale, void* stream);

/* fp32 variants of the same two operators for the fp32 PARITY mode (precision='fp32'): exact FFMA tiles with expf /
 * logf, same fused formulation (no [B,heads,N,M] matrix; q / kv read in place; lse saved for the backward), so the
 * parity mode runs through the C-ABI too instead of library batched GEMMs + softmax
 * (models/backbones/mix_transformer.py:150-160).  All tensors f32; grad_kv is [B,M,2*heads*64] (dK | dV). */
int rf_sr_attention_f32_fwd(const float* q, const float* kv, float* out, float* lse, int B, int N, int M, int heads,
                            int head_dim /* 64 or 32 */, float scale, void* stream);
int64_t rf_sr_attention_f32_bwd_workspace_bytes(int B, int N, int heads);
int rf_sr_attention_f32_bwd(const float* q, const float* kv, const float* out, const float* grad_out, const float* lse,
                            float* grad_q, float* grad_kv, void* workspace, int B, int N, int M, int heads, int head_dim,
                            float scale, void* stream);

/* ---- stage-1 OverlapPatchEmbed: 7x7/s4 conv (3 -> 32|64 channels) + LayerNorm ---- */
/* Replaces OverlapPatchEmbed.forward for patch_embed1 (models/backbones/mix_transformer.py:236-242,
 * LayerNorm eps 1e-5 :234): conv + NCHW->NLC transpose + LayerNorm in one pass.
 *   x : f32 [B,3,H,W];  weight : f32 [cout,3,7,7];  bias, gamma, beta : f32 [cout]
 *   y : f32 [B, Ho*Wo, cout] tokens (Ho = (H-1)/4 + 1);  pre_norm : f32 same shape or NULL (the
 *   convolution output before the LayerNorm, needed by the backward);  mean, rstd : f32 [B*Ho*Wo] or NULL. */
int rf_patch_embed_ln_fwd(const float* x, const float* weight, const float* bias, const float* gamma,
                          const float* beta, float* pre_norm, float* y, float* mean, float* rstd,
                          int B, int H, int W, int cout, float eps, void* stream);

/* ---- residual add + LayerNorm ------------------------------------------- */
/* Replaces the LayerNorms of the MiT encoder and the residual adds in front of
 * them (models/backbones/mix_transformer.py:203-207 Block.forward, :135/:148 SR
 * norm, :234/:240 patch-embed norm, :378-426 stage norms; drop-path scaling
 * models/modules.py:587-596):
 *   xn = x + scale[row / rows_per_sample] * branch     (branch may be NULL: xn = x)
 *   y  = (xn - mean) * rstd * gamma + beta,  mean/rstd over the C channels, fp32
 * x: [rows,C] dtype x_dtype (0 = f32, 1 = bf16; f32 when branch != NULL);
 * branch: [rows,C] dtype branch_dtype or NULL; scale: f32 [rows/rows_per_sample] or NULL (= 1);
 * xn_out: f32 [rows,C] or NULL; y: [rows,C] dtype y_dtype; mean, rstd: f32 [rows] or NULL.
 * C % 32 == 0, C <= 512. */
int rf_add_layernorm_fwd(const void* x, const void* branch, const float* scale,
                         const float* gamma, const float* beta, float* xn_out, void* y,
                         float* mean, float* rstd, int64_t rows, int C,
                         int64_t rows_per_sample, float eps, int x_dtype, int branch_dtype,
                         int y_dtype, void* stream);
/* Backward: dxn = dxn_in (or 0 when NULL) + LN'(dy);  dbranch = scale * dxn (or NULL);
 * dgamma, dbeta: f32 [C], zeroed by the call (unless accumulate != 0, see
 * rf_dwconv3x3_nhwc_bwd_weight) then accumulated with fp32 atomics.
 * xn: the forward's xn (dtype xn_dtype), dy dtype dy_dtype, dbranch dtype branch_dtype. */
int rf_add_layernorm_bwd(const void* xn, const void* dy, const float* dxn_in,
                         const float* mean, const float* rstd, const float* gamma,
                         const float* scale, float* dxn, void* dbranch, float* dgamma,
                         float* dbeta, int64_t rows, int C, int64_t rows_per_sample,
                         int xn_dtype, int dy_dtype, int branch_dtype, int accumulate,
                         void* stream);

/* ---- optimiser-side multi-tensor ops on flat buffers ------------------- */
/* Replaces update_momentum_encoder (segmentation_model.py:680-689):
 *   ema = ema*m + live*(1-m)   over one flat f32 buffer.  momentum is a double
 * because the reference forms (1. - m) in Python double precision before the
 * tensor multiply rounds it to binary32; both factors are rounded here the same
 * way so the update is bit-identical to the reference's. */
int rf_ema_update(float* ema, const float* live, int64_t n, double momentum, void* stream);

/* AdamW step (torch.optim.AdamW semantics: decoupled weight decay, bias
 * correction) over flat f32 buffers split into `nseg` contiguous segments with
 * their own lr / weight_decay -- the four param groups of
 * segmentation_model.py:390-419.  seg_end[i] = exclusive end offset.
 * seg_end / seg_lr / seg_wd are small HOST arrays (nseg <= 8), copied into the
 * kernel arguments -- the one exception to the device-pointer convention.
 * grad_scale multiplies the gradient first (1/world_size for a summed
 * all-reduce). */
int rf_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                  int64_t n, int nseg, const int64_t* seg_end, const float* seg_lr,
                  const float* seg_wd, float beta1, float beta2, float eps, int step,
                  float grad_scale, void* stream);

/* ---- training-mode BatchNorm (+ ReLU), channels-last ----------------------- */
/* Replaces the BatchNorm2d / SyncBatchNorm + ReLU that follows every convolution of the DAFormer head
 * (models/modules.py:16-56 ConvModule, models/heads/daformer.py:26-35,102-108) on [rows = B*H*W, C]
 * channels-last activations (dtype 0 = f32, 1 = bf16; C % 8 == 0; statistics and parameters f32).
 *   rf_bn_stats      sums[0..C) = sum x, sums[C..2C) = sum x^2            (zeroed by the call)
 *   -- SyncBatchNorm: the caller all-reduces `sums` and passes the global sample count --
 *   rf_bn_finalize   mean, rstd, scale = gamma * rstd, shift = beta - mean * scale; running statistics
 *                    (momentum update with the unbiased variance) when running_mean/var != NULL
 *   rf_bn_apply      y = x * scale + shift, then ReLU when relu != 0
 *   rf_bn_bwd_reduce sums[0..C) = sum g, sums[C..2C) = sum g * xhat, g = grad_y * (ReLU mask recomputed
 *                    from x); these are also the bias / weight gradients         (zeroed by the call)
 *   -- SyncBatchNorm: all-reduce `sums` --
 *   rf_bn_bwd_apply  grad_x = scale * (g - sums[c]/count - xhat * sums[C+c]/count) */
int rf_bn_stats(const void* x, float* sums, int64_t rows, int C, int dtype, void* stream);
int rf_bn_finalize(const float* sums, const float* gamma, const float* beta, float* mean, float* rstd,
                   float* scale, float* shift, float* running_mean, float* running_var, int C,
                   double count, float eps, float momentum, void* stream);
int rf_bn_apply(const void* x, const float* scale, const float* shift, void* y, int64_t rows, int C,
                int relu, int dtype, void* stream);
int rf_bn_bwd_reduce(const void* x, const void* grad_y, const float* mean, const float* rstd,
                     const float* scale, const float* shift, float* sums, int64_t rows, int C,
                     int relu, int dtype, void* stream);
int rf_bn_bwd_apply(const void* x, const void* grad_y, const float* mean, const float* rstd,
                    const float* scale, const float* shift, const float* sums, void* grad_x,
                    int64_t rows, int C, double count, int relu, int dtype, void* stream);

/* ---- helpers around the library GEMMs of the Linear layers ----------------- */
/* Bias gradient of a Linear layer (autograd of nn.Linear in mix_transformer.py / modules.py:59-68):
 * out[c] = sum_r g[r,c];  g: [rows,cols] dtype 0 = f32 / 1 = bf16, cols % 8 == 0; out f32 [cols]
 * (zeroed by the call unless accumulate != 0). */
int rf_colsum(const void* g, float* out, int64_t rows, int cols, int dtype, int accumulate,
              void* stream);
/* Space-to-depth of a channels-last token grid for the MiT spatial-reduction conv
 * (mix_transformer.py:133-134,147-149: Conv2d(dim, dim, kernel_size=sr, stride=sr), which the host side
 * runs as one GEMM on the packed tokens):  packed[b,hs,ws,i,j,:] = img[b, hs*s+i, ws*s+j, :], rows of
 * row_bytes bytes (C * element size, a multiple of 16).  inverse != 0 scatters packed -> img (the input
 * gradient).  src and dst must not alias. */
int rf_space_to_depth(const void* src, void* dst, int B, int H, int W, int row_bytes, int s,
                      int inverse, void* stream);
/* y = act(y + bias[channel]) in place for a conv output [N,C,H,W] held either channels-last
 * (chan_inner = 1) or NCHW-contiguous (chan_inner = H*W): the bias add + activation that follow the frozen
 * library convolutions of the alignment network (VGG.forward, models/backbones/vgg.py:108-120; the BN-folded
 * ConvBNReLU blocks of models/modules.py:16-56).  act: 0 none, 1 ReLU, 2 LeakyReLU(slope).
 * dtype 0 = f32 (4-element vectors) / 1 = bf16 (8-element vectors); a scalar kernel serves channel runs
 * that are not a multiple of the vector width. */
int rf_bias_act(void* y, const float* bias, int64_t numel, int C, int64_t chan_inner, int act,
                float slope, int dtype, void* stream);

/* ---- DAFormer head feature fusion -------------------------------------------------------------- */
/* y[b, :, :, off_i : off_i + E_i] = bilinear(src_i, size = (H, W), align_corners = False) for the n <= 4
 * embedded stage features src_i [B, h_i, w_i, E_i] (bf16, channels-last = the token layout [B, h_i*w_i, E_i]),
 * written once as the concatenated channels-last tensor y [B, H, W, sum E_i] (bf16).  Restates the resize + cat
 * of DAFormerHead.forward (models/heads/daformer.py:203-221: F.interpolate(..., mode='bilinear',
 * align_corners=False) per stage, then torch.cat(dim=1)).  A source that already has the output size is copied.
 * src / h / w / E are HOST arrays of n entries; E_i % 8 == 0. */
int rf_upsample_concat_fwd(const void* const* src, const int* h, const int* w, const int* E, int n,
                           void* y, int B, int H, int W, void* stream);
/* Gradients of the sources from grad_y [B, H, W, sum E_i] (gather over the output pixels that read each source
 * pixel, no atomics); grad_src[i] may be NULL to skip a source. */
int rf_upsample_concat_bwd(const void* grad_y, void* const* grad_src, const int* h, const int* w,
                           const int* E, int n, int B, int H, int W, void* stream);
/* ---- loss tail of the student forwards ----------------------------------------------------------- */
/* loss_sum[0] = sum over the B*H*W label pixels of  w_i * (logsumexp_k z_ik - z_i,target_i),  z = the bilinear
 * up-sampling (align_corners = False) of the low-resolution logits [B, K, h, w] (f32, NCHW) to (H, W); pixels
 * whose label is ignore_index (or outside [0, K)) contribute 0; pixel_weight [B, H, W] may be NULL (= 1).
 * Restates F.interpolate + PixelWeightedCrossEntropyLoss of the train step (models/segmentation_model.py:160-170,
 * 228-240; models/losses.py:10-22) without materialising the [B, K, H, W] tensor; the caller divides by B*H*W
 * (the reference's mean runs over ALL pixels).  target is int64 [B, H, W]; 2 <= K <= 32, H >= h, W >= w. */
int rf_upsample_ce_fwd(const float* logits, const int64_t* target, const float* pixel_weight, float* loss_sum,
                       int B, int K, int h, int w, int H, int W, int ignore_index, void* stream);
/* grad_logits [B, K, h, w] = grad_loss[0] / (B*H*W) * d loss_sum / d logits (gather, no atomics; the
 * full-resolution softmax is recomputed).  grad_loss is a DEVICE scalar. */
int rf_upsample_ce_bwd(const float* logits, const int64_t* target, const float* pixel_weight,
                       const float* grad_loss, float* grad_logits, int B, int K, int h, int w, int H, int W,
                       int ignore_index, void* stream);
/* out [planes, H, W] = bilinear(in [planes, h, w]), align_corners = False, f32 (the teacher logits:
 * models/segmentation_model.py:206-208); W % 4 == 0, out 16-byte aligned. */
int rf_upsample_bilinear_f32(const float* in, float* out, int64_t planes, int h, int w, int H, int W, void* stream);
/* fp32 -> bf16 copy of a flat parameter buffer (the bf16 shadow weights read by the tensor-core
 * GEMMs; replaces the per-tensor autocast casts of the reference's AMP path). */
int rf_cast_bf16(const float* src, void* dst, int64_t n, void* stream);

/* CUDA-graph friendly variants: the step-dependent scalars are read from a DEVICE block
 *   hyper[0..7] = per-segment learning rate, hyper[8] = 1 - beta1^t, hyper[9] = sqrt(1 - beta2^t),
 *   hyper[10] = EMA momentum m, hyper[11] = 1 - m   (all f32, written by the host before the replay)
 * so that a captured train step can be replayed while the schedule advances. */
int rf_ema_update_dev(float* ema, const float* live, int64_t n, const float* hyper, void* stream);
int rf_adamw_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                      int64_t n, int nseg, const int64_t* seg_end, const float* seg_wd,
                      float beta1, float beta2, float eps, float grad_scale, const float* hyper,
                      void* stream);

/* ---- bf16 tensor-core GEMM with fused epilogues (Linear layers: forward, dgrad, wgrad) ---------------------- */
/* Replaces the library GEMMs behind nn.Linear in the MiT encoder / DAFormer head
 * (models/backbones/mix_transformer.py:96-103,137-164; models/modules.py:59-68):
 *     out[m,n] (+)= sum_k A(m,k) * B(n,k) (+ bias[n])
 *   a : bf16, stored [M,K] (a_mn_major = 0) or [K,M] (a_mn_major = 1);  b : bf16, stored [N,K] (0) or [K,N] (1)
 *   out : bf16 [M,N] (out_f32 = 0) or f32 [M,N] (out_f32 = 1); accumulate = 1 (f32 only, no bias): out += product, the
 *         contraction is split over CTAs and partial sums are added with red.global (weight gradients into the flat
 *         gradient buffer); bias : f32 [N] or NULL.  All row pitches must be multiples of 8 elements.
 *   forward y = x W^T + b : (a = x, b = W);   dgrad dx = dy W : (a = dy, b = W, b_mn_major = 1);
 *   wgrad dW += dy^T x : (a = dy, a_mn_major = 1, b = x, b_mn_major = 1, out_f32 = 1, accumulate = 1).
 *   colsum : f32 [M] or NULL (accumulating calls only): colsum[m] += sum_k A(m,k) -- the bias gradient db = sum_t dy[t,:]
 *            of the same Linear layer, reduced by the tensor core inside the weight-gradient GEMM (one extra N = 16 MMA
 *            per k-step against a tile of ones) instead of a separate column-sum launch; needs the weight-gradient
 *            layouts a_mn_major = b_mn_major = 1. */
int rf_gemm_bf16(const void* a, const void* b, const float* bias, void* out, int M, int N, int K, int a_mn_major,
                 int b_mn_major, int out_f32, int accumulate, float* colsum, void* stream);

/* ---- 3x3 convolution (stride 1, padding = dilation) as an implicit GEMM on the same tcgen05 kernel ------------- */
/* Replaces the library convolutions of the DAFormer bottleneck (models/heads/daformer.py:102-108, modules.py:16-56),
 * the VGG-16 alignment backbone (models/backbones/vgg.py:108-120) and the BN-folded flow decoders / dilated
 * RefinementModule of the alignment head (models/modules.py:395-477; dilation 1..16) on channels-last bf16 tensors:
 *   x   : bf16 [B,H,W,Cin]   w : bf16 [Cout,3,3,Cin] (the channels-last filter)   bias : f32 [Cout] or NULL
 *   out : bf16 or f32 [B,H,W,Cout];  act: 0 none, 1 ReLU, 2 LeakyReLU(slope) fused in the epilogue
 * The input gradient is the same call on dy with the flipped, transposed filter [Cin,3,3,Cout].
 * rf_conv3x3_wgrad_bf16: dw f32 [Cout,3,3,Cin] += sum over pixels dy[..,co] * x[.. shifted ..,ci] (split over the
 * pixels, partial sums reduce-added).  Cin, Cout multiples of 8. */
int rf_conv3x3_bf16(const void* x, const void* w, const float* bias, void* out, int B, int H, int W, int Cin, int Cout,
                    int out_f32, int act, float slope, int dilation, void* stream);
int rf_conv3x3_wgrad_bf16(const void* dy, const void* x, float* dw, int B, int H, int W, int Cin, int Cout, void* stream);

/* ---- UncertaintyModule patch CNN (models/modules.py:534-561: the four "valid" 3x3 convolutions over the B*H*W
 * single-channel s x s patches of the correlation volume, eval-mode BatchNorm folded, LeakyReLU(slope), 2x2 max-pool after
 * conv_0 when s = 16) fused into one kernel, every intermediate in shared memory:
 *   corr   : f32 [B, s*s, H, W] (the displacement planes as produced by the correlation layers), s = search_size in {9, 16}
 *   params : rf_uncertainty_cnn_param_bytes() bytes, 16-byte aligned: f32 w0[9][32] b0[32] b1[32] b2[16] w3[6][9][16] b3[8],
 *            bf16 w1[32][296] (k = tap*32 + cin, 288 used), bf16 w2[16][296]  (built by UncertaintyModule._fused_params)
 *   out    : bf16 [B, H, W, 6]  (= predict_uncertainty's output, channels-last) */
int rf_uncertainty_cnn_param_bytes(void);
int rf_uncertainty_cnn_fwd(const float* corr, const void* params, void* out_bf16, int B, int H, int W, int search_size,
                           float slope, void* stream);

/* nn.MaxPool2d(2, 2) of VGG.forward (models/backbones/vgg.py:108-120) on a channels-last bf16 tensor:
 * x [B,H,W,C] -> y [B,H/2,W/2,C] (floor); C a multiple of 8. */
int rf_maxpool2x2_nhwc_bf16(const void* x, void* y, int B, int H, int W, int C, void* stream);

/* ---- DACS strong transform (class mix + colour jitter + gaussian blur) ---------------------------------- */
/* Replaces get_dacs_mix's per-image loop over helpers/dacs_transforms.py strong_transform
 * (models/segmentation_model.py:552-570; dacs_transforms.py:14-112; kornia 0.5.8 ColorJitter / GaussianBlur2d).
 *   rf_dacs_count : count_u64[0] = #{prob >= threshold} over n pixels (zeroed by the call)
 *   rf_dacs_mix   : mix_mask u8 [B,H,W] (1 = source pixel); img_* f32 [B,3,H,W]; gt_src / pseudo_label i64 [B,H,W];
 *                   out_weight = mask ? 1 : count / (B*H*W) (0 in the first ignore_top / last ignore_bottom rows);
 *                   params f32 [B,64] on the device: [0] jitter on, [1..4] order of (brightness, contrast, saturation,
 *                   hue), [5] brightness shift, [6] contrast factor, [7] saturation factor, [8] hue shift (radians),
 *                   [9] blur on, [10] / [11] tap radius along y / x (<= 16), [12..28] / [29..45] half kernels
 *   rf_dacs_blur  : separable gaussian with reflect border on img f32 [B,3,H,W], in place (tmp = same-size scratch);
 *                   images whose blur flag is 0 are left untouched */
int rf_dacs_count(const float* prob, int64_t n, float threshold, void* count_u64, void* stream);
int rf_dacs_mix(const float* img_src, const float* img_trg, const int64_t* gt_src, const int64_t* pseudo_label,
                const void* count_u64, const uint8_t* mix_mask, const float* params, float* out_img, int64_t* out_label,
                float* out_weight, int B, int H, int W, int ignore_top, int ignore_bottom, void* stream);
int rf_dacs_blur(float* img, float* tmp, const float* params, int B, int H, int W, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* REFIGN_B200_H_ */
