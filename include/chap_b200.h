/*
 * chap_b200.h -- C ABI of libchap_b200.so (sm_100a only).
 *
 * The reference (gardnerzhou/CHAP) has no FFI layer: its device work is ATen/cuDNN calls made
 * from nn.Module.forward and from the training script.  Each entry point below names the
 * reference interface (file:line under /root/reference) whose device work it replaces; the
 * Python host side (chap_b200/ops.py) binds them with ctypes, see INTEGRATION.md.
 *
 * Conventions
 *  - Plain pointers and sizes only.  Every device buffer (workspaces included) is owned by the
 *    caller; the library never allocates or frees device memory and keeps no pointer after
 *    return.  Descriptors are PODs copied on entry.
 *  - Activations are fp32, channels-last: [N, (D,) H, W, C] (a torch tensor of logical shape
 *    [N, C, (D,) H, W] in torch.channels_last / channels_last_3d memory format).  For 2D, D = 1.
 *  - Convolution weights / gradients use the torch parameter layout
 *    ([Cout, Cin, k..] for conv, [Cin, Cout, k..] for transposed conv); packed copies are opaque.
 *  - Every call is asynchronous on `stream` (a cudaStream_t passed as void*), performs no host
 *    synchronisation and is CUDA-graph capturable.
 *  - Return 0 on success, a negative chap_status on failure; chap_last_error() returns a
 *    thread-local message.  No exceptions cross the ABI.  There is no CPU fallback.
 */
#ifndef CHAP_B200_H
#define CHAP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CHAP_ABI_VERSION 1

typedef enum {
    CHAP_OK = 0,
    CHAP_ERR_BAD_ARG = -1,      /* null pointer, bad shape, unsupported combination */
    CHAP_ERR_ALIGNMENT = -2,    /* pointer / channel count not aligned as required */
    CHAP_ERR_WORKSPACE = -3,    /* workspace too small */
    CHAP_ERR_CUDA = -4,         /* a CUDA runtime / driver call failed */
    CHAP_ERR_ARCH = -5          /* device is not sm_100 */
} chap_status;

const char* chap_last_error(void);
int chap_abi_version(void);
/* 0 when the current device is a compute-capability 10.x part, CHAP_ERR_ARCH otherwise. */
int chap_check_device(void);
/* Number of kernels this library has launched in this process (bench.py `gpu_launches`). */
uint64_t chap_launch_count(void);
void chap_reset_launch_count(void);
/* 1: route every convolution through the CUDA-core kernels (debug aid), 0: tcgen05 where supported. */
/* Live timing of kernel families with CUDA events on the launching stream (bench.py roofline).
 * enable(1) resets the record; report() synchronises the device and writes lines
 * "name launches total_ms algorithmic_flops algorithmic_bytes" into buf. */
void chap_timing_enable(int on);
int chap_timing_report(char* buf, size_t cap);
void chap_set_force_simt(int flag);
int chap_get_force_simt(void);
/* Programmatic dependent launch between the kernels of this library.  Only effective in the experiment build (`make pdl`,
 * -DCHAP_PDL_INSN): measured without gain inside a replayed CUDA graph, so the default build compiles it out and these two calls only
 * record the flag.  There, every kernel is
 * launched with cudaLaunchAttributeProgrammaticStreamSerialization and waits (griddepcontrol.wait) before its first global-memory
 * access, so results are identical; only the launch latency between dependent kernels overlaps.  No counterpart in the reference
 * (its kernels are launched by ATen / cuDNN in plain stream order, e.g. every nn.Module call of code/networks/unet.py:49-57). */
void chap_set_pdl(int flag);
int chap_get_pdl(void);
/* Arithmetic of the tensor-core convolutions (forward and data gradient).
 *   0 (default): plain TF32 -- operands rounded to nearest by the TMA unit, fp32 accumulation in TMEM; per layer this is
 *       exactly cuDNN's TF32 arithmetic (reference default: torch.backends.cudnn.allow_tf32, code/train_ours_2D.py:542-547).
 *   c > 0: layers with max(Cin, Cout) <= c run the split-operand "3xTF32" mode: x = x_hi + x_lo and w = w_hi + w_lo with
 *       TF32 halves, x_hi*w_hi + x_lo*w_hi + x_hi*w_lo accumulated in fp32 (relative error ~2^-21 per product instead of
 *       2^-11).  c >= 1024 covers every layer ("precise mode", logits within 1e-3 of an fp32 evaluation); c = 32 covers
 *       the thin, HBM-bound layers that cuDNN itself runs in fp32 even when TF32 is allowed.
 * The weight gradient always uses plain TF32.  Env CHAP_PRECISE_MAX_C sets the initial value. */
void chap_set_conv_precision(int max_channels);
int chap_get_conv_precision(void);

/* ------------------------------------------------------------------ convolutions
 * Replaces nn.Conv2d/Conv3d/ConvTranspose2d/ConvTranspose3d forward + backward inside
 *   unet.ConvBlock code/networks/unet.py:44-60, unet.UpBlock :78-99, unet.Decoder.out_conv :168,
 *   vnet.ConvBlock code/networks/vnet.py:8-34, DownsamplingConvBlock :70-94,
 *   Upsampling_function :97-125, vnet.Decoder.out_conv :189.
 */
typedef enum {
    CHAP_CONV_K3 = 0,     /* k=3, stride 1, pad 1 */
    CHAP_CONV_K1 = 1,     /* k=1, stride 1, pad 0 */
    CHAP_CONV_DOWN2 = 2,  /* k=2, stride 2, pad 0 (out spatial = in/2) */
    CHAP_CONV_UP2 = 3     /* transposed, k=2, stride 2 (out spatial = 2*in) */
} chap_conv_kind;

typedef struct {
    int32_t kind;       /* chap_conv_kind */
    int32_t nd;         /* 2 or 3 spatial dims */
    int32_t n;          /* batch */
    int32_t in_d, in_h, in_w;   /* INPUT spatial size (in_d = 1 for nd == 2) */
    int32_t cin, cout;
} chap_conv_desc;

/* Per-channel statistics produced by a convolution epilogue are spread over CHAP_STAT_SLOTS partial
 * buffers (fewer colliding atomics); layout double[CHAP_STAT_SLOTS][2*cout], summed by chap_bn_finalize. */
#define CHAP_STAT_SLOTS 16

/* number of floats in one packed weight buffer (same for the fwd and the dgrad packing): taps * max(cin, 16) * max(cout, 16)
 * -- channel counts below 16 are zero-padded for the tensor-core operands of the thin heads */
size_t chap_conv_packed_elems(const chap_conv_desc* d);
/* torch-layout weight -> packed forward operand and packed data-gradient operand (either may be NULL) */
int chap_conv_pack_weights(const chap_conv_desc* d, const float* w, float* w_fwd, float* w_dgrad, void* stream);
/* The same for many layers in one or two launches (a trainer re-packs every conv weight once per optimiser step):
 * items[i] = {w, w_fwd, w_dgrad (either output nullable), desc}; the item table is read on the host during the call. */
typedef struct {
    const float* w;
    float* w_fwd;
    float* w_dgrad;
    chap_conv_desc desc;
} chap_pack_item;
int chap_conv_pack_weights_batched(const chap_pack_item* items, int32_t n, void* stream);
/* y = conv(x) + bias.  ch_sums (nullable): double[CHAP_STAT_SLOTS][2*cout], receives partial per-channel
 * sum(y), sum(y*y) (zeroed by the call) -- the BatchNorm batch statistics of the following layer. */
int chap_conv_fwd(const chap_conv_desc* d, const float* x, const float* w_fwd, const float* bias,
                  float* y, double* ch_sums, void* stream);
/* Convolution followed by a train-mode BatchNorm (reference: nn.ConvNd -> nn.BatchNormNd, code/networks/unet.py:50-51,
 * vnet.py:19-21): y = conv(x) + bias, batch statistics in the epilogue, and the BatchNorm "finalize" step of chap_bn_finalize
 * (mean / invstd / scale / shift, running statistics update) done by the LAST thread block of the same kernel -- no separate
 * launch between the convolution and chap_bn_act_fwd.  ch_sums must hold CHAP_STAT_SLOTS*2*cout + 1 doubles (the extra one is
 * the block ticket counter).  Shapes the tensor-core kernel does not take run conv + statistics + finalize as three launches. */
typedef struct chap_bn_train_args {
    const float* gamma;
    const float* beta;
    float eps, momentum;
    float* running_mean;             /* nullable trio: running statistics are updated when given */
    float* running_var;
    int64_t* num_batches_tracked;
    float* mean_invstd;              /* out: float[2*cout] */
    float* scale_shift;              /* out: float[2*cout] */
    int32_t stats_persistent;        /* 1: ch_sums is a persistent per-layer buffer that is ALL ZERO on entry; the call leaves it all
                                      * zero again (the finalizing block clears it) and launches no zero-fill.  0: zeroed by the call. */
    int32_t reserved_;
} chap_bn_train_args;
int chap_conv_bn_fwd(const chap_conv_desc* d, const float* x, const float* w_fwd, const float* bias, float* y,
                     double* ch_sums, const chap_bn_train_args* bn, void* stream);
/* Inference form of conv -> BatchNorm(eval) -> (Leaky)ReLU (+ additive skip), the layer pattern of code/networks/vnet.py:19-28,
 * 76-83,103-112 (and unet.py:49-57) under net.eval(): y = act(scale[c] * (conv(x) + bias[c]) + shift[c]) + residual, with
 * scale_shift = float[2*cout] from chap_bn_eval_params and residual (nullable) a tensor of y's shape.  On the tensor-core path
 * the affine map, the activation and the skip add run in the convolution epilogue -- no separate BatchNorm pass over y (the
 * sliding-window inference of code/test_3D_util.py:61-64 spends a quarter of its time there otherwise). */
int chap_conv_bn_act_fwd(const chap_conv_desc* d, const float* x, const float* w_fwd, const float* bias,
                         const float* scale_shift, float slope, const float* residual, float* y, void* stream);
/* dx = conv^T(dy) (data gradient), dx has the input shape */
int chap_conv_dgrad(const chap_conv_desc* d, const float* dy, const float* w_dgrad, float* dx, void* stream);
/* Data gradient of a convolution whose input was torch.cat([a, b], dim=1) (the U-Net skip connection, reference
 * code/networks/unet.py:97-98): channels [0, ca) of dx go to dx_a (row stride ca), [ca, cin) to dx_b (row stride cin - ca),
 * written by the tensor-core epilogue -- no separate split pass.  chap_conv_dgrad_split_supported() returns 1 when the
 * shape takes that path (ca and cin - ca multiples of 16, tensor-core data gradient); otherwise the call fails with
 * CHAP_ERR_BAD_ARG and the caller uses chap_conv_dgrad + chap_split_channels. */
int chap_conv_dgrad_split_supported(const chap_conv_desc* d, int32_t ca);
int chap_conv_dgrad_split(const chap_conv_desc* d, const float* dy, const float* w_dgrad, float* dx_a, int32_t ca,
                          float* dx_b, void* stream);
/* dw (torch layout, overwritten) and dbias (nullable) from x and dy.
 * workspace: caller-owned scratch of chap_conv_wgrad_workspace_bytes(d) bytes: 2*cout doubles for the bias gradient plus
 * taps*cin*cout floats in which the tensor-core kernel accumulates with 128-bit reductions before a small kernel writes the
 * torch layout.  NULL / a smaller buffer is accepted when dbias is NULL (the kernel then reduces with scalar atomics). */
size_t chap_conv_wgrad_workspace_bytes(const chap_conv_desc* d);
int chap_conv_wgrad(const chap_conv_desc* d, const float* x, const float* dy, float* dw, float* dbias,
                    void* workspace, size_t workspace_bytes, void* stream);
/* Gradient-sink form (the trainer's flat gradient arena, code/train_ours_2D.py:381-383 `zero_grad(); backward(); step()`):
 * dw_acc += dW and dbias_acc += db (dbias_acc nullable) instead of overwriting, so a parameter that is used by several network
 * passes of one iteration needs no autograd accumulation kernels and no gather copy.  zeroed_scratch (nullable): taps*cin*cout
 * floats that are ALL ZERO on entry and are left all zero on return (a persistent per-layer scratch: the tensor-core kernel
 * reduces into it with 128-bit atomics and a small kernel adds it into dw_acc in the torch layout).  workspace:
 * chap_conv_wgrad_workspace_bytes(d) bytes (bias sums; scratch of the pair-packed 16 -> 16 path); >= 2*cout doubles is the minimum,
 * a smaller-than-full workspace only disables the pair-packed path. */
int chap_conv_wgrad_acc(const chap_conv_desc* d, const float* x, const float* dy, float* dw_acc, float* dbias_acc,
                        void* workspace, size_t workspace_bytes, float* zeroed_scratch, void* stream);

/* ------------------------------------------------------------------ BatchNorm + activation (+dropout, +skip add)
 * Replaces BatchNorm2d/3d (train and eval) + LeakyReLU(0.01)/ReLU + Dropout + the additive skip of
 * vnet.Decoder (code/networks/unet.py:51-56, code/networks/vnet.py:21-28,202-215).
 */
/* per-channel sum / sum of squares of y[rows, c] into double[2c] (zeroed by the call) */
int chap_channel_stats(const float* y, int64_t rows, int32_t c, double* sums, void* stream);
/* train mode: batch mean / biased var from sums[slots][2c] (slots summed) -> mean_invstd[2c], scale_shift[2c]
 * (scale = gamma*invstd, shift = beta - mean*scale); if running_mean != NULL update
 * running = (1-m)*running + m*stat (unbiased var) and ++(*num_batches_tracked). */
int chap_bn_finalize(const double* sums, int32_t slots, int64_t count, const float* gamma, const float* beta,
                     float eps, float momentum, float* running_mean, float* running_var,
                     int64_t* num_batches_tracked, float* mean_invstd, float* scale_shift,
                     int32_t c, void* stream);
/* eval mode: scale/shift (and mean_invstd) from the running statistics */
int chap_bn_eval_params(const float* gamma, const float* beta, const float* running_mean,
                        const float* running_var, float eps, float* mean_invstd, float* scale_shift,
                        int32_t c, void* stream);
/* out = act(y*scale+shift) * drop + residual;   act(z) = z>0 ? z : slope*z
 * drop_nc (nullable): float[n*c] per-(sample,channel) factor (Dropout2d/3d);
 * drop_el (nullable): float[n*rows_per_sample*c] elementwise factor (nn.Dropout); residual nullable. */
int chap_bn_act_fwd(const float* y, const float* scale_shift, float slope, const float* drop_nc,
                    const float* drop_el, const float* residual, int32_t n, int64_t rows_per_sample,
                    int32_t c, float* out, void* stream);
/* backward of the above.  Pass 1 accumulates sums[2c] = (sum dz, sum dz*xhat) (doubles, zeroed by the
 * call); pass 2 writes dy and (train) dgamma, dbeta.  train = 0: BN constants are fixed (eval). */
int chap_bn_act_bwd(const float* dout, const float* y, const float* scale_shift, const float* mean_invstd,
                    const float* gamma, float slope, const float* drop_nc, const float* drop_el,
                    int32_t n, int64_t rows_per_sample, int32_t c, int32_t train,
                    double* sums, float* dy, float* dgamma, float* dbeta, void* stream);
/* same, with the BatchNorm parameter gradients ADDED into dgamma_acc / dbeta_acc (gradient-sink form, see chap_conv_wgrad_acc) */
/* sums_persistent = 1: `sums` holds 2c + 1 doubles, is ALL ZERO on entry and is left all zero (cleared by the last block of the
 * apply kernel): no zero-fill launch per call.  0: `sums` (2c doubles) is zeroed by the call. */
/* dgamma_acc / dbeta_acc may both be NULL (data gradient only, e.g. the feature-gradient probe of VAT2d) */
int chap_bn_act_bwd_acc(const float* dout, const float* y, const float* scale_shift, const float* mean_invstd, float slope,
                        const float* drop_nc, const float* drop_el, int32_t n, int64_t rows_per_sample, int32_t c, int32_t train,
                        double* sums, int32_t sums_persistent, float* dy, float* dgamma_acc, float* dbeta_acc, void* stream);
/* nn.Dropout(p) of ConvBlock (code/networks/unet.py:53, after the first LeakyReLU) WITHOUT a mask tensor: the 1/(1-p)-scaled keep
 * mask is generated inside the kernels (Philox4x32-10, counter = (vector index, subsequence), key = (seed, *epoch_dev)) and
 * regenerated identically by the backward given the same description.  epoch_dev (nullable) is an int64 in DEVICE memory read at
 * run time -- a trainer's iteration counter -- so that a replayed CUDA graph draws new masks every iteration; subsequence separates
 * the layers / passes inside one iteration.  Replaces `empty_like().bernoulli_(1-p).div_(1-p)` (two launches and three passes over
 * an activation-sized tensor per dropout) plus the mask reads of the forward and both backward passes.  Needs c % 4 == 0 and
 * 256 % (c / 4) == 0 (every layer of the two networks).  The draws are NOT torch's Bernoulli stream (masks are random either way;
 * parity tests pass explicit masks through drop_el). */
typedef struct {
    float p;                  /* drop probability, 0 < p < 1 */
    uint64_t seed;
    uint64_t subsequence;
    const int64_t* epoch_dev; /* nullable */
} chap_dropout_rng;
int chap_bn_act_fwd_rng(const float* y, const float* scale_shift, float slope, const float* drop_nc, const chap_dropout_rng* rng,
                        const float* residual, int32_t n, int64_t rows_per_sample, int32_t c, float* out, void* stream);
/* backward; accumulate = 1 adds dgamma / dbeta into their targets (gradient-sink form); dgamma / dbeta may both be NULL;
 * sums / sums_persistent as in chap_bn_act_bwd_acc */
int chap_bn_act_bwd_rng(const float* dout, const float* y, const float* scale_shift, const float* mean_invstd, float slope,
                        const float* drop_nc, const chap_dropout_rng* rng, int32_t n, int64_t rows_per_sample, int32_t c, int32_t train,
                        double* sums, int32_t sums_persistent, float* dy, float* dgamma, float* dbeta, int32_t accumulate, void* stream);

/* ------------------------------------------------------------------ pooling / upsampling / concat
 * MaxPool2d(2) code/networks/unet.py:69; Upsample(x2, bilinear|trilinear, align_corners=True)
 * unet.py:87-88, vnet.py:105; torch.cat([skip, up], 1) unet.py:98.
 */
int chap_maxpool2_fwd(const float* x, int32_t n, int32_t h, int32_t w, int32_t c, float* y, void* stream);
int chap_maxpool2_bwd(const float* x, const float* dy, int32_t n, int32_t h, int32_t w, int32_t c, float* dx, void* stream);
/* nd = 2: d must be 1.  out spatial = 2 * in spatial. */
int chap_upsample2x_fwd(const float* x, int32_t nd, int32_t n, int32_t d, int32_t h, int32_t w, int32_t c, float* y, void* stream);
int chap_upsample2x_bwd(const float* dy, int32_t nd, int32_t n, int32_t d, int32_t h, int32_t w, int32_t c, float* dx, void* stream);
/* out[r, 0:ca] = a[r, :], out[r, ca:ca+cb] = b[r, :] */
int chap_concat_channels(const float* a, const float* b, int64_t rows, int32_t ca, int32_t cb, float* out, void* stream);
/* inverse: a (nullable), b (nullable) <- slices of in */
int chap_split_channels(const float* in, int64_t rows, int32_t ca, int32_t cb, float* a, float* b, void* stream);
/* out[n, r, c] = x[n, r, c] * scale_nc[n, c]  (Dropout2d of FilterDropout.perform_dropout,
 * code/networks/FilterDropout.py:45-89, and unet.UNet.perform_dropout unet.py:532-552) */
int chap_channel_scale(const float* x, const float* scale_nc, int32_t n, int64_t rows_per_sample, int32_t c, float* out, void* stream);
/* out = a + alpha*b (flat) */
/* perform_dropout of code/networks/FilterDropout.py:45-89 for one pyramid level, fused: feat [N, rps, C] -> the two decoder
 * inputs out{1,2} [N + nu, rps, C] = cat(feat, feat[N - nu:] * m{1,2}) with per-(sample, channel) factors m{1,2} [nu, C]
 * (NULL = 1: level not perturbed).  One pass: feat read once, both outputs written.  bwd: dfeat from the two output gradients
 * (either may be NULL). */
int chap_feature_dropout_fwd(const float* feat, const float* m1, const float* m2, int32_t n, int32_t nu, int64_t rows_per_sample,
                             int32_t c, float* out1, float* out2, void* stream);
int chap_feature_dropout_bwd(const float* d1, const float* d2, const float* m1, const float* m2, int32_t n, int32_t nu,
                             int64_t rows_per_sample, int32_t c, float* dfeat, void* stream);
int chap_axpy(const float* a, const float* b, float alpha, int64_t elems, float* out, void* stream);
/* out = a*m + b*(1-m), m int64/broadcast spatial mask [rows_per_sample] (copy-paste mixing,
 * code/train_ours_2D.py:335-336) */
int chap_mask_mix(const float* a, const float* b, const int64_t* mask, int32_t n, int64_t rows_per_sample,
                  int32_t c, float* out, void* stream);

/* ------------------------------------------------------------------ losses
 * softmax / argmax / cross-entropy / masked Dice of code/train_ours_2D.py:198-216,319-325 and the
 * 'kl' | 'dice' consistency distance of losses.VAT2d (call site :372).  logits are channels-last
 * [n, rows_per_sample, c], c <= 8.
 */
/* soft1/soft2 (nullable) = softmax, arg1/arg2 (nullable) = argmax (int64, first max wins),
 * knowledge (nullable) = CE(pre1, arg2) + CE(pre2, arg1) per position */
int chap_pseudo_label(const float* pre1, const float* pre2, int64_t rows, int32_t c,
                      float* soft1, float* soft2, int64_t* arg1, int64_t* arg2, float* knowledge, void* stream);
/* softmax over c for each row (inference, code/test_3D_util.py:64, code/val_2D.py:68-80) */
int chap_softmax(const float* logits, int64_t rows, int32_t c, float* out, void* stream);
/* argmax over c of (a + b)/2 (b nullable -> argmax of a); first max wins (code/val_2D.py:72-83) */
int chap_argmax(const float* a, const float* b, int64_t rows, int32_t c, int64_t* out, void* stream);

/* label dtype codes */
#define CHAP_LABEL_I64 0
#define CHAP_LABEL_F32 1
/* Masked Dice + CE partial sums against hard labels.
 * mask: int64 [rows_per_sample] (broadcast over n), used as m (invert=0) or 1-m (invert=1).
 * sums (double[3c+2], zeroed by the call): inter[c], sum s^2 m [c], sum t m [c], sum CE m, sum m */
int chap_dice_ce_fwd(const float* logits, const void* labels, int32_t label_dtype, const int64_t* mask,
                     int32_t invert, int32_t n, int64_t rows_per_sample, int32_t c, double* sums, void* stream);
/* dlogits (+)= d/dlogits of  sum_c [coef_i[c]*inter_c + coef_s[c]*(sum s^2 m)_c] + coef_ce * sum(CE m);
 * coef: float[2c+1] on device = (coef_i[c], coef_s[c], coef_ce); accumulate != 0 adds into dlogits */
int chap_dice_ce_bwd(const float* logits, const void* labels, int32_t label_dtype, const int64_t* mask,
                     int32_t invert, int32_t n, int64_t rows_per_sample, int32_t c, const float* coef,
                     int32_t accumulate, float* dlogits, void* stream);
/* Scalar tail of mix_loss (code/train_ours_2D.py:198-216) from the two chap_dice_ce_fwd results (image pass with mask m,
 * patch pass with 1 - m): out3 = (loss_image, loss_patch, total) with loss_x = w_x (dice_x + ce_x) / 2, total = their sum.
 * chap_mix_loss_coef turns the upstream gradient of out3 (float[3], device) into the two coef vectors of chap_dice_ce_bwd. */
int chap_mix_loss_finalize(const double* sums_img, const double* sums_patch, int32_t c, float w_img, float w_patch,
                           float* out3, void* stream);
int chap_mix_loss_coef(const double* sums_img, const double* sums_patch, int32_t c, float w_img, float w_patch,
                       const float* grad_out3, float* coef_img, float* coef_patch, void* stream);

#define CHAP_DIST_KL 0
#define CHAP_DIST_DICE 1
/* consistency distance between softmax(logits) and a soft target; mask: float [n*rows_per_sample] or NULL.
 * sums (double[3c+1], zeroed by the call): KL: sums[0] = sum m * KL(t || p);
 * DICE: inter[c], sum p^2 m[c], sum t^2 m[c]. */
int chap_consistency_fwd(const float* logits, const float* target, const float* mask, int32_t dist,
                         int64_t rows, int32_t c, double* sums, void* stream);
/* KL: dlogits = coef[0] * m * (p*sum(t) - t);  DICE: through the softmax Jacobian with
 * coef = (coef_i[c], coef_p[c]) multiplying inter_c and (sum p^2 m)_c */
int chap_consistency_bwd(const float* logits, const float* target, const float* mask, int32_t dist,
                         int64_t rows, int32_t c, const float* coef, float* dlogits, void* stream);
/* patch scores of patch.create_maskV1 (call site code/train_ours_2D.py:371):
 * score[n, patch] = mean(knowledge) + mean(arg1 != arg2) over s^nd patches */
int chap_patch_score(const float* knowledge, const int64_t* arg1, const int64_t* arg2, int32_t nd, int32_t n,
                     int32_t d, int32_t h, int32_t w, int32_t s, float* score, void* stream);
/* mask[n, voxel] = score[n, patch(voxel)] >= kth[n] */
int chap_patch_mask(const float* score, const float* kth, int32_t nd, int32_t n, int32_t d, int32_t h,
                    int32_t w, int32_t s, float* mask, void* stream);

/* Largest-connected-component filter of get_ACDC_2DLargestCC (code/train_ours_2D.py:123-144) on the device:
 * seg int64 [n, d, h, w] class map -> out float32 [n, d, h, w]: per sample and foreground class (1..n_classes-1) the
 * largest component (8- / 26-connectivity) keeps its class value, everything else is 0; among equally large components
 * the one that comes first in scan order wins (np.argmax over np.bincount of skimage labels).  nd = 2: d must be 1. */
size_t chap_largest_cc_workspace_bytes(int32_t n, int32_t d, int32_t h, int32_t w, int32_t n_classes);
int chap_largest_cc(const int64_t* seg, int32_t nd, int32_t n, int32_t d, int32_t h, int32_t w, int32_t n_classes,
                    float* out, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ perturbation generator
 * The channel-spatial hierarchical adversarial perturbation of losses.VAT2d (absent in the
 * reference; ctor code/train_ours_2D.py:290, call :372; BASELINE.json north_star).
 */
typedef struct {
    const float* g;     /* gradient w.r.t. the level's perturbation direction [n, rows, c] */
    const float* f;     /* encoder feature of the level (nullable: out = r only) */
    float* out;         /* f + eps * normalise(g) */
    int64_t rows;       /* positions per sample */
    int32_t c;
    int32_t pad_;
} chap_level;

#define CHAP_PERTURB_SAMPLE 0
#define CHAP_PERTURB_CHANNEL 1
#define CHAP_PERTURB_SPATIAL 2
#define CHAP_PERTURB_CHANNEL_SPATIAL 3
/* workspace: double[ws_elems], ws_elems >= chap_perturb_workspace_elems(levels, n_levels, n) */
size_t chap_perturb_workspace_elems(const chap_level* levels, int32_t n_levels, int32_t n);
/* g is multiplied by g_scale when loaded (chain-rule factor of the probing step) */
int chap_perturb_fwd(const chap_level* levels, int32_t n_levels, int32_t n, int32_t mode, float eps, float g_scale,
                     double* workspace, size_t ws_elems, void* stream);
/* out = base + xi * d / (||d||_2 per sample + 1e-8)  (base nullable); norms: double[n] scratch */
int chap_l2n_sample_axpy(const float* d, const float* base, float xi, int32_t n, int64_t elems_per_sample,
                         double* norms, float* out, void* stream);
/* the same for several tensors in two launches (the five levels of the VAT probe, hat_l = f_l + xi * l2n(d_l)): levels[l] =
 * {g = d_l, f = base_l (nullable), out, rows * c = elements per sample (multiple of 4)}; norms: double[n_levels * n] scratch */
int chap_l2n_sample_axpy_batched(const chap_level* levels, int32_t n_levels, int32_t n, float xi, double* norms, void* stream);

/* ------------------------------------------------------------------ 2D validation (code/val_2D.py:54-97)
 * out[s, Y, X] = in[s, iy[Y], ix[X]] for a stack of n slices [h, w] -> [out_h, out_w]; elem_bytes 4 (float image) or 8 (int64
 * label map).  With scipy.ndimage.zoom(order=0)'s index tables this is the reference's `zoom` (:60, :91) on the device. */
int chap_gather2d(const void* in, int32_t elem_bytes, const int32_t* iy, const int32_t* ix, int32_t n, int32_t h, int32_t w,
                  int32_t out_h, int32_t out_w, void* out, void* stream);
/* counts[c] = {|pred == c and gt == c|, |pred == c|, |gt == c|} as uint64 [classes][3]: the sums medpy.metric.binary.dc needs
 * (code/val_2D.py:43-51), for all classes in one pass over the two int64 label volumes.  classes <= 16. */
int chap_label_overlap(const int64_t* pred, const int64_t* gt, int64_t elems, int32_t classes, uint64_t* counts, void* stream);

/* ------------------------------------------------------------------ optimiser
 * torch.optim.SGD(momentum, weight_decay) step of code/train_ours_2D.py:278,383 on flat buffers:
 * g' = grad_scale*g + wd*p ; buf = first ? g' : mom*buf + g' ; p -= lr*buf */
/* same with the learning rate read from device memory at run time and buf assumed initialised (zeros before the first
 * step give buf = g'): the launch can be captured once in a CUDA graph and replayed with a changing learning rate */
int chap_sgd_momentum_lrdev(float* p, const float* g, float* buf, int64_t elems, const float* lr_dev, float momentum,
                            float weight_decay, float grad_scale, void* stream);
/* The per-iteration scalars of code/train_ours_2D.py computed on the device from an iteration counter in device memory
 * (read, then incremented): lr = base_lr * (1 - it / max_iterations)^0.9 (:387) and
 * cw = consistency * sigmoid_rampup(it / ramp_div, rampup) (:34-36 with the call at :356, ramp_div = 150).
 * Lets a replayed CUDA graph run without any per-iteration host write. */
int chap_schedule_step(int64_t* iter_dev, double base_lr, double max_iterations, double consistency, double rampup,
                       int64_t ramp_div, float* lr_dev, float* cw_dev, void* stream);
int chap_sgd_momentum(float* p, const float* g, float* buf, int64_t elems, float lr, float momentum,
                      float weight_decay, float grad_scale, int32_t first_step, void* stream);

/* ------------------------------------------------------------------ sliding-window aggregation
 * The score_map / cnt accumulation, division and argmax of test_single_case,
 * code/test_3D_util.py:46-72 (identical: code/val_3D.py:46-72).
 */
typedef struct {
    int32_t vol[3];      /* padded volume size (ww, hh, dd) */
    int32_t patch[3];    /* patch size (pw, ph, pd) */
    int32_t nwin[3];     /* windows per axis (sx, sy, sz) */
    int32_t stride[3];   /* (stride_xy, stride_xy, stride_z) */
    int32_t c;           /* classes */
    int32_t pad_;
} chap_sw_desc;
/* gather windows [first, first+count) of the padded volume into patches[count, pw, ph, pd] */
int chap_sw_extract(const chap_sw_desc* d, const float* volume, int32_t first, int32_t count, float* patches, void* stream);
/* win: window logits (is_prob = 0: softmax fused on load) or probabilities (is_prob = 1),
 * channels-last [nwin_total, pw, ph, pd, c], windows ordered x-major then y then z like the
 * reference loops.  Each output voxel is owned by one thread which adds its windows in the reference
 * order (bit-identical fp32 sums), divides by cnt and takes the first-max argmax.
 * score (nullable): [c, ww, hh, dd] (already divided by cnt); cnt (nullable): [ww, hh, dd]. */
int chap_sw_aggregate(const chap_sw_desc* d, const float* win, int32_t is_prob, float* score, float* cnt,
                      int64_t* label, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CHAP_B200_H */
