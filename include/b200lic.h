/*
 * b200lic.h -- flat C ABI of libb200lic.so, the sm_100a (B200) implementation of the RDO-PTQ hot path.
 *
 * The reference (Eric-qi/RDO-PTQ) is pure Python/PyTorch and has no FFI; its boundary is the Python
 * class surface of task-oriented-PTQ/quantization and light-uniform-PTQ/quant_int.  Each entry point
 * below replaces the ATen/cuDNN calls issued by the cited reference lines ("TO" =
 * task-oriented-PTQ/quantization, "LU" = light-uniform-PTQ/quant_int; compressai 1.2.4 call sites
 * are cited through the reference file that reaches them).  INTEGRATION.md shows the ctypes stub a
 * maintainer would add on the reference side.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer to fp32 (unless typed otherwise) owned by the caller; the
 *    library allocates nothing and keeps no pointer after the call returns;
 *  - all work is enqueued on `stream` (a cudaStream_t); no call synchronises the host, so every entry
 *    point is CUDA-graph capturable;
 *  - activations are NCHW contiguous, conv weights [Cout,Cin,KH,KW], transposed-conv weights
 *    [Cin,Cout,KH,KW] (PyTorch layouts);
 *  - return value 0 on success, a negative B200LIC_ERR_* otherwise (never throws, never exits);
 *    b200lic_last_error_string() describes the last failure of the calling thread;
 *  - sm_100 only: on any other device every compute entry point returns B200LIC_ERR_ARCH.
 */
#ifndef B200LIC_H_
#define B200LIC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* b200lic_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define B200LIC_API __attribute__((visibility("default")))
#else
#define B200LIC_API
#endif

enum {
  B200LIC_OK = 0,
  B200LIC_ERR_ARG = -1,         /* bad shape / null pointer / unsupported combination of flags */
  B200LIC_ERR_ARCH = -2,        /* current device is not sm_100 */
  B200LIC_ERR_CUDA = -3,        /* a CUDA runtime/driver call failed (see last_error_string) */
  B200LIC_ERR_UNSUPPORTED = -4  /* valid request this build has no kernel for */
};

/* activation fused into a conv epilogue (TO quant_layer.py:128 `activation_function`) */
enum { B200LIC_ACT_NONE = 0, B200LIC_ACT_RELU = 1, B200LIC_ACT_LEAKY_RELU = 2 };

/* conv engine selection */
enum {
  B200LIC_ENGINE_AUTO = 0,   /* tcgen05 when the shape qualifies, else SIMT */
  B200LIC_ENGINE_SIMT = 1,   /* fp32 CUDA-core implicit GEMM (exact fp32 products) */
  B200LIC_ENGINE_TC = 2      /* tcgen05 tensor cores, split-bf16 operands, fp32 TMEM accumulation */
};

B200LIC_API int b200lic_version(void);
B200LIC_API const char* b200lic_last_error_string(void);
/* 0 when the current CUDA device is sm_100, B200LIC_ERR_ARCH otherwise. */
B200LIC_API int b200lic_device_check(void);
/* Debug aid: with B200LIC_TC_DEBUG=3 in the environment the tensor-core conv kernel records a %globaltimer timeline
 * (ns) of CTA 0's first work item; this copies up to 128 stamps of the last launch to `out` (synchronises). */
B200LIC_API int b200lic_debug_timeline(unsigned long long* out, int n);
/* Scheduling knobs (A/B measurements, tests).  "streamk": 1 = stream-K scheduling of the conv engine where it pays
 * (default), 0 = whole work items only, 2 = wherever the shape is eligible.  Results are identical up to fp32 summation
 * order of the K ranges.  "pair": CTA-pair form of the conv engine (clusters of two CTAs, tcgen05 cta_group::2, M = 256 per
 * MMA, each CTA staging half of the weight tile): 1 = where the plan expects it to pay (default), 0 = off, 2 = wherever
 * eligible.  "gemm1x1": 0 = the generic engine also runs the short-K 1x1 layers. */
B200LIC_API int b200lic_set_option(const char* name, int value);
/* Activation statistics from the producer's epilogue (K8 fused into K1/K2, north_star (1): "... and the next layer's
 * activation quantizer").  b200lic_conv_stats_once(minmax): the NEXT conv / transposed-conv forward this thread launches
 * on the tcgen05 engine (any of the forward entry points; not gdn_mode, Cout <= 640) also merges the per-channel
 * (min, max) of its output -- after bias / activation / Q8.8, the values it stores -- into `minmax` (the key layout of
 * b200lic_actq_stats; initialise it with b200lic_actq_stats_init): the same keys, bit for bit, without the extra pass over
 * the tensor.  b200lic_conv_stats_pending() returns 1 and disarms when no launch has consumed the request (engine or
 * shape without the fused statistics: the caller runs b200lic_actq_stats), 0 when the keys are being written. */
B200LIC_API int b200lic_conv_stats_once(float* minmax);
B200LIC_API int b200lic_conv_stats_pending(void);
/* number of kernel launches issued by this library in this process (bench.py `gpu_launches`). */
B200LIC_API unsigned long long b200lic_launch_count(void);
/* Calls that asked for B200LIC_ENGINE_AUTO and ran on the exact-fp32 SIMT engine because the tcgen05 engine rejected the
 * shape (the first one of a process is also reported on stderr; B200LIC_QUIET=1 silences the message). */
B200LIC_API unsigned long long b200lic_simt_fallback_count(void);

/* ------------------------------------------------------------------------------------------------
 * K7  per-channel weight range + fake-quant.
 * The tensor is viewed as [outer, ch, inner] with the quantisation channel in the middle:
 * conv weight -> (1, Cout, Cin*KH*KW); tconv weight -> (Cin, Cout, KH*KW); GDN gamma -> (1, C, C);
 * per-tensor -> (1, 1, numel).
 * replaces: TO quantizer.py:233-298 (init_quantization_scale, 'max' / 'max_scale'), LU quantizer.py:192-281.
 * delta/zp: [ch].  Arithmetic is the reference's: fp64 range -> fp32 delta, zp = rint(rcp(delta) * -min).
 * ---------------------------------------------------------------------------------------------- */
B200LIC_API int b200lic_wq_init_minmax(const float* w, int outer, int ch, int inner, int n_bits, int scale_variant,
                           int symmetric, float* delta, float* zero_point, b200lic_stream_t stream);

/* Search-based and moment-based ranges of the same tensor view.
 * replaces: TO quantizer.py:300-316 ('mse': 10 shrink steps of 0.05, score mean|x-xq|^3.5), :339-370 ('l1', 'l2'),
 * :318-336 ('gaussian': mean -+ 6*var), LU quantizer.py:265-280 ('mse': 80 steps of 0.01, p = 2).
 * Candidate i uses (max, min) * fp32(1 - i*shrink); the first candidate with the strictly smallest score wins.
 * The candidates' fake-quant runs in the reference's fp32 order; scores are accumulated in fp64. */
enum { B200LIC_SCALE_MSE = 1, B200LIC_SCALE_L1 = 2, B200LIC_SCALE_L2 = 3, B200LIC_SCALE_GAUSSIAN = 4 };
B200LIC_API int b200lic_wq_init_search(const float* w, int outer, int ch, int inner, int n_bits, int method,
                           int n_steps, double shrink, float p, int symmetric, float* delta, float* zero_point,
                           b200lic_stream_t stream);

/* replaces: TO quantizer.py:175-177 (UniformAffineQuantizer.forward), LU quantizer.py:171-177.
 * Any of w_dq / codes / codes_u8 may be NULL.  codes are integer-valued fp32 in [0, n_levels-1]. */
B200LIC_API int b200lic_wq_fake_quant(const float* w, const float* delta, const float* zero_point, int outer, int ch,
                          int inner, int n_levels, float* w_dq, float* codes, uint8_t* codes_u8,
                          b200lic_stream_t stream);

/* LU quant_layer.py:121-122: w = (codes_u8 - zp) * delta */
B200LIC_API int b200lic_wq_dequant_u8(const uint8_t* codes_u8, const float* delta, const float* zero_point, int outer,
                          int ch, int inner, float* w_dq, b200lic_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K6  AdaRound (learned hard sigmoid).
 * replaces: TO quantizer.py:454-466 (init_alpha), :437-452 (forward / get_soft_targets),
 * autograd of those + layer_opt.py:160-165 (rounding regulariser) + torch.optim.Adam (layer_opt.py:254,307).
 * ---------------------------------------------------------------------------------------------- */
B200LIC_API int b200lic_adaround_init_alpha(const float* w, const float* delta, int outer, int ch, int inner,
                                float* alpha, b200lic_stream_t stream);
/* soft != 0: x_int = floor(w/d) + h(alpha); else + (alpha >= 0).  codes may be NULL. */
B200LIC_API int b200lic_adaround_fwd(const float* w, const float* alpha, const float* delta, const float* zero_point,
                         int outer, int ch, int inner, int n_levels, int soft, float* w_q, float* codes,
                         b200lic_stream_t stream);
/* One fused step: d_alpha = grad_scale * d_wq * dWq/dalpha + d(reg)/dalpha; Adam(step) on alpha, m, v
 * in place.  reg_b <= 0 disables the regulariser (warm-up).  reg_loss (device scalar, may be NULL) is
 * atomically incremented by reg_weight * sum(1 - |2h-1|^b).  d_alpha_out (may be NULL) receives the
 * gradient that was applied.  exp_avg == exp_avg_sq == NULL selects gradient-only mode (alpha untouched). */
B200LIC_API int b200lic_adaround_bwd_adam(const float* w, float* alpha, const float* delta, const float* zero_point,
                              const float* d_wq, float* exp_avg, float* exp_avg_sq, int outer, int ch,
                              int inner, int n_levels, int step, float lr, float beta1, float beta2,
                              float eps, float grad_scale, float reg_weight, float reg_b, float* reg_loss,
                              float* d_alpha_out, b200lic_stream_t stream);

/* Integer weights n = code - zero_point (integer-valued fp32, |n| <= n_levels - 1): the weight operand of
 * b200lic_conv_fwd_wq / b200lic_deconv_fwd_wq.  alpha == NULL: nearest rounding (TO quantizer.py:175-177); alpha given:
 * hardened AdaRound, floor(w/d) + (alpha >= 0) (quantizer.py:437-449 with soft_targets False).  Same arithmetic as
 * b200lic_wq_fake_quant / b200lic_adaround_fwd, so n * delta equals their dequantised weight bit for bit. */
B200LIC_API int b200lic_wq_int_weights(const float* w, const float* alpha, const float* delta, const float* zero_point,
                           int outer, int ch, int inner, int n_levels, float* w_int, b200lic_stream_t stream);

/* Learned step size (LSQ-style): d loss / d delta[ch] from dL/dWq, i.e. the autograd of the fake-quant expression with
 * delta as the leaf -- the option the reference keeps as commented-out code (TO quantizer.py:166-168,
 * layer_opt.py:259-265, block_opt.py:254-266).
 *   alpha == NULL (nearest, quantizer.py:175-177): d_delta[c] = grad_scale * sum d_wq * ((x_q - zp) - in_range * w/delta)
 *   alpha != NULL (AdaRound, quantizer.py:437-449; soft selects h(alpha) or alpha >= 0): sum d_wq * (x_q - zp)
 * d_delta may be NULL when the Adam moments exp_avg / exp_avg_sq ([ch], with step >= 1) are given: delta is then
 * updated in place (clamped at 1e-8 like quantizer.py:292); zero_point stays fixed. */
B200LIC_API int b200lic_lsq_delta_grad(const float* w, const float* alpha, float* delta, const float* zero_point,
                           const float* d_wq, int outer, int ch, int inner, int n_levels, int soft, float grad_scale,
                           float* d_delta, float* exp_avg, float* exp_avg_sq, int step, float lr, float beta1,
                           float beta2, float eps, b200lic_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K8  activation quantisers.
 * replaces: TO quantizer.py:81-121 (Handle_Parameter/ActQuant: dynamic per-channel, 4-D NCHW),
 * LU quantizer.py:120-128 (static Q8.8).
 * ---------------------------------------------------------------------------------------------- */
/* minmax: [2*C] 32-bit words holding order-preserving keys of the per-channel (min, max) over (N,H,W);
 * opaque to the caller, must be initialised by b200lic_actq_stats_init before b200lic_actq_stats. */
B200LIC_API int b200lic_actq_stats_init(float* minmax, int C, b200lic_stream_t stream);
B200LIC_API int b200lic_actq_stats(const float* x, int N, int C, int HW, float* minmax, b200lic_stream_t stream);
B200LIC_API int b200lic_actq_apply(const float* x, const float* minmax, int N, int C, int HW, int n_bits, float* out,
                       float* codes, b200lic_stream_t stream);
/* b200lic_actq_apply fused with the staging of the NEXT layer's tensor-core operand: the quantised activation (same codes
 * as b200lic_actq_apply, bit for bit) is written as the split-bf16 NHWC operand [N,HW,cpad] (x*x when square != 0, for a
 * GDN consumer) into the slot b200lic_conv_x_slot reports; `out` (may be NULL) also receives it as fp32 NCHW. */
B200LIC_API int b200lic_actq_apply_stage(const float* x, const float* minmax, int N, int C, int HW, int n_bits, int square,
                             void* x_hi, void* x_lo, int cpad, float* out, b200lic_stream_t stream);
/* The three calls above in one launch (no minmax buffer): one thread-block cluster per channel keeps the channel's
 * elements in (distributed) shared memory between the min/max reduction and the quantisation, so the activation is read
 * from HBM once.  Bit-identical to stats_init + stats + apply. */
B200LIC_API int b200lic_actq_fused(const float* x, int N, int C, int HW, int n_bits, float* out, float* codes,
                       b200lic_stream_t stream);
B200LIC_API int b200lic_fixed_point(const float* x, size_t n, int a_l, int a_r, float* out, b200lic_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K9 / K10  entropy-model likelihoods (compressai 1.2.4 semantics; reference call sites
 * TO models/nic_cvt.py:297-308, quant_model.py:72-79).
 * bits (device fp32 scalar, may be NULL) is atomically incremented by sum(-log2(lik)).
 * ---------------------------------------------------------------------------------------------- */
/* y_hat = rint(y - mu) + mu; lik = Phi((.5-|y_hat-mu|)/s) - Phi((-.5-|y_hat-mu|)/s), s = max(scale, scale_bound).
 * means may be NULL (zero-mean).  scales/means may be strided views of one tensor (chunk(2,1)):
 * element (n,c,i) lives at n*param_batch_stride + c*HW + i. */
B200LIC_API int b200lic_gaussian_lik_fwd(const float* y, const float* scales, const float* means, int N, int C, int HW,
                             long long param_batch_stride, float scale_bound, float lik_bound, float* y_hat,
                             float* lik, float* bits, b200lic_stream_t stream);
/* rint(y - mu) + mu only (GaussianConditional.quantize "dequantize"); means may be NULL. */
B200LIC_API int b200lic_round_latent(const float* y, const float* means, size_t n, float* y_hat, b200lic_stream_t stream);
/* Factorised prior, filters (3,3,3,3).  params: [C][58] = matrices 3+9+9+9+3 (raw, softplus applied inside),
 * biases 3+3+3+3+1, factors 3+3+3+3 (raw, tanh applied inside).  medians: [C]. */
/* `table` (may be NULL) = the per-channel symbol tables written by b200lic_factorized_table for the same params /
 * medians / lik_bound; with NULL every CTA rebuilds its channel's table (same values, ~3 us per CTA). */
B200LIC_API int b200lic_factorized_lik_fwd(const float* z, const float* params, const float* medians, const float* table,
                               int N, int C, int HW, float lik_bound, float* z_hat, float* lik, float* bits,
                               b200lic_stream_t stream);
/* Symbol tables of the factorised prior: table[c][0][k + R] = likelihood of the symbol medians[c] + k and
 * table[c][1][k + R] = -log2 of it for |k| <= R; b200lic_factorized_table_floats() = floats per channel (2 * (2R + 1)).
 * They depend on the prior's parameters only, which PTQ never trains (compressai EntropyBottleneck reached from TO
 * models/nic_cvt.py:297): build once per model, reuse for every forward. */
B200LIC_API int b200lic_factorized_table_floats(void);
B200LIC_API int b200lic_factorized_table(const float* params, const float* medians, int C, float lik_bound, float* table,
                             b200lic_stream_t stream);

/* Backward of the two likelihood kernels for the rate term of RateDistortionLoss (TO losses/losses.py:20-28; the
 * R + lambda*D task criterion the reference keeps commented out at layer_opt.py:146-148).
 * Upstream gradient per element: g = g_lik[i] (may be NULL) + *g_bits (device scalar, may be NULL) * d(-log2 lik)/dlik;
 * g_yhat / g_zhat (may be NULL) is the gradient arriving at the rounded latent.  LowerBound semantics of compressai:
 * a gradient passes a bound iff value >= bound or the gradient is negative.
 * ste == 0: compressai autograd (torch.round has zero gradient): d_y = 0, d_means = g_yhat, d_scales live.
 * ste != 0: straight-through rounding (round_ste, TO quantizer.py:64-68, applied to y at layer_opt.py:69):
 *           d_y = g_yhat + g*dlik/dv, d_means = -g*dlik/dv (v = y_hat - mu).
 * d_scales / d_means use grad_batch_stride (so they can be the two halves of one [N,2C,H,W] gradient tensor);
 * d_y may be NULL. */
B200LIC_API int b200lic_gaussian_lik_bwd(const float* y_hat, const float* scales, const float* means, const float* g_lik,
                             const float* g_bits, const float* g_yhat, int N, int C, int HW,
                             long long param_batch_stride, long long grad_batch_stride, float scale_bound,
                             float lik_bound, int ste, float* d_y, float* d_scales, float* d_means,
                             b200lic_stream_t stream);
/* d_z of the factorised prior (parameters are frozen in PTQ).  ste == 0 writes zeros. */
B200LIC_API int b200lic_factorized_lik_bwd(const float* z_hat, const float* params, const float* medians,
                               const float* g_lik, const float* g_bits, const float* g_zhat, int N, int C, int HW,
                               float lik_bound, int ste, float* d_z, b200lic_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K11  losses.
 * replaces: TO quantizer.py:71-79 (lp_loss), losses/losses.py:20-28, test_datasets.py:21-33.
 * ---------------------------------------------------------------------------------------------- */
/* loss += scale * sum |pred-tgt|^p ; d_pred (may be NULL) = grad_scale * p*|d|^(p-1)*sign(d). */
B200LIC_API int b200lic_lp_loss_fwd_bwd(const float* pred, const float* tgt, size_t n, float p, float scale, float grad_scale,
                            float* loss, float* d_pred, b200lic_stream_t stream);
/* out[0] += sum (a-b)^2 ; out[1] += sum (clamp(a,0,1)-b)^2 (test_datasets.py:98 clamps before PSNR). */
B200LIC_API int b200lic_sq_err_sum(const float* a, const float* b, size_t n, float* out, b200lic_stream_t stream);
/* out[0] += sum -log2(lik) */
B200LIC_API int b200lic_bits_sum(const float* lik, size_t n, float* out, b200lic_stream_t stream);

/* Token-major pieces of the Linear / LayerNorm wrappers (TO quantization/quant_layer.py:38-49,117-121; the Swin blocks of
 * quant_block.py:330-641 are built from them).  F.linear over [rows, Cin] tokens is the 1x1 convolution of
 * `rows` one-pixel images: describe it as b200lic_conv_desc{N=rows, Cin, H=W=1, Cout, Ho=Wo=1, KH=KW=1, stride 1, pad 0},
 * stage the operand with b200lic_stage_tokens into the slot b200lic_conv_x_slot reports (fp32 rows -> split-bf16 rows,
 * channels padded to cpad) and run b200lic_conv_fwd_packed with x = NULL: the output [rows, Cout, 1, 1] is the token
 * matrix [rows, Cout]. */
B200LIC_API int b200lic_stage_tokens(const float* x, size_t rows, int C, int cpad, void* x_hi, void* x_lo,
                         b200lic_stream_t stream);
/* F.layer_norm(x, (C,), gamma, beta, eps) over the last axis of x [rows, C] (gamma / beta may be NULL). */
B200LIC_API int b200lic_layernorm_fwd(const float* x, const float* gamma, const float* beta, size_t rows, int C, float eps,
                          float* y, b200lic_stream_t stream);
/* ActQuantizer (TO quantizer.py:81-121) of a token-major tensor: per LAST-axis channel over all rows (the reference's 3-D
 * branch, `x_clone[:,:,i]`), dynamic (min, max), n_bits levels; same arithmetic as b200lic_actq_apply.  minmax: 2*C words of
 * scratch for the keys. */
B200LIC_API int b200lic_actq_tokens(const float* x, size_t rows, int C, int n_bits, float* minmax, float* out,
                        b200lic_stream_t stream);
/* Window attention core of the Swin blocks (TO models/layers.py:137-166; quantised form quant_block.py:383-418), in two
 * halves because the reference's ActQuantizer sits between them.  qkv: [B_, N, 3C] (q | k | v, each nH heads of C/nH);
 * bias: [nH, N, N] relative-position bias already gathered; mask: [nW, N, N] shift mask or NULL (window b uses row
 * b % nW).  N <= 64 tokens per window, head dimension <= 48 (B200LIC_ERR_UNSUPPORTED beyond).
 *   P[b,h,i,j]   = softmax_j((q_i * scale) . k_j + bias[h,i,j] + mask[b % nW,i,j])        -> P [B_, nH, N, N]
 *   out[b,i,h*hd+d] = sum_j P[b,h,i,j] * v[j,d]   (= (attn @ v).transpose(1, 2).reshape(B_, N, C))  -> out [B_, N, C] */
B200LIC_API int b200lic_window_attn_softmax(const float* qkv, const float* bias, const float* mask, int B_, int N, int C, int nH,
                                int nW, float scale, float* P, b200lic_stream_t stream);
B200LIC_API int b200lic_window_attn_apply(const float* P, const float* qkv, int B_, int N, int C, int nH, float* out,
                              b200lic_stream_t stream);
/* nn.GELU() in its exact (erf) form. */
B200LIC_API int b200lic_gelu_fwd(const float* x, size_t n, float* y, b200lic_stream_t stream);

/* MS-SSIM, the third number the reference's entry points report next to PSNR and bpp.
 * replaces: pytorch_msssim.ms_ssim(a, b, data_range=1.) at TO losses/losses.py:26,31,49-52, LU quantize.py:13,89,
 * quant.py:86, single_test.py:59-60, dataset_test.py:60-61 (third-party package, pytorch_msssim==1.0.0).
 * One level: x, y are [planes, H, W] fp32 (planes = N*C); win11 = the normalised 11-tap Gaussian (sigma 1.5);
 * sums[2*p] += sum of the ssim map of plane p, sums[2*p+1] += sum of its cs map (VALID region (H-10) x (W-10)); the caller
 * zeroes `sums`.  c1 = (0.01 L)^2, c2 = (0.03 L)^2. */
B200LIC_API int b200lic_ssim_level(const float* x, const float* y, const float* win11, int planes, int H, int W, float c1,
                       float c2, double* sums, b200lic_stream_t stream);
/* 2x2 mean with stride 2 and zero padding pad_h / pad_w in {0,1} counted in the mean (avg_pool2d(kernel_size=2,
 * padding=size % 2) between the levels): out is [planes, (H+2*pad_h-2)/2+1, (W+2*pad_w-2)/2+1]. */
B200LIC_API int b200lic_avg_pool2(const float* x, int planes, int H, int W, int pad_h, int pad_w, float* out,
                      b200lic_stream_t stream);
/* per_plane[p] = prod_l relu(v_l[p])^w_l, v_l = cs mean of level l < levels-1, ssim mean of the last level
 * (sums: [levels][planes][2], inv_count[l] = 1 / pixels of level l's maps), weights (0.0448, 0.2856, 0.3001, 0.2363,
 * 0.1333); mean[0] = mean over planes. */
B200LIC_API int b200lic_msssim_combine(const double* sums, const double* inv_count, int levels, int planes, float* per_plane,
                           float* mean, b200lic_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K1/K2/K4/K5  convolutions as implicit GEMM.
 * replaces: F.conv2d TO quant_layer.py:28,123 / LU quant_layer.py:27,128; F.conv_transpose2d
 * quant_layer.py:36; their autograd (layer_opt.py:298-307, block_opt.py:299-309).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int N, Cin, H, W;         /* input  [N,Cin,H,W]   */
  int Cout, Ho, Wo;         /* output [N,Cout,Ho,Wo] */
  int KH, KW, stride, pad;  /* square stride/pad, dilation 1, groups 1 */
  int act;                  /* B200LIC_ACT_* applied to the output (forward only) */
  float act_slope;          /* LeakyReLU negative slope */
  int engine;               /* B200LIC_ENGINE_* */
  int in_square;            /* forward/wgrad: use x*x instead of x (GDN's 1x1 conv on x^2) */
  int gdn_mode;             /* forward: 0 plain; 1 out = gdn_x * rsqrt(acc); 2 out = gdn_x * sqrt(acc) */
  int fixed_point;          /* forward: apply LU Q8.8 round(clamp(v,-128,128)*256)/256 to the output */
  int k_taps;               /* b200lic_conv_fwd_packed, plain conv: contract only the first k_taps filter taps in raster
                               order (0 = all).  The caller asserts that the remaining taps of the packed weight are zero --
                               compressai's MaskedConv2d, mask 'A': 12 of 5x5 (TO quant_model.py:45-48 wraps it as a dense
                               conv).  Every other entry point ignores the field (same value, dense contraction). */
} b200lic_conv_desc;

/* Scratch bytes the tensor-core engine needs for `op` on this shape (operand staging: NHWC split-bf16 activations and
 * packed weights); 0 when the shape only runs on the SIMT engine.  The caller owns the workspace; with a NULL or
 * too-small workspace ENGINE_AUTO uses the SIMT engine and ENGINE_TC fails with B200LIC_ERR_UNSUPPORTED. */
enum { B200LIC_OP_CONV_FWD = 0, B200LIC_OP_DECONV_FWD = 1, B200LIC_OP_CONV_DGRAD = 2, B200LIC_OP_DECONV_DGRAD = 3,
       B200LIC_OP_CONV_WGRAD = 4, B200LIC_OP_DECONV_WGRAD = 5 };
B200LIC_API size_t b200lic_conv_workspace_bytes(const b200lic_conv_desc* d, int op);
/* How the generic tcgen05 engine would run a forward problem (host-side planning only, no device work; tests and tuning
 * scripts): info[0] eligible, [1] output-channel tile BN, [2] channel tiles, [3] pixel tiles per CTA and work item,
 * [4] pixel tiles of the largest phase, [5] work items, [6] CTA-pair form, [7] stream-K, [8] grid in CTAs, [9] shared-memory
 * stages, [10] epilogue warps, [11] tensor-memory columns.  Folded-tap and short-K 1x1 layers take other kernels (the plan
 * describes the generic engine only). */
B200LIC_API int b200lic_conv_plan_info(const b200lic_conv_desc* d, int op, int* info);

/* y = act(conv2d(x, w) + bias).  gdn_x (may be NULL unless gdn_mode) is [N,Cout,Ho,Wo]; norm_out (may be NULL)
 * receives the pre-(r)sqrt accumulator in gdn_mode. */
B200LIC_API int b200lic_conv_fwd(const b200lic_conv_desc* d, const float* x, const float* w, const float* bias,
                     const float* gdn_x, float* norm_out, float* y, void* workspace, size_t workspace_bytes,
                     b200lic_stream_t stream);
/* y = act(conv_transpose2d(x, w) + bias); w is [Cin,Cout,KH,KW]; Ho/Wo carry the output_padding. */
B200LIC_API int b200lic_deconv_fwd(const b200lic_conv_desc* d, const float* x, const float* w, const float* bias, float* y,
                       void* workspace, size_t workspace_bytes, b200lic_stream_t stream);
/* Forward of a layer whose weight is HARD-quantised per OUTPUT channel (evaluation, TO quant_layer.py:113-123 with
 * the quantiser of quantizer.py:175-177 or a hardened AdaRound): y = act(conv(x, w_int) * w_scale[co] + bias[co]), where
 * w_int = code - zero_point (b200lic_wq_int_weights, n_levels <= 256 so that |n| <= 255 is exact in bf16) and
 * w_scale = delta per output channel.  Equal to b200lic_conv_fwd on the dequantised weight up to fp32 rounding, with
 * two tensor-core passes per product instead of three (the weight has no lo slice).  Tensor-core engine only; returns
 * B200LIC_ERR_UNSUPPORTED for shapes that run on the folded-tap path (3-channel layers): call b200lic_conv_fwd there. */
B200LIC_API int b200lic_conv_fwd_wq(const b200lic_conv_desc* d, const float* x, const float* w_int, const float* w_scale,
                        const float* bias, float* y, void* workspace, size_t workspace_bytes, b200lic_stream_t stream);
B200LIC_API int b200lic_deconv_fwd_wq(const b200lic_conv_desc* d, const float* x, const float* w_int, const float* w_scale,
                          const float* bias, float* y, void* workspace, size_t workspace_bytes,
                          b200lic_stream_t stream);
/* dw[Cout,Cin,KH,KW] = sum_pixels dy (x) x.  dw is overwritten. */
B200LIC_API int b200lic_conv_wgrad(const b200lic_conv_desc* d, const float* x, const float* dy, float* dw,
                       void* workspace, size_t workspace_bytes, b200lic_stream_t stream);
/* dw[Cin,Cout,KH,KW] for y = conv_transpose2d(x, w). */
B200LIC_API int b200lic_deconv_wgrad(const b200lic_conv_desc* d, const float* x, const float* dy, float* dw,
                         void* workspace, size_t workspace_bytes, b200lic_stream_t stream);
/* Staged-operand reuse between forward and weight gradient.  The tensor-core forward leaves a split-bf16 NHWC copy of
 * x in its workspace; when that copy is directly consumable by the wgrad engine (x_hi / x_lo come back non-NULL) the
 * caller may keep the forward workspace alive and call b200lic_conv_wgrad_staged instead of b200lic_conv_wgrad /
 * b200lic_deconv_wgrad, which skips re-staging x (one read + one write of the activation per iteration).
 * `op` is B200LIC_OP_CONV_FWD or B200LIC_OP_DECONV_FWD; `d` must be the descriptor of the forward call. */
B200LIC_API int b200lic_conv_staged_view(const b200lic_conv_desc* d, int op, void* fwd_workspace, size_t workspace_bytes,
                             void** x_hi, void** x_lo);
B200LIC_API int b200lic_conv_wgrad_staged(const b200lic_conv_desc* d, int transposed, const void* x_hi, const void* x_lo,
                              const float* dy, float* dw, void* workspace, size_t workspace_bytes,
                              b200lic_stream_t stream);
/* dx[N,Cin,H,W] for y = conv2d(x, w). */
B200LIC_API int b200lic_conv_dgrad(const b200lic_conv_desc* d, const float* dy, const float* w, float* dx,
                       void* workspace, size_t workspace_bytes, b200lic_stream_t stream);
/* dx[N,Cin,H,W] for y = conv_transpose2d(x, w). */
B200LIC_API int b200lic_deconv_dgrad(const b200lic_conv_desc* d, const float* dy, const float* w, float* dx,
                         void* workspace, size_t workspace_bytes, b200lic_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K3  GDN / IGDN helpers (compressai GDN via TO quant_layer.py:142-154).
 * ---------------------------------------------------------------------------------------------- */
/* NonNegativeParametrizer forward: out = max(p, bound)^2 - pedestal. */
B200LIC_API int b200lic_gdn_reparam_fwd(const float* p, size_t n, float bound, float pedestal, float* out,
                            b200lic_stream_t stream);
/* its backward incl. LowerBound's gate: dp = [(p>=bound)|(g<0)] * g, g = d_out*2*max(p,bound). */
B200LIC_API int b200lic_gdn_reparam_bwd(const float* p, const float* d_out, size_t n, float bound, float* d_p,
                            b200lic_stream_t stream);
/* d_norm = dy * x * dfn(norm), dfn = -0.5*norm^-1.5 (GDN) or 0.5*norm^-0.5 (IGDN);
 * dx_direct (may be NULL) = dy * fn(norm). */
B200LIC_API int b200lic_gdn_bwd_prep(const float* x, const float* norm, const float* dy, size_t n, int inverse,
                         float* d_norm, float* dx_direct, b200lic_stream_t stream);
/* dx = dx_direct + 2*x*t   (t = gamma^T-contracted d_norm, produced by b200lic_conv_dgrad) */
B200LIC_API int b200lic_gdn_bwd_finish(const float* x, const float* t, const float* dx_direct, size_t n, float* dx,
                           b200lic_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * small fused elementwise helpers used by the block wrappers (TO quant_block.py:219-328) and the
 * calibration loop (layer_opt.py:289-292).
 * ---------------------------------------------------------------------------------------------- */
/* out = act(a + b) ; b may be NULL */
B200LIC_API int b200lic_add_act(const float* a, const float* b, size_t n, int act, float slope, float* out,
                    b200lic_stream_t stream);
/* d_in = d_out * act'(y) where y is the activation OUTPUT */
B200LIC_API int b200lic_act_bwd(const float* y, const float* d_out, size_t n, int act, float slope, float* d_in,
                    b200lic_stream_t stream);
/* Batch pick + QDrop mix (layer_opt.py:289-292): out[b,:] = keep ? q[idx[b],:] : fp[idx[b],:] for b < rows.
 * idx (int64, may be NULL = identity).  keep = mask[b,:] != 0 when mask (uint8) is given, else a counter-based
 * hash RNG(seed, element) < prob. */
B200LIC_API int b200lic_gather_mix(const float* q, const float* fp, const long long* idx, size_t rows, size_t row_elems,
                       float prob, unsigned long long seed, const uint8_t* mask, float* out,
                       b200lic_stream_t stream);
/* ------------------------------------------------------------------------------------------------
 * Device-resident calibration schedule: everything that changes from one AdaRound iteration to the
 * next (layer_opt.py:287-309: iteration counter -> Adam bias corrections, LinearTempDecay(count) of
 * utils.py:37-54, warm-up gate of layer_opt.py:156-158, the batch pick and the QDrop draw) lives in
 * DEVICE memory, so one captured CUDA graph of the iteration can be replayed 20 000 times with no
 * host-side scalar in any kernel argument.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int step;            /* 1-based index of the iteration being executed (0 = before the first tick) */
  float lr_over_bc1;   /* lr / (1 - beta1^step) */
  float inv_sqrt_bc2;  /* 1 / sqrt(1 - beta2^step) */
  float reg_b;         /* LinearTempDecay(step), 0 while step < warmup*iters (regulariser off) */
} b200lic_calib_sched;
/* step += 1 and recompute the derived scalars (one thread).  `sched` must be zero-initialised before the
 * first tick. */
B200LIC_API int b200lic_calib_sched_tick(b200lic_calib_sched* sched, int iters, double warmup, double b_start, double b_end,
                             float lr, float beta1, float beta2, b200lic_stream_t stream);
/* b200lic_adaround_bwd_adam with step / lr / bias corrections / reg_b read from `sched`. */
B200LIC_API int b200lic_adaround_bwd_adam_sched(const float* w, float* alpha, const float* delta, const float* zero_point,
                                    const float* d_wq, float* exp_avg, float* exp_avg_sq, int outer, int ch,
                                    int inner, int n_levels, const b200lic_calib_sched* sched, float beta1,
                                    float beta2, float eps, float grad_scale, float reg_weight, float* reg_loss,
                                    b200lic_stream_t stream);
/* b200lic_lsq_delta_grad with the Adam step on delta driven by the schedule: lr = lr_scale * (the schedule's lr). */
B200LIC_API int b200lic_lsq_delta_grad_sched(const float* w, const float* alpha, float* delta, const float* zero_point,
                                 const float* d_wq, int outer, int ch, int inner, int n_levels, int soft,
                                 float grad_scale, float* d_delta, float* exp_avg, float* exp_avg_sq,
                                 const b200lic_calib_sched* sched, float lr_scale, float beta1, float beta2, float eps,
                                 b200lic_stream_t stream);
/* b200lic_gather_mix whose batch pick and QDrop seed follow the schedule: k = (sched->step - 1) * units + unit;
 * idx = idx_table[k % table_rows][0..rows) (idx_table NULL = identity rows); seed = (seed_base + k) mod 2^48. */
B200LIC_API int b200lic_gather_mix_sched(const float* q, const float* fp, const long long* idx_table, int table_rows,
                             size_t rows, size_t row_elems, float prob, unsigned long long seed_base, int units,
                             int unit, const b200lic_calib_sched* sched, float* out, b200lic_stream_t stream);

/* b200lic_lp_loss_fwd_bwd against a target batch picked from a cache: row b of pred [rows, row_elems] is compared
 * with tgt_cache[idx_table[k % table_rows][b]], k = (sched->step - 1) * units + unit -- the same pick
 * b200lic_gather_mix_sched makes for the input, so the target batch (layer_opt.py:290 `cur_out = cached_outs[idx]`)
 * is never materialised. */
B200LIC_API int b200lic_lp_loss_fwd_bwd_sched(const float* pred, const float* tgt_cache, const long long* idx_table,
                                  int table_rows, size_t rows, size_t row_elems, int units, int unit,
                                  const b200lic_calib_sched* sched, float p, float scale, float grad_scale,
                                  float* loss, float* d_pred, b200lic_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Prepared operands: the fused form of the reference's QuantModule.forward (TO quant_layer.py:107-134: weight
 * quantiser -> conv -> activation) and of one AdaRound iteration (layer_opt.py:287-309), in which the kernels that
 * already hold the values in registers emit the tensor-core operands directly, so the two staging kernels (NHWC
 * split of the activation, weight pack) in front of every GEMM disappear:
 *   batch pick + QDrop mix (layer_opt.py:289-292)            -> activation operand   b200lic_stage_mix_sched
 *   weight quantiser (quantizer.py:175-177 / :437-449)        -> weight operand       b200lic_quant_pack_weights
 *   lp_loss value + gradient (quantizer.py:71-79)             -> dY operand of wgrad  b200lic_lp_loss_stage_sched
 *   split-K sum of dW -> STE masks + regulariser + Adam        (layer_opt.py:160-165,298-307) b200lic_conv_wgrad_adam_sched
 * A frozen layer's weight operand is built once and kept by the caller (the reference re-quantises every weight on
 * every forward).  Tensor-core engine, generic (tap-by-tap) path only: for the folded-tap 3-channel layers and the
 * SIMT engine the size / slot queries return 0 / NULL and the compute calls B200LIC_ERR_UNSUPPORTED -- use the plain
 * entry points there.
 * ---------------------------------------------------------------------------------------------- */
/* Bytes of the packed weight operand of a forward op (B200LIC_OP_CONV_FWD / _DECONV_FWD); depends on the layer only
 * (Cin, Cout, KH, KW, stride), not on N / H / W.  `packed` buffers must be 128-byte aligned. */
B200LIC_API size_t b200lic_conv_packed_weight_bytes(const b200lic_conv_desc* d, int op);
/* fp32 weight ([Cout,Cin,KH,KW], or [Cin,Cout,KH,KW] for the transposed op) -> packed operand. */
B200LIC_API int b200lic_conv_pack_weights(const b200lic_conv_desc* d, int op, const float* w, void* packed,
                              size_t packed_bytes, b200lic_stream_t stream);
/* Weight quantiser fused with the packing.  alpha == NULL: nearest rounding (b200lic_wq_fake_quant); alpha given:
 * AdaRound with soft / hard targets (b200lic_adaround_fwd).  The channel view (outer, ch, inner) is the one of K6/K7.
 * integer_mode != 0 packs n = code - zero_point (the operand of the two-pass forward: pass delta per output channel as
 * w_scale to b200lic_conv_fwd_packed; needs n_levels <= 256).  w_q (may be NULL) also receives the fp32 weight that was
 * packed (needed when a dgrad will re-pack it transposed).  Values are bit-identical to the un-fused kernels. */
B200LIC_API int b200lic_quant_pack_weights(const b200lic_conv_desc* d, int op, const float* w, const float* alpha,
                               const float* delta, const float* zero_point, int outer, int ch, int inner, int n_levels,
                               int soft, int integer_mode, void* packed, size_t packed_bytes, float* w_q,
                               b200lic_stream_t stream);
/* b200lic_conv_fwd / b200lic_deconv_fwd / b200lic_conv_fwd_wq with a prepared weight operand.  x == NULL: the
 * activation operand is already staged in the workspace slot b200lic_conv_x_slot reports.  w_scale (may be NULL)
 * selects the two-pass integer-weight form. */
B200LIC_API int b200lic_conv_fwd_packed(const b200lic_conv_desc* d, int op, const float* x, const void* packed_w,
                            const float* w_scale, const float* bias, const float* gdn_x, float* norm_out, float* y,
                            void* workspace, size_t workspace_bytes, b200lic_stream_t stream);
/* GDN / IGDN forward in one kernel over the raw fp32 NCHW tensor (replaces compressai.layers.GDN.forward as wrapped by
 * TO quant_layer.py:57-78; evaluation path): y = xq * rsqrt(beta + gamma . xq^2) (inverse: * sqrt), gamma given as the
 * prepared operand of b200lic_conv_pack_weights for the 1x1 descriptor, beta[C] the reparametrised bias.  minmax != NULL:
 * x is the un-quantised output of the previous layer and minmax its b200lic_actq_stats result -- xq is the n_bits dynamic
 * quantiser of x (n_bits <= 8; bit-identical to b200lic_actq_apply), applied on chip; minmax == NULL: xq = x.  HBM traffic
 * 8 B/element.
 * b200lic_gdn_fused_ok: 1 when the shape is eligible (HW % 4 == 0, C <= 256); otherwise B200LIC_ERR_UNSUPPORTED. */
B200LIC_API int b200lic_gdn_fused_ok(int C, int HW);
/* Test hook: counts, into *mismatches_dev (device, caller-zeroed), the pairs for which the kernel's FMA division differs
 * from the correctly rounded one over n pseudo-random (value, range) pairs and every code / levels table. */
B200LIC_API int b200lic_selftest_fast_div(unsigned long long n, unsigned long long seed, unsigned long long* mismatches_dev,
                              b200lic_stream_t stream);
B200LIC_API int b200lic_gdn_fwd_fused(const float* x, const float* minmax, int n_bits, const void* packed_gamma,
                          const float* beta, int N, int C, int HW, int inverse, float* y, b200lic_stream_t stream);
/* The two data movements of the folded-tap layers (3 -> N analysis conv, N -> 3 synthesis transposed conv; the engine
 * folds their taps into the channel axis) as entry points, so the fused AdaRound iteration can run those layers as 1x1
 * problems on prepared operands.  b200lic_im2col_stage: out[n, (ho,wo), (c,r,s)] = x[n, c, ho*stride - pad + r,
 * wo*stride - pad + s] written as the split-bf16 NHWC operand (cpad channels, zero padded) -- the activation operand of
 * the folded conv forward / its weight gradient, or (applied to dL/dy) the dY operand of the folded transposed conv's
 * weight gradient.  b200lic_col2im: y = act(bias + scatter-add of col [N, Cout*KH*KW, H, W]) -- the tail of the folded
 * transposed conv forward (autograd of F.conv2d / F.conv_transpose2d, TO layer_opt.py:298-307). */
B200LIC_API int b200lic_im2col_stage(const float* x, int N, int C, int H, int W, int KH, int KW, int stride, int pad,
                         int Ho, int Wo, void* x_hi, void* x_lo, int cpad, b200lic_stream_t stream);
B200LIC_API int b200lic_col2im(const float* col, const float* bias, int N, int Cout, int H, int W, int KH, int KW, int stride,
                   int pad, int Ho, int Wo, int act, float slope, int fixed_point, float* y, b200lic_stream_t stream);
/* ---- Entropy coding of the latents -------------------------------------------------------------
 * What the reference reaches through compressai 1.2.4 (`update()`, `compress()`, `decompress()`; task-oriented-PTQ/
 * models/nic_cvt.py:426-570, light-uniform-PTQ/models/tinylic.py:236-367).
 * b200lic_pmf_to_quantized_cdf: HOST pointers, runs on the host (table build, parameter-sized).  Row r of pmf
 * [rows, max_length] (pmf_length[r] live entries) followed by tail_mass[r] becomes the 16-bit CDF row r of cdf_out
 * [rows, max_length + 2] (pmf_length[r] + 2 live entries, zero padded), every symbol keeping a non-zero width
 * (compressai ops.cpp pmf_to_quantized_cdf + EntropyModel._pmf_to_cdf). */
B200LIC_API int b200lic_pmf_to_quantized_cdf(const float* pmf, const float* tail_mass, const int* pmf_length, int rows,
                                 int max_length, int* cdf_out);
/* symbols = round_half_even(x - means) (means: NULL, [C] with means_per_channel != 0, or like x) and the table row of
 * every element: the channel (scales == NULL: EntropyBottleneck) or GaussianConditional.build_indexes(scales) over the
 * ascending scale_table[levels] with the lower bound scale_bound.  x is [*, C, HW]-shaped, n elements. */
B200LIC_API int b200lic_rans_symbols(const float* x, const float* means, int means_per_channel, const float* scales,
                         const float* scale_table, int levels, float scale_bound, int C, int HW, size_t n, int* symbols,
                         int* indexes, b200lic_stream_t stream);
/* Chunked rANS (ryg rans64 as compressai's rans_interface.cpp drives it: 16-bit CDFs, 32-bit words, 4-bit bypass digits
 * behind the last CDF entry): chunk c = symbols [c*chunk, (c+1)*chunk) is one complete stream.  _sizes writes the word
 * count of every chunk; the caller prefix-sums them into chunk_off[n_chunks + 1]; _write stores stream c at
 * out_words[chunk_off[c] .. chunk_off[c+1]); _decode is the inverse.  cdf [rows, cdf_stride], cdf_len, offset as
 * produced by b200lic_pmf_to_quantized_cdf / the models' update(). */
B200LIC_API int b200lic_rans_encode_sizes(const int* symbols, const int* indexes, unsigned n, unsigned chunk, const int* cdf,
                              const int* cdf_len, const int* offset, int cdf_stride, unsigned* chunk_words,
                              b200lic_stream_t stream);
B200LIC_API int b200lic_rans_encode_write(const int* symbols, const int* indexes, unsigned n, unsigned chunk, const int* cdf,
                              const int* cdf_len, const int* offset, int cdf_stride, const unsigned* chunk_off,
                              unsigned* out_words, b200lic_stream_t stream);
B200LIC_API int b200lic_rans_decode(const unsigned* words, const unsigned* chunk_off, unsigned n, unsigned chunk,
                        const int* indexes, const int* cdf, const int* cdf_len, const int* offset, int cdf_stride,
                        int* symbols, b200lic_stream_t stream);
/* Where a forward workspace expects the staged activation operand ([N,H,W,cpad] bf16 hi and lo; x*x for in_square),
 * and where a weight-gradient workspace (B200LIC_OP_CONV_WGRAD / _DECONV_WGRAD) expects the staged dY
 * ([N,Ho,Wo,cpad]).  NULL / 0 when the shape stages differently. */
B200LIC_API int b200lic_conv_x_slot(const b200lic_conv_desc* d, int op, void* workspace, size_t workspace_bytes,
                        void** x_hi, void** x_lo, int* cpad);
B200LIC_API int b200lic_conv_dy_slot(const b200lic_conv_desc* d, int op, void* workspace, size_t workspace_bytes,
                         void** dy_hi, void** dy_lo, int* cpad);
/* b200lic_gather_mix_sched whose result leaves as the staged activation operand (split-bf16 NHWC, channels zero-padded
 * to cpad; x*x when square != 0) instead of fp32 NCHW; `out` (may be NULL) additionally receives the fp32 batch.  Same
 * picks and QDrop draws as b200lic_gather_mix_sched.  sched == NULL: identity rows, seed_base as the seed. */
B200LIC_API int b200lic_stage_mix_sched(const float* q, const float* fp, const long long* idx_table, int table_rows, int rows,
                            int C, int HW, float prob, unsigned long long seed_base, int units, int unit,
                            const b200lic_calib_sched* sched, int square, void* x_hi, void* x_lo, int cpad, float* out,
                            b200lic_stream_t stream);
/* b200lic_lp_loss_fwd_bwd_sched whose gradient leaves as the staged dY operand of the weight-gradient GEMM; d_pred (may
 * be NULL) additionally receives it as fp32 NCHW.  idx_table == NULL: tgt_cache is the target batch itself.  `act` is
 * the activation fused into the layer that produced `pred`: its derivative (b200lic_act_bwd, on the activation output)
 * is applied to the gradient, which is then the one the weight gradient needs.  For a GDN unit (f_gdn, TO quant_layer.py:142-154)
 * pass its input gdn_x and the pre-(r)sqrt accumulator gdn_norm (both [rows,C,HW]; NULL otherwise): the gradient is then
 * carried on to the accumulator, d_norm = dy * x * d(norm^-+1/2)/dnorm (b200lic_gdn_bwd_prep), the dY operand of gamma's
 * weight gradient.  With gdn_x given, pred may be NULL: the unit's output x * norm^-+1/2 is then recomputed on the fly (the
 * forward only has to produce the accumulator: a plain 1x1 convolution of x*x with bias beta). */
B200LIC_API int b200lic_lp_loss_stage_sched(const float* pred, const float* tgt_cache, const long long* idx_table,
                                int table_rows, int rows, int C, int HW, int units, int unit,
                                const b200lic_calib_sched* sched, float p, float scale, float grad_scale, int act,
                                float act_slope, const float* gdn_x, const float* gdn_norm, int gdn_inverse,
                                float* loss, void* dy_hi, void* dy_lo, int cpad, float* d_pred,
                                b200lic_stream_t stream);
/* b200lic_conv_wgrad_staged for an x operand staged with channel pitch x_cpad (the cpad b200lic_conv_x_slot reports:
 * any multiple of 32 >= Cin); dy == NULL: dY is already staged in the slot b200lic_conv_dy_slot reports. */
B200LIC_API int b200lic_conv_wgrad_prepared(const b200lic_conv_desc* d, int transposed, const void* x_hi, const void* x_lo,
                                int x_cpad, const float* dy, float* dw, void* workspace, size_t workspace_bytes,
                                b200lic_stream_t stream);
/* The weight gradient with its tail fused: the per-split partial sums are reduced (fixed order) straight into
 * b200lic_adaround_bwd_adam_sched's arithmetic -- dW never exists in memory unless dw_out (may be NULL) asks for it.
 * Bit-identical to b200lic_conv_wgrad_staged followed by b200lic_adaround_bwd_adam_sched. */
B200LIC_API int b200lic_conv_wgrad_adam_sched(const b200lic_conv_desc* d, int transposed, const void* x_hi, const void* x_lo,
                                  int x_cpad, const float* dy, void* workspace, size_t workspace_bytes, const float* w, float* alpha,
                                  const float* delta, const float* zero_point, float* exp_avg, float* exp_avg_sq,
                                  int outer, int ch, int inner, int n_levels, const b200lic_calib_sched* sched,
                                  float beta1, float beta2, float eps, float grad_scale, float reg_weight,
                                  float* reg_loss, float* dw_out, b200lic_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU tail of the AdaRound iteration over NVLink peer memory (the only exchange step of the path is
 * the sum of dL/dWq over the data-parallel ranks).  Replaces ncclAllReduce(dL/dWq) + b200lic_adaround_bwd_adam_sched on
 * every rank: rank r sums the ranks' gradients of ITS shard [shard_lo, shard_hi) out of their memory (peer loads, fixed
 * rank order), applies STE masks / regulariser / Adam with its shard of the moments, and stores the new alpha into every
 * rank's alpha buffer (peer stores).  grad_ptrs / alpha_ptrs / flag_ptrs are DEVICE arrays of `world` peer pointers into
 * symmetric allocations the host set up (CUDA IPC / torch symmetric memory; the library maps nothing itself): each rank's
 * local gradient, alpha, and a zero-initialised block of 2*world 32-bit flags; `state` = two zero-initialised local
 * words.  Every rank must launch the call in the same order.  Shards must be 4-element aligned.  alpha is bit-identical
 * on all ranks by construction; the kernel completes on a rank only after every rank's alpha stores have landed there. */
B200LIC_API int b200lic_xgpu_reduce_adam_sched(const float* const* grad_ptrs, float* const* alpha_ptrs,
                                   unsigned* const* flag_ptrs, unsigned* state, int rank, int world, size_t shard_lo,
                                   size_t shard_hi, const float* w, const float* delta, const float* zero_point,
                                   float* exp_avg, float* exp_avg_sq, int outer, int ch, int inner, int n_levels,
                                   const b200lic_calib_sched* sched, float beta1, float beta2, float eps,
                                   float grad_scale, float reg_weight, float* reg_loss, int exit_barrier, b200lic_stream_t stream);

/* out = a * sigmoid(b) + c   (AttentionBlock tail) */
B200LIC_API int b200lic_attn_gate(const float* a, const float* b, const float* c, size_t n, float* out,
                      b200lic_stream_t stream);
/* |x| (ScaleHyperprior h_a input) */
B200LIC_API int b200lic_abs(const float* x, size_t n, float* out, b200lic_stream_t stream);
/* PixelShuffle(r) with optional fused activation: in [N, C*r*r, H, W] -> out [N, C, H*r, W*r] */
B200LIC_API int b200lic_pixel_shuffle(const float* x, int N, int C, int H, int W, int r, int act, float slope, float* out,
                          b200lic_stream_t stream);
B200LIC_API int b200lic_pixel_unshuffle(const float* dy, int N, int C, int H, int W, int r, float* dx,
                            b200lic_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200LIC_H_ */
