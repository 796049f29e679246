/*
 * fp8fq.h -- C ABI of libfp8fq.so, the B200 (sm_100a) FP8 fake-quantization engine.
 *
 * This is the drop-in boundary for the hot path of Qualcomm-AI-research/FP8-quantization
 * (reference paths below are relative to the reference checkout).  The reference has no
 * native code at all; what a maintainer would bind is listed per entry point as the Python
 * function the call replaces.  INTEGRATION.md shows the ctypes stubs.
 *
 * Conventions
 *   - All `const float*` / `float*` arguments are DEVICE pointers owned by the caller unless the
 *     name ends in `_host`.  The library never allocates device memory, keeps no global state
 *     (other than the monotonic launch counter read by fp8fq_launch_count) and is re-entrant; every call is asynchronous on `stream` (a cudaStream_t passed as
 *     void*) and CUDA-graph capturable (no host synchronisation, no D2H reads) -- except the
 *     `*_host_*` entry points, which own their staging buffers and synchronise before returning.
 *   - Return value: 0 = ok; > 0 = a cudaError_t; < 0 = FP8FQ_ERR_* argument error.  Never throws.
 *   - Tensors are contiguous fp32.  "per-channel" means channel = dim 0 (reference
 *     fp8_quantizer.py:108-109), i.e. a [C, inner] row-major view.
 *   - The FP format is given at run time exactly as the reference carries it: `n_bits`,
 *     `sign_bits` and the float `mantissa_bits`; M = clamp(round_half_even(mantissa_bits), 1,
 *     n_bits - sign_bits), E = n_bits - sign_bits - M (fp8_quantizer.py:105-106).
 *     Supported: E <= 7 (at most 127 exponent codes), M <= 12.
 *
 * Quantisation parameters ("table")
 *   Everything the reference derives per channel from (maxval, M, E) -- `bias`
 *   (fp8_quantizer.py:110), the clamp bounds (:112-113) and, per exponent code e in
 *   [1, max(1, 2^E - 1)], the scale 2^(e - M - bias) (:130) and the |x| threshold at which
 *   floor(log2|x| + bias) steps to e (:128) -- is computed ONCE by fp8fq_prepare_f32 into a
 *   caller-owned device buffer of fp8fq_table_floats(...) floats and then consumed by the
 *   streaming kernels.  Layout (floats, per channel, stride = fp8fq_table_stride(...)):
 *     [0] maxval  [1] minval  [2] bucket base (int bits)  [3] bias
 *     [4] flags (int bits): bit 0 irregular thresholds, 1 all scales exact powers of two, 2 an unusable reciprocal,
 *         3 scales exact doublings of each other, 4 qualifies for the scaled-domain element path, 5 ... with two scale groups;
 *         bits 8..27 mantissa
 *         band of the exponent-arithmetic look-up (0xfffff = every mantissa); bits 28..31 M
 *     [5] K (int bits, low 8) | last code of the first scale group of a two-group table << 8 (0: one group)
 *     [6] tie guard of the reciprocal multiply  [7] reference point of the exponent-arithmetic look-up (int bits)
 *     [8 .. 8+KP)            thr[j], j = 0..K      (KP = K+1 rounded up to even)
 *     [8+KP .. 8+KP+2(K+1))  (scale, 1/scale)[e'], e' = 0..K
 *     [8+KP+2(K+1) .. +4)    constants of the scaled-domain element path: switching point between the two scale groups
 *                            (NaN: one group), normalised scale of the second group, tie guard, unused
 */
#ifndef FP8FQ_H
#define FP8FQ_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FP8FQ_OK 0
#define FP8FQ_ERR_BAD_ARG (-1)       /* null pointer, negative size, n != C*inner ... */
#define FP8FQ_ERR_UNSUPPORTED (-2)   /* E > 7 or M > 12 */
#define FP8FQ_ERR_ALIGNMENT (-3)     /* pointer not 4-byte aligned */
#define FP8FQ_ERR_WORKSPACE (-4)     /* workspace too small */

/* activation applied before quantisation by the fused entry points
 * (quantized_folded_bn.py:50-51, models/resnet_quantized.py:44) */
#define FP8FQ_ACT_NONE 0
#define FP8FQ_ACT_RELU 1
#define FP8FQ_ACT_RELU6 2

/* range-estimator update rule applied to (current_xmin, current_xmax) */
#define FP8FQ_EST_CURRENT 0 /* CurrentMinMaxEstimator, range_estimators.py:61-76  (overwrite) */
#define FP8FQ_EST_ALL 1     /* AllMinMaxEstimator,     range_estimators.py:83-100 (running min/max) */
#define FP8FQ_EST_RUNNING 2 /* RunningMinMaxEstimator, range_estimators.py:108-125 (EMA) */
#define FP8FQ_EST_DP_STATS 3 /* no update rule: cur_min [C] receives -min, cur_max [2C] max and, in its second half, a
                              * NaN flag (1.0 / 0.0; a NaN statistic itself travels as -inf) of THIS batch -- with
                              * cur_max = cur_min + C that is one packed [3C] buffer which one MAX all-reduce merges
                              * across data-parallel ranks; followed by fp8fq_dp_finish_prepare_f32 */

/* library / build introspection */
int fp8fq_version(void);
const char* fp8fq_build_info(void);

/* Host helpers (no CUDA): format split and table sizing. */
int fp8fq_format_split(float mantissa_bits, int n_bits, int sign_bits, int* M, int* E, int* K);
int64_t fp8fq_table_stride(float mantissa_bits, int n_bits, int sign_bits);
int64_t fp8fq_table_floats(float mantissa_bits, int n_bits, int sign_bits, int64_t C);

/* Replaces: the (M, E, bias) prologue of quantize_to_fp8_ste_MM (fp8_quantizer.py:105-113) plus
 * the per-code evaluation of :128 and :130.  `maxval` is [C] on the device. */
int fp8fq_prepare_f32(const float* maxval, int64_t C, float mantissa_bits, int n_bits, int sign_bits,
                      float* table, void* stream);

/* Replaces: FPQuantizer.set_quant_range (fp8_quantizer.py:222-240) followed by the prologue above:
 * maxval[c] = | max(|xmin[c]|, xmax[c]) |, written to `maxval_out` ([C]) and turned into `table`. */
int fp8fq_set_range_prepare_f32(const float* xmin, const float* xmax, int64_t C, float* maxval_out,
                                float mantissa_bits, int n_bits, int sign_bits, float* table,
                                void* stream);

/* Replaces: FPQuantizer.forward / quantize_to_fp8_ste_MM (fp8_quantizer.py:91-133,194-205).
 * x, y: [n] = [C, inner]; C == 1 is the per-tensor case.  y may alias x. */
int fp8fq_fake_quant_f32(const float* x, float* y, const float* table, int64_t n, int64_t C,
                         int64_t inner, float mantissa_bits, int n_bits, int sign_bits, void* stream);

/* Several per-channel tensors of the same format in ONE launch (a model's weight tensors are each too small to
 * fill the GPU): replaces the per-layer QuantizationHijacker.quantize_weights calls (hijacker.py:88-98) of one
 * forward.  `descs_host` is a HOST array; it is copied into the kernel parameters (no device allocation). */
typedef struct {
  const float* x;     /* [C, inner] device */
  float* y;           /* [C, inner] device (may alias x) */
  const float* table; /* C channel tables */
  int64_t C, inner;
} fp8fq_tensor_desc;
int fp8fq_fake_quant_multi_f32(const fp8fq_tensor_desc* descs_host, int count, float mantissa_bits, int n_bits,
                               int sign_bits, void* stream);

/* Same, additionally writing the reference's intermediate integers for parity tests:
 * codes[i] = (sign << 31) | (e << 16) | q, with e = log_scales (:128) and q = |round(xc/scales)|
 * (:132); NaN -> 0x7fffffff. */
int fp8fq_fake_quant_codes_f32(const float* x, float* y, int32_t* codes, const float* table, int64_t n,
                               int64_t C, int64_t inner, float mantissa_bits, int n_bits,
                               int sign_bits, void* stream);

/* Replaces: BNFusedHijacker.forward's epilogue (quantized_folded_bn.py:39-55): eval-mode
 * F.batch_norm -> activation -> per-tensor activation quantiser, one pass, 8 B/element.
 * x, y: [rows, hw] with channel(row) = row % Cbn (NCHW contiguous: rows = N*Cbn, hw = H*W).
 * bn_scale/bn_shift: [Cbn] from fp8fq_bn_fold_f32.  `table` is a per-tensor (C == 1) table.
 * bn_mode 0 ("affine"): bn_scale/bn_shift from fp8fq_bn_fold_f32, y = fma(x, scale, shift).
 * bn_mode 1 ("exact") : bn_scale = the packed [4 * Cbn] buffer from fp8fq_bn_pack_f32, bn_shift ignored;
 *                       y = fma(gamma * (x - mean), rsqrtf(var + eps), beta), the arithmetic of ATen's eval-mode batch
 *                       norm on CUDA -- bit-identical to F.batch_norm on the same GPU (measured, profiles/). */
int fp8fq_bn_act_quant_f32(const float* x, float* y, const float* bn_scale, const float* bn_shift,
                           int64_t rows, int64_t hw, int64_t Cbn, int act, int bn_mode, const float* table,
                           float mantissa_bits, int n_bits, int sign_bits, void* stream);

/* Replaces: the whole tail of QuantizedBlock.forward (models/resnet_quantized.py:39-46): the last BNFusedHijacker's
 * epilogue (BN -> inner activation quantiser, quantized_folded_bn.py:39-55) followed by
 * `out += residual; relu; quantize_activations(out)`:  y = Q_outer(act(Q_inner(bn(x)) + residual)), one pass,
 * 12 B/element instead of 8 + 12.  Returns FP8FQ_ERR_UNSUPPORTED for shapes the fused variant does not cover
 * (the caller then issues fp8fq_bn_act_quant_f32 + fp8fq_add_act_quant_f32). */
int fp8fq_bn_quant_add_act_quant_f32(const float* x, const float* residual, float* y, const float* bn_scale,
                                     const float* bn_shift, int64_t rows, int64_t hw, int64_t Cbn, int act,
                                     int bn_mode, const float* table_inner, float mantissa_bits_inner,
                                     int n_bits_inner, int sign_bits_inner, const float* table_outer,
                                     float mantissa_bits_outer, int n_bits_outer, int sign_bits_outer,
                                     void* stream);

/* Channel-innermost twins of the two fused epilogues above: x, residual, y are [pixels, Cbn] with
 * channel(i) = i % Cbn -- channels_last ([N, H, W, C] in memory) activations, which is the layout cuDNN's
 * tensor-core convolutions produce natively (in NCHW it brackets every convolution with two transpose kernels), and
 * the [N, C] outputs of Linear layers (quantized_folded_bn.py:39-55 for BNQLinear).  Same arithmetic, same
 * parameters (bn_scale/bn_shift or the packed buffer), same results element for element as the NCHW entry points
 * on the permuted tensor; 128-bit accesses need Cbn % 4 == 0 and 16-byte aligned pointers (scalar path otherwise). */
int fp8fq_bn_act_quant_nhwc_f32(const float* x, float* y, const float* bn_scale, const float* bn_shift,
                                int64_t pixels, int64_t Cbn, int act, int bn_mode, const float* table,
                                float mantissa_bits, int n_bits, int sign_bits, void* stream);
int fp8fq_bn_quant_add_act_quant_nhwc_f32(const float* x, const float* residual, float* y, const float* bn_scale,
                                          const float* bn_shift, int64_t pixels, int64_t Cbn, int act, int bn_mode,
                                          const float* table_inner, float mantissa_bits_inner, int n_bits_inner,
                                          int sign_bits_inner, const float* table_outer, float mantissa_bits_outer,
                                          int n_bits_outer, int sign_bits_outer, void* stream);

/* Replaces: the per-channel parameter arithmetic of the eval-mode F.batch_norm call in BNFusedHijacker.forward
 * (quantized_folded_bn.py:39-48), done once per change of the BN tensors instead of once per element.
 * Packed batch-norm parameters for bn_mode 1: [Cbn][4] = {mean, gamma (1 if NULL), rsqrtf(var + eps), beta (0 if NULL)}
 * per channel; `packed` must be 16-byte aligned. */
int fp8fq_bn_pack_f32(const float* mean, const float* var, const float* gamma, const float* beta, float eps,
                      int64_t Cbn, float* packed, void* stream);

/* Same call site (quantized_folded_bn.py:39-48), bn_mode 0: per-channel affine form of eval-mode batch norm,
 * scale = gamma * rsqrt(var + eps) (as 1/sqrt), shift = beta - mean * scale.  All [Cbn]. */
int fp8fq_bn_fold_f32(const float* mean, const float* var, const float* gamma, const float* beta,
                      float eps, int64_t Cbn, float* bn_scale, float* bn_shift, void* stream);

/* Replaces: QuantizedBlock.forward's tail (models/resnet_quantized.py:43-46) and
 * QuantizedInvertedResidual.forward (models/mobilenet_v2_quantized.py:22-24):
 * y = Q(act(a + b)), one pass, 12 B/element. */
int fp8fq_add_act_quant_f32(const float* a, const float* b, float* y, int64_t n, int act,
                            const float* table, float mantissa_bits, int n_bits, int sign_bits,
                            void* stream);

/* STE backward of the fake-quantiser: what autograd computes through quantize_to_fp8_ste_MM
 * (fp8_quantizer.py:91-133; round_ste_func, detached exponent code) for learnable ranges (:242-254).
 * grad_x[i] = ((grad_y[i] * s) / s) * (1 inside the clipping range, 0 outside, 1/2 on exact ties), s the element's
 * scale and the two roundings those of autograd's mul / div backward (the identity whenever s is a power of two);
 * a NaN x[i] gives NaN in grad_x[i] and in both sums, as in the reference;
 * acc[2c] = sum_i grad_y * d xc/d maxval (clipping term), acc[2c+1] = sum_i grad_y * (q - xc/s) * s (scale term), in
 * double; the caller finishes  grad_maxval[c] = acc[2c] + acc[2c+1] / maxval[c]  and
 * grad_mantissa_bits = ln2 * (-1 - dbias/dM) * sum_c acc[2c+1].  `acc` ([2*C] doubles) is zeroed by the call. */
int fp8fq_fake_quant_backward_f32(const float* grad_y, const float* x, float* grad_x, const float* table, int64_t n,
                                  int64_t C, int64_t inner, float mantissa_bits, int n_bits, int sign_bits,
                                  double* acc, void* stream);

/* INT uniform quantisers -- the reference's comparison baseline.
 * fp8fq_uniform_prepare_f32 replaces Asymmetric/SymmetricUniformQuantizer.set_quant_range
 * (quantization/quantizers/uniform_quantizers.py:224-246, 303-314): from (x_min [C], x_max [C]) it writes delta [C],
 * zero_float [C] (asymmetric), the `signed` flag (symmetric; 1.0 / 0.0) and the channel tables
 * (fp8fq_uniform_table_floats(C) floats).  fp8fq_uniform_quant_f32 replaces .forward (uniform_quantizers.py:107-164):
 * y = scale * (clamp(round(x / scale) + zero_point, int_min, int_max) - zero_point), scale = clamp(delta, min=eps).
 * Only scale_domain="linear" with the round-to-nearest-even discretizer.
 * aten_cuda_scalar_div: the reference computes delta = range / int_max with a Python scalar divisor; ATen performs a
 * true division on the CPU and a multiplication by the fp32 reciprocal on CUDA (1 ulp apart for some ranges).
 * 0 reproduces the reference run on the CPU bit for bit, 1 the reference run on the GPU. */
int64_t fp8fq_uniform_table_floats(int64_t C);
int fp8fq_uniform_prepare_f32(const float* xmin, const float* xmax, int64_t C, int n_bits, int symmetric,
                              int aten_cuda_scalar_div, float eps, float* delta_out, float* zero_float_out,
                              float* signed_out, float* table, void* stream);
int fp8fq_uniform_quant_f32(const float* x, float* y, const float* table, int64_t n, int64_t C, int64_t inner,
                            void* stream);

/* Data format on the caller's side of the path (no counterpart in the reference, which hands the NCHW image to
 * F.conv2d, hijacker.py:70-98 -> autoquant_utils.py:34-44): 2x2 space-to-depth of an NCHW image x [N, C <= 4, H, W] into
 * channel-innermost y [N, Hs, Ws, 16] with the convolution's zero padding applied,
 *   y[n, Y, X, c*4 + p*2 + q] = xpad[n, c, 2Y + p, 2X + q],  xpad[r, t] = x[r - pad, t - pad] or 0,  channels >= 4C zero.
 * A stride-2 k x k convolution (k odd, padding k/2) over x equals the stride-1 (k+1)/2 x (k+1)/2 convolution of y with
 * the correspondingly re-indexed (zero-extended to k+1) weights: the same products, summed in a different order.
 * Hs >= (H + 2 pad + 1) / 2 rows are written (rows / columns beyond the padded image are zero); y 16-byte aligned. */
int fp8fq_space_to_depth2_nhwc_f32(const float* x, float* y, int64_t N, int64_t C, int64_t H, int64_t W, int64_t pad,
                                   int64_t Hs, int64_t Ws, void* stream);

/* Data format on the caller's side of the path: the ToTensor + Normalize steps of the reference's input pipeline
 * (utils/imagenet_dataloaders.py:66-81: x / 255, then (x - mean[c]) / std[c]) for a uint8 NCHW image batch
 * x [N, C <= 8, HW] -> y fp32 [N, C, HW], so that images cross the host link as 1 byte per pixel.  lut [C, 256] (device)
 * holds the result for every (channel, byte value), computed by the caller with the reference's own fp32 operations:
 * y[n, c, p] = lut[c, x[n, c, p]] is then bit-identical to torchvision's output.  128-bit accesses when HW % 16 == 0
 * and x, y are 16-byte aligned (scalar path otherwise). */
int fp8fq_u8_normalize_nchw_f32(const uint8_t* x, const float* lut, float* y, int64_t N, int64_t C, int64_t HW,
                                void* stream);

/* Replaces (for channel-innermost activations): the nn.MaxPool2d that consumes the stem's quantised activation
 * (models/resnet_quantized.py:73-78 keeps torchvision's module; ATen: max_pool_forward_nhwc).  x [N, H, W, C] ->
 * y [N, Ho, Wo, C], Ho = (H + 2 ph - kh) / sh + 1 (floor mode, dilation 1, no indices); C % 4 == 0, 16-byte aligned.
 * ATen's selection rule (`val > max || isnan(val)` in row-major window order): NaN propagates, results are the same
 * bits as F.max_pool2d. */
int fp8fq_max_pool2d_nhwc_f32(const float* x, float* y, int64_t N, int64_t H, int64_t W, int64_t C, int kh, int kw,
                              int sh, int sw, int ph, int pw, void* stream);

/* Replaces: the min/max of every range estimator (range_estimators.py:73-74, 85-91, 110-116) and
 * its update rule, NaN-propagating like torch.min/max.  One pass over x (4 B/element).
 *   per-tensor (C == 1): two-stage reduce finished by the last CTA; `workspace` must hold
 *   fp8fq_minmax_workspace_bytes() bytes, zero-initialised once (the kernel leaves it zeroed).
 *   cur_min/cur_max: [C] estimator state, updated in place according to `est_mode`;
 *   `initialized` != 0 means the state already holds a previous estimate.
 *   `momentum` is a double because the reference's EMA weights are Python floats:
 *   (float)(1.0 - momentum) and (float)momentum (range_estimators.py:121-123). */
int64_t fp8fq_minmax_workspace_bytes(void);
int fp8fq_minmax_f32(const float* x, int64_t n, int64_t C, int64_t inner, float* cur_min, float* cur_max,
                     int est_mode, int initialized, double momentum, void* workspace, void* stream);

/* Calibration fast path = estimator + set_quant_range + prologue in one launch chain, all on device:
 * QuantizationManager.forward in state estimate_ranges (quantization_manager.py:114-122) up to, not
 * including, the quantiser call.  Equivalent to fp8fq_minmax_f32 + fp8fq_set_range_prepare_f32. */
int fp8fq_estimate_prepare_f32(const float* x, int64_t n, int64_t C, int64_t inner, float* cur_min,
                               float* cur_max, int est_mode, int initialized, double momentum,
                               float* maxval_out, float mantissa_bits, int n_bits, int sign_bits,
                               float* table, void* workspace, void* stream);

/* Calibration epilogue of a BN-fused layer: BNFusedHijacker.forward in state estimate_ranges
 * (quantized_folded_bn.py:39-55 -> quantization_manager.py:114-122): the per-tensor estimator statistics of
 * act(bn(x)) computed WITHOUT materialising it (the convolution output is read once, 4 B/element; the op-by-op
 * path moves 20 B/element before the quantiser starts), then the estimator update, and -- when `table` is not NULL --
 * set_quant_range + the quantiser table, exactly as fp8fq_estimate_prepare_f32.  The caller follows it with
 * fp8fq_bn_act_quant(_nhwc)_f32 on the same x.  x: NCHW ([outer = N*Cbn rows, hw]) or, with nhwc != 0,
 * channel-innermost ([outer = pixels, Cbn]; hw ignored).  bn_scale / bn_shift / bn_mode / act as in
 * fp8fq_bn_act_quant_f32.  Only 128-bit-addressable shapes (16-byte aligned x; hw % 4 == 0 resp. Cbn % 4 == 0; at most
 * one channel wrap per 4096-element tile): FP8FQ_ERR_UNSUPPORTED otherwise, and the caller composes the unfused ops.
 * maxval_out may be NULL (statistics only, e.g. before a data-parallel all-reduce). */
int fp8fq_bn_act_estimate_prepare_f32(const float* x, int64_t outer, int64_t hw, int64_t Cbn, int nhwc,
                                      const float* bn_scale, const float* bn_shift, int bn_mode, int act,
                                      float* cur_min, float* cur_max, int est_mode, int initialized, double momentum,
                                      float* maxval_out, float mantissa_bits, int n_bits, int sign_bits, float* table,
                                      void* workspace, void* stream);

/* Data-parallel calibration, second half (reference dependency order quantization_manager.py:114-122):
 * packed = [-min (C) | max (C) | NaN flag (C)] of the global batch, i.e. the result of a MAX all-reduce over every rank's
 * statistics (fp8fq_minmax_f32 / fp8fq_bn_act_estimate_prepare_f32 with est_mode FP8FQ_EST_DP_STATS writing into one
 * [3C] buffer; a set flag makes the range NaN, as torch.min / torch.max of the concatenated batch would be).
 * One launch applies the estimator's update rule to (cur_min, cur_max) -- est_mode CURRENT / ALL / RUNNING as in
 * fp8fq_minmax_f32 --, FPQuantizer.set_quant_range (fp8_quantizer.py:236-237) into maxval_out [C] (may be NULL) and,
 * when table is not NULL, the quantiser table.  Every rank ends with the range a single process would have computed on
 * the concatenated batch, bit for bit (min / max are order independent). */
int fp8fq_dp_finish_prepare_f32(const float* packed, int64_t C, float* cur_min, float* cur_max, int est_mode,
                                int initialized, double momentum, float* maxval_out, float mantissa_bits, int n_bits,
                                int sign_bits, float* table, void* stream);

/* Data-parallel calibration step as ONE launch, the exchange done by the kernel over NVLink peer memory (no collective
 * call; replaces the estimator -> all-reduce -> set_quant_range sequence of quantization_manager.py:114-122 with
 * range_estimators.py:73-98 under data parallelism): the per-tensor variants of fp8fq_estimate_prepare_f32 / fp8fq_bn_act_estimate_prepare_f32 whose last CTA, after
 * the local reduction, stores this rank's (-min, max, NaN flag) into every peer's exchange buffer (64-bit system-scope
 * stores, value + epoch in one word), waits until its own buffer holds every peer's words of this epoch, takes the MAX
 * and continues with the estimator rule, set_quant_range and the table -- every rank ends with the range of the
 * concatenated batch, bit for bit, as with fp8fq_dp_finish_prepare_f32 after an all-reduce.
 *   peer_bufs : DEVICE array [world] of pointers to each rank's exchange buffer (this rank's own included), each of
 *               fp8fq_dp_exchange_words(world) 64-bit words, peer-accessible (symmetric / IPC memory), zeroed once
 *               before the first call;  epoch : the same on every rank for the same call, >= 1, increasing by one per
 *               call (all ranks must issue the same sequence of calls);  world <= 16.
 * A peer that does not arrive within ~3 s makes the range NaN instead of hanging the device. */
int64_t fp8fq_dp_exchange_words(int world);
int fp8fq_estimate_prepare_p2p_f32(const float* x, int64_t n, float* cur_min, float* cur_max, int est_mode, int initialized,
                                   double momentum, float* maxval_out, float mantissa_bits, int n_bits, int sign_bits,
                                   float* table, void* workspace, const void* peer_bufs, int rank, int world,
                                   unsigned int epoch, void* stream);
int fp8fq_bn_act_estimate_prepare_p2p_f32(const float* x, int64_t outer, int64_t hw, int64_t Cbn, int nhwc,
                                          const float* bn_scale, const float* bn_shift, int bn_mode, int act,
                                          float* cur_min, float* cur_max, int est_mode, int initialized, double momentum,
                                          float* maxval_out, float mantissa_bits, int n_bits, int sign_bits, float* table,
                                          void* workspace, const void* peer_bufs, int rank, int world, unsigned int epoch,
                                          void* stream);

/* Replaces: the double Python loop of FP_MSE_Estimator.forward (range_estimators.py:337-347):
 * mses[m, g, c] += mean over the non-channel elements of (x - Q(x; maxval = grid[g, c], M = mbits[m]))^2.
 * grid: [G, C] device; mbits_host: [Mn] HOST floats; mses: [Mn, G, C] device, accumulated.
 * tables: caller-owned device scratch of fp8fq_mse_table_floats(...) floats. */
int64_t fp8fq_mse_table_floats(const float* mbits_host, int Mn, int n_bits, int sign_bits, int64_t G,
                               int64_t C);
int fp8fq_mse_grid_f32(const float* x, int64_t n, int64_t C, int64_t inner, const float* grid, int64_t G,
                       const float* mbits_host, int Mn, int n_bits, int sign_bits, float* mses,
                       float* tables, void* stream);

/* End-to-end entry point with HOST buffers (what a caller without device memory binds):
 * chunked H2D -> fake-quant -> D2H pipeline over internal pinned staging and `nstreams` streams.
 * maxval_host: [C].  Synchronises before returning.  The one entry point with state of its own: per device, staging
 * buffers and streams are created on first use and kept (calls for the same device are serialised by a mutex, calls
 * for different devices run concurrently -- a single process may drive several GPUs); the caller's current device is
 * restored before returning.  Per-channel rows longer than the internal chunk are processed in pieces. */
int fp8fq_fake_quant_host_f32(const float* x_host, float* y_host, const float* maxval_host, int64_t n,
                              int64_t C, int64_t inner, float mantissa_bits, int n_bits, int sign_bits,
                              int device);

/* Number of kernel launches issued by this library since load (for bench.py's gpu_launches). */
int64_t fp8fq_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* FP8FQ_H */
