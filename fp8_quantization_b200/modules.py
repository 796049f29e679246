"""Quantised-module layer: the reference's L2 API re-stated on top of the fused kernels.

Mirrors (same class names, constructor kwargs and forward semantics):
  QuantizedModule / QuantizedActivation / FP32Acts   quantization/base_quantized_classes.py:40-181
  QuantizationHijacker                                quantization/hijacker.py:32-112
  BNFusedHijacker                                     quantization/quantized_folded_bn.py:12-68
  QuantConv / QuantLinear / BNQConv / ...             quantization/autoquant_utils.py:20-122
  QuantizedActivationWrapper, Flattener               quantization/autoquant_utils.py:125-177
  fold_bn / quantize_sequential / quantize_model      quantization/autoquant_utils.py:266-381
  QuantizedModel                                      quantization/base_quantized_model.py:19-135

What is different underneath: with ranges fixed (the validation path) every quantisation site is a
single launch -- weight fake-quant per layer, and ``F.batch_norm -> ReLU/ReLU6 -> fake-quant``
(reference: 1 + 1 + 13 full passes over the activation) collapses into ``bn_act_quant`` (one read,
one write); the residual tail ``out += residual; relu; quant`` collapses into ``add_act_quant``.
The convolution / matmul itself stays a cuDNN / cuBLAS library call, as in the reference.
"""
from __future__ import annotations

import copy
import warnings

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn.modules.conv import _ConvNd
from torch.nn.modules.pooling import _AdaptiveAvgPoolNd, _AvgPoolNd

from . import dist as fq_dist
from . import ops
from .quantization_manager import QuantizationManager
from .quantizers import AsymmetricUniformQuantizer, FPQuantizer, QuantizerBase
from .range_estimators import CurrentMinMaxEstimator, RangeEstimatorBase, RunningMinMaxEstimator

# hijacker.py:15-29 minus the timm classes (timm is not a dependency of this package)
activations_set = [nn.ReLU, nn.ReLU6, nn.Hardtanh, nn.Sigmoid, nn.Tanh, nn.GELU, nn.PReLU, nn.SiLU, nn.Hardswish,
                   nn.Hardsigmoid]

# Global switches (tests flip them to compare against the unfused composition).
FUSE_EPILOGUES = True      # BN + act + quant in one launch; residual add + act + quant in one launch
FUSE_BLOCK_TAIL = True     # BN + quant + residual add + act + quant of a residual block in one launch
BATCH_WEIGHT_QUANT = True  # all per-layer weight fake-quants of a forward in one launch
FUSE_CALIBRATION = True    # estimate_ranges state of a BN-fused layer: statistics of act(bn(x)) without materialising it
STEM_SPACE_TO_DEPTH = True  # channels_last network, NCHW image: stride-2 k x k stem conv over <= 4 channels as a stride-1
                            # conv over the 2x2 space-to-depth image (same sum re-indexed; a shape cuDNN handles well)
NATIVE_MAX_POOL = True     # channels_last nn.MaxPool2d through this library's kernel (same bits as ATen's)
CACHE_QUANTIZED_WEIGHTS = False  # keep each layer's quantised weight until the weight (storage / version counter) or its
                                 # range / format changes, instead of re-quantising it on every forward as the reference
                                 # does (hijacker.py:88-98).  Same bits either way; off by default because a CUDA graph
                                 # captured with it on replays the cached tensors and does not see later weight updates.
BN_EXACT = True            # fused epilogues use ATen-CUDA's eval batch-norm arithmetic bit for bit (bn_mode 1);
                           # False: one-FMA affine form (2 fewer instructions per element, ulp-level differences
                           # from F.batch_norm before quantisation)


def _sync_batch_norm_train(x, running_mean, running_var, gamma, beta, momentum, eps):
    """Training-mode batch norm over the GLOBAL (all-rank) batch: per-channel sum and sum of squares are all-reduced
    (one collective of 2C+1 floats), so that every rank normalises with -- and records -- the statistics a single
    process would compute on the concatenated batch.  Used by BN re-estimation under data parallelism
    (utils/qat_utils.py:45-90 + SURVEY.md section 8f1); running stats follow F.batch_norm's update rule
    (unbiased variance, momentum)."""
    C = x.shape[1]
    dims = [0] + list(range(2, x.dim()))
    n_local = x.numel() // C
    stats = torch.cat([x.sum(dims, dtype=torch.float64), (x.double() * x.double()).sum(dims),
                       torch.tensor([float(n_local)], dtype=torch.float64, device=x.device)])
    fq_dist.all_reduce_sum(stats)
    n = stats[-1]
    mean = stats[:C] / n
    var = (stats[C:2 * C] / n - mean * mean).clamp_min(0.0)
    with torch.no_grad():
        m = 1.0 if momentum is None else momentum
        running_mean.mul_(1 - m).add_(mean.float(), alpha=m)
        running_var.mul_(1 - m).add_((var * (n / (n - 1))).float(), alpha=m)
    view = [1, C] + [1] * (x.dim() - 2)
    inv = torch.rsqrt(var.float() + eps)
    y = (x - mean.float().view(view)) * inv.view(view)
    if gamma is not None:
        y = y * gamma.view(view)
    if beta is not None:
        y = y + beta.view(view)
    return y


def _like(t, ref):
    """``t`` in a layout the kernels address and with the same strides as ``ref`` (the residual of a block may come
    from a different producer than the features: NCHW images into a channels_last network, a pooling layer, ...)."""
    t = ops.dense(t)
    if t.stride() != ref.stride():
        t = t.contiguous(memory_format=torch.channels_last if ops.is_channels_last(ref) and t.dim() == 4
                         else torch.contiguous_format)
    return t


def _weight_for_quantizer(w):
    """The tensor a weight quantiser receives: the live (graph-attached) weight while autograd records and the weight
    wants a gradient -- FPQuantizer then returns its STE node, quantizers._FakeQuantSTE, and the forward-only INT
    quantisers raise instead of silently freezing the layer -- otherwise a detached view."""
    return w if (torch.is_grad_enabled() and w.requires_grad) else w.detach()


def _act_code(act):
    """Kernel activation code for a fusable activation module, or None if it cannot be fused."""
    if act is None:
        return ops.ACT_NONE
    if type(act) is nn.ReLU:
        return ops.ACT_RELU
    if type(act) is nn.ReLU6:
        return ops.ACT_RELU6
    return None


def _fusable_manager(mgr) -> bool:
    """A per-tensor FP activation quantiser with fixed ranges: its table can be consumed by a fused kernel."""
    # (the fused kernels are forward-only: with autograd recording, the op-by-op path keeps the graph intact)
    return (FUSE_EPILOGUES and not torch.is_grad_enabled() and isinstance(mgr, QuantizationManager)
            and isinstance(mgr.quantizer, FPQuantizer) and not mgr.estimating() and not mgr.per_channel
            and mgr.quantizer.maxval.numel() == 1)


# ---- state switches applied with nn.Module.apply (base_quantized_classes.py:16-37) ------------------
def _set_layer_learn_ranges(layer):
    if isinstance(layer, QuantizationManager) and layer.quantizer.is_initialized:
        layer.learn_ranges()


def _set_layer_fix_ranges(layer):
    if isinstance(layer, QuantizationManager) and layer.quantizer.is_initialized:
        layer.fix_ranges()


def _set_layer_estimate_ranges(layer):
    if isinstance(layer, QuantizationManager):
        layer.estimate_ranges()


def _set_layer_estimate_ranges_train(layer):
    if isinstance(layer, QuantizationManager) and layer.quantizer.is_initialized:
        layer.estimate_ranges_train()


class QuantizedModule(nn.Module):
    """base_quantized_classes.py:40-153.  The ``quant_params`` kwargs are the configuration contract
    (utils/click_options.py:490-508); classes are selected by passing them as ``method`` /
    ``act_method`` / ``weight_range_method`` / ``act_range_method``."""

    def __init__(self, *args, method: QuantizerBase = AsymmetricUniformQuantizer, act_method=None,
                 weight_range_method: RangeEstimatorBase = CurrentMinMaxEstimator,
                 act_range_method: RangeEstimatorBase = RunningMinMaxEstimator, n_bits=8, n_bits_act=None,
                 per_channel_weights=False, percentile=None, weight_range_options=None, act_range_options=None,
                 scale_domain="linear", act_quant_kwargs={}, weight_quant_kwargs={}, quantize_input=False,
                 fp8_kwargs=None, **kwargs):
        kwargs.pop("act_quant_dict", None)
        kwargs.pop("quant_setup", None)
        super().__init__(*args, **kwargs)
        self.method = method
        self.act_method = act_method or method
        self.n_bits = n_bits
        self.n_bits_act = n_bits_act or n_bits
        self.per_channel_weights = per_channel_weights
        self.percentile = percentile
        self.weight_range_method = weight_range_method
        self.weight_range_options = weight_range_options if weight_range_options else {}
        self.act_range_method = act_range_method
        self.act_range_options = act_range_options if act_range_options else {}
        self.scale_domain = scale_domain
        self.quantize_input = quantize_input
        self.fp8_kwargs = fp8_kwargs or {}
        self.quant_params = None
        # state-dict compatible flags (base_quantized_classes.py:84-85) + host mirrors so that the
        # forward never reads a device tensor (no sync, CUDA-graph capturable)
        self.register_buffer("_quant_w", torch.BoolTensor([False]))
        self.register_buffer("_quant_a", torch.BoolTensor([False]))
        self._qw = False
        self._qa = False
        self.act_qparams = dict(n_bits=self.n_bits_act, scale_domain=self.scale_domain, **act_quant_kwargs,
                                **self.fp8_kwargs)
        self.weight_qparams = dict(n_bits=self.n_bits, scale_domain=self.scale_domain, **weight_quant_kwargs,
                                   **self.fp8_kwargs)

    def quantized_weights(self):
        self._quant_w = torch.BoolTensor([True])
        self._qw = True

    def full_precision_weights(self):
        self._quant_w = torch.BoolTensor([False])
        self._qw = False

    def quantized_acts(self):
        self._quant_a = torch.BoolTensor([True])
        self._qa = True

    def full_precision_acts(self):
        self._quant_a = torch.BoolTensor([False])
        self._qa = False

    def quantized(self):
        self.quantized_weights()
        self.quantized_acts()

    def full_precision(self):
        self.full_precision_weights()
        self.full_precision_acts()

    def get_quantizer_status(self):
        return dict(quant_a=self._qa, quant_w=self._qw)

    def set_quantizer_status(self, quantizer_status):
        self.quantized_acts() if quantizer_status["quant_a"] else self.full_precision_acts()
        self.quantized_weights() if quantizer_status["quant_w"] else self.full_precision_weights()

    def _load_from_state_dict(self, state_dict, prefix, *a, **k):
        super()._load_from_state_dict(state_dict, prefix, *a, **k)
        self._qw = bool(self._quant_w.item())
        self._qa = bool(self._quant_a.item())

    def learn_ranges(self):
        self.apply(_set_layer_learn_ranges)

    def fix_ranges(self):
        self.apply(_set_layer_fix_ranges)

    def estimate_ranges(self):
        self.apply(_set_layer_estimate_ranges)

    def estimate_ranges_train(self):
        self.apply(_set_layer_estimate_ranges_train)

    def extra_repr(self):
        quant_state = "weight_quant={}, act_quant={}".format(self._qw, self._qa)
        parent_repr = super().extra_repr()
        return "{},\n{}".format(parent_repr, quant_state) if parent_repr else quant_state


class QuantizedActivation(QuantizedModule):
    """base_quantized_classes.py:156-173."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.activation_quantizer = QuantizationManager(qmethod=self.act_method, qparams=self.act_qparams,
                                                        init=self.act_range_method,
                                                        range_estim_params=self.act_range_options)

    def quantize_activations(self, x):
        return self.activation_quantizer(x) if self._qa else x

    def add_act_quantize(self, a, b, act):
        """Q(act(a + b)): the residual tail (models/resnet_quantized.py:43-46,
        models/mobilenet_v2_quantized.py:22-24) as one kernel when the ranges are fixed."""
        code = _act_code(act)
        mgr = self.activation_quantizer
        if self._qa and code is not None and _fusable_manager(mgr) and ops.on_device(a):
            q = mgr.quantizer
            a = ops.dense(a)
            b = _like(b, a)
            table, _ = q.table_for(a)
            return ops.add_act_quant(a, b, code, table, q._mbits_host, q.n_bits, q.sign_bits)
        out = a + b
        if act is not None:
            out = act(out)
        return self.quantize_activations(out)

    def block_tail(self, features, x, residual_fn, act):
        """``Q(act(features(x) + residual))`` for a residual block whose ``features`` end in a BNFusedHijacker
        without activation (models/resnet_quantized.py:39-46).  With fixed ranges the last layer's epilogue and the
        block's add/act/quant run as ONE kernel: Q_outer(act(Q_inner(bn(conv)) + residual)), 12 B/element."""
        last = features[-1] if isinstance(features, nn.Sequential) and len(features) > 0 else None
        residual = residual_fn()
        code = _act_code(act)
        if (FUSE_BLOCK_TAIL and FUSE_EPILOGUES and isinstance(last, BNFusedHijacker) and last.activation_function is None
                and self._qa and code is not None and _fusable_manager(self.activation_quantizer)):
            h = x
            for m in list(features)[:-1]:
                h = m(h)
            res = last.conv_only(h)
            if last._fused_epilogue_ok(res) and ops.on_device(residual) and residual.shape == res.shape:
                qi, qo = last.activation_quantizer.quantizer, self.activation_quantizer.quantizer
                res = ops.dense(res)
                residual = _like(residual, res)
                ti, _ = qi.table_for(res)
                to, _ = qo.table_for(res)
                scale, shift, mode = last.folded_bn()
                out = ops.bn_quant_add_act_quant(res, residual, scale, shift, code, ti,
                                                 (qi._mbits_host, qi.n_bits, qi.sign_bits), to,
                                                 (qo._mbits_host, qo.n_bits, qo.sign_bits), bn_mode=mode)
                if out is not None:
                    return out
            return self.add_act_quantize(last.epilogue(res), residual, act)
        return self.add_act_quantize(features(x), residual, act)

    def forward(self, x):
        return self.quantize_activations(x)


class FP32Acts(nn.Module):
    """base_quantized_classes.py:176-181."""

    def forward(self, x):
        return x

    def reset_ranges(self):
        pass


class QuantizationHijacker(QuantizedModule):
    """hijacker.py:32-112: mixin that quantises the weights and the output (or input) activations of
    the nn.Module it is combined with."""

    def __init__(self, *args, activation: nn.Module = None, **kwargs):
        super().__init__(*args, **kwargs)
        if activation:
            assert isinstance(activation, tuple(activations_set)), str(activation)
        self.activation_function = copy.deepcopy(activation) if activation else None
        self.activation_quantizer = QuantizationManager(qmethod=self.act_method, init=self.act_range_method,
                                                        qparams=self.act_qparams,
                                                        range_estim_params=self.act_range_options)
        # hijacker.py:57-60 always ends up forwarding weight_range_options
        self.weight_quantizer = QuantizationManager(qmethod=self.method, init=self.weight_range_method,
                                                    per_channel=self.per_channel_weights,
                                                    qparams=self.weight_qparams,
                                                    range_estim_params=self.weight_range_options)
        if self.weight_quantizer.range_estimator is not None:   # replicated weights: no collective for their ranges
            self.weight_quantizer.range_estimator.replicated_input = True

    def forward(self, x, offsets=None):
        if self.quantize_input and self._qa:
            x = self.activation_quantizer(x)
        weight, bias = self.get_params()
        res = self.run_forward(x, weight, bias, offsets=offsets)
        if self.activation_function is not None:
            res = self.activation_function(res)
        if not self.quantize_input and self._qa:
            res = self.activation_quantizer(res)
        return res

    def get_params(self):
        weight, bias = self.get_weight_bias()
        if self._qw:
            stash = self.__dict__.pop("_wq_stash", None)  # set by QuantizedModel.prequantize_weights for THIS forward
            # (stamped with the weight's storage and version: a layer skipped by one forward must not hand a stale
            # result to a later one after the weights have changed)
            if stash is not None and stash[0] == (weight.data_ptr(), weight._version):
                weight = stash[1]
            else:
                weight = self.quantize_weights(weight)
        return weight, bias

    def quantize_weights(self, weights):
        # hijacker.py:96 hands the live Parameter to the quantiser so that the STE gradient reaches ``self.weight``;
        # without autograd recording the detached view keeps the forward-only kernels on their fast path
        return self.weight_quantizer(_weight_for_quantizer(weights))

    def get_weight_bias(self):
        bias = self.bias if hasattr(self, "bias") else None
        return self.weight, bias

    def run_forward(self, x, weight, bias, offsets=None):
        raise NotImplementedError()

    def extra_repr(self):
        activation = "input" if self.quantize_input else "output"
        return f"{super().extra_repr()}-{activation}"


class BNFusedHijacker(QuantizationHijacker):
    """quantized_folded_bn.py:12-68: weight layer + batch norm (kept in fp32, not folded into the
    weights) + activation + activation quantiser."""

    def __init__(self, *args, **kwargs):
        kwargs.pop("bias", None)
        momentum = kwargs.pop("momentum", 0.1)
        super().__init__(*args, **kwargs, bias=False)
        bn_dim = self.get_bn_dim()
        self.register_buffer("running_mean", torch.zeros(bn_dim))
        self.register_buffer("running_var", torch.ones(bn_dim))
        self.momentum = momentum
        self.gamma = nn.Parameter(torch.ones(bn_dim))
        self.beta = nn.Parameter(torch.zeros(bn_dim))
        self.epsilon = kwargs.get("eps", 1e-5)
        self.bias = None
        self._bn_key = None
        self._bn_folded = None

    def folded_bn(self):
        """Eval-mode batch norm in the form the fused kernels consume: (params, shift | None, bn_mode).  Recomputed
        (one tiny launch) only when one of the four BN tensors or eps changed -- they are constants of the validate
        pass; BN re-estimation or a state-dict load bumps their version counters / storage and invalidates it."""
        ts = (self.running_mean, self.running_var, self.gamma, self.beta)
        mode = 1 if BN_EXACT else 0
        # keyed on the tensor OBJECTS (kept alive in _bn_src, so neither an id nor a storage address can be recycled
        # by a re-bound tensor, e.g. reestimate_BN_stats' fresh running statistics) and their version counters
        src = self.__dict__.get("_bn_src")
        key = tuple((t.data_ptr(), t._version) for t in ts) + (float(self.epsilon), mode)
        if key != self._bn_key or src is None or any(a is not b for a, b in zip(src, ts)):
            self.__dict__["_bn_src"] = ts
            if mode == 1:
                self._bn_folded = (ops.bn_pack(self.running_mean, self.running_var, self.gamma.detach(),
                                               self.beta.detach(), self.epsilon), None, 1)
            else:
                self._bn_folded = ops.bn_fold(self.running_mean, self.running_var, self.gamma.detach(),
                                              self.beta.detach(), self.epsilon) + (0,)
            self._bn_key = key
        return self._bn_folded

    def _load_from_state_dict(self, state_dict, prefix, *a, **k):
        super()._load_from_state_dict(state_dict, prefix, *a, **k)
        self._bn_key = None   # loaded statistics: re-pack on the next forward

    def _fused_epilogue_ok(self, res) -> bool:
        return (self._qa and not self.quantize_input and not self.training and ops.on_device(res) and res.dim() >= 2
                and res.dtype == torch.float32 and _act_code(self.activation_function) is not None
                and _fusable_manager(self.activation_quantizer))

    def _fused_calibration_ok(self, res) -> bool:
        mgr = self.activation_quantizer
        return (FUSE_EPILOGUES and FUSE_CALIBRATION and self._qa and not self.quantize_input and not self.training
                and not torch.is_grad_enabled() and ops.on_device(res) and res.dim() >= 2 and res.dtype == torch.float32
                and _act_code(self.activation_function) is not None and isinstance(mgr, QuantizationManager)
                and mgr.estimating() and not mgr.per_channel and mgr.device_resident_calibration())

    def conv_only(self, x):
        """Input quantisation (if configured), weight quantisation and the linear operation -- everything of
        forward() up to, not including, the BN / activation / activation-quantiser epilogue."""
        if self.quantize_input and self._qa:
            x = self.activation_quantizer(x)
        weight, bias = self.get_params()
        return self.run_forward(x, weight, bias)

    def epilogue(self, res):
        """quantized_folded_bn.py:39-55: F.batch_norm -> activation -> activation quantiser; ONE launch when the
        ranges are fixed."""
        if self._fused_epilogue_ok(res):
            q = self.activation_quantizer.quantizer
            res = ops.dense(res)
            table, _ = q.table_for(res)
            scale, shift, mode = self.folded_bn()
            out = ops.bn_act_quant(res, scale, shift, _act_code(self.activation_function), table, q._mbits_host,
                                   q.n_bits, q.sign_bits, bn_mode=mode)
            if out is not None:     # None: beyond the fused kernels' 32-bit index range -> op-by-op composition below
                return out
        elif self._fused_calibration_ok(res):
            # calibration: statistics of act(bn(conv)) + range + table in one launch (4 B/element), then the fused
            # epilogue -- instead of F.batch_norm, the activation, the estimator and the quantiser as separate passes
            mgr = self.activation_quantizer
            q = mgr.quantizer
            res = ops.dense(res)
            scale, shift, mode = self.folded_bn()
            code = _act_code(self.activation_function)
            if mgr.range_estimator.fused_bn_estimate_prepare(res, q, scale, shift, mode, code):
                table, _ = q.table_for(res)
                out = ops.bn_act_quant(res, scale, shift, code, table, q._mbits_host, q.n_bits, q.sign_bits,
                                       bn_mode=mode)
                if out is not None:
                    return out
        if self.training and fq_dist.active():
            res = _sync_batch_norm_train(res, self.running_mean, self.running_var, self.gamma, self.beta,
                                         self.momentum, self.epsilon)
        else:
            res = F.batch_norm(res, self.running_mean, self.running_var, self.gamma, self.beta, self.training,
                               self.momentum, self.epsilon)
        if self.activation_function is not None:
            res = self.activation_function(res)
        if not self.quantize_input and self._qa:
            res = self.activation_quantizer(res)
        return res

    def forward(self, x):
        return self.epilogue(self.conv_only(x))

    def get_bn_dim(self):
        if isinstance(self, nn.Linear):
            return self.out_features
        if isinstance(self, _ConvNd):
            return self.out_channels
        raise NotImplementedError(f"Unsupported type used: {self}. Must be a linear or (transpose)-convolutional "
                                  f"nn.Module")


# ---- concrete hijacked layers (autoquant_utils.py:20-122) ---------------------------------------------
class _Conv1dForward:
    def run_forward(self, x, weight, bias, offsets=None):
        return F.conv1d(x.contiguous(), weight.contiguous(), bias=bias, stride=self.stride, padding=self.padding,
                        dilation=self.dilation, groups=self.groups)


def _space_to_depth_weight(w):
    """[O, C, k, k] (k odd) -> [O, 16, (k+1)/2, (k+1)/2], channels_last: the stride-2 kernel zero-extended to k+1 and
    split into its four (row parity, column parity) phases, channel order c*4 + p*2 + q as ops.space_to_depth2."""
    O, C, k, _ = w.shape
    a = (k + 1) // 2
    w = F.pad(w, (0, 1, 0, 1)).reshape(O, C, a, 2, a, 2).permute(0, 1, 3, 5, 2, 4).reshape(O, C * 4, a, a)
    if C * 4 < 16:
        w = F.pad(w, (0, 0, 0, 0, 0, 16 - C * 4))
    return w.contiguous(memory_format=torch.channels_last)


class _Conv2dForward:
    def _stem_s2d_ok(self, x, weight) -> bool:
        k = self.kernel_size[0]
        return (STEM_SPACE_TO_DEPTH and not torch.is_grad_enabled() and ops.on_device(x) and x.dtype == torch.float32
                and x.dim() == 4 and x.is_contiguous() and x.shape[1] <= 4 and x.shape[2] % 2 == 0
                and x.shape[3] % 2 == 0 and ops.is_channels_last(weight) and self.groups == 1
                and tuple(self.kernel_size) == (k, k) and k % 2 == 1 and k >= 3 and tuple(self.stride) == (2, 2)
                and tuple(self.padding) == (k // 2, k // 2) and tuple(self.dilation) == (1, 1)
                and self.padding_mode == "zeros")

    def run_forward(self, x, weight, bias, offsets=None):
        if self._stem_s2d_ok(x, weight):
            # NCHW image into a channels_last network: cuDNN's kernels for a stride-2 convolution over 3 input channels
            # run at a fraction of their roofline; the same sum as a stride-1 convolution over the space-to-depth
            # image (one HBM-bound gather + a 16-channel convolution) is 2x faster.  Exact re-indexing: only the
            # summation order inside the convolution changes, as between any two cuDNN algorithms.
            k = self.kernel_size[0]
            a = (k + 1) // 2
            hs, ws = x.shape[2] // 2 + a - 1, x.shape[3] // 2 + a - 1
            return F.conv2d(ops.space_to_depth2(x, k // 2, hs, ws), _space_to_depth_weight(weight), bias=bias)
        # (the reference forces NCHW here; a channels_last network -- model.to(memory_format=torch.channels_last) --
        # keeps its layout so that cuDNN runs its NHWC tensor-core kernels without the transposes around them)
        return F.conv2d(ops.dense(x), ops.dense(weight), bias=bias, stride=self.stride, padding=self.padding,
                        dilation=self.dilation, groups=self.groups)


class _LinearForward:
    def run_forward(self, x, weight, bias, offsets=None):
        return F.linear(x.contiguous(), weight.contiguous(), bias=bias)


class QuantConv1d(_Conv1dForward, QuantizationHijacker, nn.Conv1d):
    pass


class QuantConv(_Conv2dForward, QuantizationHijacker, nn.Conv2d):
    pass


class QuantLinear(_LinearForward, QuantizationHijacker, nn.Linear):
    pass


class BNQConv1d(_Conv1dForward, BNFusedHijacker, nn.Conv1d):
    pass


class BNQConv(_Conv2dForward, BNFusedHijacker, nn.Conv2d):
    pass


class BNQLinear(_LinearForward, BNFusedHijacker, nn.Linear):
    pass


class QuantConvTransposeBase(QuantizationHijacker):
    """autoquant_utils.py:46-58: per-channel quantisation applies to the OUT channels, which are dim 1
    of a transposed-conv weight."""

    def quantize_weights(self, weights):
        if self.per_channel_weights:
            weights = weights.transpose(1, 0).contiguous()
        weights = self.weight_quantizer(_weight_for_quantizer(weights))
        if self.per_channel_weights:
            weights = weights.transpose(1, 0).contiguous()
        return weights


class QuantConvTranspose1d(QuantConvTransposeBase, nn.ConvTranspose1d):
    def run_forward(self, x, weight, bias, offsets=None):
        return F.conv_transpose1d(x.contiguous(), weight.contiguous(), bias=bias, stride=self.stride,
                                  padding=self.padding, output_padding=self.output_padding, dilation=self.dilation,
                                  groups=self.groups)


class QuantConvTranspose(QuantConvTransposeBase, nn.ConvTranspose2d):
    def run_forward(self, x, weight, bias, offsets=None):
        return F.conv_transpose2d(x.contiguous(), weight.contiguous(), bias=bias, stride=self.stride,
                                  padding=self.padding, output_padding=self.output_padding, dilation=self.dilation,
                                  groups=self.groups)


class QuantLayerNorm(QuantizationHijacker, nn.LayerNorm):
    def run_forward(self, x, weight, bias, offsets=None):
        return F.layer_norm(input=x.contiguous(), normalized_shape=self.normalized_shape, weight=weight.contiguous(),
                            bias=bias.contiguous(), eps=self.eps)


class QuantizedActivationWrapper(QuantizedActivation):
    """autoquant_utils.py:125-163: wraps a parameter-free layer (pooling) and quantises its output,
    optionally re-using ("tying") the quantiser of the layer that feeds it, without a range update."""

    def __init__(self, layer, tie_activation_quantizers=False, input_quantizer: QuantizationManager = None, *args,
                 **kwargs):
        super().__init__(*args, **kwargs)
        self.tie_activation_quantizers = tie_activation_quantizers
        if input_quantizer:
            assert isinstance(input_quantizer, QuantizationManager)
            self.activation_quantizer = input_quantizer
        self.layer = layer

    def quantize_activations_no_range_update(self, x):
        return self.activation_quantizer.quantizer(x) if self._qa else x

    def forward(self, x):
        x = self.layer(x)
        if self.tie_activation_quantizers:
            return self.quantize_activations_no_range_update(x)
        return self.quantize_activations(x)

    def extra_repr(self):
        return f"tie_activation_quantizers={self.tie_activation_quantizers}"


def _pair(v):
    return (int(v), int(v)) if not isinstance(v, (tuple, list)) else (int(v[0]), int(v[1]))


class NativeMaxPool2d(nn.MaxPool2d):
    """nn.MaxPool2d (what models/resnet_quantized.py:73-78 keeps from torchvision) that pools channels_last
    activations with this library's kernel: one HBM-bound pass instead of ATen's max_pool_forward_nhwc (measured
    440 -> 85 us for the ResNet-18 stem at batch 128).  Same bits as F.max_pool2d; every other case (NCHW input, dilation,
    ceil mode, indices, autograd) goes to F.max_pool2d."""

    @classmethod
    def from_module(cls, m: nn.MaxPool2d):
        return cls(m.kernel_size, m.stride, m.padding, m.dilation, m.return_indices, m.ceil_mode)

    def forward(self, x):
        k, s, p = _pair(self.kernel_size), _pair(self.stride if self.stride is not None else self.kernel_size), \
            _pair(self.padding)
        if (NATIVE_MAX_POOL and not torch.is_grad_enabled() and ops.on_device(x) and x.dtype == torch.float32 and x.dim() == 4
                and ops.is_channels_last(x) and x.shape[1] % 4 == 0 and _pair(self.dilation) == (1, 1)
                and not self.ceil_mode and not self.return_indices and x.numel() > 0
                and x.shape[2] + 2 * p[0] >= k[0] and x.shape[3] + 2 * p[1] >= k[1]):
            return ops.max_pool2d_channels_last(x, k, s, p)
        return super().forward(x)


class Flattener(nn.Module):
    def forward(self, x):
        return x.view(x.shape[0], -1)


# These dicts are the injection point of the reference (autoquant_utils.py:183-194): plain, mutable.
non_bn_module_map = {
    nn.Conv1d: QuantConv1d,
    nn.Conv2d: QuantConv,
    nn.ConvTranspose1d: QuantConvTranspose1d,
    nn.ConvTranspose2d: QuantConvTranspose,
    nn.Linear: QuantLinear,
    nn.LayerNorm: QuantLayerNorm,
}
non_param_modules = (_AdaptiveAvgPoolNd, _AvgPoolNd)
bn_module_map = {nn.Conv1d: BNQConv1d, nn.Conv2d: BNQConv, nn.Linear: BNQLinear}
quant_conv_modules = (QuantConv1d, QuantConv, BNQConv1d, BNQConv)


def _layer_kwargs(mod, act):
    """Constructor kwargs that rebuild ``mod`` as its hijacked twin (autoquant_utils.py:219-263)."""
    if isinstance(mod, _ConvNd):
        kw = dict(in_channels=mod.in_channels, out_channels=mod.out_channels, kernel_size=mod.kernel_size,
                  stride=mod.stride, padding=mod.padding, dilation=mod.dilation, groups=mod.groups,
                  bias=mod.bias is not None)
        if isinstance(mod, (nn.ConvTranspose1d, nn.ConvTranspose2d)):
            kw["output_padding"] = mod.output_padding
    elif isinstance(mod, nn.Linear):
        kw = dict(in_features=mod.in_features, out_features=mod.out_features, bias=mod.bias is not None)
    elif isinstance(mod, nn.LayerNorm):
        kw = dict(normalized_shape=mod.normalized_shape, eps=mod.eps)
    else:
        raise ValueError(f"cannot quantise module of type {type(mod)}")
    kw["activation"] = act
    return kw


def _is_bn(m):
    return isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d))


def _is_act(m):
    return isinstance(m, tuple(activations_set))


def fold_bn(module, i, **quant_params):
    """Builds the hijacked layer for ``module[i]`` absorbing a following BN and/or activation
    (autoquant_utils.py:266-289).  Returns (new_module, index of the next unconsumed child)."""
    layer = module[i]
    has_bn = len(module) > i + 1 and _is_bn(module[i + 1])
    act = None
    j = i + 1 + int(has_bn)
    if len(module) > j and _is_act(module[j]):
        act = module[j]
    cls = (bn_module_map if has_bn else non_bn_module_map)[type(layer)]
    new = cls(**_layer_kwargs(layer, act), **quant_params)
    new.weight.data = layer.weight.data.clone()
    if has_bn:
        bn = module[i + 1]
        new.gamma.data = bn.weight.data.clone()
        new.beta.data = bn.bias.data.clone()
        new.running_mean.data = bn.running_mean.data.clone()
        new.running_var.data = bn.running_var.data.clone()
        if layer.bias is not None:
            new.running_mean.data -= layer.bias.data
            print("Warning: bias in conv/linear before batch normalization.")
        new.epsilon = bn.eps
    elif layer.bias is not None:
        new.bias.data = layer.bias.data.clone()
    return new, i + 1 + int(has_bn) + int(act is not None)


def quantize_sequential(model, specials=None, tie_activation_quantizers=False, **quant_params):
    """autoquant_utils.py:292-345."""
    specials = specials or dict()
    out = []
    i = 0
    while i < len(model):
        child = model[i]
        if isinstance(child, QuantizedModule):
            out.append(child)
        elif type(child) in non_bn_module_map:
            new, i = fold_bn(model, i, **quant_params)
            out.append(new)
            continue
        elif type(child) in specials:
            out.append(specials[type(child)](child, **quant_params))
        elif isinstance(child, non_param_modules):
            feeder = None
            if out and isinstance(out[-1], QuantizedModule):
                feeder = out[-1]
            elif out and isinstance(out[-1], nn.Sequential) and isinstance(out[-1][-1], QuantizedModule):
                feeder = out[-1][-1]
            if feeder is not None and tie_activation_quantizers:
                print(f"Tying input quantizer {i-1}^th layer of type {type(feeder)} to the quantized "
                      f"{type(child)} following it")
                out.append(QuantizedActivationWrapper(child, tie_activation_quantizers=tie_activation_quantizers,
                                                      input_quantizer=feeder.activation_quantizer, **quant_params))
            else:
                out.append(QuantizedActivationWrapper(child, **quant_params))
                if tie_activation_quantizers:
                    warnings.warn("Input quantizer not found, so we do not tie quantizers")
        else:
            out.append(quantize_model(child, specials=specials, **quant_params))
        i += 1
    return nn.Sequential(*out)


def quantize_model(model, specials=None, tie_activation_quantizers=False, **quant_params):
    """autoquant_utils.py:348-381."""
    specials = specials or dict()
    if isinstance(model, nn.Sequential):
        return quantize_sequential(model, specials, tie_activation_quantizers, **quant_params)
    if type(model) in specials:
        return specials[type(model)](model, **quant_params)
    if isinstance(model, non_param_modules):
        return QuantizedActivationWrapper(model, **quant_params)
    if type(model) is nn.MaxPool2d:   # the reference deep-copies it (autoquant_utils.py:376-381); same module, own kernel
        return NativeMaxPool2d.from_module(model)
    if type(model) in non_bn_module_map:
        new = non_bn_module_map[type(model)](**_layer_kwargs(model, None), **quant_params)
        new.weight.data = model.weight.data
        if getattr(model, "bias", None) is not None:
            new.bias.data = model.bias.data
        return new
    new = copy.deepcopy(model)
    for name, child in new._modules.items():
        q = quantize_model(child, specials=specials, **quant_params)
        if q is not None:
            setattr(new, name, q)
    return new


class QuantizedModel(nn.Module):
    """base_quantized_model.py:19-135: whole-model switches."""

    def __init__(self, input_size=(1, 3, 224, 224)):
        super().__init__()
        self.input_size = input_size

    def prequantize_weights(self):
        """Quantises the weights of every hijacked layer whose weight range is fixed in ONE launch per format
        and hands each layer its result for the forward that follows (same work as the per-layer
        QuantizationHijacker.quantize_weights calls, hijacker.py:88-98, without 21..53 tiny launches).  Layers that
        are still estimating ranges, or whose weights need a transposed layout, keep the per-layer path."""
        if not BATCH_WEIGHT_QUANT or torch.is_grad_enabled():
            return
        groups = {}
        for m in self.modules():
            if not isinstance(m, QuantizationHijacker) or not m._qw or isinstance(m, QuantConvTransposeBase):
                continue
            mgr = m.weight_quantizer
            q = mgr.quantizer
            w = m.weight
            if (not isinstance(mgr, QuantizationManager) or not isinstance(q, FPQuantizer) or mgr.estimating()
                    or not ops.on_device(w) or w.dtype != torch.float32
                    or not (w.is_contiguous() or ops.is_channels_last(w))):
                continue
            w = w.detach()
            table, C = q.table_for(w)
            key = ((m.weight.data_ptr(), m.weight._version), q._table_key)
            if CACHE_QUANTIZED_WEIGHTS:
                hit = m.__dict__.get("_wq_cache")
                if hit is not None and hit[0] == key:
                    m.__dict__["_wq_stash"] = (key[0], hit[1])
                    continue
            groups.setdefault((q._mbits_host, q.n_bits, q.sign_bits), []).append((m, w, table, C, key))
        for (mb, nb, sb), items in groups.items():
            outs = ops.fake_quant_multi([it[1] for it in items], [it[2] for it in items], [it[3] for it in items],
                                        mb, nb, sb)
            for (m, _, _, _, key), out in zip(items, outs):
                m.__dict__["_wq_stash"] = (key[0], out)
                if CACHE_QUANTIZED_WEIGHTS:
                    m.__dict__["_wq_cache"] = (key, out)
                else:
                    m.__dict__.pop("_wq_cache", None)

    def load_state_dict(self, state_dict, strict: bool = True):
        flags = {k: v for k, v in state_dict.items() if k.endswith("_quant_a") or k.endswith("_quant_w")}
        if not flags:
            raise ValueError("The quantization states of activations or weights should be included in the state dict ")
        super().load_state_dict(flags, strict=False)
        device = next(self.parameters()).device
        with torch.no_grad():
            self.forward(torch.rand(*self.input_size, device=device))  # materialise range buffers
        return super().load_state_dict(state_dict, strict)

    def _each(self, fn_name):
        def _fn(layer):
            if isinstance(layer, QuantizedModule):
                getattr(layer, fn_name)()

        self.apply(_fn)

    def quantized_weights(self):
        self._each("quantized_weights")

    def full_precision_weights(self):
        self._each("full_precision_weights")

    def quantized_acts(self):
        self._each("quantized_acts")

    def full_precision_acts(self):
        self._each("full_precision_acts")

    def quantized(self):
        self._each("quantized")

    def full_precision(self):
        self._each("full_precision")

    def estimate_ranges(self):
        self.apply(_set_layer_estimate_ranges)

    def estimate_ranges_train(self):
        self.apply(_set_layer_estimate_ranges_train)

    def set_quant_state(self, weight_quant, act_quant):
        self.quantized_acts() if act_quant else self.full_precision_acts()
        self.quantized_weights() if weight_quant else self.full_precision_weights()

    def grad_scaling(self, grad_scaling=True):
        def _fn(module):
            if isinstance(module, QuantizerBase):
                module.grad_scaling = grad_scaling

        self.apply(_fn)

    def learn_ranges(self):
        self.apply(_set_layer_learn_ranges)

    def fix_ranges(self):
        self.apply(_set_layer_fix_ranges)
