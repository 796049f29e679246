"""The two quantised CNN workloads of the reference, built on this package's module layer, plus the
validate-quantized flow.  (They are the *callers* of the hot path -- SURVEY.md section 8 rows a12,
a15, a16 -- needed on the GPU box where the reference checkout does not exist.)

Mirrors:
  QuantizedBlock / QuantizedResNet / resnet18_quantized   models/resnet_quantized.py:14-167
  MobileNetV2 (fp32 definition)                            models/mobilenet_v2.py:16-132
  QuantizedInvertedResidual / QuantizedMobileNetV2         models/mobilenet_v2_quantized.py:15-113
  pass_data_for_range_estimation                           quantization/utils.py:74-115
  validate_quantized (flow)                                image_net.py:48-96
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn

from . import dist as fq_dist
from .modules import (BNQConv, FP32Acts, Flattener, QuantizedActivation, QuantizedActivationWrapper, QuantizedModel,
                      quantize_model, quantize_sequential)
from .quantizers import FPQuantizer
from .range_estimators import AllMinMaxEstimator, CurrentMinMaxEstimator


# ---------------------------------------------------------------------------------------------------
# quant_setup: per-layer deviations from the uniform configuration
# ---------------------------------------------------------------------------------------------------
# The reference applies them as if/elif ladders at the end of the model constructors (models/resnet_quantized.py:
# 94-124, models/mobilenet_v2_quantized.py:45-84).  Here they are data: setup name -> edits, an edit being
# (module path below the model, field, value) with the fields
#   "w_bits" / "a_bits"  n_bits of the layer's weight / activation quantiser
#   "fp32_acts"          replace the layer's activation quantiser by FP32Acts (no quantisation)
#   "fp32_block_acts"    the same for every QuantizedActivation below the path (block outputs, tied wrappers)
#   "dw_w_bits"          w_bits of every depthwise BNQConv below the path
# Negative indices count from the end of a Sequential, as in the reference's ``self.features[-1][-1]``.
_RESNET_SETUPS = {
    "all": (),
    "LSQ": (("features.0", "w_bits", 8), ("features.-1.-1", "a_bits", 8), ("features.-1.-1.features.-1", "a_bits", 8),
            ("fc", "w_bits", 8), ("fc", "fp32_acts", None)),
    "LSQ_paper": (("features.0", "fp32_acts", None), ("features.0", "w_bits", 8), ("fc", "a_bits", 8), ("fc", "w_bits", 8),
                  ("features", "fp32_block_acts", None)),
    "FP_logits": (("fc", "fp32_acts", None),),
    "fc4": (("features.0", "w_bits", 8), ("fc", "w_bits", 4)),
}
_MOBILENETV2_SETUPS = {
    "all": (),
    "FP_logits": (("classifier.1", "fp32_acts", None),),
    "fc4": (("features.0.0", "w_bits", 8), ("classifier.1", "w_bits", 4)),
    "fc4_dw8": (("features.0.0", "w_bits", 8), ("classifier.1", "w_bits", 4), ("", "dw_w_bits", 8)),
    "LSQ": (("features.0.0", "w_bits", 8), ("features.-2.0", "a_bits", 8), ("classifier.1", "w_bits", 8),
            ("classifier.1", "fp32_acts", None)),
    "LSQ_paper": (("features.0.0", "fp32_acts", None), ("features.0.0", "w_bits", 8), ("classifier.1", "w_bits", 8),
                  ("classifier.1", "a_bits", 8), ("features", "fp32_block_acts", None)),
}


def _resolve(root, path):
    mod = root
    for part in (path.split(".") if path else ()):
        mod = mod[int(part)] if part.lstrip("-").isdigit() else getattr(mod, part)
    return mod


def apply_quant_setup(model, family, table, quant_setup):
    """Applies the edits of ``table[quant_setup]`` to a freshly built quantised model (None = "all")."""
    if quant_setup is None:
        return
    if quant_setup not in table:
        raise ValueError("Quantization setup '{}' not supported for {}".format(quant_setup, family))
    for path, field, value in table[quant_setup]:
        target = _resolve(model, path)
        if field == "w_bits":
            target.weight_quantizer.quantizer.n_bits = value
        elif field == "a_bits":
            target.activation_quantizer.quantizer.n_bits = value
        elif field == "fp32_acts":
            target.activation_quantizer = FP32Acts()
        elif field == "fp32_block_acts":
            for layer in target.modules():
                if isinstance(layer, QuantizedActivation):
                    layer.activation_quantizer = FP32Acts()
        elif field == "dw_w_bits":
            for layer in target.modules():
                if isinstance(layer, BNQConv) and layer.groups == layer.in_channels:
                    layer.weight_quantizer.quantizer.n_bits = value
        else:  # pragma: no cover
            raise KeyError(field)


# ---------------------------------------------------------------------------------------------------
# ResNet
# ---------------------------------------------------------------------------------------------------
class QuantizedBlock(QuantizedActivation):
    """models/resnet_quantized.py:14-46."""

    def __init__(self, block, **quant_params):
        super().__init__(**quant_params)
        from torchvision.models.resnet import BasicBlock, Bottleneck

        if isinstance(block, Bottleneck):
            features = nn.Sequential(block.conv1, block.bn1, block.relu, block.conv2, block.bn2, block.relu,
                                     block.conv3, block.bn3)
        elif isinstance(block, BasicBlock):
            features = nn.Sequential(block.conv1, block.bn1, block.relu, block.conv2, block.bn2)
        else:
            raise ValueError(f"unsupported residual block {type(block)}")
        self.features = quantize_model(features, **quant_params)
        self.downsample = quantize_model(block.downsample, **quant_params) if block.downsample else None
        self.relu = block.relu

    def forward(self, x):
        # residual = x | downsample(x); out = features(x); out += residual; relu; quantise.
        # With fixed ranges the last BN + its quantiser + add + relu + quantiser run as one kernel.
        return self.block_tail(self.features, x, lambda: x if self.downsample is None else self.downsample(x), self.relu)


class QuantizedResNet(QuantizedModel):
    """models/resnet_quantized.py:49-133."""

    def __init__(self, resnet, input_size=(1, 3, 224, 224), quant_setup=None, **quant_params):
        super().__init__(input_size)
        from torchvision.models.resnet import BasicBlock, Bottleneck

        specials = {BasicBlock: QuantizedBlock, Bottleneck: QuantizedBlock}
        stem = [resnet.conv1, resnet.bn1, resnet.relu]
        if hasattr(resnet, "maxpool"):
            stem.append(resnet.maxpool)
        features = nn.Sequential(*stem, resnet.layer1, resnet.layer2, resnet.layer3, resnet.layer4)
        self.features = quantize_model(features, specials=specials, **quant_params)

        if quant_setup and quant_setup == "LSQ_paper":
            self.avgpool = resnet.avgpool
        else:
            self.avgpool = QuantizedActivationWrapper(
                resnet.avgpool, tie_activation_quantizers=True,
                input_quantizer=self.features[-1][-1].activation_quantizer, **quant_params)
        self.flattener = Flattener()
        self.fc = quantize_model(resnet.fc, **quant_params)

        apply_quant_setup(self, "Resnet", _RESNET_SETUPS, quant_setup)

    def forward(self, x):
        self.prequantize_weights()
        x = self.features(x)
        x = self.avgpool(x)
        x = self.flattener(x)
        return self.fc(x)


def resnet18_quantized(pretrained=False, model_dir=None, load_type="fp32", **qparams):
    """models/resnet_quantized.py:136-150.  No network on the GPU box -> random init unless a
    state dict is given (``pretrained`` must stay False)."""
    from torchvision.models import resnet18

    if pretrained:
        raise ValueError("pretrained weights cannot be downloaded here; load a state dict explicitly")
    model = QuantizedResNet(resnet18(), **qparams)
    if load_type == "quantized":
        model.load_state_dict(torch.load(model_dir))
    elif load_type != "fp32":
        raise ValueError("wrong load_type specified")
    return model


def resnet50_quantized(pretrained=False, model_dir=None, load_type="fp32", **qparams):
    from torchvision.models import resnet50

    if pretrained:
        raise ValueError("pretrained weights cannot be downloaded here; load a state dict explicitly")
    model = QuantizedResNet(resnet50(), **qparams)
    if load_type == "quantized":
        model.load_state_dict(torch.load(model_dir))
    elif load_type != "fp32":
        raise ValueError("wrong load_type specified")
    return model


# ---------------------------------------------------------------------------------------------------
# MobileNetV2
# ---------------------------------------------------------------------------------------------------
def _conv_bn_relu6(cin, cout, k, stride, groups=1):
    return [nn.Conv2d(cin, cout, k, stride, (k - 1) // 2, groups=groups, bias=False), nn.BatchNorm2d(cout),
            nn.ReLU6(inplace=True)]


class InvertedResidual(nn.Module):
    """models/mobilenet_v2.py:28-69: (1x1 expand) -> 3x3 depthwise -> 1x1 linear projection."""

    def __init__(self, inp, oup, stride, expand_ratio):
        super().__init__()
        assert stride in (1, 2)
        self.stride = stride
        hidden = round(inp * expand_ratio)
        self.use_res_connect = stride == 1 and inp == oup
        layers = []
        if expand_ratio != 1:
            layers += _conv_bn_relu6(inp, hidden, 1, 1)
        layers += _conv_bn_relu6(hidden, hidden, 3, stride, groups=hidden)
        layers += [nn.Conv2d(hidden, oup, 1, 1, 0, bias=False), nn.BatchNorm2d(oup)]
        self.conv = nn.Sequential(*layers)

    def forward(self, x):
        return x + self.conv(x) if self.use_res_connect else self.conv(x)


class MobileNetV2(nn.Module):
    """models/mobilenet_v2.py:72-132 (the tonylins/pytorch-mobilenet-v2 layout the reference uses)."""

    SETTINGS = [  # expansion t, channels c, repeats n, stride s
        (1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1)]

    def __init__(self, n_class=1000, input_size=224, width_mult=1.0, dropout=0.0):
        super().__init__()
        assert input_size % 32 == 0
        cin = int(32 * width_mult)
        self.last_channel = int(1280 * width_mult) if width_mult > 1.0 else 1280
        features = [nn.Sequential(*_conv_bn_relu6(3, cin, 3, 2))]
        for t, c, n, s in self.SETTINGS:
            cout = int(c * width_mult)
            for i in range(n):
                features.append(InvertedResidual(cin, cout, s if i == 0 else 1, expand_ratio=t))
                cin = cout
        features.append(nn.Sequential(*_conv_bn_relu6(cin, self.last_channel, 1, 1)))
        features.append(nn.AvgPool2d(input_size // 32))
        self.features = nn.Sequential(*features)
        self.classifier = nn.Sequential(nn.Dropout(dropout), nn.Linear(self.last_channel, n_class))
        self._initialize_weights()

    def forward(self, x):
        x = self.features(x)
        x = F.adaptive_avg_pool2d(x, 1).squeeze()
        return self.classifier(x)

    def _initialize_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2.0 / n))
                if m.bias is not None:
                    m.bias.data.zero_()
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
            elif isinstance(m, nn.Linear):
                m.weight.data.normal_(0, 0.01)
                m.bias.data.zero_()


class QuantizedInvertedResidual(QuantizedActivation):
    """models/mobilenet_v2_quantized.py:15-26."""

    def __init__(self, inv_res_orig, **quant_params):
        super().__init__(**quant_params)
        self.use_res_connect = inv_res_orig.use_res_connect
        self.conv = quantize_sequential(inv_res_orig.conv, **quant_params)

    def forward(self, x):
        if self.use_res_connect:
            return self.block_tail(self.conv, x, lambda: x, None)  # Q(x + conv(x)); the conv ends in BN + quantiser
        return self.conv(x)


class QuantizedMobileNetV2(QuantizedModel):
    """models/mobilenet_v2_quantized.py:29-92."""

    def __init__(self, model_fp, input_size=(1, 3, 224, 224), quant_setup=None, **quant_params):
        super().__init__(input_size)
        specials = {InvertedResidual: QuantizedInvertedResidual}
        quantize_input = quant_setup and quant_setup == "LSQ_paper"
        self.features = quantize_sequential(model_fp.features, tie_activation_quantizers=not quantize_input,
                                            specials=specials, **quant_params)
        self.flattener = Flattener()
        self.classifier = quantize_model(model_fp.classifier, **quant_params)

        apply_quant_setup(self, "MobilenetV2", _MOBILENETV2_SETUPS, quant_setup)

    def forward(self, x):
        self.prequantize_weights()
        x = self.features(x)
        x = self.flattener(x)
        return self.classifier(x)


def mobilenetv2_quantized(pretrained=False, model_dir=None, load_type="fp32", **qparams):
    """models/mobilenet_v2_quantized.py:95-113; random init unless a state dict path is given."""
    fp_model = MobileNetV2()
    if load_type == "fp32":
        if pretrained:
            fp_model.load_state_dict(torch.load(model_dir))
        return QuantizedMobileNetV2(fp_model, **qparams)
    if load_type == "quantized":
        model = QuantizedMobileNetV2(fp_model, **qparams)
        model.load_state_dict(torch.load(model_dir), strict=False)
        return model
    raise ValueError("wrong load_type specified")


# ---------------------------------------------------------------------------------------------------
# configuration + flow
# ---------------------------------------------------------------------------------------------------
def readme_quant_params(mantissa_bits: int, *, method=FPQuantizer, weight_range_method=CurrentMinMaxEstimator,
                        act_range_method=AllMinMaxEstimator, mse_include_mantissa_bits=False, act_range_options=None,
                        weight_range_options=None, allow_unsigned=False):
    """The kwargs dict utils/click_options.py:477-510 produces for the README command line
    (README.md:63-68): --n-bits 8 --per-channel --quant-setup all --qmethod fp_quantizer
    --fp8-mantissa-bits=M --fp8-set-maxval --no-fp8-mse-include-mantissa-bits
    --weight-quant-method=current_minmax --act-quant-method=allminmax."""
    return dict(
        method=method, n_bits=8, n_bits_act=None, act_method=method, per_channel_weights=True, quant_setup="all",
        weight_range_method=weight_range_method, weight_range_options=weight_range_options or {},
        act_range_method=act_range_method, act_range_options=act_range_options or {}, quantize_input=False,
        fp8_kwargs=dict(maxval=None, mantissa_bits=mantissa_bits, set_maxval=True, learn_maxval=False,
                        learn_mantissa_bits=False, mse_include_mantissa_bits=mse_include_mantissa_bits,
                        allow_unsigned=allow_unsigned))


def pass_data_for_range_estimation(loader, model, act_quant, weight_quant, max_num_batches=20, cross_entropy_layer=None,
                                   inp_idx=0):
    """quantization/utils.py:74-115: calibration forward passes in eval mode under no_grad.  ``loader`` yields what the
    reference's loaders yield -- ``(x, y)`` tuples / lists (``inp_idx`` selects the input), dicts of keyword tensors --
    or bare input tensors; inputs are moved to the model's device.  Returns the inputs that were passed (the reference
    returns host numpy copies of them, :103; that device-to-host copy per batch is not made here).  The cross-entropy
    range estimator the reference can install here (:82-94) is not part of the FP8 path."""
    if cross_entropy_layer is not None:
        raise NotImplementedError("the cross-entropy range estimator is outside the FP8 fake-quantisation path")
    from . import ops

    model.set_quant_state(weight_quant, act_quant)
    model.eval()  # BN EMA must not be updated
    device = next(model.parameters()).device
    passed = []
    with torch.no_grad(), ops.nvtx_range("fp8fq.pass_data_for_range_estimation"):
        for i, data in enumerate(loader):
            if isinstance(data, dict):
                model(**{k: v.to(device=device) for k, v in data.items()})
            else:
                x = data[inp_idx] if isinstance(data, (tuple, list)) else data
                x = x.to(device=device)
                passed.append(x)
                model(x)
            if i >= max_num_batches - 1 or not act_quant:
                break
    return passed


class _BatchStatisticsMode:
    """Context in which every BNFusedHijacker of ``model`` (the layer itself, not its quantiser children) normalises with
    -- and, momentum being 1, records as its running statistics -- the statistics of the current batch."""

    def __init__(self, model):
        from .modules import BNFusedHijacker

        self.layers = [m for m in model.modules() if isinstance(m, BNFusedHijacker)]
        self.saved = None

    def __enter__(self):
        self.saved = [(m.momentum, m.training) for m in self.layers]
        for m in self.layers:
            m.momentum, m.training = 1.0, True
        return self.layers

    def __exit__(self, *exc):
        for m, (momentum, training) in zip(self.layers, self.saved):
            m.momentum, m.training = momentum, training


def _keep_ema_statistics(layer):
    """``store_ema_stats``: the statistics being replaced stay available as buffers (and thus in the state dict)."""
    for name in ("running_mean", "running_var"):
        ema = getattr(layer, name).detach().clone()
        if name + "_ema" in layer._buffers:
            setattr(layer, name + "_ema", ema)
        else:
            layer.register_buffer(name + "_ema", ema)


@torch.no_grad()
def reestimate_BN_stats(model, data_loader, num_batches=50, store_ema_stats=False):
    """utils/qat_utils.py:45-90: re-estimate the batch-norm statistics of the QUANTISED network as the mean, over up to
    ``num_batches`` batches, of the per-batch statistics each fused layer sees (quantisers stay in their current state).
    Under data parallelism (fp8_quantization_b200.dist active) the batch statistics are those of the global batch
    (per-channel sum / sum-of-squares all-reduce, modules._sync_batch_norm_train), so every rank ends with the
    statistics a single process would compute on the concatenated batches.  Returns the number of batches used."""
    from . import ops

    model.eval()
    device = next(model.parameters()).device
    batches = 0
    with ops.nvtx_range("fp8fq.reestimate_BN_stats"), _BatchStatisticsMode(model) as layers:
        if store_ema_stats:
            for layer in layers:
                _keep_ema_statistics(layer)
        totals = [(torch.zeros_like(m.running_mean), torch.zeros_like(m.running_var)) for m in layers]
        for data in data_loader:    # the reference's loaders yield (x, y) (:72); bare input tensors are accepted too
            model((data[0] if isinstance(data, (tuple, list)) else data).to(device))
            for layer, (mean_sum, var_sum) in zip(layers, totals):
                mean_sum += layer.running_mean
                var_sum += layer.running_var
            batches += 1
            if batches == num_batches:
                break
        for layer, (mean_sum, var_sum) in zip(layers, totals):
            layer.running_mean = mean_sum / batches
            layer.running_var = var_sum / batches
            layer._bn_key = None   # fresh statistics tensors: the packed batch-norm parameters are stale
    model.eval()
    return batches


class ReestimateBNStats:
    """utils/qat_utils.py:33-42 (callable handler)."""

    def __init__(self, model, data_loader, num_batches=50):
        self.model = model
        self.data_loader = data_loader
        self.num_batches = num_batches

    def __call__(self, engine=None):
        print("-- Reestimate current BN statistics --")
        reestimate_BN_stats(self.model, self.data_loader, self.num_batches)


@torch.no_grad()
def validate(model, batches, labels=None):
    """image_net.py:72-96 without ignite: top-1 / top-5 / mean CE loss over ``batches``; under data
    parallelism the four counters are summed across ranks with one all-reduce."""
    stats = None
    for i, x in enumerate(batches):
        logits = model(x)
        y = labels[i] if labels is not None else torch.zeros(x.shape[0], dtype=torch.long, device=x.device)
        top5 = logits.topk(5, dim=1).indices
        c1 = (top5[:, 0] == y).sum()
        c5 = (top5 == y[:, None]).any(dim=1).sum()
        loss = F.cross_entropy(logits, y, reduction="sum")
        cur = torch.stack([c1.float(), c5.float(), loss.float(), torch.tensor(float(x.shape[0]), device=x.device)])
        stats = cur if stats is None else stats + cur
    if fq_dist.world_size() > 1:
        fq_dist.all_reduce_sum(stats)
    c1, c5, loss, n = stats.tolist()
    return dict(top_1_accuracy=c1 / n, top_5_accuracy=c5 / n, loss=loss / n, count=int(n))


# ---------------------------------------------------------------------------------------------------
# empirical quantisation error (the data-parallel half of compute_quant_error.py)
# ---------------------------------------------------------------------------------------------------
@torch.no_grad()
def estimate_rounding_error_empirical(W, quantizer, range_min, range_max) -> float:
    """quantization/quant_error_estimator.py:68-75: mean((Q(W) - W)^2) with the given range."""
    quantizer.set_quant_range(range_min, range_max)
    return torch.mean(((quantizer(W) - W) ** 2).flatten()).item()


@torch.no_grad()
def estimate_dot_prod_error_empirical(x, y, quantizer_x, quantizer_y, x_range_min, x_range_max, y_range_min,
                                      y_range_max) -> float:
    """quant_error_estimator.py:78-89: mean((x*y - Q(x)*Q(y))^2)."""
    quantizer_x.set_quant_range(x_range_min, x_range_max)
    quantizer_y.set_quant_range(y_range_min, y_range_max)
    return torch.mean((torch.mul(x, y) - torch.mul(quantizer_x(x), quantizer_y(y))) ** 2).item()


@torch.no_grad()
def compute_quant_error_empirical(sample, sample_y=None, n_bits=8, num_candidates=1000, exp_bits_list=(5, 4, 3, 2, 0)):
    """The empirical flow of compute_quant_error.py:18-57 on a device-resident sample: for every format of the script
    (E5M2 .. E2M5 with FPQuantizer, E = 0 with SymmetricUniformQuantizer) the clipping range found by the line search
    (all candidates in one launch of the MSE-grid kernel for the FP formats), the empirical rounding MSE / SQNR and the
    empirical dot-product MSE / SQNR (``sample_y``: the second operand; default the sample itself, reversed).  The
    analytic expectations the script prints next to them (scipy integrals over the format's grid,
    quant_error_estimator.py:32-66, utils/distributions.py) are CPU-side analytics and out of scope."""
    import math

    from .quantizers import SymmetricUniformQuantizer
    from .range_estimators import LineSearchEstimator

    if sample_y is None:
        sample_y = sample.flip(0).contiguous()
    rows = []
    for exp_bits in exp_bits_list:
        mantissa_bits = n_bits - 1 - exp_bits

        def make():
            if exp_bits > 0:
                return FPQuantizer(n_bits=n_bits, mantissa_bits=mantissa_bits, set_maxval=True)
            return SymmetricUniformQuantizer(n_bits=n_bits)

        quant = make()
        rmin, rmax = LineSearchEstimator(quantizer=quant, num_candidates=num_candidates).forward(sample)
        mse = estimate_rounding_error_empirical(sample, quant, rmin, rmax)
        dot = estimate_dot_prod_error_empirical(sample, sample_y, make(), make(), rmin, rmax, rmin, rmax)
        rows.append(dict(exp_bits=exp_bits, mantissa_bits=mantissa_bits, range_min=float(rmin.reshape(-1)[0]),
                         range_max=float(rmax.reshape(-1)[0]), mse=mse,
                         sqnr=-10.0 * math.log10(mse) if mse > 0 else float("inf"), dot_prod_mse=dot,
                         dot_prod_sqnr=-10.0 * math.log10(dot) if dot > 0 else float("inf")))
    return rows


# ---------------------------------------------------------------------------------------------------
# validate forward as one CUDA graph
# ---------------------------------------------------------------------------------------------------
IMAGENET_MEAN = (0.485, 0.456, 0.406)   # utils/imagenet_dataloaders.py:66
IMAGENET_STD = (0.229, 0.224, 0.225)


class U8Normalize:
    """ToTensor + Normalize of the reference's input pipeline (utils/imagenet_dataloaders.py:66-81) for a uint8 NCHW
    image batch that is already on the device: one HBM-bound launch through a per-(channel, byte value) table built with
    the reference's own fp32 operations -- bit-identical to torchvision, and the images cross the host link as 1 byte
    per pixel instead of 4.  Use as ``GraphedForward(model, example_u8, preprocess=U8Normalize(device=...))``."""

    def __init__(self, mean=IMAGENET_MEAN, std=IMAGENET_STD, device=None):
        from . import ops

        self.lut = ops.normalize_lut(mean, std, device if device is not None else ops.default_device())

    def __call__(self, x_u8):
        from . import ops

        return ops.normalize_u8(x_u8, self.lut)


class GraphedForward:
    """``model(x)`` for one input shape captured in a CUDA graph and replayed.

    The validate pass of the reference (image_net.py:72-96) calls the model once per batch from Python: 51..117
    quantiser calls, each a handful of eager launches.  With fixed ranges nothing in this package's forward touches
    the host (no ``.item()``, tables and batch-norm parameters device-resident), so the whole forward -- cuDNN
    convolutions, the fused epilogues, the per-forward weight re-quantisation -- is capturable and replays with no
    Python or launch overhead.

    Valid only while the model's state does not change: ranges fixed, eval mode, same weights / batch-norm tensors
    (storage, not values -- in-place updates are seen).  Call :meth:`capture` again after changing any of that.
    """

    def __init__(self, model, example: torch.Tensor, warmup: int = 3, preprocess=None):
        """``preprocess``: an optional device-side callable applied to the input inside the graph (e.g. U8Normalize for
        uint8 images); ``example`` then has the dtype / shape of what the callable takes."""
        self.model = model
        self.preprocess = preprocess
        self.graph = None
        self.static_in = None
        self.static_out = None
        self.capture(example, warmup)

    def _forward(self):
        x = self.static_in if self.preprocess is None else self.preprocess(self.static_in)
        return self.model(x)

    def _check_state(self):
        from .quantization_manager import QuantizationManager

        if self.model.training:
            raise RuntimeError("GraphedForward: model.eval() first (training-mode batch norm updates host-visible state)")
        for m in self.model.modules():
            if isinstance(m, QuantizationManager) and m.estimating():
                raise RuntimeError("GraphedForward: ranges must be fixed (model.fix_ranges()) before capture: range "
                                   "estimation mutates quantiser state on every call")

    @torch.no_grad()
    def capture(self, example: torch.Tensor, warmup: int = 3):
        self._check_state()
        if not example.is_cuda:
            raise ValueError("GraphedForward: the example input must live on the GPU")
        self.static_in = example.detach().clone()
        side = torch.cuda.Stream(device=example.device)
        side.wait_stream(torch.cuda.current_stream(example.device))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):  # cuDNN algorithm selection, table builds, batch-norm packing
                self._forward()
        torch.cuda.current_stream(example.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = self._forward()
        return self

    def replay(self):
        """Runs the forward on whatever ``static_in`` currently holds; returns ``static_out`` (overwritten by the
        next replay)."""
        from . import ops

        with ops.nvtx_range("fp8fq.GraphedForward.replay"):
            self.graph.replay()
        return self.static_out

    def __call__(self, x: torch.Tensor):
        if tuple(x.shape) != tuple(self.static_in.shape):
            raise ValueError(f"GraphedForward was captured for shape {tuple(self.static_in.shape)}, got {tuple(x.shape)}")
        self.static_in.copy_(x, non_blocking=True)
        return self.replay()

    @torch.no_grad()
    def run_pipelined(self, host_batches, host_out: torch.Tensor):
        """Validate loop over batches that live in (pinned) HOST memory, results written to ``host_out``
        ([len(host_batches), *output shape], pinned): the H2D copy of batch k+1 and the D2H copy of batch k-1 overlap
        the forward of batch k (copy-in stream, compute stream, copy-out stream; two staging buffers each way).
        Synchronises before returning."""
        dev = self.static_in.device
        cur = torch.cuda.current_stream(dev)
        s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        stage = [torch.empty_like(self.static_in) for _ in range(2)]
        d_out = [torch.empty_like(self.static_out) for _ in range(2)]
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_free = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(2)]
        ev_host = [torch.cuda.Event() for _ in range(2)]
        s_in.wait_stream(cur)
        s_out.wait_stream(cur)
        for it, hb in enumerate(host_batches):
            k = it & 1
            with torch.cuda.stream(s_in):
                if it >= 2:
                    s_in.wait_event(ev_free[k])      # the forward that read stage[k] two batches ago has consumed it
                stage[k].copy_(hb, non_blocking=True)
                ev_in[k].record(s_in)
            cur.wait_event(ev_in[k])
            self.static_in.copy_(stage[k])
            ev_free[k].record(cur)
            self.graph.replay()
            if it >= 2:
                cur.wait_event(ev_host[k])           # d_out[k] has left for the host
            d_out[k].copy_(self.static_out)
            ev_out[k].record(cur)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_out[k])
                host_out[it].copy_(d_out[k], non_blocking=True)
                ev_host[k].record(s_out)
        cur.wait_stream(s_in)
        cur.wait_stream(s_out)
        torch.cuda.synchronize(dev)
        return host_out
