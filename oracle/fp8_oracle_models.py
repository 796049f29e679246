"""Oracle-composed quantised networks.  TEST INFRASTRUCTURE ONLY (see oracle/fp8_oracle.py).

The reference's quantised ResNet and MobileNetV2 (``quant_setup="all"``) restated as plain functions over the fp32
network's own tensors: every weight layer is ``F.conv2d`` / ``F.linear`` on an oracle-quantised weight, every batch
norm is ``F.batch_norm`` in eval mode, every activation quantiser is ``oracle.fake_quant`` behind an oracle range
estimator -- no module of the product package takes part, so a wiring error there (a wrong tied quantiser, a wrong
residual operand, a missing range update) cannot cancel out.  Device agnostic: with CPU tensors this is the
reference's CPU run (pinned against the real reference's golden ranges and logits in tests/test_oracle_golden.py);
with CUDA tensors it is "the reference as shipped, run with --cuda on the same B200", the thing the fused forward of
the product must equal bit for bit (tests/test_gpu_model_parity.py).

Restated flows (paths relative to /root/reference):
  QuantizationHijacker.forward            quantization/hijacker.py:70-98
  BNFusedHijacker.forward                 quantization/quantized_folded_bn.py:30-56
  QuantizedBlock / QuantizedResNet        models/resnet_quantized.py:14-133
  QuantizedInvertedResidual / MobileNetV2 models/mobilenet_v2_quantized.py:15-92, models/mobilenet_v2.py:72-118
  QuantizedActivationWrapper (tied)       quantization/autoquant_utils.py:125-163
  QuantizationManager.forward             quantization/quantization_manager.py:114-122
Site names are the reference's ``named_modules()`` paths of the ``FPQuantizer`` objects
(``features.2.0.features.1.weight_quantizer.quantizer`` ...), so ranges can be compared by name.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn.functional as F

from . import fp8_oracle as O

_ESTIMATORS = {
    "current_minmax": O.OracleCurrentMinMax,
    "allminmax": O.OracleAllMinMax,
    "running_minmax": O.OracleRunningMinMax,
    "MSE": O.OracleFPMSE,
}


class OracleSites:
    """All quantisation sites of one network: name -> (estimator, quantiser), created on first use with the README's
    parameters (utils/click_options.py:477-510: n_bits 8, per-channel weights, set_maxval, maxval=None)."""

    def __init__(self, mantissa_bits: int, act_estimator="allminmax", weight_estimator="current_minmax", n_bits=8,
                 mse_include_mantissa_bits=False, allow_unsigned=False):
        self.mantissa_bits = mantissa_bits
        self.n_bits = n_bits
        self.act_estimator = act_estimator
        self.weight_estimator = weight_estimator
        self.mse_include_mantissa_bits = mse_include_mantissa_bits
        self.allow_unsigned = allow_unsigned
        self.sites = OrderedDict()
        self.estimating = True    # state estimate_ranges (quantization_manager.py:131-136); False = fix_ranges
        self.trace = None         # optional dict name -> quantiser input (for per-site checks)

    def _site(self, name, per_channel):
        if name not in self.sites:
            qz = O.OracleFPQuantizer(self.n_bits, per_channel, mantissa_bits=self.mantissa_bits, maxval=None,
                                     set_maxval=True, mse_include_mantissa_bits=self.mse_include_mantissa_bits,
                                     allow_unsigned=self.allow_unsigned)
            est = _ESTIMATORS[self.weight_estimator if per_channel else self.act_estimator](per_channel=per_channel,
                                                                                            quantizer=qz)
            self.sites[name] = (est, qz)
        return self.sites[name]

    def declare(self, name, per_channel=False):
        """A quantiser the reference constructs but whose forward this flow never calls (it keeps its default range)."""
        self._site(name, per_channel)

    def fix_ranges(self):
        self.estimating = False

    def weight(self, path, w):
        est, qz = self._site(path + ".weight_quantizer.quantizer", True)
        return O.manager_forward(est, qz, w, self.estimating)       # hijacker.py:88-98

    def act(self, path, x):
        name = path + ".activation_quantizer.quantizer"
        est, qz = self._site(name, False)
        if self.trace is not None:
            self.trace[name] = x.detach().clone()
        return O.manager_forward(est, qz, x, self.estimating)       # quantization_manager.py:114-122

    def act_tied(self, path, x):
        """autoquant_utils.py:147-160: the feeding layer's quantiser, no range update."""
        return self._site(path + ".activation_quantizer.quantizer", False)[1](x)

    def maxvals(self):
        return OrderedDict((n, qz.maxval.detach().reshape(-1).clone()) for n, (_, qz) in self.sites.items())

    def mantissa_widths(self):
        return OrderedDict((n, float(qz.mantissa_bits.reshape(-1)[0])) for n, (_, qz) in self.sites.items())


def _conv_bn_act(S, path, x, conv, bn, act):
    """BNFusedHijacker.forward (quantized_folded_bn.py:30-56) / QuantizationHijacker.forward (hijacker.py:70-98) for
    a convolution: weight quantiser -> F.conv2d on contiguous operands (autoquant_utils.py:34-44) -> eval-mode
    F.batch_norm -> activation -> activation quantiser."""
    w = S.weight(path, conv.weight.detach())
    bias = conv.bias.detach() if conv.bias is not None and bn is None else None
    y = F.conv2d(x.contiguous(), w.contiguous(), bias=bias, stride=conv.stride, padding=conv.padding,
                 dilation=conv.dilation, groups=conv.groups)
    if bn is not None:
        mean = bn.running_mean
        if conv.bias is not None:   # autoquant_utils.py:283-285: a conv bias in front of BN moves into the mean
            mean = mean - conv.bias.detach()
        y = F.batch_norm(y, mean, bn.running_var, bn.weight.detach(), bn.bias.detach(), False, bn.momentum, bn.eps)
    if act == "relu":
        y = torch.relu(y)
    elif act == "relu6":
        y = F.relu6(y)
    return S.act(path, y)


def _linear(S, path, x, lin):
    w = S.weight(path, lin.weight.detach())
    y = F.linear(x.contiguous(), w.contiguous(), bias=lin.bias.detach() if lin.bias is not None else None)
    return S.act(path, y)


@torch.no_grad()
def resnet_forward(S: OracleSites, net, x):
    """QuantizedResNet.forward (models/resnet_quantized.py:126-133) over a torchvision BasicBlock ResNet ``net``."""
    h = _conv_bn_act(S, "features.0", x, net.conv1, net.bn1, "relu")
    h = F.max_pool2d(h, net.maxpool.kernel_size, net.maxpool.stride, net.maxpool.padding)  # features.1: not quantised
    last = None
    for li, layer in enumerate((net.layer1, net.layer2, net.layer3, net.layer4)):
        for bi, blk in enumerate(layer):
            p = f"features.{li + 2}.{bi}"
            # models/resnet_quantized.py:39-46
            if blk.downsample is not None:
                residual = _conv_bn_act(S, p + ".downsample.0", h, blk.downsample[0], blk.downsample[1], None)
            else:
                residual = h
            out = _conv_bn_act(S, p + ".features.0", h, blk.conv1, blk.bn1, "relu")
            out = _conv_bn_act(S, p + ".features.1", out, blk.conv2, blk.bn2, None)
            out = out + residual
            out = torch.relu(out)
            h = S.act(p, out)
            last = p
    h = F.adaptive_avg_pool2d(h, 1)
    h = S.act_tied(last, h)                   # :86-91, tied to features[-1][-1].activation_quantizer
    h = h.view(h.shape[0], -1)
    return _linear(S, "fc", h, net.fc)


@torch.no_grad()
def mobilenetv2_forward(S: OracleSites, net, x):
    """QuantizedMobileNetV2.forward (models/mobilenet_v2_quantized.py:87-92) over the fp32 MobileNetV2 definition
    ``net`` (models/mobilenet_v2.py:72-118 layout: features = [conv_bn, InvertedResidual x17, conv_1x1_bn, AvgPool2d],
    classifier = [Dropout, Linear])."""
    feats = list(net.features)
    h = _conv_bn_act(S, "features.0.0", x, feats[0][0], feats[0][1], "relu6")
    for i, blk in enumerate(feats[1:-2], start=1):
        p = f"features.{i}"
        S.declare(p + ".activation_quantizer.quantizer")   # every QuantizedInvertedResidual owns one (:17)
        mods = list(blk.conv)
        out = h
        j = k = 0
        while j < len(mods):               # quantize_sequential (autoquant_utils.py:292-345): conv [+ bn] [+ act]
            conv, bn = mods[j], mods[j + 1]
            has_act = j + 2 < len(mods) and isinstance(mods[j + 2], torch.nn.ReLU6)
            out = _conv_bn_act(S, f"{p}.conv.{k}", out, conv, bn, "relu6" if has_act else None)
            j += 3 if has_act else 2
            k += 1
        if blk.use_res_connect:            # :21-26
            h = S.act(p, h + out)
        else:
            h = out
    n = len(feats)
    h = _conv_bn_act(S, f"features.{n - 2}.0", h, feats[-2][0], feats[-2][1], "relu6")
    pool = feats[-1]
    h = F.avg_pool2d(h, pool.kernel_size, pool.stride, pool.padding)
    h = S.act_tied(f"features.{n - 2}.0", h)   # tie_activation_quantizers=True (:33-38)
    h = h.view(h.shape[0], -1)
    # classifier[0] is Dropout (identity in eval mode); classifier[1] the QuantLinear
    return _linear(S, "classifier.1", h, net.classifier[1])
