"""Golden vectors at the sizes SURVEY.md section 8(c) asks for, produced by the REAL reference on the CPU:

    python tests/golden/make_golden_survey_sizes.py          (build container only: needs /root/reference)

  * quantiser: M in 1..7 x sign_bits in {0,1} x sigma in {1e-3, 1, 1e3} on N(0,1)*sigma, seed 10 (README.md:64),
    per tensor n = 2^20 and per channel [C, inner] with inner in {1, 9, 27, 147, 576} (the weight-row lengths of the two
    workloads).  A million-element output per case would be ~100 MB of fixtures, so what is stored per case is the
    SHA-256 of the reference's output bit patterns (NaN canonicalised) and of its exponent-code / mantissa-integer
    planes, plus the range -- the inputs are regenerated from the seed.  tests/test_oracle_golden.py re-derives the
    digests with the oracle (this is the pin of the oracle at full size); the GPU test compares the kernel with the
    oracle on the same tensors.
  * FP_MSE_Estimator (range_estimators.py:285-369) with the internal mantissa sweep on a per-tensor [8,64,56,56]
    activation and a per-channel [128,64,3,3] weight: the [6,111,C] MSE tables, the voted mantissa width, the selected
    ranges (small enough to store outright).
For every case the oracle restatement is run on the same input and must equal the reference bit for bit before anything
is written."""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import fp8_oracle as O  # noqa: E402
from oracle.reference_loader import load_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
INNERS = (1, 9, 27, 147, 576)
SIGMAS = (1e-3, 1.0, 1e3)


def digest(t):
    """SHA-256 of the fp32 bit patterns with every NaN replaced by the canonical quiet NaN."""
    t = t.detach().contiguous().float()
    t = torch.where(torch.isnan(t), torch.full_like(t, float("nan")), t)
    return hashlib.sha256(t.numpy().tobytes()).hexdigest()


def quantizer_case(M, sb, sigma, inner):
    """The seeded input of one case: per tensor [2^20] (inner == 0) or per channel [C, inner] with ~2^20 elements."""
    seed = 10 + 1000 * M + 100 * sb + 10 * SIGMAS.index(sigma) + (INNERS.index(inner) + 1 if inner else 0)
    g = torch.Generator().manual_seed(seed)
    if inner == 0:
        return torch.randn(1 << 20, generator=g) * sigma, False
    C = max(32, ((1 << 20) // inner) // 32 * 32)
    C = min(C, 4096)
    x = torch.randn(C, inner, generator=g) * sigma * torch.linspace(0.25, 4.0, C).view(-1, 1)
    return x, True


def mse_case(key):
    g = torch.Generator().manual_seed(12)
    if key == "act_8x64x56x56":
        x = torch.randn(8, 64, 56, 56, generator=g)
        return torch.relu(x) + 0.05 * torch.randn(8, 64, 56, 56, generator=g), False
    return torch.randn(128, 64, 3, 3, generator=g) * 0.05, True


def main():
    torch.set_num_threads(1)
    R = load_reference()
    RE = R.range_estimators
    out = {}
    n = 0
    for M in range(1, 8):
        for sb in (0, 1):
            for sigma in SIGMAS:
                for inner in (0,) + INNERS:
                    x, pc = quantizer_case(M, sb, sigma, inner)
                    q = R.FPQuantizer(8, per_channel=pc, mantissa_bits=M, set_maxval=True)
                    q.sign_bits = sb
                    mn, mx = O.minmax(x, pc)
                    q.set_quant_range(mn * 0.9, mx * 0.9)
                    y_ref = q(x)
                    y, e, qq = O.fake_quant(x, 8, q.maxval, q.mantissa_bits, sb, return_codes=True)
                    assert digest(y) == digest(y_ref), ("oracle != reference", M, sb, sigma, inner)
                    key = f"q_M{M}_s{sb}_sig{SIGMAS.index(sigma)}_in{inner}"
                    out[key] = np.array([digest(y_ref), digest(e), digest(qq)])
                    out[key + "_maxval_digest"] = np.array(digest(q.maxval))
                    out[key + "_shape"] = np.array(x.shape)
                    n += 1
    print(n, "quantiser cases")
    for key in ("act_8x64x56x56", "weight_128x64x3x3"):
        x, pc = mse_case(key)
        q = R.FPQuantizer(8, per_channel=pc, mantissa_bits=4, set_maxval=True, mse_include_mantissa_bits=True)
        est = RE.FP_MSE_Estimator(per_channel=pc, quantizer=q)
        oq = O.OracleFPQuantizer(8, per_channel=pc, mantissa_bits=4, set_maxval=True, mse_include_mantissa_bits=True)
        oest = O.OracleFPMSE(per_channel=pc, quantizer=oq)
        mn, mx = est(x)
        omn, omx = oest(x)
        assert digest(mx.float()) == digest(omx.float()) and digest(est.mses) == digest(oest.mses), key
        assert float(q.mantissa_bits) == float(oq.mantissa_bits)
        out["mse_" + key + "_mses"] = est.mses.numpy()
        out["mse_" + key + "_grid"] = est.search_grid.numpy()
        out["mse_" + key + "_xmax"] = mx.float().numpy()
        out["mse_" + key + "_best_m"] = np.array(float(q.mantissa_bits))
        print(key, "best M", float(q.mantissa_bits))
    np.savez_compressed(os.path.join(OUT, "survey_sizes.npz"), **out)
    print("survey_sizes.npz", os.path.getsize(os.path.join(OUT, "survey_sizes.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
