"""Generates tests/golden/quant_error.npz from the REAL reference (build container only; needs /root/reference):
the empirical half of compute_quant_error.py:18-57 -- for three sample distributions and the five formats of the
script (E5M2, E4M3, E3M4, E2M5 with FPQuantizer, E0 with SymmetricUniformQuantizer): the range found by
estimate_range_line_search (range_estimators.py:372-379), estimate_rounding_error_empirical and
estimate_dot_prod_error_empirical (quant_error_estimator.py:68-89), at sizes the CPU reference finishes in a minute
(8192 samples, 200 candidates instead of 5 M / 1000).  The analytic half (scipy integrals) is out of scope.

    python tests/golden/make_golden_quant_error.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.reference_loader import load_reference  # noqa: E402

load_reference()
from quantization.quant_error_estimator import (estimate_dot_prod_error_empirical,  # noqa: E402
                                                estimate_rounding_error_empirical)
from quantization.quantizers.fp8_quantizer import FPQuantizer  # noqa: E402
from quantization.quantizers.uniform_quantizers import SymmetricUniformQuantizer  # noqa: E402
from quantization.range_estimators import LineSearchEstimator  # noqa: E402

N, NCAND = 8192, 200
rng = np.random.default_rng(10)
samples = {
    "uniform": rng.uniform(-1.0, 1.0, N),
    "gauss": np.clip(rng.standard_normal(N), -10.0, 10.0),
    "student_t": np.clip(rng.standard_t(8.0, N), -100.0, 100.0),
    "half_gauss": np.abs(rng.standard_normal(N)),          # one-sided: x_min stays 0
}
out = {"names": np.array(list(samples)), "num_candidates": np.int64(NCAND), "exp_bits": np.array([5, 4, 3, 2, 0])}
torch.set_num_threads(1)
for name, s in samples.items():
    x = torch.tensor(s.astype(np.float32))
    y = torch.tensor(rng.permutation(s).astype(np.float32))
    out[name + "_x"], out[name + "_y"] = x.numpy(), y.numpy()
    for eb in (5, 4, 3, 2, 0):
        M = 8 - 1 - eb
        mk = (lambda: FPQuantizer(n_bits=8, mantissa_bits=M, set_maxval=True)) if eb > 0 else \
            (lambda: SymmetricUniformQuantizer(n_bits=8))
        quant = mk()
        est = LineSearchEstimator(quantizer=quant, num_candidates=NCAND)
        xmin, xmax = est.forward(x)
        mse = estimate_rounding_error_empirical(x, quant, xmin, xmax)
        qx, qy = mk(), mk()
        dot = estimate_dot_prod_error_empirical(x, y, qx, qy, xmin, xmax, xmin, xmax)
        key = f"{name}_e{eb}"
        out[key + "_xmin"], out[key + "_xmax"] = xmin.numpy(), xmax.numpy()
        out[key + "_loss"] = est.loss_array.astype(np.float64)
        out[key + "_mse"], out[key + "_dot"] = np.float64(mse), np.float64(dot)
        print(key, float(xmin[0]), float(xmax[0]), mse, dot, flush=True)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "quant_error.npz"), **out)
