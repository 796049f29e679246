"""GPU: next row f5 -- the empirical flow of compute_quant_error.py (workloads.compute_quant_error_empirical) against
the real reference's golden vectors (tests/golden/quant_error.npz, made by tests/golden/make_golden_quant_error.py).
The same flow runs on the host simulation in tests/test_host_sim_models.py.  Kept in its own module, collected after
the other GPU modules: it was written after round 1's GPU budget was spent, so its first run on a device is the
driver's round-end run."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_next_row_f5_empirical_quant_error_flow_vs_reference_golden():
    """compute_quant_error.py:18-57, empirical half (workloads.compute_quant_error_empirical), vs the real reference run
    on the CPU (tests/golden/make_golden_quant_error.py): four sample distributions x the script's five formats.  The
    loss curves are flat near their minimum, so the chosen threshold may be another point of the same plateau: it must
    be optimal for the REFERENCE's loss; the empirical errors agree to the backends' ulp-level differences."""
    from fp8_quantization_b200 import workloads

    g = load_golden("quant_error.npz")
    ncand = int(g["num_candidates"])
    for name in g["names"]:
        x, y = torch.from_numpy(g[f"{name}_x"]).to(DEV), torch.from_numpy(g[f"{name}_y"]).to(DEV)
        rows = workloads.compute_quant_error_empirical(x, y, n_bits=8, num_candidates=ncand)
        assert [r["exp_bits"] for r in rows] == list(g["exp_bits"])
        for r in rows:
            key = f"{name}_e{r['exp_bits']}"
            ref_loss = g[key + "_loss"][0]
            step = float(g[key + "_xmax"][0]) / max(int(np.argmin(ref_loss)), 1)
            ours_i = int(round(r["range_max"] / step))
            assert 1 <= ours_i <= ncand and ref_loss[ours_i] <= ref_loss.min() * (1 + 1e-3), (key, ours_i)
            assert (r["range_min"] == 0.0) == (float(g[key + "_xmin"][0]) == 0.0), key
            np.testing.assert_allclose(r["mse"], float(g[key + "_mse"]), rtol=5e-3, err_msg=key)
            np.testing.assert_allclose(r["dot_prod_mse"], float(g[key + "_dot"]), rtol=5e-3, err_msg=key)
