"""Helpers for the -m gpu parity tests: build the product model (ctypes -> libdsdgp.so) from a synthetic problem."""
import numpy as np

from doubly_stochastic_dgp import settings
from doubly_stochastic_dgp.dgp import DGP_Base
from doubly_stochastic_dgp.kernels import RBF, Matern52
from doubly_stochastic_dgp.layers import SVGP_Layer
from doubly_stochastic_dgp.likelihoods import Gaussian, MultiClass
from doubly_stochastic_dgp.mean_functions import Identity, Linear, Zero


from workloads import build_model  # noqa: F401,E402


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300))


def record(test, **vals):
    """Append measured parity errors to gpurun_out/parity_errors.jsonl (when run on the GPU box through gpurun), so the
    tolerances stated in the tests can be compared with what the hardware actually delivers (summarised in DESIGN.md)."""
    import json
    import os
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_errors.jsonl"), "a") as f:
            f.write(json.dumps(dict(test=test, **{k: float(v) for k, v in vals.items()})) + "\n")
    except OSError:
        pass
