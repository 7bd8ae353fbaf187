"""Kernel descriptors standing in for gpflow.kernels.RBF / Matern52 / White and `k1 + k2` (reference call sites
layers.py:161,171,184,213; `.input_dim` layer_initializations.py:27-28; `RBF(...) + White(...)` demos/run_regression.py:65-66,
demos/demo_step_function.ipynb:111).  The Gram arithmetic itself lives in the CUDA kernels (csrc/layer_simt.cu gram_stage,
csrc/layer_tc.cu, csrc/small_matrix.cu k_prepA)."""
import copy

import numpy as np

from .params import Parameter, Parameterized


class Kernel(Parameterized):
    def __add__(self, other):
        return Sum([self, other])


class Stationary(Kernel):
    code = -1

    def __init__(self, input_dim, variance=1.0, lengthscales=None, ARD=False):
        self.input_dim = int(input_dim)
        if lengthscales is None:
            lengthscales = np.ones(input_dim) if ARD else 1.0
        ls = np.asarray(lengthscales, dtype=np.float64)
        self.ARD = bool(ARD or (ls.ndim > 0 and ls.size > 1))
        if self.ARD:
            ls = np.broadcast_to(ls, (self.input_dim,)).copy()
        else:
            ls = ls.reshape(())
        self.variance = Parameter(variance)
        self.lengthscales = Parameter(ls)
        self.white_variance = None          # set by Sum([stationary, White])


class RBF(Stationary):
    code = 0


SquaredExponential = RBF


class Matern52(Stationary):
    code = 1


class White(Kernel):
    """gpflow.kernels.White: K(X) = variance I, K(X, X2) = 0, Kdiag = variance."""
    def __init__(self, input_dim, variance=1.0):
        self.input_dim = int(input_dim)
        self.variance = Parameter(variance)


def Sum(kern_list):
    """gpflow.kernels.Sum for the one combination the reference uses: a stationary kernel plus a White term.  Returns a copy of
    the stationary descriptor carrying `white_variance` (the device adds it to diag(Kuu) and to Kdiag; K(Z, X) has no White
    part).  `.kern_list` keeps the gpflow attribute."""
    flat = []
    for k in kern_list:
        flat += getattr(k, "kern_list", [k])
    stat = [k for k in flat if isinstance(k, Stationary)]
    whites = [k for k in flat if isinstance(k, White)]
    if len(stat) != 1 or len(stat) + len(whites) != len(flat):
        raise NotImplementedError("Sum kernels: exactly one stationary kernel (RBF / Matern52) plus White terms are on the "
                                  "accelerated path")
    base = stat[0]
    out = copy.copy(base)
    out.variance = Parameter(base.variance._value)
    out.lengthscales = Parameter(base.lengthscales._value)
    w = sum(float(k.variance._value) for k in whites) + (float(base.white_variance._value) if base.white_variance is not None else 0.0)
    out.white_variance = Parameter(w)
    object.__setattr__(out, "kern_list", flat)
    return out
