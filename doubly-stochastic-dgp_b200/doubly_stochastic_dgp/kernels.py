"""Kernel descriptors standing in for gpflow.kernels.RBF / Matern52 (reference call sites
layers.py:161,171,184,213; `.input_dim` layer_initializations.py:27-28).  The Gram arithmetic
itself lives in the CUDA kernels (csrc/layer_simt.cu gram_stage, csrc/small_matrix.cu k_prepA)."""
import numpy as np

from .params import Parameter, Parameterized


class Stationary(Parameterized):
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

    def __add__(self, other):
        raise NotImplementedError("Sum kernels (e.g. RBF + White) are not on the accelerated path yet "
                                  "(SURVEY.md section 8(f) rank 4)")


class RBF(Stationary):
    code = 0


SquaredExponential = RBF


class Matern52(Stationary):
    code = 1
