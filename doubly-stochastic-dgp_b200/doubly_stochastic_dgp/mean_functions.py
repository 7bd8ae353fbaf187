"""gpflow.mean_functions stand-ins (reference call site layers.py:219)."""
import numpy as np

from .params import Parameter, Parameterized


class MeanFunction(Parameterized):
    code = -1


class Zero(MeanFunction):
    code = 0


class Identity(MeanFunction):
    code = 1


class Linear(MeanFunction):
    """X @ A + b; fixed inside a DGP (layer_initializations.py:41-42 `mf.set_trainable(False)`)."""
    code = 2

    def __init__(self, A=None, b=None):
        A = np.ones((1, 1)) if A is None else np.asarray(A, dtype=np.float64)
        b = np.zeros(A.shape[1]) if b is None else np.asarray(b, dtype=np.float64)
        self.A = Parameter(A, trainable=False)
        self.b = Parameter(b, trainable=False)

    def set_trainable(self, flag):
        if flag:
            raise NotImplementedError("trainable Linear mean functions are not supported on the accelerated path")
