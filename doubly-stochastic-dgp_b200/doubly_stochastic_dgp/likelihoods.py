"""gpflow.likelihoods stand-ins (reference call sites utils.py:88-121).  variational_expectations
(the training path) runs on the device (csrc/lik_adam.cu); predict_mean_and_var / predict_density
are the prediction-side epilogues applied to the (S,N,D) device outputs."""
import math

import numpy as np
from scipy.special import erf

from .params import Parameter, Parameterized


class Likelihood(Parameterized):
    code = -1


class Gaussian(Likelihood):
    code = 0

    def __init__(self, variance=1.0):
        self.variance = Parameter(variance)

    def predict_mean_and_var(self, Fmu, Fvar):
        return Fmu, Fvar + float(self.variance.value)

    def predict_density(self, Fmu, Fvar, Y):
        v = Fvar + float(self.variance.value)
        return -0.5 * math.log(2 * math.pi) - 0.5 * np.log(v) - 0.5 * (Y - Fmu) ** 2 / v


class MultiClass(Likelihood):
    """MultiClass(K) with the default RobustMax(epsilon=1e-3) link."""
    code = 1

    def __init__(self, num_classes, epsilon=1e-3):
        self.num_classes = int(num_classes)
        if epsilon != 1e-3:
            raise NotImplementedError("RobustMax epsilon is fixed at 1e-3 in the device kernel")
        self.epsilon = float(epsilon)

    def _prob_is_largest(self, Yi, mu, var):
        gh_x, gh_w = np.polynomial.hermite.hermgauss(20)
        K = self.num_classes
        oh = np.eye(K)[Yi.astype(int).reshape(-1)]
        mu_sel = np.sum(oh * mu, 1, keepdims=True)
        var_sel = np.sum(oh * var, 1, keepdims=True)
        X = mu_sel + np.sqrt(2.0 * np.clip(var_sel, 1e-10, np.inf)) * gh_x[None, :]
        dist = (X[:, None, :] - mu[:, :, None]) / np.sqrt(np.clip(var, 1e-10, np.inf))[:, :, None]
        cdfs = 0.5 * (1.0 + erf(dist / math.sqrt(2.0))) * (1 - 2e-4) + 1e-4
        off = (1.0 - oh)[:, :, None]
        cdfs = cdfs * off + (1.0 - off)
        return np.prod(cdfs, 1) @ (gh_w / math.sqrt(math.pi))[:, None]

    def predict_mean_and_var(self, Fmu, Fvar):
        shp = Fmu.shape
        mu, var = Fmu.reshape(-1, shp[-1]), Fvar.reshape(-1, shp[-1])
        ps = np.concatenate([self._prob_is_largest(np.full(mu.shape[0], k), mu, var)
                             for k in range(self.num_classes)], 1)
        return ps.reshape(shp), (ps - ps ** 2).reshape(shp)

    def predict_density(self, Fmu, Fvar, Y):
        shp = Fmu.shape
        mu, var = Fmu.reshape(-1, shp[-1]), Fvar.reshape(-1, shp[-1])
        Yb = np.broadcast_to(Y, shp[:-1] + (1,)).reshape(-1)
        p = self._prob_is_largest(Yb, mu, var)
        eps = self.epsilon
        return np.log(p * (1 - eps) + (1.0 - p) * (eps / (self.num_classes - 1.0))).reshape(shp[:-1] + (1,))
