"""Host mirror of doubly_stochastic_dgp/utils.py (reference).  `reparameterize` (utils.py:22-51) runs on the device:
the diagonal branch is fused into the forward kernels' epilogue (csrc/layer_simt.cu, csrc/layer_tc.cu), the full_cov
branch is the batched Cholesky draw of csrc/full_cov.cu (k_fc_chol_draw); both are reached through
`model.propagate(..., zs=...)` / `layer.sample_from_conditional(X, z)`.  The function here is the diagonal formula as a
convenience on host arrays that were already returned (it is not called by any product path).  `BroadcastingLikelihood` (utils.py:54-121)
keeps the wrapper object so that `model.likelihood.likelihood.variance` (and the reference demo's
`model.likelihood.variance`) both work."""
import numpy as np

from . import settings
from .likelihoods import Gaussian
from .params import Parameter, Parameterized


def reparameterize(mean, var, z, full_cov=False):
    if var is None:
        return mean
    if full_cov:
        raise NotImplementedError("the full_cov draw runs on the device: use model.propagate(X, full_cov=True, zs=...) or "
                                  "layer.sample_from_conditional(X, z, full_cov=True)")
    return mean + z * (var + settings.jitter) ** 0.5


class BroadcastingLikelihood(Parameterized):
    def __init__(self, likelihood):
        object.__setattr__(self, "likelihood", likelihood)
        object.__setattr__(self, "needs_broadcasting", not isinstance(likelihood, Gaussian))

    def __setattr__(self, name, value):
        # demos/run_regression.py:74 sets `model.likelihood.variance` on the wrapper; forward it
        if name == "variance" and hasattr(self.likelihood, "variance"):
            self.likelihood.variance = value
        else:
            Parameterized.__setattr__(self, name, value)

    @property
    def variance(self):
        return self.likelihood.variance

    def parameters(self):
        return self.likelihood.parameters()

    # ---- utils.py:88-121: the wrapped likelihood applied to (S,N,D) marginals, Y of shape (N,D_y).  The arithmetic is the
    # device epilogue of csrc/lik_adam.cu (dsdgp_likelihood_apply), reached through the owning model's context; a wrapper
    # used stand-alone builds a private one-layer context of the right width.
    def _ctx(self, Fmu):
        Fmu = np.asarray(Fmu)
        if Fmu.ndim != 3:
            raise ValueError(f"expected (S, N, D) marginals, got shape {Fmu.shape}")
        S, N, D = Fmu.shape
        model = self.__dict__.get("_model")
        if model is None or model.layers[-1].num_outputs != D:
            from .dgp import _likelihood_model
            model = _likelihood_model(self.likelihood, D)
            if self.__dict__.get("_model") is None:
                object.__setattr__(self, "_private_model", model)
        return model._ensure_ctx(N, S) if len(model.layers) > 1 else model._ensure_ctx(S * N, 1)

    def variational_expectations(self, Fmu, Fvar, Y):
        return self._ctx(Fmu).likelihood_apply(0, Fmu, Fvar, Y).astype(np.float64)

    def predict_mean_and_var(self, Fmu, Fvar):
        m, v = self._ctx(Fmu).likelihood_apply(1, Fmu, Fvar)
        return m.astype(np.float64), v.astype(np.float64)

    def predict_density(self, Fmu, Fvar, Y):
        return self._ctx(Fmu).likelihood_apply(2, Fmu, Fvar, Y).astype(np.float64)
