"""Host mirror of doubly_stochastic_dgp/utils.py (reference).  `reparameterize` (utils.py:22-41, diagonal
branch) is fused into the forward kernel's epilogue on the device (csrc/layer_simt.cu); the function here
is the same formula for host-side use on returned arrays.  `BroadcastingLikelihood` (utils.py:54-121)
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
        raise NotImplementedError("full_cov reparameterisation is not on the accelerated path yet "
                                  "(SURVEY.md section 8(f) rank 3)")
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

    def predict_mean_and_var(self, Fmu, Fvar):
        raise NotImplementedError("the likelihood epilogues run on the device: use model.predict_y(Xnew, num_samples)")

    def predict_density(self, Fmu, Fvar, Y):
        raise NotImplementedError("the likelihood epilogues run on the device: use "
                                  "model.predict_density(Xnew, Ynew, num_samples)")
