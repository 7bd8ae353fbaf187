"""Host mirror of doubly_stochastic_dgp/layers.py: `SVGP_Layer` with the reference constructor
(layers.py:122-165) and attribute names (q_mu, q_sqrt, feature.Z, kern, mean_function, num_outputs, white,
num_inducing).  The layer is a parameter container; conditional_ND / conditional_SND /
sample_from_conditional / KL (layers.py:46-119,178-246) are evaluated by the owning model's device context
(a layer used stand-alone builds a private one-layer context)."""
import numpy as np

from . import settings
from .params import Parameter, Parameterized


class InducingPoints(Parameterized):
    """gpflow.features.InducingPoints stand-in: holds Z (layers.py:153)."""
    def __init__(self, Z):
        self.Z = Parameter(np.asarray(Z, dtype=np.float64))

    def __len__(self):
        return self.Z.shape[0]


class Layer(Parameterized):
    def __init__(self, input_prop_dim=None, **kwargs):
        # layers.py:39-44: the first input_prop_dim input columns are concatenated in front of the layer's samples / mean
        # (zeros in front of the variance), layers.py:105-117 -- done by the forward kernel's epilogue on the device
        self.input_prop_dim = input_prop_dim


class SVGP_Layer(Layer):
    def __init__(self, kern, Z, num_outputs, mean_function, white=False, input_prop_dim=None, **kwargs):
        Layer.__init__(self, input_prop_dim, **kwargs)
        Z = np.asarray(Z, dtype=np.float64)
        self.num_inducing = Z.shape[0]
        self.q_mu = Parameter(np.zeros((self.num_inducing, num_outputs)))                     # layers.py:146-147
        self.q_sqrt = Parameter(np.tile(np.eye(self.num_inducing)[None], [num_outputs, 1, 1]))  # layers.py:149-151
        self.feature = InducingPoints(Z)
        self.kern = kern
        self.mean_function = mean_function
        self.num_outputs = num_outputs
        self.white = white
        self._model = None
        if not self.white:
            # layers.py:160-163 "initialize to prior": q_sqrt = chol(K(Z,Z) + jitter I).  This runs once at
            # construction on M x M host data (NumPy in the reference too: np.linalg.cholesky, layers.py:162).
            Ku = _host_K(kern, Z)
            Lu = np.linalg.cholesky(Ku + np.eye(Z.shape[0]) * settings.jitter)
            self.q_sqrt = np.tile(Lu[None], [num_outputs, 1, 1])

    # ---- evaluation through the device
    def _ctx_model(self):
        if self._model is None:
            from .dgp import _single_layer_model
            object.__setattr__(self, "_model", _single_layer_model(self))
        return self._model

    def conditional_ND(self, X, full_cov=False):
        """layers.py:178-219: mean (N, num_outputs) and var (N, num_outputs), or (N, N, num_outputs) with full_cov."""
        return self._ctx_model()._layer_conditional(self, np.asarray(X), full_cov=full_cov)

    def conditional_SND(self, X, full_cov=False):
        """layers.py:52-74: independent over the S samples; var is (S,N,D_out), or (S,N,N,D_out) with full_cov."""
        X = np.asarray(X)
        S, N, D = X.shape
        if full_cov:        # tf.map_fn over the samples (layers.py:66-69)
            ms, vs = zip(*[self.conditional_ND(X[s], full_cov=True) for s in range(S)])
            return np.stack(ms), np.stack(vs)
        m, v = self.conditional_ND(X.reshape(S * N, D))
        return m.reshape(S, N, self.num_outputs), v.reshape(S, N, self.num_outputs)

    def sample_from_conditional(self, X, z=None, full_cov=False):
        """layers.py:76-119: returns samples, mean (S,N,num_outputs) and var (S,N,num_outputs) or (S,N,N,num_outputs);
        conditional and draw both run on the device (z=None: host standard normals handed to the kernel)."""
        X = np.asarray(X)
        S, N, D = X.shape
        Do = self.num_outputs
        z = np.random.randn(S, N, Do) if z is None else np.asarray(z, dtype=np.float64).reshape(S, N, Do)
        model = self._ctx_model()
        if full_cov:
            outs = [model._layer_propagate(self, X[s], S=1, full_cov=True, zs=[z[s][None]]) for s in range(S)]
            samples, mean, var = (np.concatenate([o[0][0] for o in outs]), np.concatenate([o[1][0] for o in outs]),
                                  np.concatenate([o[2][0] for o in outs]))
        else:
            Fs, Fm, Fv = model._layer_propagate(self, X.reshape(S * N, D), S=1, zs=[z.reshape(1, S * N, Do)])
            samples, mean, var = Fs[0].reshape(S, N, Do), Fm[0].reshape(S, N, Do), Fv[0].reshape(S, N, Do)
        if self.input_prop_dim:          # layers.py:105-117 (inside a DGP the forward kernel's epilogue does this)
            Xp = np.asarray(X, dtype=np.float64)[:, :, :self.input_prop_dim]
            samples = np.concatenate([Xp, samples], 2)
            mean = np.concatenate([Xp, mean], 2)
            zeros = np.zeros((S, N, N, self.input_prop_dim)) if full_cov else np.zeros_like(Xp)
            var = np.concatenate([zeros, var], 3 if full_cov else 2)
        return samples, mean, var

    def KL(self):
        """layers.py:221-246."""
        return self._ctx_model()._layer_KL(self)


def _host_K(kern, Z):
    """K(Z,Z) for the construction-time q_sqrt initialisation only (M x M, once); a White term adds to the diagonal."""
    K = _host_K_stationary(kern, Z)
    if getattr(kern, "white_variance", None) is not None:
        K = K + float(kern.white_variance.value) * np.eye(Z.shape[0])
    return K


def _host_K_stationary(kern, Z):
    ls = np.asarray(kern.lengthscales.value, dtype=np.float64)
    d = (Z[:, None, :] - Z[None, :, :]) / ls
    r2 = np.sum(d * d, -1)
    var = float(kern.variance.value)
    if kern.code == 0:
        return var * np.exp(-0.5 * r2)
    r = np.sqrt(r2 + 1e-12)
    return var * (1 + np.sqrt(5.0) * r + 5.0 / 3.0 * r * r) * np.exp(-np.sqrt(5.0) * r)
