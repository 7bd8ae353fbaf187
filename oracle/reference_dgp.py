"""
ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

A float64 CPU restatement (torch-CPU, so that autograd supplies the gradient
oracle the reference gets from TensorFlow autodiff) of the doubly-stochastic
DGP hot path of UCL-SML/Doubly-Stochastic-DGP.  Every function cites the
reference file:line it follows (paths relative to /root/reference).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` leg may import this module, and only as the checker (or
as the timed CPU baseline) -- never from the product package.

PARITY STATUS: "parity unpinned at the GPflow/TF boundary".  The reference
runs on gpflow==1.1.1 + tensorflow==1.8 (README.md:4), neither of which is in
/root/reference nor installable here (Python 3.12, no network), and the
reference's tests hold no golden vectors (they are differential tests against
GPflow's own SVGP/GPR, tests/test_dgp.py:66-117).  The GPflow pieces (kernels,
Kuu/Kuf, likelihoods, mean functions, transforms) are therefore restated from
their published definitions (GPflow 1.1.1: kernels.py `Stationary.square_dist`,
`RBF.K`, `Matern52.K`; likelihoods.py `Gaussian`, `MultiClass`/`RobustMax`,
`Bernoulli`; mean_functions.py) and the restatement is re-pinned by the same
identities the reference's tests use, evaluated against an INDEPENDENT
closed-form SVGP / GPR written with different linear algebra
(oracle/closed_form.py; see tests/test_oracle_identities.py):
  I1  L=1 DGP == SVGP                       (tests/test_dgp.py:66-117)
  I2  L=2 DGP with a 1e-24-variance inner layer == SVGP   (same, L=2 branch)
  I6  NatGrad gamma=1 on the last layer == SGPR optimum  (tests/test_collapsed.py:57-104)
  I8  reparameterize literal formula        (tests/test_utils.py:181-206)
Since round 2 the restatement of the REFERENCE-OWNED files (dgp.py, layers.py, utils.py,
layer_initializations.py) is additionally pinned by outputs of those files themselves: they are
imported unmodified from /root/reference and executed over a torch stand-in for the TF ops / GPflow
classes they touch (oracle/tf_gpflow_shim.py); the fixtures tests/golden/refshim_*.npz (generator:
tests/golden/make_from_reference_shim.py) hold what the reference computed -- propagate, KL, ELBO,
predict_*, full_cov, DGP_Quad, input propagation, Sum(RBF, White), Bernoulli broadcasting -- and
tests/test_refshim_cpu.py requires this module to reproduce them to 1e-9.  GPflow itself is still
not run anywhere, hence "unpinned at the GPflow/TF boundary" stands.

Two execution modes:
  faithful=True   materialises exactly what the reference graph materialises
                  (S-tiled first layer dgp.py:63; D_out-tiled Ku/Lu/A
                  layers.py:173-174,192; B = SK @ A_tiled layers.py:204).  This
                  is the form timed as the CPU baseline.
  faithful=False  same arithmetic without the replicated temporaries (used for
                  the big parity cases so the oracle finishes in seconds).
Both give identical results to rounding (checked in tests).
"""
import math

import numpy as np
import torch

DT = torch.float64
_LOG2PI = math.log(2.0 * math.pi)


def _t(x):
    if isinstance(x, torch.Tensor):
        return x.to(DT)
    return torch.as_tensor(np.asarray(x, dtype=np.float64), dtype=DT)


class Settings:
    """gpflow.settings stand-in: numerics.jitter_level (default 1e-6), float64."""
    jitter = 1e-6


settings = Settings()


# ----------------------------------------------------------------------------
# GPflow 1.1.1 kernels (restated; call sites layers.py:161,171,184,209,213)
# ----------------------------------------------------------------------------
class Stationary:
    def __init__(self, input_dim, variance=1.0, lengthscales=None, ARD=False):
        self.input_dim = int(input_dim)
        if lengthscales is None:
            lengthscales = np.ones(input_dim) if ARD else 1.0
        ls = np.asarray(lengthscales, dtype=np.float64)
        self.ARD = bool(ARD or ls.ndim > 0 and ls.size > 1)
        self.variance = _t(variance).clone()
        self.lengthscales = _t(ls).clone()

    def parameters(self):
        return [self.variance, self.lengthscales]

    def square_dist(self, X, X2=None):
        # gpflow Stationary.square_dist: scale, then -2XX2^T + |X|^2 + |X2|^2 (no clamp)
        X = X / self.lengthscales
        Xs = torch.sum(X * X, 1)
        if X2 is None:
            return -2.0 * X @ X.T + Xs[:, None] + Xs[None, :]
        X2 = X2 / self.lengthscales
        X2s = torch.sum(X2 * X2, 1)
        return -2.0 * X @ X2.T + Xs[:, None] + X2s[None, :]

    def Kdiag(self, X):
        return self.variance * torch.ones(X.shape[0], dtype=DT)


class RBF(Stationary):
    def K(self, X, X2=None):
        return self.variance * torch.exp(-self.square_dist(X, X2) / 2.0)


class Matern52(Stationary):
    def K(self, X, X2=None):
        r = torch.sqrt(self.square_dist(X, X2) + 1e-12)
        s5 = math.sqrt(5.0)
        return self.variance * (1.0 + s5 * r + 5.0 / 3.0 * r * r) * torch.exp(-s5 * r)


class White:
    """gpflow.kernels.White: K(X) = variance I, K(X, X2) = 0, Kdiag = variance."""
    def __init__(self, input_dim, variance=1.0):
        self.input_dim = int(input_dim)
        self.variance = _t(variance).clone()

    def parameters(self):
        return [self.variance]

    def K(self, X, X2=None):
        if X2 is None:
            return self.variance * torch.eye(X.shape[0], dtype=DT)
        return torch.zeros(X.shape[0], X2.shape[0], dtype=DT)

    def Kdiag(self, X):
        return self.variance * torch.ones(X.shape[0], dtype=DT)


class Sum:
    """gpflow.kernels.Sum (`k1 + k2`, demos/run_regression.py:65-66): K and Kdiag add."""
    def __init__(self, kern_list):
        self.kern_list = list(kern_list)
        self.input_dim = self.kern_list[0].input_dim

    @property
    def variance(self):
        return self.kern_list[0].variance

    def parameters(self):
        return [p for k in self.kern_list for p in k.parameters()]

    def K(self, X, X2=None):
        return sum(k.K(X, X2) for k in self.kern_list)

    def Kdiag(self, X):
        return sum(k.Kdiag(X) for k in self.kern_list)


# ----------------------------------------------------------------------------
# GPflow mean functions (call site layers.py:219)
# ----------------------------------------------------------------------------
class Zero:
    def __call__(self, X):
        return torch.zeros(X.shape[0], 1, dtype=DT)


class Identity:
    def __call__(self, X):
        return X


class Linear:
    def __init__(self, A, b=None):
        self.A = _t(A)
        self.b = _t(np.zeros(self.A.shape[1]) if b is None else b)

    def __call__(self, X):
        return X @ self.A + self.b


# ----------------------------------------------------------------------------
# GPflow likelihoods (call sites utils.py:88-121)
# ----------------------------------------------------------------------------
def _gh(n):
    x, w = np.polynomial.hermite.hermgauss(n)
    return _t(x), _t(w)


class Gaussian:
    def __init__(self, variance=1.0):
        self.variance = _t(variance).clone()

    def parameters(self):
        return [self.variance]

    def variational_expectations(self, Fmu, Fvar, Y):
        return (-0.5 * _LOG2PI - 0.5 * torch.log(self.variance)
                - 0.5 * ((Y - Fmu) ** 2 + Fvar) / self.variance)

    def predict_mean_and_var(self, Fmu, Fvar):
        return Fmu, Fvar + self.variance

    def predict_density(self, Fmu, Fvar, Y):
        v = Fvar + self.variance
        return -0.5 * _LOG2PI - 0.5 * torch.log(v) - 0.5 * (Y - Fmu) ** 2 / v


class MultiClass:
    """gpflow MultiClass(K) with the default RobustMax(K, epsilon=1e-3) link."""
    num_gauss_hermite_points = 20

    def __init__(self, num_classes, epsilon=1e-3):
        self.num_classes = int(num_classes)
        self.epsilon = float(epsilon)

    def parameters(self):
        return []

    def _prob_is_largest(self, Y, mu, var):
        # RobustMax.prob_is_largest: Y (R,) int, mu/var (R,K)
        gh_x, gh_w = _gh(self.num_gauss_hermite_points)
        K = self.num_classes
        Yi = Y.reshape(-1).long()
        oh = torch.nn.functional.one_hot(Yi, K).to(DT)
        mu_sel = torch.sum(oh * mu, 1, keepdim=True)
        var_sel = torch.sum(oh * var, 1, keepdim=True)
        X = mu_sel + torch.sqrt(2.0 * torch.clamp(var_sel, 1e-10, np.inf)) * gh_x[None, :]   # R,H
        dist = (X[:, None, :] - mu[:, :, None]) / torch.sqrt(torch.clamp(var, 1e-10, np.inf))[:, :, None]
        cdfs = 0.5 * (1.0 + torch.erf(dist / math.sqrt(2.0)))
        cdfs = cdfs * (1 - 2e-4) + 1e-4
        oh_off = (1.0 - oh)[:, :, None]
        cdfs = cdfs * oh_off + (1.0 - oh_off)
        return torch.prod(cdfs, 1) @ (gh_w / math.sqrt(math.pi))[:, None]      # R,1

    def variational_expectations(self, Fmu, Fvar, Y):
        p = self._prob_is_largest(Y, Fmu, Fvar)
        eps = self.epsilon
        return p * math.log(1 - eps) + (1.0 - p) * math.log(eps / (self.num_classes - 1.0))

    def predict_mean_and_var(self, Fmu, Fvar):
        ps = []
        for k in range(self.num_classes):
            Yk = torch.full((Fmu.shape[0],), k, dtype=torch.long)
            ps.append(self._prob_is_largest(Yk, Fmu, Fvar))
        ps = torch.cat(ps, 1)
        return ps, ps - ps ** 2

    def predict_density(self, Fmu, Fvar, Y):
        p = self._prob_is_largest(Y, Fmu, Fvar)
        eps = self.epsilon
        return torch.log(p * (1 - eps) + (1.0 - p) * (eps / (self.num_classes - 1.0)))


class Bernoulli:
    """gpflow Bernoulli with probit link; 20-pt Gauss-Hermite variational expectations."""
    def parameters(self):
        return []

    @staticmethod
    def _probit(x):
        return 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0))) * (1 - 2e-3) + 1e-3

    def logp(self, F, Y):
        p = self._probit(F)
        return torch.log(torch.where(Y == 1, p, 1 - p))

    def variational_expectations(self, Fmu, Fvar, Y):
        gh_x, gh_w = _gh(20)
        X = Fmu[..., None] + torch.sqrt(2.0 * Fvar)[..., None] * gh_x
        lp = self.logp(X, Y[..., None].expand_as(X))
        return (lp * gh_w).sum(-1) / math.sqrt(math.pi)

    def predict_mean_and_var(self, Fmu, Fvar):
        p = 0.5 * (1.0 + torch.erf(Fmu / torch.sqrt(1 + Fvar) / math.sqrt(2.0))) * (1 - 2e-3) + 1e-3
        return p, p - p ** 2

    def predict_density(self, Fmu, Fvar, Y):
        p, _ = self.predict_mean_and_var(Fmu, Fvar)
        return torch.log(torch.where(Y == 1, p, 1 - p))


# ----------------------------------------------------------------------------
# utils.py
# ----------------------------------------------------------------------------
def reparameterize(mean, var, z, full_cov=False):
    """utils.py:22-51."""
    if var is None:
        return mean
    if full_cov is False:
        return mean + z * (var + settings.jitter) ** 0.5          # utils.py:41
    S, N, D = mean.shape
    mean = mean.permute(0, 2, 1)
    var = var.permute(0, 3, 1, 2)
    I = settings.jitter * torch.eye(N, dtype=DT)[None, None]
    chol = torch.linalg.cholesky(var + I)
    z_SDN1 = z.permute(0, 2, 1)[:, :, :, None]
    f = mean + (chol @ z_SDN1)[:, :, :, 0]
    return f.permute(0, 2, 1)


class BroadcastingLikelihood:
    """utils.py:54-121."""
    def __init__(self, likelihood):
        self.likelihood = likelihood
        self.needs_broadcasting = not isinstance(likelihood, Gaussian)   # utils.py:66-69

    def _broadcast(self, f, vars_SND, vars_ND):
        if not self.needs_broadcasting:
            return f(vars_SND, [v[None] for v in vars_ND])             # utils.py:72-73
        S, N, D = vars_SND[0].shape
        tiled = [x[None].repeat(S, 1, 1) for x in vars_ND]              # utils.py:77
        flat_SND = [x.reshape(S * N, D) for x in vars_SND]
        flat_tiled = [x.reshape(S * N, -1) for x in tiled]
        res = f(flat_SND, flat_tiled)
        if isinstance(res, tuple):
            return [x.reshape(S, N, -1) for x in res]
        return res.reshape(S, N, -1)

    def variational_expectations(self, Fmu, Fvar, Y):
        return self._broadcast(lambda a, b: self.likelihood.variational_expectations(a[0], a[1], b[0]),
                               [Fmu, Fvar], [Y])

    def predict_mean_and_var(self, Fmu, Fvar):
        return self._broadcast(lambda a, b: self.likelihood.predict_mean_and_var(a[0], a[1]),
                               [Fmu, Fvar], [])

    def predict_density(self, Fmu, Fvar, Y):
        return self._broadcast(lambda a, b: self.likelihood.predict_density(a[0], a[1], b[0]),
                               [Fmu, Fvar], [Y])


# ----------------------------------------------------------------------------
# layers.py
# ----------------------------------------------------------------------------
class SVGP_Layer:
    """layers.py:122-246."""
    faithful = True

    def __init__(self, kern, Z, num_outputs, mean_function, white=False, input_prop_dim=None):
        Z = np.asarray(Z, dtype=np.float64)
        self.input_prop_dim = input_prop_dim
        self.num_inducing = Z.shape[0]
        self.q_mu = torch.zeros(self.num_inducing, num_outputs, dtype=DT)           # layers.py:146-147
        self.q_sqrt = torch.eye(self.num_inducing, dtype=DT)[None].repeat(num_outputs, 1, 1)  # :149-151
        self.Z = _t(Z).clone()
        self.kern = kern
        self.mean_function = mean_function
        self.num_outputs = num_outputs
        self.white = white
        if not self.white:                                                           # layers.py:160-163
            with torch.no_grad():
                Ku = self.kern.K(self.Z)
                Lu = torch.linalg.cholesky(Ku + torch.eye(Z.shape[0], dtype=DT) * settings.jitter)
            self.q_sqrt = Lu[None].repeat(num_outputs, 1, 1).clone()

    def parameters(self):
        return [self.Z, self.q_mu, self.q_sqrt] + self.kern.parameters()

    def _chol(self):
        # layers.py:167-175 (recomputed per call here; TF caches it per graph)
        Ku = self.kern.K(self.Z) + settings.jitter * torch.eye(self.num_inducing, dtype=DT)
        Lu = torch.linalg.cholesky(Ku)
        return Ku, Lu

    def conditional_ND(self, X, full_cov=False):
        """layers.py:178-219, op for op."""
        Ku, Lu = self._chol()
        D, M = self.num_outputs, self.num_inducing
        Kuf = self.kern.K(self.Z, X)                                                # :184
        A = torch.linalg.solve_triangular(Lu, Kuf, upper=False)                     # :186
        if not self.white:
            A = torch.linalg.solve_triangular(Lu.T, A, upper=True)                  # :188
        mean = A.T @ self.q_mu                                                      # :190
        q_sqrt = torch.tril(self.q_sqrt)
        if self.faithful or full_cov:
            A_tiled = A[None].repeat(D, 1, 1)                                       # :192
            I = torch.eye(M, dtype=DT)[None]
            SK = -I if self.white else -Ku[None].repeat(D, 1, 1)                    # :195-198
            SK = SK + q_sqrt @ q_sqrt.transpose(1, 2)                               # :200-201
            B = SK @ A_tiled                                                        # :204
            if full_cov:
                delta_cov = A_tiled.transpose(1, 2) @ B                             # :208
                Kff = self.kern.K(X)
            else:
                delta_cov = torch.sum(A_tiled * B, 1)                               # :212
                Kff = self.kern.Kdiag(X)
        else:
            # identical arithmetic without the D_out-tiled temporaries
            P = torch.eye(M, dtype=DT) if self.white else Ku
            base = -torch.sum(A * (P @ A), 0)
            C = torch.einsum('dij,ir->djr', q_sqrt, A)
            delta_cov = base[None] + torch.sum(C * C, 1)
            Kff = self.kern.Kdiag(X)
        var = Kff[None] + delta_cov                                                 # :216
        var = var.permute(*reversed(range(var.dim())))                              # tf.transpose, :217
        return mean + self.mean_function(X), var                                    # :219

    def conditional_SND(self, X, full_cov=False):
        """layers.py:52-74."""
        if full_cov:
            ms, vs = zip(*[self.conditional_ND(x, full_cov=True) for x in X])
            return torch.stack(ms), torch.stack(vs)
        S, N, D = X.shape
        mean, var = self.conditional_ND(X.reshape(S * N, D))
        return mean.reshape(S, N, self.num_outputs), var.reshape(S, N, self.num_outputs)

    def sample_from_conditional(self, X, z=None, full_cov=False):
        """layers.py:76-119."""
        mean, var = self.conditional_SND(X, full_cov=full_cov)
        if z is None:
            z = torch.randn(mean.shape, dtype=DT)
        samples = reparameterize(mean, var, z, full_cov=full_cov)
        if self.input_prop_dim:
            X_prop = X[:, :, :self.input_prop_dim]
            samples = torch.cat([X_prop, samples], 2)
            mean = torch.cat([X_prop, mean], 2)
            if full_cov:
                zeros = torch.zeros(X.shape[0], X.shape[1], X.shape[1], self.input_prop_dim, dtype=DT)
                var = torch.cat([zeros, var], 3)
            else:
                var = torch.cat([torch.zeros_like(X_prop), var], 2)
        return samples, mean, var

    def KL(self):
        """layers.py:221-246."""
        Ku, Lu = self._chol()
        D, M = self.num_outputs, self.num_inducing
        q_sqrt = torch.tril(self.q_sqrt)
        KL = -0.5 * D * M
        KL = KL - 0.5 * torch.sum(torch.log(torch.diagonal(q_sqrt, dim1=1, dim2=2) ** 2))
        if not self.white:
            KL = KL + torch.sum(torch.log(torch.diagonal(Lu))) * D
            KL = KL + 0.5 * torch.sum(torch.linalg.solve_triangular(
                Lu[None].repeat(D, 1, 1), q_sqrt, upper=False) ** 2)
            Kinv_m = torch.cholesky_solve(self.q_mu, Lu)
            KL = KL + 0.5 * torch.sum(self.q_mu * Kinv_m)
        else:
            KL = KL + 0.5 * torch.sum(q_sqrt ** 2)
            KL = KL + 0.5 * torch.sum(self.q_mu ** 2)
        return KL


# ----------------------------------------------------------------------------
# layer_initializations.py:16-52
# ----------------------------------------------------------------------------
def init_layers_linear(X, Y, Z, kernels, num_outputs=None, mean_function=None, white=False,
                       W_list=None):
    """W_list (optional): explicit projection matrices for the dim-changing layers,
    because the SVD sign of layer_initializations.py:35-36 is ambiguous."""
    X = np.asarray(X, dtype=np.float64)
    num_outputs = num_outputs or Y.shape[1]
    mean_function = Zero() if mean_function is None else mean_function
    layers = []
    X_running, Z_running = X.copy(), np.asarray(Z, dtype=np.float64).copy()
    wi = 0
    for kern_in, kern_out in zip(kernels[:-1], kernels[1:]):
        dim_in, dim_out = kern_in.input_dim, kern_out.input_dim
        if dim_in == dim_out:
            mf = Identity()
        else:
            if W_list is not None:
                W = np.asarray(W_list[wi], dtype=np.float64); wi += 1
            elif dim_in > dim_out:
                _, _, V = np.linalg.svd(X_running, full_matrices=False)
                W = V[:dim_out, :].T
            else:
                W = np.concatenate([np.eye(dim_in), np.zeros((dim_in, dim_out - dim_in))], 1)
            mf = Linear(W)
        layers.append(SVGP_Layer(kern_in, Z_running, dim_out, mf, white=white))
        if dim_in != dim_out:
            Z_running = Z_running.dot(W)
            X_running = X_running.dot(W)
    layers.append(SVGP_Layer(kernels[-1], Z_running, num_outputs, mean_function, white=white))
    return layers


# ----------------------------------------------------------------------------
# dgp.py
# ----------------------------------------------------------------------------
class DGP_Base:
    """dgp.py:35-126 (no Minibatch iterator: the caller passes the minibatch explicitly)."""
    def __init__(self, X, Y, likelihood, layers, minibatch_size=None, num_samples=1, num_data=None):
        self.num_samples = num_samples
        self.num_data = num_data or X.shape[0]
        self.X, self.Y = _t(X), _t(Y)
        self.likelihood = BroadcastingLikelihood(likelihood)
        self.layers = layers

    def parameters(self):
        ps = []
        for l in self.layers:
            ps += l.parameters()
        return ps + self.likelihood.likelihood.parameters()

    def propagate(self, X, full_cov=False, S=1, zs=None):
        """dgp.py:61-76."""
        sX = _t(X)[None].repeat(S, 1, 1)                                # dgp.py:63
        Fs, Fmeans, Fvars = [], [], []
        F = sX
        zs = zs or [None] * len(self.layers)
        for layer, z in zip(self.layers, zs):
            F, Fmean, Fvar = layer.sample_from_conditional(F, z=None if z is None else _t(z),
                                                           full_cov=full_cov)
            Fs.append(F); Fmeans.append(Fmean); Fvars.append(Fvar)
        return Fs, Fmeans, Fvars

    def _build_predict(self, X, full_cov=False, S=1, zs=None):
        Fs, Fmeans, Fvars = self.propagate(X, full_cov=full_cov, S=S, zs=zs)
        return Fmeans[-1], Fvars[-1]

    def E_log_p_Y(self, X, Y, zs=None):
        """dgp.py:83-90."""
        Fmean, Fvar = self._build_predict(X, full_cov=False, S=self.num_samples, zs=zs)
        var_exp = self.likelihood.variational_expectations(Fmean, Fvar, _t(Y))
        return torch.mean(var_exp, 0)

    def elbo(self, X=None, Y=None, zs=None):
        """dgp.py:92-98  (_build_likelihood)."""
        X = self.X if X is None else _t(X)
        Y = self.Y if Y is None else _t(Y)
        L = torch.sum(self.E_log_p_Y(X, Y, zs=zs))
        KL = sum(layer.KL() for layer in self.layers)
        scale = float(self.num_data) / float(X.shape[0])
        return L * scale - KL

    def compute_log_likelihood(self, zs=None):
        with torch.no_grad():
            return float(self.elbo(zs=zs))

    def elbo_and_grad(self, X=None, Y=None, zs=None):
        ps = self.parameters()
        for p in ps:
            p.requires_grad_(True)
            p.grad = None
        e = self.elbo(X, Y, zs)
        e.backward()
        grads = [p.grad.clone() if p.grad is not None else torch.zeros_like(p) for p in ps]
        for p in ps:
            p.requires_grad_(False)
            p.grad = None
        return float(e.detach()), grads

    # dgp.py:100-126
    def predict_f(self, Xnew, num_samples, zs=None):
        with torch.no_grad():
            return self._build_predict(_t(Xnew), full_cov=False, S=num_samples, zs=zs)

    def predict_f_full_cov(self, Xnew, num_samples, zs=None):
        with torch.no_grad():
            return self._build_predict(_t(Xnew), full_cov=True, S=num_samples, zs=zs)

    def predict_all_layers(self, Xnew, num_samples, zs=None):
        with torch.no_grad():
            return self.propagate(_t(Xnew), full_cov=False, S=num_samples, zs=zs)

    def predict_y(self, Xnew, num_samples, zs=None):
        with torch.no_grad():
            Fmean, Fvar = self._build_predict(_t(Xnew), full_cov=False, S=num_samples, zs=zs)
            return self.likelihood.predict_mean_and_var(Fmean, Fvar)

    def predict_density(self, Xnew, Ynew, num_samples, zs=None):
        with torch.no_grad():
            Fmean, Fvar = self._build_predict(_t(Xnew), full_cov=False, S=num_samples, zs=zs)
            l = self.likelihood.predict_density(Fmean, Fvar, _t(Ynew))
            return torch.logsumexp(l - math.log(num_samples), 0)


def mvhermgauss(H, D):
    """gpflow.quadrature.mvhermgauss (GPflow 1.1.1, restated): tensor-product Gauss-Hermite nodes (H**D, D) and
    weights (H**D,).  Call site dgp.py:143."""
    import itertools
    gh_x, gh_w = np.polynomial.hermite.hermgauss(H)
    x = np.array(list(itertools.product(*(gh_x,) * D)))
    w = np.prod(np.array(list(itertools.product(*(gh_w,) * D))), 1)
    return x, w


class DGP_Quad(DGP_Base):
    """dgp.py:129-166: quadrature over the inner layers' whitened draws instead of Monte-Carlo sampling."""
    def __init__(self, *args, H=100, **kwargs):
        DGP_Base.__init__(self, *args, **kwargs)
        self.H = H
        self.D_quad = sum(layer.q_mu.shape[1] for layer in self.layers[:-1])           # dgp.py:142
        gh_x, gh_w = mvhermgauss(H, self.D_quad)
        gh_x = gh_x * 2. ** 0.5                                                          # dgp.py:144
        self.gh_w = _t(gh_w * np.pi ** (-0.5 * self.D_quad))                             # dgp.py:145
        s, e = 0, 0
        self.gh_x = []
        for layer in self.layers[:-1]:                                                   # dgp.py:149-154
            e += layer.q_mu.shape[1]
            self.gh_x.append(_t(gh_x[:, None, s:e]))
            s += layer.q_mu.shape[1]
        self.gh_x.append(torch.zeros(1, 1, 1, dtype=DT))                                 # dgp.py:157

    def E_log_p_Y(self, X, Y, zs=None):
        """dgp.py:159-166."""
        _, Fmeans, Fvars = self.propagate(X, zs=self.gh_x, full_cov=False, S=self.H ** self.D_quad)
        var_exp = self.likelihood.variational_expectations(Fmeans[-1], Fvars[-1], _t(Y))
        return torch.sum(var_exp * self.gh_w[:, None, None], 0)


class DGP(DGP_Base):
    """dgp.py:169-192."""
    def __init__(self, X, Y, Z, kernels, likelihood, num_outputs=None, mean_function=None,
                 white=False, W_list=None, **kwargs):
        layers = init_layers_linear(X, Y, Z, kernels, num_outputs=num_outputs,
                                    mean_function=mean_function, white=white, W_list=W_list)
        DGP_Base.__init__(self, X, Y, likelihood, layers, **kwargs)


# ----------------------------------------------------------------------------
# Training step of the reference = TF autodiff + Adam on GPflow's unconstrained
# variables (demos/run_regression.py:83).  GPflow transforms:
# positive -> softplus(+1e-6 lower) ; q_sqrt -> packed lower triangle.
# ----------------------------------------------------------------------------
POS_LOWER = 1e-6


def softplus_fwd(free):
    return torch.nn.functional.softplus(free) + POS_LOWER


def softplus_inv(x):
    y = x - POS_LOWER
    return y + torch.log(-torch.expm1(-y))


class AdamState:
    """tf.train.AdamOptimizer(lr) semantics: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); eps outside sqrt-hat."""
    def __init__(self, model, lr=0.01, b1=0.9, b2=0.999, eps=1e-8):
        self.model, self.lr, self.b1, self.b2, self.eps = model, lr, b1, b2, eps
        self.t = 0
        self.kinds = []
        for l in model.layers:
            self.kinds += ['id', 'id', 'tril', 'pos', 'pos']
        self.kinds += ['pos'] * len(model.likelihood.likelihood.parameters())
        self.free = []
        for p, k in zip(model.parameters(), self.kinds):
            self.free.append(softplus_inv(p.detach().clone()) if k == 'pos' else p.detach().clone())
        self.m = [torch.zeros_like(f) for f in self.free]
        self.v = [torch.zeros_like(f) for f in self.free]

    def step(self, X=None, Y=None, zs=None):
        """One minimize step of objective = -ELBO. Returns the ELBO before the update."""
        model = self.model
        elbo, grads = model.elbo_and_grad(X, Y, zs)
        self.t += 1
        lr_t = self.lr * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        for i, (p, g, k) in enumerate(zip(model.parameters(), grads, self.kinds)):
            g = -g
            if k == 'pos':
                g = g * torch.sigmoid(self.free[i])
            elif k == 'tril':
                g = torch.tril(g)
            self.m[i] = self.b1 * self.m[i] + (1 - self.b1) * g
            self.v[i] = self.b2 * self.v[i] + (1 - self.b2) * g * g
            self.free[i] = self.free[i] - lr_t * self.m[i] / (torch.sqrt(self.v[i]) + self.eps)
            with torch.no_grad():
                if k == 'pos':
                    p.copy_(softplus_fwd(self.free[i]))
                elif k == 'tril':
                    p.copy_(torch.tril(self.free[i]))
                else:
                    p.copy_(self.free[i])
        return elbo


# ----------------------------------------------------------------------------
# Natural-gradient step on (q_mu, q_sqrt) of chosen layers.  The reference calls
# gpflow.training.NatGradOptimizer(gamma).minimize(model, var_list=[[q_mu, q_sqrt]], maxiter=1)
# (tests/test_collapsed.py:99-100, demos/using_natural_gradients.ipynb "ng_action",
# demos/demo_regression_UCI.ipynb:357-366).  GPflow 1.1.1 is not in /root/reference; this is a
# restatement of its published algorithm (training/natgrad_optimizer.py, default XiNat
# parameterisation; Salimbeni, Eleftheriadis & Hensman 2018, "Natural gradients in practice"):
#   eta = (m, S + m m^T),  theta = (S^-1 m, -1/2 S^-1)
#   dL/d eta  by back-propagating (dL/dq_mu, dL/dq_sqrt) through eta -> (eta1, chol(eta2 - eta1 eta1^T))
#   theta <- theta - gamma dL/d eta           (L = -ELBO)
#   (q_mu, q_sqrt) <- natural_to_meanvarsqrt(theta): C = chol(-2 theta2), V = C^-1, S = V^T V,
#                      m = S theta1, q_sqrt = chol(S)
# Pinned by identity I6 (gamma = 1 on a Gaussian-likelihood last layer lands on the SGPR optimum,
# tests/test_collapsed.py:57-104) in tests/test_oracle_identities.py / tests/test_natgrad_cpu.py.
# ----------------------------------------------------------------------------
def natgrad_step(model, layer_ids, gamma, X=None, Y=None, zs=None):
    """One NatGrad step on the listed layers; returns the ELBO before the update."""
    elbo, grads = model.elbo_and_grad(X, Y, zs)
    ps = model.parameters()
    for li in layer_ids:
        layer = model.layers[li]
        i_mu = next(i for i, p in enumerate(ps) if p is layer.q_mu)
        i_sq = next(i for i, p in enumerate(ps) if p is layer.q_sqrt)
        dL_dmean = -grads[i_mu]                        # (M, D)
        dL_dsqrt = -torch.tril(grads[i_sq])            # (D, M, M)  LowerTriangular transform: packed lower triangle
        q_mu = layer.q_mu.detach()
        q_sqrt = torch.tril(layer.q_sqrt.detach())
        D, M = q_sqrt.shape[0], q_sqrt.shape[1]
        new_mu = torch.empty_like(q_mu)
        new_sqrt = torch.empty_like(q_sqrt)
        for d in range(D):
            m = q_mu[:, d:d + 1]
            Lq = q_sqrt[d]
            S = Lq @ Lq.T
            eta1 = m.clone().requires_grad_(True)
            eta2 = (S + m @ m.T).clone().requires_grad_(True)
            mean = eta1
            varsqrt = torch.linalg.cholesky(eta2 - eta1 @ eta1.T)
            g1, g2 = torch.autograd.grad([mean, varsqrt], [eta1, eta2],
                                         grad_outputs=[dL_dmean[:, d:d + 1], dL_dsqrt[d]])
            g2 = 0.5 * (g2 + g2.T)
            Sinv = torch.cholesky_inverse(torch.linalg.cholesky(S))
            nat1 = Sinv @ m - gamma * g1
            nat2 = -0.5 * Sinv - gamma * g2
            C = torch.linalg.cholesky(-2.0 * nat2)
            V = torch.linalg.solve_triangular(C, torch.eye(M, dtype=DT), upper=False)
            S_new = V.T @ V
            new_mu[:, d:d + 1] = S_new @ nat1
            new_sqrt[d] = torch.linalg.cholesky(S_new)
        with torch.no_grad():
            layer.q_mu.copy_(new_mu)
            layer.q_sqrt.copy_(new_sqrt)
    return elbo
