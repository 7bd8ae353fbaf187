"""Host mirror of doubly_stochastic_dgp/dgp.py: `DGP_Base` / `DGP` with the reference's constructor signatures
(dgp.py:42-45,184-187) and methods (propagate dgp.py:61-76, compute_log_likelihood = _build_likelihood
dgp.py:92-98, predict_* dgp.py:100-126).  Everything numeric runs on the device through _lib.Context; this
module only owns parameters, minibatching and shapes.  The training step the reference gets from
`AdamOptimizer().minimize(model)` (demos/run_regression.py:83) is `model.adam_init(lr)` + `model.train_step()`.
"""
import numpy as np

from . import _lib, settings
from .layer_initializations import init_layers_linear
from .likelihoods import Gaussian, MultiClass
from .mean_functions import Linear, Zero
from .params import Parameter, Parameterized
from .utils import BroadcastingLikelihood


class _Minibatch:
    """gpflow.params.Minibatch(X, batch_size, seed) stand-in: repeat -> shuffle(buffer=N, seed) -> batch
    (SURVEY App. C.5).  X and Y built with the same seed stay aligned (dgp.py:51-52)."""
    def __init__(self, data, batch_size, seed=0):
        self.data = np.asarray(data)
        self.batch_size = int(batch_size)
        self.rng = np.random.RandomState(seed)
        self.perm = self.rng.permutation(len(self.data))
        self.pos = 0

    def next(self):
        n = len(self.data)
        idx = []
        while len(idx) < self.batch_size:
            if self.pos >= n:
                self.perm = self.rng.permutation(n)
                self.pos = 0
            take = min(self.batch_size - len(idx), n - self.pos)
            idx.extend(self.perm[self.pos:self.pos + take])
            self.pos += take
        return self.data[np.asarray(idx)]


class DGP_Base(Parameterized):
    """dgp.py:35-126."""
    def __init__(self, X, Y, likelihood, layers, minibatch_size=None, num_samples=1, num_data=None,
                 device=0, **kwargs):
        X = np.asarray(X, dtype=np.float64)
        Y = np.asarray(Y, dtype=np.float64)
        self.num_samples = num_samples
        self.num_data = num_data or X.shape[0]
        self.minibatch_size = minibatch_size
        if minibatch_size:
            self._Xmb = _Minibatch(X, minibatch_size, seed=0)      # dgp.py:51-52
            self._Ymb = _Minibatch(Y, minibatch_size, seed=0)
        self.X, self.Y = X, Y
        self.likelihood = BroadcastingLikelihood(likelihood)        # dgp.py:57
        object.__setattr__(self.likelihood, "_model", self)
        self.layers = list(layers)                                   # dgp.py:59
        self._device = device
        self._ctx = None
        self._host_dirty = True
        self._trainable_dirty = True
        self._device_newer = False
        self._seed = 0x5D6A1
        self._adam = None
        self._comm = None
        for p in self.parameters():
            p._owner = self
        for l in self.layers:
            object.__setattr__(l, "_model", self)

    # ------------------------------------------------------------------ parameter plumbing
    def parameters(self):
        out = []
        for l in self.layers:
            out += [l.feature.Z, l.q_mu, l.q_sqrt, l.kern.lengthscales, l.kern.variance]
            if getattr(l.kern, "white_variance", None) is not None:
                out.append(l.kern.white_variance)
            if isinstance(l.mean_function, Linear):
                out += [l.mean_function.A, l.mean_function.b]
        lik = self.likelihood.likelihood
        if isinstance(lik, Gaussian):
            out.append(lik.variance)
        return out

    def _mark_host_dirty(self):
        self._host_dirty = True

    def _mark_trainable_dirty(self):
        self._trainable_dirty = True

    def _param_fields(self):
        """(layer index, C-ABI field, Parameter) of every trainable-capable parameter."""
        out = []
        for i, l in enumerate(self.layers):
            out += [(i, _lib.F_Z, l.feature.Z), (i, _lib.F_Q_MU, l.q_mu), (i, _lib.F_Q_SQRT, l.q_sqrt),
                    (i, _lib.F_LENGTHSCALES, l.kern.lengthscales), (i, _lib.F_VARIANCE, l.kern.variance)]
            if getattr(l.kern, "white_variance", None) is not None:
                out.append((i, _lib.F_WHITE_VARIANCE, l.kern.white_variance))
        lik = self.likelihood.likelihood
        if isinstance(lik, Gaussian):
            out.append((-1, _lib.F_LIK_VARIANCE, lik.variance))
        return out

    def _refresh_from_device(self):
        if not self._device_newer or self._ctx is None:
            return
        self._device_newer = False
        c = self._ctx
        for i, l in enumerate(self.layers):
            l.feature.Z._value = c.get_param(i, _lib.F_Z, l.feature.Z.shape)
            l.q_mu._value = c.get_param(i, _lib.F_Q_MU, l.q_mu.shape)
            l.q_sqrt._value = c.get_param(i, _lib.F_Q_SQRT, l.q_sqrt.shape)
            l.kern.lengthscales._value = c.get_param(i, _lib.F_LENGTHSCALES, l.kern.lengthscales.shape)
            l.kern.variance._value = c.get_param(i, _lib.F_VARIANCE, ())
            if getattr(l.kern, "white_variance", None) is not None:
                l.kern.white_variance._value = c.get_param(i, _lib.F_WHITE_VARIANCE, ())
        lik = self.likelihood.likelihood
        if isinstance(lik, Gaussian):
            lik.variance._value = c.get_param(-1, _lib.F_LIK_VARIANCE, ())

    def _layer_descs(self):
        descs = []
        for l in self.layers:
            Z = l.feature.Z._value
            descs.append((Z.shape[0], Z.shape[1], l.num_outputs, l.kern.code, int(l.kern.ARD), int(bool(l.white)),
                          l.mean_function.code, int(getattr(l.kern, "white_variance", None) is not None),
                          # (a final layer never feeds another one: a stand-alone layer's propagation is the host-side concat
                          # of layers.sample_from_conditional)
                          0 if l is self.layers[-1] else int(l.input_prop_dim or 0)))
            if l.kern.input_dim != Z.shape[1]:
                raise ValueError(f"kernel input_dim {l.kern.input_dim} != Z columns {Z.shape[1]}")
        return descs

    def _ensure_ctx(self, N, S):
        if self._ctx is not None and (N > self._ctx.N_max or S > self._ctx.S_max):
            if self._adam is not None:
                raise RuntimeError(f"call needs N={N}, S={S} beyond the device workspaces "
                                   f"(N_max={self._ctx.N_max}, S_max={self._ctx.S_max}) after training started; "
                                   "predict in chunks or construct the model with larger minibatch/num_samples")
            self._refresh_from_device()
            N, S = max(N, self._ctx.N_max), max(S, self._ctx.S_max)
            self._ctx.close()
            self._ctx = None
        if self._ctx is None:
            lik = self.likelihood.likelihood
            K = lik.num_classes if isinstance(lik, MultiClass) else 0
            D_y = 1 if K else self.layers[-1].num_outputs
            N_max = max(N, self.minibatch_size or min(self.X.shape[0], 4096), 128)
            S_max = max(S, self.num_samples, 8)
            self._ctx = _lib.Context(self._layer_descs(), lik.code, K, D_y, float(settings.jitter), N_max, S_max,
                                     device=self._device)
            self._trainable_dirty = True
            if self._comm is not None:
                self._ctx.comm_init(*self._comm)
            self._host_dirty = True
        if self._host_dirty:
            c = self._ctx
            for i, l in enumerate(self.layers):
                c.set_param(i, _lib.F_Z, l.feature.Z._value)
                c.set_param(i, _lib.F_Q_MU, l.q_mu._value)
                c.set_param(i, _lib.F_Q_SQRT, l.q_sqrt._value)
                c.set_param(i, _lib.F_LENGTHSCALES, l.kern.lengthscales._value)
                c.set_param(i, _lib.F_VARIANCE, l.kern.variance._value)
                if getattr(l.kern, "white_variance", None) is not None:
                    c.set_param(i, _lib.F_WHITE_VARIANCE, l.kern.white_variance._value)
                if isinstance(l.mean_function, Linear):
                    c.set_param(i, _lib.F_MEAN_W, l.mean_function.A._value)
                    c.set_param(i, _lib.F_MEAN_B, l.mean_function.b._value)
            lik = self.likelihood.likelihood
            if isinstance(lik, Gaussian):
                c.set_param(-1, _lib.F_LIK_VARIANCE, lik.variance._value)
            self._host_dirty = False
        if self._trainable_dirty:
            for i, field, p in self._param_fields():
                self._ctx.set_trainable(i, field, p.trainable)
            self._trainable_dirty = False
        return self._ctx

    def _next_seed(self):
        self._seed = (self._seed * 6364136223846793005 + 1442695040888963407) % (1 << 64)
        return self._seed

    def _default_zs(self, N):
        """Draws used when the caller passes zs=None: None = in-kernel Philox normals (DGP_Quad overrides this)."""
        return None

    def _minibatch(self):
        if self.minibatch_size:
            return self._Xmb.next(), self._Ymb.next()
        return self.X, self.Y

    # ------------------------------------------------------------------ reference API
    def _plan(self, N, S, full_cov=False):
        """Row and sample chunks of one prediction call.  Normally a single chunk (the context is re-created with larger
        workspaces when a call outgrows them); once training has started the optimiser state lives in the context, so a
        larger prediction (the reference predicts with S=100 after training with num_samples=1, demos/run_regression.py:
        108-113) is evaluated in chunks instead -- exact, since with full_cov=False every (sample, row) is independent
        (layers.py:71-74), and with full_cov=True every sample is (layers.py:66-69)."""
        c = self._ctx
        if c is None or self._adam is None or (N <= c.N_max and S <= c.S_max):
            return [(0, N)], [(0, S)]
        if full_cov and N > c.N_max:
            raise RuntimeError(f"full_cov prediction needs all N={N} rows at once (N_max={c.N_max} after training started)")
        rows = [(a, min(a + c.N_max, N)) for a in range(0, N, c.N_max)]
        samples = [(a, min(a + c.S_max, S)) for a in range(0, S, c.S_max)]
        return rows, samples

    @staticmethod
    def _cut(zs, s0, s1, n0, n1):
        """The [s0:s1, n0:n1] block of every layer's draws; axes of length 1 broadcast (DGP_Quad's (S,1,D) nodes)."""
        if zs is None:
            return None
        out = []
        for z in zs:
            if z is not None:
                z = np.asarray(z)
                if z.ndim != 3:
                    raise ValueError(f"zs entries must be (S, N, D) arrays, got shape {z.shape}")
                z = z[(slice(None) if z.shape[0] == 1 else slice(s0, s1)), (slice(None) if z.shape[1] == 1 else slice(n0, n1))]
            out.append(z)
        return out

    def propagate(self, X, full_cov=False, S=1, zs=None):
        """dgp.py:61-76 -> (Fs, Fmeans, Fvars), lists of (S,N,D_l) float64 arrays ((S,N,N,D_l) variances with full_cov)."""
        X = np.asarray(X, dtype=np.float64)
        N = X.shape[0]
        rows, samples = self._plan(N, S, full_cov)
        parts = []
        for s0, s1 in samples:
            line = []
            for n0, n1 in rows:
                ctx = self._ensure_ctx(n1 - n0, s1 - s0)
                fn = ctx.propagate_full_cov if full_cov else ctx.propagate     # full_cov: csrc/full_cov.cu (float64 pipeline)
                line.append(fn(X[n0:n1], s1 - s0, zs=self._cut(zs, s0, s1, n0, n1), seed=self._next_seed()))
            parts.append(line)
        L = len(self.layers)
        out = []
        for which in range(3):
            out.append([np.concatenate([np.concatenate([cell[which][l] for cell in line], axis=1) for line in parts], axis=0)
                        .astype(np.float64) for l in range(L)])
        for l, layer in enumerate(self.layers):
            ipd = 0 if l == L - 1 else (layer.input_prop_dim or 0)
            if ipd and not full_cov:
                # layers.py:105-117: mean = concat([X_prop, mean]), var = concat([0, var]); the samples left the device
                # already concatenated.  X_prop is the layer's own input, i.e. the first ipd columns of its Fs
                Xp = out[0][l][:, :, :ipd]
                out[1][l] = np.concatenate([Xp, out[1][l]], axis=2)
                out[2][l] = np.concatenate([np.zeros_like(Xp), out[2][l]], axis=2)
        return out[0], out[1], out[2]

    def _build_predict(self, X, full_cov=False, S=1, zs=None):
        """dgp.py:78-81."""
        Fs, Fmeans, Fvars = self.propagate(X, full_cov=full_cov, S=S, zs=zs)
        return Fmeans[-1], Fvars[-1]

    def compute_log_likelihood(self, zs=None, X=None, Y=None):
        """gpflow Model.compute_log_likelihood -> _build_likelihood (dgp.py:92-98): the ELBO on the (next)
        minibatch.  `zs` (list of (S,N,D_l) arrays) fixes the draws, like propagate(zs=...)."""
        if X is None:
            X, Y = self._minibatch()
        ctx = self._ensure_ctx(X.shape[0], self.num_samples)
        zs = self._default_zs(X.shape[0]) if zs is None else zs
        return ctx.elbo(X, Y, self.num_samples, self.num_data, zs=zs, seed=self._next_seed())

    def compute_log_likelihood_and_grad(self, zs=None, X=None, Y=None):
        """ELBO and dELBO/dparam for every trainable (what tf.gradients gives the reference's optimisers)."""
        if X is None:
            X, Y = self._minibatch()
        ctx = self._ensure_ctx(X.shape[0], self.num_samples)
        zs = self._default_zs(X.shape[0]) if zs is None else zs
        e = ctx.elbo_grad(X, Y, self.num_samples, self.num_data, zs=zs, seed=self._next_seed())
        grads = []
        for i, l in enumerate(self.layers):
            grads.append(dict(Z=ctx.get_grad(i, _lib.F_Z, l.feature.Z.shape),
                              q_mu=ctx.get_grad(i, _lib.F_Q_MU, l.q_mu.shape),
                              q_sqrt=ctx.get_grad(i, _lib.F_Q_SQRT, l.q_sqrt.shape),
                              lengthscales=ctx.get_grad(i, _lib.F_LENGTHSCALES, l.kern.lengthscales.shape),
                              variance=ctx.get_grad(i, _lib.F_VARIANCE, ())))
            if getattr(l.kern, "white_variance", None) is not None:
                grads[-1]["white_variance"] = ctx.get_grad(i, _lib.F_WHITE_VARIANCE, ())
        lik_grad = None
        if isinstance(self.likelihood.likelihood, Gaussian):
            lik_grad = ctx.get_grad(-1, _lib.F_LIK_VARIANCE, ())
        return e, grads, lik_grad

    def predict_f(self, Xnew, num_samples):
        return self._build_predict(Xnew, full_cov=False, S=num_samples)

    def predict_f_full_cov(self, Xnew, num_samples):
        return self._build_predict(Xnew, full_cov=True, S=num_samples)

    def predict_all_layers(self, Xnew, num_samples):
        return self.propagate(Xnew, full_cov=False, S=num_samples)

    def predict_all_layers_full_cov(self, Xnew, num_samples):
        return self.propagate(Xnew, full_cov=True, S=num_samples)

    def predict_y(self, Xnew, num_samples, zs=None):
        """dgp.py:116-119: likelihood.predict_mean_and_var of the last layer's marginals, (S,N,D) each; the likelihood
        epilogue runs on the device (csrc/lik_adam.cu)."""
        Xnew = np.asarray(Xnew, dtype=np.float64)
        rows, samples = self._plan(Xnew.shape[0], num_samples)
        means, vars_ = [], []
        for s0, s1 in samples:
            ms, vs = [], []
            for n0, n1 in rows:
                ctx = self._ensure_ctx(n1 - n0, s1 - s0)
                m, v = ctx.predict_y(Xnew[n0:n1], s1 - s0, zs=self._cut(zs, s0, s1, n0, n1), seed=self._next_seed())
                ms.append(m); vs.append(v)
            means.append(np.concatenate(ms, 1)); vars_.append(np.concatenate(vs, 1))
        return np.concatenate(means, 0).astype(np.float64), np.concatenate(vars_, 0).astype(np.float64)

    def predict_density(self, Xnew, Ynew, num_samples, zs=None):
        """dgp.py:121-126: logsumexp_S(likelihood.predict_density(Fmean, Fvar, Ynew) - log S), on the device.  When the
        samples have to be evaluated in chunks (see _plan) the per-chunk log-mean-exps are merged here."""
        Xnew = np.asarray(Xnew, dtype=np.float64)
        Ynew = np.asarray(Ynew, dtype=np.float64)
        rows, samples = self._plan(Xnew.shape[0], num_samples)
        cols = []
        for n0, n1 in rows:
            acc = None
            for s0, s1 in samples:
                ctx = self._ensure_ctx(n1 - n0, s1 - s0)
                d = ctx.predict_density(Xnew[n0:n1], Ynew[n0:n1], s1 - s0, zs=self._cut(zs, s0, s1, n0, n1),
                                        seed=self._next_seed()).astype(np.float64)
                d = d + np.log((s1 - s0) / float(num_samples))          # chunk's share of the mean over all S
                acc = d if acc is None else np.logaddexp(acc, d)
            cols.append(acc)
        return np.concatenate(cols, 0)

    # ------------------------------------------------------------------ training (AdamOptimizer.minimize)
    def adam_init(self, lr=0.01, beta1=0.9, beta2=0.999, eps=1e-8):
        N = self.minibatch_size or self.X.shape[0]
        ctx = self._ensure_ctx(N, self.num_samples)
        ctx.adam_init(lr, beta1, beta2, eps)
        self._adam = (lr, beta1, beta2, eps)

    def train_step(self, X=None, Y=None, zs=None):
        """one session.run(minimize_op): minibatch draw + ELBO fwd + bwd + Adam update; returns the ELBO."""
        if self._adam is None:
            self.adam_init()
        if X is None:
            X, Y = self._minibatch()
        ctx = self._ensure_ctx(X.shape[0], self.num_samples)
        zs = self._default_zs(X.shape[0]) if zs is None else zs
        try:
            e = ctx.train_step(_lib.f32(X), _lib.f32(Y), X.shape[0], self.num_samples, self.num_data, self._next_seed(),
                               zs=zs)
        finally:
            # (a step that fails with NOT_PD leaves the device parameters untouched -- csrc/lik_adam.cu k_adam -- but the
            # host copy is re-read either way)
            self._device_newer = True
        return e

    def natgrad_step(self, var_list=None, gamma=1.0, X=None, Y=None, zs=None):
        """NatGradOptimizer(gamma).minimize(model, var_list=var_list, maxiter=1) (tests/test_collapsed.py:99-100):
        one ELBO+gradient pass on the (next) minibatch, then the natural-gradient update of every [q_mu, q_sqrt] pair
        in var_list (default: the last layer's).  Returns the ELBO before the update."""
        ids = self._natgrad_layers(var_list)
        if X is None:
            X, Y = self._minibatch()
        ctx = self._ensure_ctx(X.shape[0], self.num_samples)
        zs = self._default_zs(X.shape[0]) if zs is None else zs
        try:
            e = ctx.natgrad_step(_lib.f32(X), _lib.f32(Y), X.shape[0], self.num_samples, self.num_data, self._next_seed(),
                                 ids, gamma, zs=zs)
        finally:
            self._device_newer = True
        return e

    def _natgrad_layers(self, var_list):
        if var_list is None:
            return [len(self.layers) - 1]
        ids = []
        for pair in var_list:
            if len(pair) != 2:
                raise ValueError("var_list entries must be [q_mu, q_sqrt] pairs")
            hit = [i for i, l in enumerate(self.layers) if pair[0] is l.q_mu and pair[1] is l.q_sqrt]
            if not hit:
                raise ValueError("var_list entry is not the (q_mu, q_sqrt) of a layer of this model")
            ids.append(hit[0])
        return ids

    def minimize(self, maxiter=1000, lr=0.01):
        if self._adam is None:
            self.adam_init(lr)
        e = None
        for _ in range(maxiter):
            e = self.train_step()
        return e

    # ------------------------------------------------------------------ multi-GPU
    def comm_init(self, id_bytes, rank, world):
        """Attach an NCCL communicator: X,Y given to this rank are its shard of the minibatch; ELBO and gradient
        are all-reduced inside the step (csrc/api.cu)."""
        self._comm = (id_bytes, rank, world)
        if self._ctx is not None:
            self._ctx.comm_init(id_bytes, rank, world)

    # ------------------------------------------------------------------ per-layer services for layers.py
    def _layer_snapshot(self, layer):
        """A private one-layer model holding a copy of `layer`'s current parameters: layers.py:46-119 work on any layer, also
        one that belongs to a model (its own context is sized and wired for the whole chain)."""
        import copy
        self._refresh_from_device()
        kern = copy.copy(layer.kern)
        kern.variance = Parameter(layer.kern.variance._value)
        kern.lengthscales = Parameter(layer.kern.lengthscales._value)
        if getattr(layer.kern, "white_variance", None) is not None:
            kern.white_variance = Parameter(layer.kern.white_variance._value)
        mf = layer.mean_function
        if isinstance(mf, Linear):
            mf = Linear(mf.A._value, mf.b._value)
        from .layers import SVGP_Layer
        clone = SVGP_Layer(kern, layer.feature.Z._value, layer.num_outputs, mf, white=layer.white,
                           input_prop_dim=layer.input_prop_dim)
        clone.q_mu = layer.q_mu._value
        clone.q_sqrt = layer.q_sqrt._value
        return _single_layer_model(clone)

    def _layer_conditional(self, layer, X, full_cov=False):
        return self._layer_snapshot(layer)._layer_conditional(layer, X, full_cov=full_cov)

    def _layer_propagate(self, layer, X, **kw):
        return self._layer_snapshot(layer)._layer_propagate(layer, X, **kw)

    def _layer_KL(self, layer):
        ctx = self._ensure_ctx(1, 1)
        return ctx.kl()[self.layers.index(layer)]


def mvhermgauss(H, D):
    """gpflow.quadrature.mvhermgauss: tensor-product Gauss-Hermite nodes (H**D, D) and weights (H**D,) (dgp.py:143).
    Construction-time constants (host), like the reference."""
    import itertools
    gh_x, gh_w = np.polynomial.hermite.hermgauss(H)
    x = np.array(list(itertools.product(*(gh_x,) * D))).reshape(H ** D, D)
    w = np.prod(np.array(list(itertools.product(*(gh_w,) * D))).reshape(H ** D, D), 1)
    return x, w


class DGP_Quad(DGP_Base):
    """dgp.py:129-166: Gauss-Hermite quadrature over the inner layers' whitened draws instead of Monte-Carlo samples.
    The H**D_quad nodes are handed to the device as the layers' z (dsdgp_* `zs`), the node weights as per-sample
    likelihood weights (dsdgp_set_sample_weights); everything else is the DGP_Base hot path."""
    def __init__(self, *args, H=100, **kwargs):
        DGP_Base.__init__(self, *args, **kwargs)
        self.H = H
        self.D_quad = sum(l.q_mu.shape[1] for l in self.layers[:-1])                    # dgp.py:142
        gh_x, gh_w = mvhermgauss(H, self.D_quad)
        gh_x = gh_x * 2. ** 0.5                                                          # dgp.py:144
        self.gh_w = gh_w * np.pi ** (-0.5 * self.D_quad)                                 # dgp.py:145
        self.gh_x, s = [], 0
        for l in self.layers[:-1]:                                                       # dgp.py:149-154
            e = s + l.q_mu.shape[1]
            self.gh_x.append(gh_x[:, None, s:e])
            s = e
        self.gh_x.append(None)            # the final layer is never sampled (dgp.py:156-157)
        self.num_samples = H ** self.D_quad                                              # dgp.py:164 S=H**D_quad
        self._zs_cache = {}
        self._weights_ctx = None

    def _ensure_ctx(self, N, S):
        ctx = DGP_Base._ensure_ctx(self, N, S)
        if self._weights_ctx is not ctx:
            ctx.set_sample_weights(self.gh_w)
            self._weights_ctx = ctx
        return ctx

    def _default_zs(self, N):
        if N not in self._zs_cache:       # (S,1,D) nodes broadcast over the N rows (dgp.py:147-148)
            self._zs_cache = {N: [None if g is None else np.ascontiguousarray(
                np.broadcast_to(g, (g.shape[0], N, g.shape[2])), dtype=np.float32) for g in self.gh_x]}
        return self._zs_cache[N]


class DGP(DGP_Base):
    """dgp.py:169-192."""
    def __init__(self, X, Y, Z, kernels, likelihood, num_outputs=None, mean_function=None, white=False, **kwargs):
        mean_function = Zero() if mean_function is None else mean_function
        layers = init_layers_linear(X, Y, Z, kernels, num_outputs=num_outputs, mean_function=mean_function,
                                    white=white)
        DGP_Base.__init__(self, X, Y, likelihood, layers, **kwargs)


class _SingleLayer(DGP_Base):
    """Private one-layer model so that a stand-alone SVGP_Layer can evaluate conditional_ND / KL on the device."""
    def _layer_conditional(self, layer, X, full_cov=False):
        Fs, Fmeans, Fvars = self.propagate(X, S=1, full_cov=full_cov)
        return Fmeans[0][0], Fvars[0][0]

    def _layer_propagate(self, layer, X, **kw):
        return self.propagate(X, **kw)


def _likelihood_model(likelihood, D):
    """Private one-layer model of output width D whose context evaluates `likelihood`'s epilogues (utils.py:88-121) for a
    BroadcastingLikelihood used outside a model."""
    from .kernels import RBF
    from .layers import SVGP_Layer
    layer = SVGP_Layer(RBF(1), np.zeros((1, 1)), D, Zero())
    K = likelihood.num_classes if isinstance(likelihood, MultiClass) else 0
    m = _SingleLayer.__new__(_SingleLayer)
    DGP_Base.__init__(m, np.zeros((1, 1)), np.zeros((1, 1 if K else D)), likelihood, [layer])
    return m


def _single_layer_model(layer):
    D = layer.num_outputs
    Din = layer.feature.Z.shape[1]
    return _SingleLayer(np.zeros((1, Din)), np.zeros((1, D)), Gaussian(), [layer])
