"""TEST INFRASTRUCTURE: a stand-in for doubly_stochastic_dgp._lib.Context whose "device" is the float64 oracle.

It exists so that the HOST logic of the product package (dgp.py, layers.py, training.py: parameter plumbing, dirty flags,
trainable flags, var_list -> layer mapping, DGP_Quad's nodes/weights, shapes of the full_cov / predict paths) is exercised
by `-m "not gpu"` tests in a container without a GPU.  It is installed by monkeypatching `_lib.Context` inside those tests
only; nothing in the product package can reach it (the product path fails loudly without libdsdgp.so + a CUDA device)."""
import math

import numpy as np
import torch

from doubly_stochastic_dgp import _lib
from oracle import reference_dgp as R


class FakeContext:
    def __init__(self, layer_descs, likelihood, num_classes, D_y, jitter, N_max, S_max, device=0):
        self.descs = [tuple(d) for d in layer_descs]
        self.L = len(self.descs)
        self.lik_code, self.K, self.D_y, self.jitter = likelihood, num_classes, D_y, jitter
        self.N_max, self.S_max = N_max, S_max
        self.p = {}
        for l, (M, Din, Dout, kern, ard, white, mean, kwhite, ipd) in enumerate(self.descs):
            self.p[(l, _lib.F_WHITE_VARIANCE)] = np.ones(()) if kwhite else None
            self.p[(l, _lib.F_Z)] = np.zeros((M, Din))
            self.p[(l, _lib.F_Q_MU)] = np.zeros((M, Dout))
            self.p[(l, _lib.F_Q_SQRT)] = np.tile(np.eye(M)[None], (Dout, 1, 1))
            self.p[(l, _lib.F_LENGTHSCALES)] = np.ones(Din if ard else ())
            self.p[(l, _lib.F_VARIANCE)] = np.ones(())
            self.p[(l, _lib.F_MEAN_W)] = np.zeros((Din, Dout))
            self.p[(l, _lib.F_MEAN_B)] = np.zeros(Dout)
        self.p[(-1, _lib.F_LIK_VARIANCE)] = np.ones(())
        self.trainable = {}
        self.grads = {}
        self.weights = None
        self.adam = None
        self.closed = False
        self.calls = []

    # ---- parameters
    def _key(self, layer, field):
        return (-1 if field == _lib.F_LIK_VARIANCE else layer, field)

    def set_param(self, layer, field, value):
        k = self._key(layer, field)
        v = np.asarray(value, dtype=np.float32).astype(np.float64).reshape(self.p[k].shape)
        if field == _lib.F_Q_SQRT:
            v = np.tril(v)
        self.p[k] = v
        self.adam = None if self.adam is None else self.adam     # (the real ctx re-derives its free variables)

    def get_param(self, layer, field, shape):
        return self.p[self._key(layer, field)].reshape(shape).copy()

    def get_grad(self, layer, field, shape):
        return self.grads[self._key(layer, field)].reshape(shape).copy()

    def set_trainable(self, layer, field, flag):
        self.trainable[self._key(layer, field)] = bool(flag)

    def set_sample_weights(self, w):
        self.weights = None if w is None else np.asarray(w, dtype=np.float64)

    # ---- oracle model from the current parameters
    def _model(self, S, num_data=1.0, X=None, Y=None):
        R.settings.jitter = self.jitter
        layers = []
        for l, (M, Din, Dout, kern, ard, white, mean, kwhite, ipd) in enumerate(self.descs):
            kcls = R.RBF if kern == 0 else R.Matern52
            k = kcls(Din, variance=float(self.p[(l, _lib.F_VARIANCE)]), lengthscales=self.p[(l, _lib.F_LENGTHSCALES)])
            if kwhite:
                k = R.Sum([k, R.White(Din, variance=float(self.p[(l, _lib.F_WHITE_VARIANCE)]))])
            mf = [R.Zero(), R.Identity(), None][mean] or R.Linear(self.p[(l, _lib.F_MEAN_W)], self.p[(l, _lib.F_MEAN_B)])
            lay = R.SVGP_Layer(k, self.p[(l, _lib.F_Z)], Dout, mf, white=bool(white), input_prop_dim=ipd or None)
            lay.q_mu = torch.as_tensor(self.p[(l, _lib.F_Q_MU)]).clone()
            lay.q_sqrt = torch.as_tensor(self.p[(l, _lib.F_Q_SQRT)]).clone()
            layers.append(lay)
        lik = R.Gaussian(float(self.p[(-1, _lib.F_LIK_VARIANCE)])) if self.lik_code == 0 else R.MultiClass(self.K)
        X = np.zeros((1, self.descs[0][1])) if X is None else X
        Y = np.zeros((1, self.D_y)) if Y is None else Y
        return R.DGP_Base(X, Y, lik, layers, num_samples=S, num_data=num_data)

    def _pull(self, m):
        for l, lay in enumerate(m.layers):
            self.p[(l, _lib.F_Z)] = lay.Z.detach().numpy().copy()
            self.p[(l, _lib.F_Q_MU)] = lay.q_mu.detach().numpy().copy()
            self.p[(l, _lib.F_Q_SQRT)] = np.tril(lay.q_sqrt.detach().numpy())
            self.p[(l, _lib.F_LENGTHSCALES)] = lay.kern.lengthscales.detach().numpy().copy()
            self.p[(l, _lib.F_VARIANCE)] = lay.kern.variance.detach().numpy().copy()
        if self.lik_code == 0:
            self.p[(-1, _lib.F_LIK_VARIANCE)] = m.likelihood.likelihood.variance.detach().numpy().copy()

    def _zs(self, zs, N, S, seed):
        rng = np.random.default_rng(seed % (1 << 32))
        out = []
        for l, d in enumerate(self.descs):
            z = None if zs is None or l >= len(zs) else zs[l]
            out.append(rng.normal(size=(S, N, d[2])) if z is None else np.asarray(z, dtype=np.float64))
        return out

    def _elbo_tensor(self, m, X, Y, zs):
        if self.weights is None:
            return m.elbo(X, Y, zs)
        Fm, Fv = m._build_predict(R._t(X), S=len(self.weights), zs=zs)
        ve = m.likelihood.variational_expectations(Fm, Fv, R._t(Y))
        Lsum = torch.sum(ve * R._t(self.weights)[:, None, None])
        return Lsum * (float(m.num_data) / X.shape[0]) - sum(l.KL() for l in m.layers)

    # ---- compute entry points (same signatures as _lib.Context)
    def elbo(self, X, Y, S, num_data, zs=None, seed=0):
        self.calls.append(("elbo", S))
        m = self._model(S, num_data)
        with torch.no_grad():
            return float(self._elbo_tensor(m, np.asarray(X, dtype=np.float64), np.asarray(Y, dtype=np.float64),
                                           self._zs(zs, len(X), S, seed)))

    def elbo_grad(self, X, Y, S, num_data, zs=None, seed=0, _model=None):
        self.calls.append(("elbo_grad", S))
        m = _model or self._model(S, num_data)
        ps = m.parameters()
        for p in ps:
            p.requires_grad_(True)
        e = self._elbo_tensor(m, np.asarray(X, dtype=np.float64), np.asarray(Y, dtype=np.float64),
                              self._zs(zs, len(X), S, seed))
        e.backward()
        g = [p.grad.numpy().copy() if p.grad is not None else np.zeros(p.shape) for p in ps]
        for p in ps:
            p.requires_grad_(False)
            p.grad = None
        for l in range(self.L):
            Z, q_mu, q_sqrt, var, ls = g[5 * l:5 * l + 5]
            self.grads[(l, _lib.F_Z)], self.grads[(l, _lib.F_Q_MU)] = Z, q_mu
            self.grads[(l, _lib.F_Q_SQRT)] = np.tril(q_sqrt)
            self.grads[(l, _lib.F_VARIANCE)], self.grads[(l, _lib.F_LENGTHSCALES)] = var, ls
        if self.lik_code == 0:
            self.grads[(-1, _lib.F_LIK_VARIANCE)] = g[5 * self.L]
        return float(e.detach())

    def adam_init(self, lr=0.01, beta1=0.9, beta2=0.999, eps=1e-8):
        self.adam = dict(lr=lr, t=0)

    def train_step(self, X, Y, N, S, num_data, seed, flags=0, want_elbo=True, zs=None):
        """Plain gradient ascent on the trainable fields (enough for host-logic tests: it moves exactly the trainable
        parameters and leaves the rest alone; the real optimiser is tested on the GPU against oracle AdamState)."""
        assert self.adam is not None, "call adam_init first"
        e = self.elbo_grad(X, Y, S, num_data, zs=zs, seed=seed)
        for k, g in self.grads.items():
            if self.trainable.get(k, True) and k[1] not in (_lib.F_LENGTHSCALES, _lib.F_VARIANCE, _lib.F_LIK_VARIANCE):
                self.p[k] = self.p[k] + 1e-6 * np.sign(g).reshape(self.p[k].shape)
        return e

    def natgrad_step(self, X, Y, N, S, num_data, seed, layers, gamma, flags=0, zs=None):
        self.calls.append(("natgrad", tuple(layers), gamma))
        m = self._model(S, num_data, X, Y)
        zz = self._zs(zs, N, S, seed)
        if self.weights is None:
            e = R.natgrad_step(m, list(layers), gamma, zs=zz)
        else:            # DGP_Quad: same update, weighted likelihood
            fake = self

            class _M(R.DGP_Base):
                def elbo(self_, X=None, Y=None, zs=None):
                    return fake._elbo_tensor(self_, np.asarray(self_.X), np.asarray(self_.Y), zz)
            m.__class__ = _M
            e = R.natgrad_step(m, list(layers), gamma)
        self._pull(m)
        return e

    def propagate(self, X, S, zs=None, seed=0, want=(True, True, True), flags=0):
        m = self._model(S)
        with torch.no_grad():
            out = m.propagate(np.asarray(X, dtype=np.float64), S=S, zs=self._zs(zs, len(X), S, seed))
        return [[a.numpy().astype(np.float32) for a in lst] for lst in out]

    def propagate_full_cov(self, X, S, zs=None, seed=0):
        m = self._model(S)
        with torch.no_grad():
            out = m.propagate(np.asarray(X, dtype=np.float64), full_cov=True, S=S, zs=self._zs(zs, len(X), S, seed))
        return [[a.numpy().astype(np.float32) for a in lst] for lst in out]

    def predict_y(self, X, S, zs=None, seed=0):
        m = self._model(S)
        a, b = m.predict_y(np.asarray(X, dtype=np.float64), S, zs=self._zs(zs, len(X), S, seed))
        return a.numpy().astype(np.float32), b.numpy().astype(np.float32)

    def predict_density(self, X, Y, S, zs=None, seed=0):
        m = self._model(S)
        return m.predict_density(np.asarray(X, dtype=np.float64), np.asarray(Y, dtype=np.float64), S,
                                 zs=self._zs(zs, len(X), S, seed)).numpy().astype(np.float32)

    def kl(self):
        m = self._model(1)
        with torch.no_grad():
            return np.array([float(l.KL()) for l in m.layers])

    def comm_init(self, *a):
        pass

    def set_option(self, *a):
        pass

    def sync(self):
        pass

    def launch_count(self):
        return 0

    def close(self):
        self.closed = True
