"""Synthetic workloads (seeds 1000*config + replicate, X~N(0,1), lengthscale sqrt(D_in), ...) and the
product-model builder for them.  Shared by bench.py, tools/, __graft_entry__.smoke() and the tests; NumPy + the product
package only (the oracle-side builder lives in oracle/problems.py, test infrastructure)."""
import os
import sys

import numpy as np

_PKG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "doubly-stochastic-dgp_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)


def _gram(kern, Z, ls, var):
    d = (Z[:, None, :] - Z[None, :, :]) / ls
    r2 = np.sum(d * d, -1)
    if kern == 'rbf':
        return var * np.exp(-0.5 * r2)
    r = np.sqrt(r2 + 1e-12)
    return var * (1 + np.sqrt(5) * r + 5 / 3 * r * r) * np.exp(-np.sqrt(5) * r)


def make_problem(dims, N, M, S, seed, kern='rbf', white=False, ard=False, inner_var=0.05,
                 final_var=1.0, lik_var=0.05, jitter=1e-6, inner_q_scale=1e-5, num_data=None,
                 n_classes=0, max_cond=3e4):
    """dims = [D_in, D_1, ..., D_L].  Returns a dict of float64 arrays (per-layer lists).
    Lengthscales start at sqrt(D_in) and are shrunk until cond(Kuu + jitter I) <= max_cond, the
    conditioning regime of the north-star config (1.7e4): the per-row kernels are fp32, whose error scales as
    eps_fp32 * cond(Kuu) (DESIGN.md "Numerics"), so parity problems are kept where fp32 is meaningful."""
    rng = np.random.default_rng(seed)
    L = len(dims) - 1
    X = rng.normal(size=(N, dims[0]))
    if n_classes:
        Y = rng.integers(0, n_classes, size=(N, 1)).astype(np.float64)
    else:
        Y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.normal(size=(N, 1))
        Y = np.tile(Y, (1, dims[-1])) + 0.01 * rng.normal(size=(N, dims[-1]))
    idx = rng.choice(N, size=M, replace=M > N)
    Z0 = X[idx] + 0.3 * rng.normal(size=(M, dims[0]))
    layers = []
    Zrun = Z0
    for l in range(L):
        din, dout = dims[l], dims[l + 1]
        last = l == L - 1
        ls = np.sqrt(din) * (np.ones(din) * (1.0 + 0.1 * rng.uniform(size=din)) if ard else 1.0)
        var = final_var if last else inner_var
        W = None
        if last:
            mean = 'zero'
        elif din == dout:
            mean = 'identity'
        else:
            mean = 'linear'
            if din > dout:
                Q, _ = np.linalg.qr(rng.normal(size=(din, dout)))
                W = Q
            else:
                W = np.concatenate([np.eye(din), np.zeros((din, dout - din))], 1)
        Z = Zrun.copy()
        for _ in range(40):
            if max_cond is None or np.linalg.cond(_gram(kern, Z, ls, var) + jitter * np.eye(M)) <= max_cond:
                break
            ls = ls * 0.85
        q_mu = 0.3 * rng.normal(size=(M, dout))
        layers.append(dict(kern=kern, Z=Z, q_mu=q_mu, ls=ls, var=var, white=white, mean=mean, W=W,
                           din=din, dout=dout, last=last))
        if W is not None:
            Zrun = Zrun @ W
    # q_sqrt needs Kuu -> filled by the caller-independent helper below
    for lay in layers:
        K = _gram(kern, lay['Z'], lay['ls'], lay['var'])
        Lu = np.linalg.cholesky(K + jitter * np.eye(M))
        if lay['last']:
            q = np.tril(0.1 * rng.normal(size=(lay['dout'], M, M))) + 0.3 * np.eye(M)[None]
        else:
            base = np.eye(M) if white else Lu
            q = np.tile((inner_q_scale * base)[None], (lay['dout'], 1, 1))
        lay['q_sqrt'] = q
    zs = [rng.normal(size=(S, N, lay['dout'])) for lay in layers]
    return dict(X=X, Y=Y, layers=layers, zs=zs, lik_var=lik_var, jitter=jitter, S=S, N=N, M=M,
                num_data=num_data or N, dims=list(dims), kern=kern, white=white, n_classes=n_classes)


def well_conditioned_q(prob, seed=0):
    """Replace the last layer's q_sqrt (make_problem: tril(0.1 N(0,1)) + 0.3 I, whose S = q q^T has a condition number growing
    exponentially with M -- 5e10 at M=100, not numerically positive definite at M=512) by 0.3 I + tril(N(0,1)) 0.1/sqrt(M).
    A natural-gradient step with gamma < 1 forms S^-1 (as GPflow's does), which needs S to be numerically positive definite."""
    rng = np.random.default_rng(seed)
    lay = prob['layers'][-1]
    D, M, _ = lay['q_sqrt'].shape
    lay['q_sqrt'] = 0.3 * np.eye(M)[None] + np.tril(rng.normal(size=(D, M, M)), -1) * (0.1 / np.sqrt(M))
    return prob


def round_f32(prob):
    """Round every input the device sees in fp32 to fp32 (returned as float64), so that the
    oracle and the CUDA path evaluate the SAME problem."""
    f = lambda a: None if a is None else np.asarray(a, dtype=np.float32).astype(np.float64)
    out = dict(prob)
    out['X'], out['Y'] = f(prob['X']), f(prob['Y'])
    out['zs'] = [f(z) for z in prob['zs']]
    out['lik_var'] = float(np.float32(prob['lik_var']))
    out['layers'] = []
    for lay in prob['layers']:
        l2 = dict(lay)
        for k in ('Z', 'q_mu', 'q_sqrt', 'W'):
            l2[k] = f(lay[k])
        l2['ls'] = f(lay['ls']) if np.ndim(lay['ls']) else float(np.float32(lay['ls']))
        l2['var'] = float(np.float32(lay['var']))
        out['layers'].append(l2)
    return out


def build_model(prob, **kw):
    """The product model (ctypes -> libdsdgp.so) for a problem dict of make_problem()."""
    from doubly_stochastic_dgp import settings
    from doubly_stochastic_dgp.dgp import DGP_Base
    from doubly_stochastic_dgp.kernels import RBF, Matern52
    from doubly_stochastic_dgp.layers import SVGP_Layer
    from doubly_stochastic_dgp.likelihoods import Gaussian, MultiClass
    from doubly_stochastic_dgp.mean_functions import Identity, Linear, Zero
    settings.jitter = prob['jitter']
    from doubly_stochastic_dgp.kernels import White
    from doubly_stochastic_dgp.likelihoods import Bernoulli
    kcls = RBF if prob['kern'] == 'rbf' else Matern52
    layers = []
    for lay in prob['layers']:
        kern = kcls(lay['din'], variance=lay['var'], lengthscales=lay['ls'])
        if lay.get('wvar') is not None:           # Sum(kernel, White)  (demos/run_regression.py:65-66)
            kern = kern + White(lay['din'], variance=lay['wvar'])
        mf = {'zero': Zero, 'identity': Identity}.get(lay['mean'], None)
        mf = mf() if mf else Linear(lay['W'])
        layer = SVGP_Layer(kern, lay['Z'], lay['dout'], mf, white=lay['white'], input_prop_dim=lay.get('ipd'))
        layer.q_mu = lay['q_mu']
        layer.q_sqrt = lay['q_sqrt']
        layers.append(layer)
    if prob.get('lik') == 'bernoulli':
        lik = Bernoulli()
    else:
        lik = MultiClass(prob['n_classes']) if prob['n_classes'] else Gaussian(prob['lik_var'])
    return DGP_Base(prob['X'], prob['Y'], lik, layers, num_samples=prob['S'], num_data=prob['num_data'], **kw)


