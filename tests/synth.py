"""Synthetic problem generator shared by tests, bench.py and the golden-vector script
(seeds 1000*config + replicate, X~N(0,1), lengthscale sqrt(D_in), ...)."""
import numpy as np


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


def build_oracle(prob, faithful=False):
    """Instantiate oracle/reference_dgp.py for a problem dict."""
    import torch
    from oracle import reference_dgp as R
    R.settings.jitter = prob['jitter']
    R.SVGP_Layer.faithful = faithful
    kcls = R.RBF if prob['kern'] == 'rbf' else R.Matern52
    mfs = {'zero': lambda l: R.Zero(), 'identity': lambda l: R.Identity(), 'linear': lambda l: R.Linear(l['W'])}
    layers = []
    for lay in prob['layers']:
        kern = kcls(lay['din'], variance=lay['var'], lengthscales=lay['ls'])
        layer = R.SVGP_Layer(kern, lay['Z'], lay['dout'], mfs[lay['mean']](lay), white=lay['white'])
        layer.q_mu = torch.as_tensor(lay['q_mu']).clone()
        layer.q_sqrt = torch.as_tensor(lay['q_sqrt']).clone()
        layers.append(layer)
    lik = R.MultiClass(prob['n_classes']) if prob['n_classes'] else R.Gaussian(prob['lik_var'])
    return R.DGP_Base(prob['X'], prob['Y'], lik, layers, num_samples=prob['S'], num_data=prob['num_data'])
