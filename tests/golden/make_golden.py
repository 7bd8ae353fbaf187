"""Generates tests/golden/*.npz: seeded inputs and the float64 oracle's outputs (ELBO, per-layer propagate outputs,
gradients).  The reference (TF 1.8 / GPflow 1.1.1) cannot run here or on the GPU box, and its tests store no golden
vectors; these vectors come from oracle/reference_dgp.py, which is pinned by the reference's own test identities
(tests/test_oracle_identities.py).  Run from the repo root:   python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.synth import build_oracle, make_problem, round_f32  # noqa: E402

CASES = {
    "svgp_rbf_nonwhite": dict(seed=11, dims=[8, 1], N=100, M=10, S=1, num_data=1000),                 # BASELINE configs[0]
    "svgp_rbf_white": dict(seed=12, dims=[8, 1], N=100, M=10, S=1, white=True, num_data=1000),
    "dgp2_rbf": dict(seed=13, dims=[8, 8, 1], N=64, M=32, S=4, inner_q_scale=0.3, num_data=640),
    "dgp3_matern_ard_linear": dict(seed=14, dims=[5, 3, 4, 2], N=48, M=24, S=3, kern='matern52', ard=True,
                                   inner_q_scale=0.3, num_data=480),
    "dgp2_white": dict(seed=15, dims=[4, 4, 1], N=40, M=16, S=5, white=True, inner_q_scale=0.3, num_data=400),
}


def pack_problem(prob):
    d = dict(X=prob['X'], Y=prob['Y'], lik_var=prob['lik_var'], jitter=prob['jitter'], S=prob['S'], num_data=prob['num_data'],
             kern=prob['kern'], white=prob['white'], dims=np.array(prob['dims']), L=len(prob['layers']))
    for l, lay in enumerate(prob['layers']):
        for k in ('Z', 'q_mu', 'q_sqrt', 'ls', 'var'):
            d[f"l{l}_{k}"] = np.asarray(lay[k])
        d[f"l{l}_mean"] = lay['mean']
        d[f"l{l}_W"] = np.zeros((0, 0)) if lay['W'] is None else lay['W']
        d[f"z{l}"] = prob['zs'][l]
    return d


def unpack_problem(g):
    L = int(g['L'])
    dims = [int(x) for x in g['dims']]
    layers = []
    for l in range(L):
        W = g[f"l{l}_W"]
        ls = g[f"l{l}_ls"]
        layers.append(dict(kern=str(g['kern']), Z=g[f"l{l}_Z"], q_mu=g[f"l{l}_q_mu"], q_sqrt=g[f"l{l}_q_sqrt"],
                           ls=ls if ls.ndim else float(ls), var=float(g[f"l{l}_var"]), white=bool(g['white']),
                           mean=str(g[f"l{l}_mean"]), W=None if W.size == 0 else W, din=dims[l], dout=dims[l + 1],
                           last=l == L - 1))
    return dict(X=g['X'], Y=g['Y'], layers=layers, zs=[g[f"z{l}"] for l in range(L)], lik_var=float(g['lik_var']),
                jitter=float(g['jitter']), S=int(g['S']), N=g['X'].shape[0], M=layers[0]['Z'].shape[0],
                num_data=float(g['num_data']), dims=dims, kern=str(g['kern']), white=bool(g['white']), n_classes=0)


def oracle_outputs(prob):
    o = build_oracle(prob)
    Fs, Fm, Fv = o.propagate(prob['X'], S=prob['S'], zs=prob['zs'])
    e, grads = o.elbo_and_grad(zs=prob['zs'])
    out = dict(elbo=e)
    for l in range(len(Fs)):
        out[f"F{l}"], out[f"Fmean{l}"], out[f"Fvar{l}"] = Fs[l].numpy(), Fm[l].numpy(), Fv[l].numpy()
    i = 0
    for l in range(len(Fs)):
        Z, q_mu, q_sqrt, var, ls = [x.numpy() for x in grads[i:i + 5]]
        i += 5
        out[f"g{l}_Z"], out[f"g{l}_q_mu"], out[f"g{l}_q_sqrt"] = Z, q_mu, np.tril(q_sqrt)
        out[f"g{l}_variance"], out[f"g{l}_lengthscales"] = var, ls
    out["g_lik_variance"] = grads[i].numpy()
    return out


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    for name, kw in CASES.items():
        prob = round_f32(make_problem(**kw))          # inputs exactly representable in fp32
        d = pack_problem(prob)
        d.update({"out_" + k: v for k, v in oracle_outputs(prob).items()})
        np.savez_compressed(os.path.join(here, name + ".npz"), **d)
        print(name, d["out_elbo"])
