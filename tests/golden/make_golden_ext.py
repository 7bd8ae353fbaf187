"""Generates tests/golden/ext/*.npz: golden vectors for the paths added after the first fixture set -- NatGrad step,
full_cov propagate, prediction epilogues, DGP_Quad -- from the float64 oracle on the SAME seeded problems the GPU parity
tests use (tests/test_gpu_natgrad.py, test_gpu_full_cov.py, test_gpu_predict.py, test_gpu_quad.py).  As for
make_golden.py, the reference itself cannot run here; the oracle is pinned by the reference's identities.
Run from the repo root:   python tests/golden/make_golden_ext.py"""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "doubly-stochastic-dgp_b200"))
from oracle import reference_dgp as R  # noqa: E402
from tests.golden.make_golden import pack_problem  # noqa: E402
from tests.synth import build_oracle, make_problem, round_f32  # noqa: E402
from tests.test_natgrad_cpu import well_conditioned_q  # noqa: E402

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ext")


def natgrad_case(kw, ids, gamma):
    prob = round_f32(well_conditioned_q(make_problem(**kw)))
    o = build_oracle(prob)
    e0 = R.natgrad_step(o, ids, gamma, zs=prob['zs'])
    out = dict(ids=np.array(ids), gamma=gamma, elbo_before=e0, elbo_after=o.compute_log_likelihood(zs=prob['zs']))
    for l in ids:
        out[f"q_mu{l}"] = o.layers[l].q_mu.numpy()
        out[f"q_sqrt{l}"] = o.layers[l].q_sqrt.numpy()
    return prob, out


def full_cov_case(kw):
    prob = round_f32(make_problem(**kw))
    o = build_oracle(prob)
    Fs, Fm, Fv = o.propagate(prob['X'], full_cov=True, S=prob['S'], zs=prob['zs'])
    out = {}
    for l in range(len(Fs)):
        out[f"F{l}"], out[f"Fmean{l}"], out[f"Fvar{l}"] = Fs[l].numpy(), Fm[l].numpy(), Fv[l].numpy()
    return prob, out


def predict_case(kw):
    prob = round_f32(make_problem(**kw))
    o = build_oracle(prob)
    ym, yv = o.predict_y(prob['X'], prob['S'], zs=prob['zs'])
    dens = o.predict_density(prob['X'], prob['Y'], prob['S'], zs=prob['zs'])
    return prob, dict(y_mean=ym.numpy(), y_var=yv.numpy(), density=dens.numpy())


CASES = {
    # name: (builder, args)  -- same problems as the GPU parity tests
    "natgrad_gamma1_last_layer": (natgrad_case, (dict(seed=411, dims=[8, 8, 1], N=256, M=100, S=4, inner_q_scale=0.3,
                                                      num_data=2560), [1], 1.0)),
    "natgrad_small_gamma_all_layers": (natgrad_case, (dict(seed=412, dims=[3, 3, 3, 2], N=70, M=37, S=2, kern='matern52',
                                                           inner_q_scale=0.3, num_data=700), [0, 1, 2], 0.005)),
    "full_cov_dgp2": (full_cov_case, (dict(seed=901, dims=[3, 3, 2], N=40, M=12, S=3, inner_q_scale=0.3),)),
    "predict_gauss_dgp2": (predict_case, (dict(seed=800, dims=[8, 8, 2], N=150, M=40, S=5, inner_q_scale=0.3),)),
}


def pack_extra(prob):
    d = pack_problem(prob)
    d['n_classes'] = prob['n_classes']
    return d


if __name__ == "__main__":
    os.makedirs(HERE, exist_ok=True)
    for name, (fn, args) in CASES.items():
        prob, out = fn(*args)
        d = pack_extra(prob)
        d.update({"out_" + k: v for k, v in out.items()})
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, {k: (v if np.ndim(v) == 0 else np.shape(v)) for k, v in out.items()})
