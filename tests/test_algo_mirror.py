"""The kernels' algorithm (tests/algo_mirror.py: dedup'd layer 1, whitened projections, hand-derived
adjoints) == the reference restatement + autograd, in float64 on CPU."""
import numpy as np
import pytest
from numpy.testing import assert_allclose

from tests import algo_mirror as A
from tests.synth import build_oracle, make_problem


def _mirror_layers(prob):
    return [A.LayerP(prob['kern'], l['Z'], l['q_mu'], l['q_sqrt'], l['ls'], l['var'], l['white'], l['mean'],
                     W=l['W'], bvec=None if l['W'] is None else np.zeros(l['dout'])) for l in prob['layers']]


CASES = [
    dict(dims=[3, 1], N=17, M=6, S=1),
    dict(dims=[3, 3, 1], N=13, M=5, S=3),
    dict(dims=[3, 3, 3, 2], N=11, M=7, S=2, kern='matern52'),
    dict(dims=[4, 2, 3, 1], N=9, M=5, S=2, ard=True),
]


@pytest.mark.parametrize("white", [False, True])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_mirror_matches_oracle_autograd(case, white):
    kw = dict(CASES[case])
    prob = make_problem(seed=100 + case, white=white, inner_q_scale=0.3, num_data=50, **kw)
    m = build_oracle(prob)
    e_ref, g_ref = m.elbo_and_grad(zs=prob['zs'])
    e, grads, lvbar = A.elbo_and_grad(_mirror_layers(prob), prob['X'], prob['Y'], prob['lik_var'], prob['S'],
                                      prob['zs'], prob['num_data'], prob['jitter'])
    assert_allclose(e, e_ref, rtol=1e-10)
    i = 0
    for l, g in enumerate(grads):
        Z, q_mu, q_sqrt, var, ls = [x.numpy() for x in g_ref[i:i + 5]]
        i += 5
        assert_allclose(g['Z'], Z, rtol=1e-6, atol=1e-8, err_msg=f"Z l={l}")
        assert_allclose(g['q_mu'], q_mu, rtol=1e-6, atol=1e-8, err_msg=f"q_mu l={l}")
        assert_allclose(g['q_sqrt'], np.tril(q_sqrt), rtol=1e-6, atol=1e-8, err_msg=f"q_sqrt l={l}")
        assert_allclose(g['var'], var, rtol=1e-6, atol=1e-8, err_msg=f"var l={l}")
        assert_allclose(g['ls'], ls, rtol=1e-6, atol=1e-8, err_msg=f"ls l={l}")
    assert_allclose(lvbar, g_ref[i].numpy(), rtol=1e-8)
