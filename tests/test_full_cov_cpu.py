"""full_cov=True on CPU: the kernel-order NumPy mirror of csrc/full_cov.cu (tests/algo_mirror.py::full_cov_propagate) equals
the oracle's restatement of layers.py:52-74,178-219 (full_cov branch) and utils.py:43-51."""
import numpy as np
import pytest
from numpy.testing import assert_allclose

from tests import algo_mirror as A
from tests.synth import build_oracle, make_problem
from tests.test_algo_mirror import _mirror_layers

CASES = [
    dict(dims=[3, 1], N=17, M=6, S=2),
    dict(dims=[3, 3, 2], N=13, M=5, S=3),
    dict(dims=[4, 2, 3, 1], N=9, M=5, S=2, ard=True, kern='matern52'),
]


@pytest.mark.parametrize("white", [False, True])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_full_cov_mirror_matches_oracle(case, white):
    prob = make_problem(seed=600 + case, white=white, inner_q_scale=0.3, **CASES[case])
    o = build_oracle(prob)
    oFs, oFm, oFv = o.propagate(prob['X'], full_cov=True, S=prob['S'], zs=prob['zs'])
    Fs, Fm, Fv = A.full_cov_propagate(_mirror_layers(prob), prob['X'], prob['S'], prob['zs'], prob['jitter'])
    for l in range(len(Fs)):
        assert oFv[l].shape == Fv[l].shape
        assert_allclose(Fm[l], oFm[l].numpy(), rtol=1e-8, atol=1e-9, err_msg=f"mean l={l}")
        assert_allclose(Fv[l], oFv[l].numpy(), rtol=1e-7, atol=1e-9, err_msg=f"var l={l}")
        assert_allclose(Fs[l], oFs[l].numpy(), rtol=1e-7, atol=1e-8, err_msg=f"F l={l}")
    # the diagonal of the full covariance is the full_cov=False variance (layers.py:206-213)
    dFs, dFm, dFv = o.propagate(prob['X'], full_cov=False, S=prob['S'], zs=prob['zs'])
    assert_allclose(np.einsum('snnd->snd', Fv[0]), dFv[0].numpy(), rtol=1e-7, atol=1e-9)
