"""NatGrad on CPU: (i) the oracle's restatement of GPflow's NatGradOptimizer is pinned by the reference's identity I6
(gamma = 1 on a Gaussian last layer == SGPR optimum, tests/test_collapsed.py:57-104); (ii) the direct update the CUDA
kernels perform from the row-reduced accumulators (tests/algo_mirror.py::natgrad_update, csrc/natgrad.cu) equals the
oracle's theta-space update for any gamma, any layer, white or not."""
import numpy as np
import pytest
import torch
from numpy.testing import assert_allclose

from oracle import closed_form as cf
from oracle import reference_dgp as R
from tests import algo_mirror as A
from tests.synth import build_oracle, make_problem
from tests.test_algo_mirror import _mirror_layers


def test_I6_oracle_natgrad_gamma1_reaches_sgpr_optimum():
    rng = np.random.default_rng(5)
    N, D, M = 20, 2, 7
    R.settings.jitter = 1e-6
    X = rng.uniform(size=(N, D)); Y = rng.uniform(size=(N, 1)); Z = X[:M] + 0.01
    lik_var = 0.1
    m = R.DGP(X, Y, Z, [R.RBF(D, lengthscales=0.5)], R.Gaussian(lik_var), num_samples=1)
    m.layers[0].q_mu = torch.as_tensor(rng.normal(size=(M, 1)))
    m.layers[0].q_sqrt = torch.as_tensor(np.tril(rng.normal(size=(1, M, M))) + 2 * np.eye(M)[None])
    R.natgrad_step(m, [0], 1.0)
    m_opt, S_opt = cf.optimal_q_gaussian('rbf', 1.0, 0.5, Z, X, Y, lik_var, 1e-6)
    assert_allclose(m.layers[0].q_mu.numpy(), m_opt, rtol=1e-6, atol=1e-8)
    Lq = m.layers[0].q_sqrt.numpy()[0]
    assert_allclose(Lq @ Lq.T, S_opt, rtol=1e-6, atol=1e-9)
    # stationary: a second gamma=1 step does not move
    before = m.layers[0].q_mu.clone()
    R.natgrad_step(m, [0], 1.0)
    assert_allclose(m.layers[0].q_mu.numpy(), before.numpy(), rtol=1e-6, atol=1e-9)


def positive_diag(prob):
    """Flip the sign of q_sqrt columns whose diagonal entry is negative (S = q_sqrt q_sqrt^T is unchanged).  GPflow
    back-propagates dL/dq_sqrt through chol(S), which is the chain rule at q_sqrt only when q_sqrt IS chol(S), i.e. has a
    positive diagonal -- always true once NatGrad manages the variable (it writes chol(S')).  The direct update is the
    exact natural gradient either way; parity with GPflow's route is defined on positive-diagonal factors."""
    for lay in prob['layers']:
        q = lay['q_sqrt']
        sg = np.sign(np.einsum('dii->di', q))
        sg[sg == 0] = 1.0
        lay['q_sqrt'] = q * sg[:, None, :]
    return prob


from workloads import well_conditioned_q  # noqa: E402,F401  (re-exported: other test modules import it from here)


CASES = [
    dict(dims=[3, 1], N=17, M=6, S=1),
    dict(dims=[3, 3, 1], N=13, M=5, S=3),
    dict(dims=[3, 3, 3, 2], N=11, M=7, S=2, kern='matern52'),
]


@pytest.mark.parametrize("gamma", [1.0, 0.02])
@pytest.mark.parametrize("white", [False, True])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_direct_update_equals_theta_space_update(case, white, gamma):
    prob = make_problem(seed=300 + case, white=white, inner_q_scale=0.3, num_data=40, **CASES[case])
    L = len(prob['layers'])
    positive_diag(prob)
    layers = _mirror_layers(prob)
    aux = {}
    A.elbo_and_grad(layers, prob['X'], prob['Y'], prob['lik_var'], prob['S'], prob['zs'], prob['num_data'],
                    prob['jitter'], aux=aux)
    o = build_oracle(prob)
    # gamma = 1 is only meaningful where the likelihood term is concave in f (the Gaussian last layer, as the reference
    # uses it); a small step is taken on every layer
    ids = [L - 1] if gamma == 1.0 else list(range(L))
    R.natgrad_step(o, ids, gamma, zs=prob['zs'])
    for l in ids:
        mu, sq = A.natgrad_update(layers[l], aux[l]['Kinv'], aux[l]['Pd'], aux[l]['qmubar'], gamma)
        assert_allclose(mu, o.layers[l].q_mu.numpy(), rtol=1e-6, atol=1e-9, err_msg=f"q_mu l={l}")
        assert_allclose(sq, o.layers[l].q_sqrt.numpy(), rtol=1e-6, atol=1e-9, err_msg=f"q_sqrt l={l}")
