"""DGP_Quad (dgp.py:129-166) on the device: Gauss-Hermite nodes as the layers' z, node weights in the likelihood kernels.
Parity against the oracle's restatement, and the reference's own use of it: tests/test_collapsed.py:57-104 runs
NatGradOptimizer(gamma=1).minimize(m_ng, var_list=[[last q_mu, q_sqrt]], maxiter=1) on a DGP_Quad."""
import numpy as np
import pytest

from tests.synth import make_problem, round_f32
from tests.test_natgrad_cpu import well_conditioned_q

pytestmark = pytest.mark.gpu


def _build(prob, H):
    from doubly_stochastic_dgp import settings
    from doubly_stochastic_dgp.dgp import DGP_Quad
    from doubly_stochastic_dgp.kernels import RBF
    from doubly_stochastic_dgp.layers import SVGP_Layer
    from doubly_stochastic_dgp.likelihoods import Gaussian
    from doubly_stochastic_dgp.mean_functions import Identity, Zero
    from oracle import reference_dgp as R
    import torch
    settings.jitter = R.settings.jitter = prob['jitter']
    layers, olayers = [], []
    for lay in prob['layers']:
        mf, omf = (Identity(), R.Identity()) if lay['mean'] == 'identity' else (Zero(), R.Zero())
        l = SVGP_Layer(RBF(lay['din'], variance=lay['var'], lengthscales=lay['ls']), lay['Z'], lay['dout'], mf)
        l.q_mu = lay['q_mu']; l.q_sqrt = lay['q_sqrt']
        ol = R.SVGP_Layer(R.RBF(lay['din'], variance=lay['var'], lengthscales=lay['ls']), lay['Z'], lay['dout'], omf)
        ol.q_mu = torch.as_tensor(lay['q_mu']).clone(); ol.q_sqrt = torch.as_tensor(lay['q_sqrt']).clone()
        layers.append(l); olayers.append(ol)
    m = DGP_Quad(prob['X'], prob['Y'], Gaussian(prob['lik_var']), layers, H=H, num_data=prob['num_data'])
    o = R.DGP_Quad(prob['X'], prob['Y'], R.Gaussian(prob['lik_var']), olayers, H=H, num_data=prob['num_data'])
    return m, o


@pytest.mark.parametrize("dims,H", [([2, 1, 1], 30), ([2, 2, 1], 7)])
def test_quad_elbo_grad_and_natgrad_match_oracle(dims, H):
    from oracle import reference_dgp as R
    prob = round_f32(well_conditioned_q(make_problem(seed=1500 + H, dims=dims, N=40, M=10, S=1, inner_q_scale=0.3,
                                                     num_data=40)))
    m, o = _build(prob, H)
    assert m.num_samples == H ** dims[1]
    e = m.compute_log_likelihood()
    e_ref = o.compute_log_likelihood()
    assert abs(e - e_ref) <= 1e-4 * abs(e_ref), (e, e_ref)
    e2, grads, glik = m.compute_log_likelihood_and_grad()
    e2_ref, g_ref = o.elbo_and_grad()
    assert abs(e2 - e2_ref) <= 1e-4 * abs(e2_ref)
    for l, g in enumerate(grads):
        Z, q_mu, q_sqrt, var, ls = [x.numpy() for x in g_ref[5 * l:5 * l + 5]]
        for name, got, ref in (("Z", g['Z'], Z), ("q_mu", g['q_mu'], q_mu), ("q_sqrt", g['q_sqrt'], np.tril(q_sqrt))):
            sc = np.max(np.abs(ref)) + 1e-12
            np.testing.assert_allclose(got, ref, atol=1e-2 * sc, rtol=0, err_msg=f"{name} l={l}")
    # the reference's NatGrad use (tests/test_collapsed.py:99-100)
    from doubly_stochastic_dgp.training import NatGradOptimizer
    p = [[m.layers[-1].q_mu, m.layers[-1].q_sqrt]]
    NatGradOptimizer(gamma=1.).minimize(m, var_list=p, maxiter=1)
    R.natgrad_step(o, [len(o.layers) - 1], 1.0)
    e3, e3_ref = m.compute_log_likelihood(), o.compute_log_likelihood()
    assert e3_ref > e_ref
    assert abs(e3 - e3_ref) <= 1e-3 * abs(e3_ref), (e3, e3_ref)


def test_I3_quadrature_equals_mean_of_sampled_elbos_on_the_device():
    """Identity I3 of the reference (tests/test_dgp.py:120-174), on the device for BOTH models: the Gauss-Hermite ELBO of DGP_Quad
    lies within 3 standard errors of the mean of the Monte-Carlo ELBOs of DGP_Base (in-kernel Philox draws, a fresh seed per
    evaluation) with the same parameters, and the quadrature value is deterministic.  Fixture as in the reference: N = 2 points,
    two RBF(1, lengthscales=0.1) layers, Gaussian(0.01), H = 300 nodes, 100 samples per evaluation; the reference first fits
    (q_mu, q_sqrt) with L-BFGS, which the identity does not need -- random well-conditioned q here."""
    from doubly_stochastic_dgp.dgp import DGP_Base, DGP_Quad
    from doubly_stochastic_dgp.kernels import RBF
    from doubly_stochastic_dgp.layer_initializations import init_layers_linear
    from doubly_stochastic_dgp.likelihoods import Gaussian
    from doubly_stochastic_dgp import settings
    settings.jitter = 1e-6
    np.random.seed(0)
    N = 2
    X = np.random.uniform(size=(N, 1))
    Y = np.sin(20 * X) + np.random.randn(*X.shape) * 0.001
    rng = np.random.default_rng(4)
    q_mu = [0.5 * rng.normal(size=(N, 1)) for _ in range(2)]
    q_sqrt = [np.tril(0.2 * rng.normal(size=(1, N, N))) + 0.4 * np.eye(N)[None] for _ in range(2)]

    def layers():
        ls = init_layers_linear(X, Y, X, [RBF(1, lengthscales=0.1), RBF(1, lengthscales=0.1)])
        for l, mu, sq in zip(ls, q_mu, q_sqrt):
            l.q_mu = mu
            l.q_sqrt = sq
        return ls

    m_quad = DGP_Quad(X, Y, Gaussian(0.01), layers(), H=300)
    m_mc = DGP_Base(X, Y, Gaussian(0.01), layers(), num_samples=100)
    Lq = [m_quad.compute_log_likelihood() for _ in range(2)]
    np.testing.assert_allclose(Lq[0], Lq[1], rtol=1e-12)               # quadrature is deterministic (up to the order of the fp64 atomics)
    Ls = np.array([m_mc.compute_log_likelihood() for _ in range(1000)])
    mean, se = Ls.mean(), Ls.std() / np.sqrt(len(Ls))
    assert abs(Lq[0] - mean) < 3 * se + 1e-4 * abs(mean), (Lq[0], mean, se)   # 99.73 % CI (+ the fp32 path's own 1e-4)
