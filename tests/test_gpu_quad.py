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
