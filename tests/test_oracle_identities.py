"""Pins oracle/reference_dgp.py (the restatement) with the identities the reference's own tests
use, against INDEPENDENT closed forms (oracle/closed_form.py) standing in for GPflow SVGP/GPR.
Reference: tests/test_dgp.py:28-117 (I1, I2), tests/test_collapsed.py:57-104 (I6),
tests/test_utils.py:181-206 (I8)."""
import numpy as np
import pytest
import torch
from numpy.testing import assert_allclose

from oracle import closed_form as cf
from oracle import reference_dgp as R


@pytest.fixture(autouse=True)
def _jitter():
    old = R.settings.jitter
    yield
    R.settings.jitter = old


def _fixture():
    # tests/test_dgp.py:28-36
    Ns, N, D_X, D_Y = 20, 19, 2, 3
    np.random.seed(0)
    X = np.random.uniform(size=(N, D_X))
    Xs = np.random.uniform(size=(Ns, D_X))
    q_mu = np.random.randn(N, D_Y)
    q_sqrt = 0.001 * np.eye(N)[None, :, :] * np.ones((D_Y, 1, 1))
    Y = np.random.randn(N, D_Y)
    Ys = np.random.randn(Ns, D_Y)
    return X, Xs, q_mu, q_sqrt, Y, Ys


@pytest.mark.parametrize("white", [True, False])
@pytest.mark.parametrize("L", [1, 2])
def test_I1_I2_dgp_equals_svgp_gaussian(L, white):
    R.settings.jitter = 1e-18                       # tests/test_dgp.py:7-8
    jit = 1e-18
    if L == 2:
        # our float64 Cholesky of the 1e-24-variance inner Kuu needs the jitter to dominate rounding
        pass
    X, Xs, q_mu, q_sqrt, Y, Ys = _fixture()
    lik_var = 0.01
    kerns = [R.Matern52(2, variance=1e-24, lengthscales=0.5) for _ in range(L - 1)]
    kerns.append(R.Matern52(2, lengthscales=0.5))
    m = R.DGP(X, Y, X, kerns, R.Gaussian(lik_var), white=white, num_samples=2)
    m.layers[-1].q_mu = torch.as_tensor(q_mu)
    m.layers[-1].q_sqrt = torch.as_tensor(q_sqrt)
    L_dgp = m.compute_log_likelihood()
    L_svgp = cf.svgp_elbo_gaussian('matern52', 1.0, 0.5, X, q_mu, q_sqrt, X, Y, lik_var, white, jit)
    tol = 1e-7 if L == 1 else 1e-6                  # tests/test_dgp.py:101-106
    assert_allclose(L_dgp, L_svgp, rtol=tol, atol=tol)

    pm, pv = m.predict_f(Xs, 1)
    sm, sv = cf.svgp_predict_f('matern52', 1.0, 0.5, X, q_mu, q_sqrt, Xs, white, jit)
    assert_allclose(pm[0].numpy(), sm, rtol=tol, atol=tol)
    assert_allclose(pv[0].numpy(), sv, rtol=tol, atol=tol)
    ym, yv = m.predict_y(Xs, 1)
    assert_allclose(ym[0].numpy(), sm, rtol=tol, atol=tol)
    assert_allclose(yv[0].numpy(), sv + lik_var, rtol=tol, atol=tol)
    dens = m.predict_density(Xs, Ys, 1)
    ref = -0.5 * np.log(2 * np.pi) - 0.5 * np.log(sv + lik_var) - 0.5 * (Ys - sm) ** 2 / (sv + lik_var)
    assert_allclose(dens.numpy(), ref, rtol=tol, atol=tol)
    fm, fv = m.predict_f_full_cov(Xs, 1)
    cm, cv = cf.svgp_predict_f('matern52', 1.0, 0.5, X, q_mu, q_sqrt, Xs, white, jit, full_cov=True)
    assert_allclose(fm[0].numpy(), cm, rtol=tol, atol=tol)
    assert_allclose(fv[0].numpy(), cv, rtol=tol, atol=tol)


def test_I1_rbf_minibatch_scale():
    rng = np.random.default_rng(3)
    N, M, D = 30, 7, 4
    X = rng.normal(size=(N, D)); Y = rng.normal(size=(N, 2)); Z = rng.normal(size=(M, D))
    q_mu = rng.normal(size=(M, 2)); q_sqrt = np.tril(rng.normal(size=(2, M, M))) * 0.3 + np.eye(M)
    for white in (True, False):
        m = R.DGP(X, Y, Z, [R.RBF(D, variance=0.7, lengthscales=1.9)], R.Gaussian(0.2), white=white,
                  num_samples=3, num_data=500)
        m.layers[0].q_mu = torch.as_tensor(q_mu); m.layers[0].q_sqrt = torch.as_tensor(q_sqrt)
        ref = cf.svgp_elbo_gaussian('rbf', 0.7, 1.9, Z, q_mu, q_sqrt, X, Y, 0.2, white, 1e-6, num_data=500)
        assert_allclose(m.compute_log_likelihood(), ref, rtol=1e-9)


def test_faithful_equals_deduplicated():
    rng = np.random.default_rng(5)
    N, M, D = 25, 6, 3
    X = rng.normal(size=(N, D)); Y = rng.normal(size=(N, 1)); Z = rng.normal(size=(M, D))
    zs = [rng.normal(size=(4, N, D)), rng.normal(size=(4, N, 1))]
    vals = []
    for faithful in (True, False):
        R.SVGP_Layer.faithful = faithful
        m = R.DGP(X, Y, Z, [R.RBF(D, lengthscales=1.5), R.Matern52(D, lengthscales=1.5)], R.Gaussian(0.1),
                  num_samples=4)
        m.layers[0].q_mu = torch.as_tensor(rng.normal(size=(M, D)) * 0 + 0.3)
        vals.append(m.compute_log_likelihood(zs=zs))
    R.SVGP_Layer.faithful = True
    assert_allclose(vals[0], vals[1], rtol=1e-12)


def test_I6_natgrad_optimum_maximises_elbo():
    """The closed-form optimal q(u) (what NatGrad gamma=1 reaches, tests/test_collapsed.py:99-104)
    is a stationary point of the restated ELBO and gives log marginal of the exact GP when Z=X."""
    rng = np.random.default_rng(100)
    N, D = 12, 2
    X = rng.uniform(size=(N, D)); Y = rng.uniform(size=(N, 1))
    lik_var = 0.1
    m_opt, S_opt = cf.optimal_q_gaussian('rbf', 1.0, 0.3, X, X, Y, lik_var, 1e-6)
    m = R.DGP(X, Y, X, [R.RBF(D, lengthscales=0.3)], R.Gaussian(lik_var), num_samples=1)
    m.layers[0].q_mu = torch.as_tensor(m_opt)
    m.layers[0].q_sqrt = torch.as_tensor(np.linalg.cholesky(S_opt)[None])
    e, grads = m.elbo_and_grad()
    # Z=X, optimal q  =>  bound is tight up to jitter (test_collapsed.py:30-54 uses 1e-5)
    assert_allclose(e, cf.gpr_log_marginal('rbf', 1.0, 0.3, X, Y, lik_var), rtol=1e-5, atol=1e-5)
    assert float(grads[1].abs().max()) < 1e-6          # d/dq_mu
    assert float(torch.tril(grads[2]).abs().max()) < 1e-5   # d/dq_sqrt


def test_I8_reparameterize():
    # tests/test_utils.py:181-206
    S, N, D = 4, 3, 2
    rng = np.random.default_rng(0)
    mean = rng.normal(size=(S, N, D)); var = rng.normal(size=(S, N, D)) ** 2; z = rng.normal(size=(S, N, D))
    f = mean + z * (var + 1e-6) ** 0.5
    assert_allclose(f, R.reparameterize(torch.as_tensor(mean), torch.as_tensor(var), torch.as_tensor(z)).numpy())
    U = rng.normal(size=(S, N, N, D))
    var = np.einsum('SnNd,SmNd->Snmd', U, U) + np.eye(N)[None, :, :, None] * 1e-6
    var_flat = np.reshape(np.transpose(var, [0, 3, 1, 2]), [S * D, N, N])
    L_flat = np.linalg.cholesky(var_flat + np.eye(N)[None] * 1e-6)
    Lc = np.transpose(np.reshape(L_flat, [S, D, N, N]), [0, 2, 3, 1])
    f = mean + np.einsum('SnNd,SNd->Snd', Lc, z)
    got = R.reparameterize(torch.as_tensor(mean), torch.as_tensor(var), torch.as_tensor(z), full_cov=True)
    assert_allclose(f, got.numpy(), rtol=1e-10)


def test_multiclass_matches_monte_carlo():
    """RobustMax variational expectation vs brute-force MC of its definition."""
    rng = np.random.default_rng(1)
    K, Rn = 4, 6
    mu = rng.normal(size=(Rn, K)); var = rng.uniform(0.2, 1.0, size=(Rn, K)); Y = rng.integers(0, K, size=(Rn, 1))
    lik = R.MultiClass(K)
    ve = lik.variational_expectations(torch.as_tensor(mu), torch.as_tensor(var), torch.as_tensor(Y)).numpy()[:, 0]
    f = mu[None] + np.sqrt(var)[None] * rng.normal(size=(400000, Rn, K))
    correct = (np.argmax(f, -1) == Y[:, 0][None])
    eps = 1e-3
    mc = np.mean(np.where(correct, np.log(1 - eps), np.log(eps / (K - 1))), 0)
    assert_allclose(ve, mc, atol=2e-2)
