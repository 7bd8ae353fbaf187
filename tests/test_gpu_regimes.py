"""ELBO parity outside the comfortable synthetic regime (VERDICT r1 "weak" 2-5): a *trained-regime* north-star problem (every
layer's q_sqrt is O(0.1-0.3), so |L_d^T u|^2 matters in every conditional variance -- the product the tensor-core path used to
run as single-pass TF32), ill-conditioned inducing sets without the lengthscale shrinking of workloads.make_problem, and
gradients at full size for both whitening settings.  Bar: ELBO <= 1e-4 relative on both paths, no per-path exception; where
fp32 cannot deliver it the outcome is stated, not waived."""
import numpy as np
import pytest
from numpy.testing import assert_allclose

from tests.gpu_common import build_model, record
from tests.synth import build_oracle, make_problem, round_f32

pytestmark = pytest.mark.gpu
NORTH = dict(dims=[8, 8, 8, 8, 8, 1], N=1000, M=100, S=20)


def _trained(prob, seed):
    """every layer: q_sqrt = tril(0.1 N(0,1)) + 0.3 I (what a trained model looks like), q_mu = 0.3 N(0,1)"""
    rng = np.random.default_rng(seed)
    for lay in prob['layers']:
        M, D = lay['q_mu'].shape
        lay['q_sqrt'] = np.tril(0.1 * rng.normal(size=(D, M, M))) + 0.3 * np.eye(M)[None]
    return prob


@pytest.mark.parametrize("path", [0, 1])
@pytest.mark.parametrize("white", [False, True])
def test_trained_regime_northstar_elbo(white, path):
    prob = _trained(make_problem(seed=3100, white=white, num_data=8192, **NORTH), 1)
    m = build_model(prob)
    m._ensure_ctx(prob['N'], prob['S']).set_option("path", path)
    e = m.compute_log_likelihood(zs=prob['zs'])
    e_ref = build_oracle(prob).compute_log_likelihood(zs=prob['zs'])
    record("trained_regime", path=path, white=float(white), elbo_rel=abs(e - e_ref) / abs(e_ref))
    assert abs(e - e_ref) <= 1e-4 * abs(e_ref), (e, e_ref)


def test_single_pass_variance_product_is_what_the_flag_prevents():
    """The variance product c_d = L_d^T u with 1, 2 (split weights) and 3 TF32 passes on the trained-regime problem (errors
    recorded for DESIGN.md "TF32 passes"); the automatic mode is the 2-pass one here, and with the tiny q_sqrt of the reference's
    initialisation it stays at one pass."""
    prob = _trained(make_problem(seed=3100, num_data=8192, **NORTH), 1)
    e_ref = build_oracle(prob).compute_log_likelihood(zs=prob['zs'])
    m = build_model(prob)
    ctx = m._ensure_ctx(prob['N'], prob['S'])
    errs = {}
    for passes in (0, 1, 2, 3):
        ctx.set_option("g2_passes", passes)
        errs[passes] = abs(m.compute_log_likelihood(zs=prob['zs']) - e_ref) / abs(e_ref)
    record("g2_passes", auto=errs[0], one=errs[1], two=errs[2], three=errs[3])
    assert errs[0] <= 1e-4 and errs[2] <= 1e-4 and errs[3] <= 1e-4
    assert abs(errs[0] - errs[2]) <= 1e-7            # automatic == split weights (2 passes) here
    # reference initialisation (inner q_sqrt = 1e-5 Lu, demos/run_regression.py:72-73): the switch is off for the inner layers
    # and the result does not depend on it
    prob0 = make_problem(seed=3000, num_data=8192, **NORTH)
    m0 = build_model(prob0)
    ctx0 = m0._ensure_ctx(prob0['N'], prob0['S'])
    e_auto = m0.compute_log_likelihood(zs=prob0['zs'])
    ctx0.set_option("g2_passes", 3)
    e_3 = m0.compute_log_likelihood(zs=prob0['zs'])
    assert abs(e_auto - e_3) <= 2e-6 * abs(e_3)


@pytest.mark.parametrize("path", [0, 1])
@pytest.mark.parametrize("target_cond", [1e6, 1e8])
def test_ill_conditioned_inducing_set(target_cond, path):
    """No lengthscale shrinking: the lengthscale is GROWN until cond(Kuu + jitter I) reaches the target (up to ~1e8
    with the default jitter).  The per-row kernels are fp32 / 3xTF32: their error scales as eps_fp32 * cond.  Outcome stated:
    at cond 1e6 the ELBO bar of 1e-4 holds on both paths; at cond 1e8 fp32 cannot resolve sigma^2 - |b|^2 any more, the call must
    still return a finite ELBO within 5e-2 (the float64 reference handles this regime; a float32 hot path does not -- DESIGN.md)."""
    from workloads import _gram
    prob = make_problem(seed=3200, dims=[8, 8, 1], N=500, M=100, S=5, num_data=4000, inner_q_scale=0.3, max_cond=None)
    for lay in prob['layers']:
        ls = float(np.mean(lay['ls']))
        for _ in range(60):
            c = np.linalg.cond(_gram('rbf', lay['Z'], ls, lay['var']) + prob['jitter'] * np.eye(prob['M']))
            if c >= target_cond:
                break
            ls *= 1.1
        lay['ls'] = ls
        K = _gram('rbf', lay['Z'], ls, lay['var'])
        Lu = np.linalg.cholesky(K + prob['jitter'] * np.eye(prob['M']))
        if not lay['last']:
            lay['q_sqrt'] = np.tile((0.3 * Lu)[None], (lay['dout'], 1, 1))
    m = build_model(prob)
    m._ensure_ctx(prob['N'], prob['S']).set_option("path", path)
    e = m.compute_log_likelihood(zs=prob['zs'])
    e_ref = build_oracle(prob).compute_log_likelihood(zs=prob['zs'])
    rel = abs(e - e_ref) / abs(e_ref)
    record("ill_conditioned", path=path, cond=target_cond, elbo_rel=rel)
    assert np.isfinite(e)
    assert rel <= (1e-4 if target_cond <= 1e6 else 5e-2), (e, e_ref, rel)


@pytest.mark.parametrize("white", [False, True])
def test_gradients_northstar_shape_both_whitenings(white):
    """Full-size (N=1000, M=100, S=20, L=5) gradient parity on the benchmarked path for white=False AND white=True, trained-regime
    q_sqrt: every tensor within 5e-3 of its max |g| (kernel hyper-parameters: 2e-2 -- cancellation-heavy sums over all
    (row, inducing point) pairs accumulated from TF32 products)."""
    prob = round_f32(_trained(make_problem(seed=3300, white=white, num_data=8192, **NORTH), 2))
    m = build_model(prob)
    e, grads, glik = m.compute_log_likelihood_and_grad(zs=prob['zs'])
    o = build_oracle(prob)
    e_ref, g_ref = o.elbo_and_grad(zs=prob['zs'])
    assert abs(e - e_ref) <= 1e-4 * abs(e_ref)
    i, worst = 0, {}
    for l, g in enumerate(grads):
        Z, q_mu, q_sqrt, var, ls = [x.numpy() for x in g_ref[i:i + 5]]
        i += 5
        for name, got, ref in (("Z", g['Z'], Z), ("q_mu", g['q_mu'], q_mu), ("q_sqrt", g['q_sqrt'], np.tril(q_sqrt)),
                               ("variance", g['variance'], var), ("lengthscales", g['lengthscales'], ls)):
            sc = np.max(np.abs(ref)) + 1e-12
            err = float(np.max(np.abs(np.asarray(got) - ref)) / sc)
            worst[name] = max(worst.get(name, 0.0), err)
            assert err <= (2e-2 if name in ("variance", "lengthscales") else 5e-3), (name, l, err)
    record("grad_northstar", white=float(white), **worst)
    assert_allclose(glik, g_ref[i].numpy(), rtol=5e-3)
