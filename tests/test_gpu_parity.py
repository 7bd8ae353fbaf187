"""GPU parity: the CUDA path (through the C-ABI) vs the float64 oracle on the same seeded inputs.
Tolerances are stated per test; the bar from BASELINE.json is ELBO within 1e-4 relative."""
import numpy as np
import pytest
from numpy.testing import assert_allclose

from tests.synth import build_oracle, make_problem, round_f32

pytestmark = pytest.mark.gpu


def _model(prob, path=1):
    from tests.gpu_common import build_model
    m = build_model(prob)
    m._ensure_ctx(prob['N'], prob['S']).set_option("path", path)
    return m


# path 0 = fp32 SIMT row kernels everywhere; path 1 = tcgen05 kernels where supported (3xTF32 projections,
# 1xTF32 variance / gradient GEMMs).  Tolerances are stated per path.
PATHS = [0, 1]


SMALL = [
    dict(dims=[3, 1], N=17, M=6, S=1),
    dict(dims=[3, 3, 1], N=13, M=5, S=3),
    dict(dims=[3, 3, 3, 2], N=70, M=37, S=2, kern='matern52'),
    dict(dims=[4, 2, 3, 1], N=33, M=20, S=4, ard=True),
    dict(dims=[8, 8, 8, 1], N=150, M=100, S=5),
]


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("white", [False, True])
@pytest.mark.parametrize("case", range(len(SMALL)))
def test_propagate_matches_oracle(case, white, path):
    """Per-layer Fmean / Fvar / F with injected z (propagate(zs=...), dgp.py:62-70).  fp32 row kernels, errors
    compound through the layers at ~eps_fp32*cond(Kuu): tolerance 5e-4 of the layer's scale (1e-3 for the draw)."""
    prob = round_f32(make_problem(seed=200 + case, white=white, inner_q_scale=0.3, **SMALL[case]))
    m = _model(prob, path)
    tol = 1.0 if path == 0 else 6.0      # single-pass TF32 on the |L_d^T u|^2 GEMM: ~2^-11 per product
    Fs, Fm, Fv = m.propagate(prob['X'], S=prob['S'], zs=prob['zs'])
    o = build_oracle(prob)
    oFs, oFm, oFv = o.propagate(prob['X'], S=prob['S'], zs=prob['zs'])
    for l in range(len(Fs)):
        sc = max(1.0, float(np.abs(oFm[l].numpy()).max()))
        assert_allclose(Fm[l], oFm[l].numpy(), atol=5e-4 * sc * tol, rtol=0, err_msg=f"Fmean l={l}")
        assert_allclose(Fv[l], oFv[l].numpy(), atol=5e-4 * sc * tol, rtol=0, err_msg=f"Fvar l={l}")
        assert_allclose(Fs[l], oFs[l].numpy(), atol=1e-3 * sc * tol, rtol=0, err_msg=f"F l={l}")


CONFIGS = {
    # BASELINE.json configs 1-3 (SURVEY 8(d) shapes)
    "cfg1": dict(dims=[8, 1], N=100, M=10, S=1),
    "cfg2": dict(dims=[8, 8, 1], N=1000, M=100, S=20),
    "cfg3": dict(dims=[8, 8, 8, 8, 8, 1], N=1000, M=100, S=20),
}


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("white", [False, True])
@pytest.mark.parametrize("name", list(CONFIGS))
def test_elbo_within_1e4(name, white, path):
    """ELBO vs the float64 oracle evaluated on the ORIGINAL float64 inputs: rel err <= 1e-4 (north_star)."""
    prob = make_problem(seed=1000 * (1 + list(CONFIGS).index(name)), white=white, num_data=8192, **CONFIGS[name])
    m = _model(prob, path)
    e = m.compute_log_likelihood(zs=prob['zs'])
    o = build_oracle(prob)
    e_ref = o.compute_log_likelihood(zs=prob['zs'])
    assert abs(e - e_ref) <= 1e-4 * abs(e_ref), (e, e_ref)


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("white", [False, True])
@pytest.mark.parametrize("case", range(len(SMALL)))
def test_gradients_match_autograd(case, white, path):
    """dELBO/dparam vs oracle autograd; tolerance 2e-3 (fp32 SIMT) / 1e-2 (TF32 tensor-core GEMMs) of each
    tensor's max |g| (5e-2 relative for the scalar kernel hyper-parameters on the TF32 path: they are
    cancellation-heavy sums over all (row, inducing point) pairs); ELBO 1e-5 / 1e-4."""
    prob = round_f32(make_problem(seed=300 + case, white=white, inner_q_scale=0.3, num_data=500, **SMALL[case]))
    m = _model(prob, path)
    gtol = 2e-3 if path == 0 else 1e-2
    e, grads, glik = m.compute_log_likelihood_and_grad(zs=prob['zs'])
    o = build_oracle(prob)
    e_ref, g_ref = o.elbo_and_grad(zs=prob['zs'])
    assert abs(e - e_ref) <= (1e-5 if path == 0 else 1e-4) * abs(e_ref)
    i = 0
    for l, g in enumerate(grads):
        Z, q_mu, q_sqrt, var, ls = [x.numpy() for x in g_ref[i:i + 5]]
        i += 5
        for name, got, ref in (("Z", g['Z'], Z), ("q_mu", g['q_mu'], q_mu), ("q_sqrt", g['q_sqrt'], np.tril(q_sqrt)),
                               ("variance", g['variance'], var), ("lengthscales", g['lengthscales'], ls)):
            sc = np.max(np.abs(ref)) + 1e-12
            tol_ = 5e-2 if (path == 1 and name in ("variance", "lengthscales")) else gtol
            assert_allclose(got, ref, atol=tol_ * sc, rtol=0, err_msg=f"{name} l={l}")
    assert_allclose(glik, g_ref[i].numpy(), rtol=gtol)


def test_gradients_northstar_shape():
    """config 3 shape (N=1000, M=100, S=20, L=5): gradient parity at full size, 5e-3 of max |g|."""
    prob = round_f32(make_problem(seed=3001, num_data=8192, inner_q_scale=0.1, **CONFIGS["cfg3"]))
    m = _model(prob)
    e, grads, glik = m.compute_log_likelihood_and_grad(zs=prob['zs'])
    o = build_oracle(prob)
    e_ref, g_ref = o.elbo_and_grad(zs=prob['zs'])
    assert abs(e - e_ref) <= 1e-4 * abs(e_ref)
    i = 0
    for l, g in enumerate(grads):
        Z, q_mu, q_sqrt, var, ls = [x.numpy() for x in g_ref[i:i + 5]]
        i += 5
        for name, got, ref in (("Z", g['Z'], Z), ("q_mu", g['q_mu'], q_mu), ("q_sqrt", g['q_sqrt'], np.tril(q_sqrt)),
                               ("variance", g['variance'], var), ("lengthscales", g['lengthscales'], ls)):
            sc = np.max(np.abs(ref)) + 1e-12
            assert_allclose(got, ref, atol=5e-3 * sc, rtol=0, err_msg=f"{name} l={l}")


def test_I1_dgp_equals_closed_form_svgp():
    """Reference tests/test_dgp.py:66-117 fixture (Z = X, M = 19, Matern52 l=0.5, jitter 1e-18) against the
    independent closed-form SVGP.  The reference tolerance 1e-7 is a float64 statement; fp32 rows give ~1e-4."""
    from oracle import closed_form as cf
    Ns, N, D_X, D_Y = 20, 19, 2, 3
    np.random.seed(0)
    X = np.random.uniform(size=(N, D_X)); Xs = np.random.uniform(size=(Ns, D_X))
    q_mu = np.random.randn(N, D_Y); q_sqrt = 0.001 * np.eye(N)[None] * np.ones((D_Y, 1, 1))
    Y = np.random.randn(N, D_Y)
    from doubly_stochastic_dgp import settings
    from doubly_stochastic_dgp.dgp import DGP
    from doubly_stochastic_dgp.kernels import Matern52
    from doubly_stochastic_dgp.likelihoods import Gaussian
    for white in (True, False):
        with settings.temp_settings(jitter_level=1e-10):
            m = DGP(X, Y, X, [Matern52(2, lengthscales=0.5)], Gaussian(0.01), white=white, num_samples=2)
            m.layers[-1].q_mu = q_mu
            m.layers[-1].q_sqrt = q_sqrt
            L_dgp = m.compute_log_likelihood()
            pm, pv = m.predict_f(Xs, 1)
            ym, yv = m.predict_y(Xs, 1)
        L_svgp = cf.svgp_elbo_gaussian('matern52', 1.0, 0.5, X, q_mu, q_sqrt, X, Y, 0.01, white, 1e-10)
        sm, sv = cf.svgp_predict_f('matern52', 1.0, 0.5, X, q_mu, q_sqrt, Xs, white, 1e-10)
        assert abs(L_dgp - L_svgp) <= 2e-4 * abs(L_svgp), (L_dgp, L_svgp)
        assert_allclose(pm[0], sm, atol=2e-3)
        assert_allclose(pv[0], sv, atol=2e-3)
        assert_allclose(yv[0], sv + 0.01, atol=2e-3)


def test_multiclass_elbo_and_grad():
    prob = round_f32(make_problem(seed=501, dims=[6, 4, 5], N=60, M=15, S=3, n_classes=5, inner_q_scale=0.3,
                                  num_data=300))
    m = _model(prob)
    e, grads, _ = m.compute_log_likelihood_and_grad(zs=prob['zs'])
    o = build_oracle(prob)
    e_ref, g_ref = o.elbo_and_grad(zs=prob['zs'])
    assert abs(e - e_ref) <= 1e-4 * abs(e_ref)
    i = 0
    for l, g in enumerate(grads):
        Z, q_mu, q_sqrt, var, ls = [x.numpy() for x in g_ref[i:i + 5]]
        i += 5
        for name, got, ref in (("Z", g['Z'], Z), ("q_mu", g['q_mu'], q_mu), ("q_sqrt", g['q_sqrt'], np.tril(q_sqrt))):
            sc = np.max(np.abs(ref)) + 1e-12
            assert_allclose(got, ref, atol=3e-3 * sc, rtol=0, err_msg=f"{name} l={l}")


def test_adam_steps_match_reference_optimiser():
    """5 Adam steps on GPflow's unconstrained variables (softplus / lower-tri) with injected z vs the oracle's
    AdamState (tf.train.AdamOptimizer semantics).  Parameters after the steps agree to 1e-4 of their scale."""
    from oracle import reference_dgp as R
    prob = round_f32(make_problem(seed=601, dims=[4, 4, 1], N=40, M=12, S=3, inner_q_scale=0.3, num_data=200))
    m = _model(prob, path=0)      # Adam is scale-free: element-wise parity needs the fp32-exact gradient path
    o = build_oracle(prob)
    st = R.AdamState(o, lr=0.01)
    m.adam_init(0.01)
    for it in range(5):
        e_ref = st.step(zs=prob['zs'])
        e = m.train_step(prob['X'], prob['Y'], zs=prob['zs'])
        assert abs(e - e_ref) <= 1e-4 * abs(e_ref), (it, e, e_ref)
    for l, (lay, olay) in enumerate(zip(m.layers, o.layers)):
        for name, got, ref in (("Z", lay.feature.Z.value, olay.Z.numpy()), ("q_mu", lay.q_mu.value, olay.q_mu.numpy()),
                               ("q_sqrt", lay.q_sqrt.value, np.tril(olay.q_sqrt.numpy())),
                               ("ls", lay.kern.lengthscales.value, olay.kern.lengthscales.numpy()),
                               ("var", lay.kern.variance.value, olay.kern.variance.numpy())):
            sc = np.max(np.abs(ref)) + 1e-12
            assert_allclose(got, ref, atol=1e-4 * sc + 2e-5, rtol=0, err_msg=f"{name} l={l}")
    assert_allclose(m.likelihood.likelihood.variance.value, o.likelihood.likelihood.variance.numpy(), rtol=1e-4)


def test_train_step_improves_elbo_and_keeps_constraints():
    prob = make_problem(seed=701, dims=[4, 4, 1], N=200, M=20, S=5, inner_q_scale=1e-3, num_data=200)
    m = _model(prob)
    e0 = np.mean([m.compute_log_likelihood() for _ in range(5)])
    m.adam_init(0.01)
    for _ in range(200):
        m.train_step()
    e1 = np.mean([m.compute_log_likelihood() for _ in range(5)])
    assert e1 > e0 + 1.0, (e0, e1)
    for l in m.layers:
        assert np.all(l.kern.lengthscales.value > 0) and l.kern.variance.value > 0
        q = l.q_sqrt.value
        assert np.all(np.triu(q, 1) == 0)
    assert m.likelihood.likelihood.variance.value > 0


def test_philox_draws_are_standard_normal_and_shard_invariant():
    prob = make_problem(seed=801, dims=[3, 3, 1], N=256, M=8, S=16, inner_q_scale=0.5)
    m = _model(prob)
    ctx = m._ensure_ctx(prob['N'], prob['S'])
    Fs, Fm, Fv = ctx.propagate(prob['X'], prob['S'], seed=1234)
    z = (Fs[0] - Fm[0]) / np.sqrt(Fv[0] + prob['jitter'])
    assert abs(z.mean()) < 0.03 and abs(z.std() - 1) < 0.03
    # same seed -> same draws; different seed -> different
    Fs2, _, _ = ctx.propagate(prob['X'], prob['S'], seed=1234)
    assert np.array_equal(Fs[0], Fs2[0])
    Fs3, _, _ = ctx.propagate(prob['X'], prob['S'], seed=1235)
    assert not np.array_equal(Fs[0], Fs3[0])
    # shard invariance: rows [64:128) evaluated alone with n_offset=64 reproduce the same samples
    ctx.set_option("n_offset", 64)
    Fs4, _, _ = ctx.propagate(prob['X'][64:128], prob['S'], seed=1234)
    ctx.set_option("n_offset", -1)
    assert_allclose(Fs4[0], Fs[0][:, 64:128], rtol=0, atol=1e-6)


def test_not_positive_definite_is_reported():
    from doubly_stochastic_dgp import _lib
    prob = make_problem(seed=901, dims=[2, 1], N=10, M=4, S=1, white=True)
    prob['layers'][0]['Z'][1] = prob['layers'][0]['Z'][0]       # duplicate inducing point
    prob['jitter'] = 0.0
    m = _model(prob)
    with pytest.raises(_lib.DsdgpError) as ei:
        m.compute_log_likelihood()
    assert ei.value.code == _lib.ERR_NOT_PD


def test_ragged_and_tiny_shapes():
    """N not a multiple of the row tile, M not a multiple of 4/16, D_out=1, N=1 (tests/test_dgp.py:176-183 uses N=1)."""
    for N, M, S, dims in [(1, 1, 1, [1, 2, 1]), (65, 3, 2, [2, 2, 1]), (129, 17, 3, [5, 5, 2])]:
        prob = round_f32(make_problem(seed=1000 + N, dims=dims, N=N, M=M, S=S, inner_q_scale=0.3))
        m = _model(prob)
        e = m.compute_log_likelihood(zs=prob['zs'])
        e_ref = build_oracle(prob).compute_log_likelihood(zs=prob['zs'])
        assert abs(e - e_ref) <= 1e-4 * abs(e_ref) + 1e-4, (N, M, e, e_ref)


@pytest.mark.parametrize("white", [False, True])
def test_tcgen05_forward_matches_simt_forward(white):
    """The tensor-core forward (3xTF32 projections + 1xTF32 variance GEMM) against the fp32 SIMT forward of the
    same library and the oracle, per layer.  TF32 single-pass on |c_d|^2 costs ~1e-4 relative on Fvar."""
    prob = round_f32(make_problem(seed=1234, dims=[8, 8, 8, 1], N=300, M=100, S=3, white=white, inner_q_scale=0.3))
    m = _model(prob)
    ctx = m._ensure_ctx(prob['N'], prob['S'])
    out = {}
    for path in (0, 1):
        ctx.set_option("path", path)
        out[path] = m.propagate(prob['X'], S=prob['S'], zs=prob['zs'])
    o = build_oracle(prob).propagate(prob['X'], S=prob['S'], zs=prob['zs'])
    for l in range(3):
        for which, name in ((1, "Fmean"), (2, "Fvar"), (0, "F")):
            ref = o[which][l].numpy()
            sc = max(1.0, float(np.abs(ref).max()))
            assert_allclose(out[1][which][l], out[0][which][l], atol=1e-3 * sc, rtol=0, err_msg=f"tc vs simt {name} l={l}")
            assert_allclose(out[1][which][l], ref, atol=1e-3 * sc, rtol=0, err_msg=f"tc vs oracle {name} l={l}")


def test_sample_sharding_is_exact():
    """S-sharded data parallelism (options s_world / s_offset): two shards of S samples each, drawn from the global
    Philox sample indices [0,S) and [S,2S), reproduce the single evaluation with 2S samples:
    e_full = e_0 + e_1 + KL  (each shard's ELBO carries the likelihood/(2S) part and, without a communicator, a full KL)."""
    prob = make_problem(seed=4321, dims=[4, 4, 1], N=96, M=24, S=3, inner_q_scale=0.3, num_data=960)
    m = _model(prob)
    ctx = m._ctx
    N, S = prob['N'], prob['S']
    e_full = ctx.elbo(prob['X'], prob['Y'], 2 * S, prob['num_data'], seed=99) if False else None
    m2 = _model(prob)
    c2 = m2._ensure_ctx(N, 2 * S)
    e_full = c2.elbo(prob['X'], prob['Y'], 2 * S, prob['num_data'], seed=99)
    kl = float(np.sum(c2.kl()))
    ctx.set_option("s_world", 2)
    parts = []
    for r in range(2):
        ctx.set_option("s_offset", r * S)
        parts.append(ctx.elbo(prob['X'], prob['Y'], S, prob['num_data'], seed=99))
    assert abs((parts[0] + parts[1] + kl) - e_full) <= 2e-6 * abs(e_full), (parts, kl, e_full)
