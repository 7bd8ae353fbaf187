"""Host logic of the product package on CPU (no GPU, no compute in libdsdgp.so): `_lib.Context` is monkeypatched with
tests/fake_ctx.FakeContext, whose "device" is the oracle.  What is checked is the Python side of the boundary: parameter
upload / read-back and dirty flags, set_trainable plumbing, NatGrad var_list -> layer ids, the training.* optimiser
shapes, DGP_Quad's nodes and weights, full_cov / predict shapes, stand-alone layer services -- i.e. that dgp.py, layers.py
and training.py call the C-ABI the way the reference's API implies.  Numerics of the real kernels: tests/test_gpu_*.py."""
import numpy as np
import pytest
import torch
from numpy.testing import assert_allclose

from tests.fake_ctx import FakeContext
from tests.synth import build_oracle, make_problem, round_f32
from tests.test_natgrad_cpu import well_conditioned_q


@pytest.fixture(autouse=True)
def fake_device(monkeypatch):
    from doubly_stochastic_dgp import _lib
    monkeypatch.setattr(_lib, "Context", FakeContext)


def _model(prob, **kw):
    from tests.gpu_common import build_model
    return build_model(prob, **kw)


def test_parameters_reach_the_device_and_come_back():
    prob = round_f32(make_problem(seed=1, dims=[3, 3, 1], N=20, M=6, S=2, inner_q_scale=0.3))
    m = _model(prob)
    e = m.compute_log_likelihood(zs=prob['zs'])
    assert abs(e - build_oracle(prob).compute_log_likelihood(zs=prob['zs'])) <= 1e-9 * abs(e)
    # assignment after the context exists marks the host copy dirty and is re-uploaded
    m.layers[0].q_mu = np.zeros_like(prob['layers'][0]['q_mu'])
    prob['layers'][0]['q_mu'] = np.zeros_like(prob['layers'][0]['q_mu'])
    e2 = m.compute_log_likelihood(zs=prob['zs'])
    assert abs(e2 - build_oracle(prob).compute_log_likelihood(zs=prob['zs'])) <= 1e-9 * abs(e2)
    assert e2 != e
    # device-side updates are read back lazily
    m.adam_init(0.01)
    z0 = m.layers[0].feature.Z.value
    m.train_step(zs=prob['zs'])
    assert np.abs(m.layers[0].feature.Z.value - z0).max() > 0
    # the training entry points refuse to outgrow the workspaces once training has started (Adam state lives on the device)
    with pytest.raises(RuntimeError):
        m.compute_log_likelihood(X=np.zeros((10000, 3)), Y=np.zeros((10000, 1)))


def test_set_trainable_is_forwarded_and_respected():
    from doubly_stochastic_dgp import _lib
    prob = round_f32(make_problem(seed=2, dims=[3, 3, 1], N=20, M=6, S=2, inner_q_scale=0.3))
    m = _model(prob)
    m.layers[-1].q_mu.set_trainable(False)
    m.layers[-1].q_sqrt.set_trainable(False)
    m.adam_init(0.01)
    ctx = m._ctx
    assert ctx.trainable[(1, _lib.F_Q_MU)] is False and ctx.trainable[(1, _lib.F_Q_SQRT)] is False
    assert ctx.trainable[(0, _lib.F_Q_MU)] is True and ctx.trainable[(-1, _lib.F_LIK_VARIANCE)] is True
    q0, z0 = m.layers[-1].q_mu.value, m.layers[-1].feature.Z.value
    m.train_step(zs=prob['zs'])
    assert_allclose(m.layers[-1].q_mu.value, q0, rtol=0, atol=0)
    assert np.abs(m.layers[-1].feature.Z.value - z0).max() > 0
    m.layers[-1].q_mu.set_trainable(True)
    m.train_step(zs=prob['zs'])
    assert ctx.trainable[(1, _lib.F_Q_MU)] is True
    assert np.abs(m.layers[-1].q_mu.value - q0).max() > 0


def test_natgrad_optimizer_api_and_var_list_mapping():
    from doubly_stochastic_dgp.training import AdamOptimizer, Loop, NatGradOptimizer
    from oracle import reference_dgp as R
    prob = round_f32(well_conditioned_q(make_problem(seed=3, dims=[3, 3, 1], N=30, M=6, S=2, inner_q_scale=0.3)))
    m = _model(prob)
    p = [[m.layers[-1].q_mu, m.layers[-1].q_sqrt]]
    e0 = m.natgrad_step(var_list=p, gamma=1.0, zs=prob['zs'])
    o = build_oracle(prob)
    e0_ref = R.natgrad_step(o, [1], 1.0, zs=prob['zs'])
    assert abs(e0 - e0_ref) <= 1e-9 * abs(e0_ref)
    assert m._ctx.calls[-1] == ("natgrad", (1,), 1.0)
    assert_allclose(m.layers[-1].q_mu.value, o.layers[-1].q_mu.numpy(), rtol=1e-6, atol=1e-9)
    assert_allclose(m.layers[-1].q_sqrt.value, o.layers[-1].q_sqrt.numpy(), rtol=1e-6, atol=1e-9)
    # default var_list = last layer; both layers by explicit list; foreign parameters are rejected
    m.natgrad_step(gamma=0.01, zs=prob['zs'])
    assert m._ctx.calls[-1][1] == (1,)
    both = [[l.q_mu, l.q_sqrt] for l in m.layers]
    NatGradOptimizer(gamma=0.001).minimize(m, var_list=both, maxiter=2)
    assert m._ctx.calls[-1] == ("natgrad", (0, 1), 0.001)
    with pytest.raises(ValueError):
        m.natgrad_step(var_list=[[m.layers[0].q_mu, m.layers[1].q_sqrt]])
    # the notebook's loop (demos/using_natural_gradients.ipynb)
    for v in p[0]:
        v.set_trainable(False)
    ng = NatGradOptimizer(gamma=1.).make_optimize_action(m, var_list=p)
    ad = AdamOptimizer(0.001).make_optimize_action(m)
    out = Loop([ng, ad], stop=3)()
    assert np.isfinite(out)
    assert AdamOptimizer(0.001).minimize(m, maxiter=2) is not None


def test_dgp_quad_nodes_weights_and_elbo():
    from doubly_stochastic_dgp.dgp import DGP_Quad
    from tests.test_gpu_quad import _build
    prob = round_f32(well_conditioned_q(make_problem(seed=1507, dims=[2, 2, 1], N=12, M=5, S=1, inner_q_scale=0.3,
                                                     num_data=12)))
    m, o = _build(prob, 5)
    assert isinstance(m, DGP_Quad) and m.num_samples == 25 and m.D_quad == 2
    assert_allclose(m.gh_w, o.gh_w.numpy())
    assert_allclose(m.gh_x[0], o.gh_x[0].numpy())
    e, e_ref = m.compute_log_likelihood(), o.compute_log_likelihood()
    assert abs(e - e_ref) <= 1e-6 * abs(e_ref), (e, e_ref)          # nodes travel as float32
    assert_allclose(m._ctx.weights, m.gh_w)
    e2, grads, _ = m.compute_log_likelihood_and_grad()
    _, g_ref = o.elbo_and_grad()
    assert_allclose(grads[0]['q_mu'], g_ref[1].numpy(), rtol=1e-4, atol=1e-7)
    m.natgrad_step(gamma=1.0)
    assert m.compute_log_likelihood() > e


def test_prediction_and_full_cov_shapes():
    prob = round_f32(make_problem(seed=5, dims=[3, 2, 2], N=15, M=6, S=3, inner_q_scale=0.3))
    m = _model(prob)
    o = build_oracle(prob)
    Xs = prob['X'][:7]
    mean, var = m.predict_f(Xs, 4)
    assert mean.shape == var.shape == (4, 7, 2)
    fm, fv = m.predict_f_full_cov(Xs, 4)
    assert fm.shape == (4, 7, 2) and fv.shape == (4, 7, 7, 2)
    Fs, Fm, Fv = m.predict_all_layers_full_cov(Xs, 2)
    assert [v.shape for v in Fv] == [(2, 7, 7, 2), (2, 7, 7, 2)]
    Fs, Fm, Fv = m.predict_all_layers(Xs, 2)
    assert [v.shape for v in Fv] == [(2, 7, 2), (2, 7, 2)]
    zs = [z[:, :7] for z in prob['zs']]
    ym, yv = m.predict_y(Xs, 3, zs=zs)
    om, ov = o.predict_y(Xs, 3, zs=zs)
    assert_allclose(ym, om.numpy(), rtol=1e-5, atol=1e-6)
    d = m.predict_density(Xs, prob['Y'][:7], 3, zs=zs)
    assert d.shape == (7, 2)
    assert_allclose(d, o.predict_density(Xs, prob['Y'][:7], 3, zs=zs).numpy(), rtol=1e-4, atol=1e-5)


def test_standalone_layer_services():
    from doubly_stochastic_dgp.kernels import RBF
    from doubly_stochastic_dgp.layers import SVGP_Layer
    from doubly_stochastic_dgp.mean_functions import Zero
    from oracle import reference_dgp as R
    rng = np.random.default_rng(3)
    N, M, Din, D, S = 12, 5, 2, 2, 3
    Z = np.float32(rng.normal(size=(M, Din))).astype(np.float64)
    X = np.float32(rng.normal(size=(S, N, Din))).astype(np.float64)
    z = np.float32(rng.normal(size=(S, N, D))).astype(np.float64)
    lay = SVGP_Layer(RBF(Din, lengthscales=1.5), Z, D, Zero())
    R.settings.jitter = 1e-6
    olay = R.SVGP_Layer(R.RBF(Din, lengthscales=1.5), Z, D, R.Zero())
    olay.q_sqrt = torch.as_tensor(np.float32(lay.q_sqrt.value).astype(np.float64))
    mean, var = lay.conditional_ND(X[0])
    assert mean.shape == var.shape == (N, D)
    mean, var = lay.conditional_ND(X[0], full_cov=True)
    assert var.shape == (N, N, D)
    ms, vs = lay.conditional_SND(X, full_cov=True)
    assert ms.shape == (S, N, D) and vs.shape == (S, N, N, D)
    f, fm, fv = lay.sample_from_conditional(X, z, full_cov=True)
    of, _, ofv = olay.sample_from_conditional(torch.as_tensor(X), z=torch.as_tensor(z), full_cov=True)
    assert_allclose(f, of.numpy(), atol=1e-5)
    assert_allclose(fv, ofv.numpy(), atol=1e-5)
    f2, fm2, fv2 = lay.sample_from_conditional(X, z)
    of2, _, ofv2 = olay.sample_from_conditional(torch.as_tensor(X), z=torch.as_tensor(z))
    assert_allclose(f2, of2.numpy(), atol=1e-5)
    assert abs(lay.KL() - float(olay.KL())) < 1e-8
    f3, _, _ = lay.sample_from_conditional(X)           # z=None draws on the host and still goes through the device
    assert f3.shape == (S, N, D)


def test_predictions_larger_than_the_workspaces_after_training_are_chunked():
    """demos/run_regression.py:108-113 predicts with S=100 in batches of 1000 rows after training with num_samples=1: once
    Adam state lives in the context it cannot be re-created, so the host evaluates the call in (row, sample) chunks."""
    prob = round_f32(make_problem(seed=6, dims=[3, 3, 2], N=20, M=6, S=2, inner_q_scale=0.3))
    m = _model(prob)
    m.adam_init(0.01)
    m.train_step(zs=prob['zs'])
    ctx = m._ctx
    ctx.N_max, ctx.S_max = 16, 4                    # pretend small workspaces
    rng = np.random.default_rng(0)
    N, S = 37, 10
    Xs = np.float32(rng.normal(size=(N, 3))).astype(np.float64)
    Ys = np.float32(rng.normal(size=(N, 2))).astype(np.float64)
    zs = [np.float32(rng.normal(size=(S, N, 3))).astype(np.float64), np.float32(rng.normal(size=(S, N, 2))).astype(np.float64)]
    assert m._plan(N, S) == ([(0, 16), (16, 32), (32, 37)], [(0, 4), (4, 8), (8, 10)])
    # reference values: the same (device-side) parameters in an unchunked oracle model
    o = ctx._model(S)
    oFs, oFm, oFv = o.propagate(Xs, S=S, zs=zs)
    Fs, Fm, Fv = m.propagate(Xs, S=S, zs=zs)
    assert m._ctx is ctx                             # not re-created
    for l in range(2):
        assert Fs[l].shape == (S, N, prob['dims'][l + 1])
        assert_allclose(Fm[l], oFm[l].numpy(), rtol=1e-5, atol=1e-6)
        assert_allclose(Fs[l], oFs[l].numpy(), rtol=1e-5, atol=1e-6)
    ym, yv = m.predict_y(Xs, S, zs=zs)
    om, ov = o.predict_y(Xs, S, zs=zs)
    assert_allclose(ym, om.numpy(), rtol=1e-5, atol=1e-6)
    assert_allclose(yv, ov.numpy(), rtol=1e-5, atol=1e-6)
    d = m.predict_density(Xs, Ys, S, zs=zs)
    assert_allclose(d, o.predict_density(Xs, Ys, S, zs=zs).numpy(), rtol=1e-4, atol=1e-5)
    # full_cov: samples may be chunked, rows may not
    fm, fv = m.predict_f_full_cov(Xs[:12], S)
    assert fm.shape == (S, 12, 2) and fv.shape == (S, 12, 12, 2)
    with pytest.raises(RuntimeError):
        m.predict_f_full_cov(Xs, S)
