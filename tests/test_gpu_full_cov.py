"""GPU parity of the full_cov=True path (dsdgp_propagate_full_cov, csrc/full_cov.cu: float64 on the device, float32 at
the boundary) against the oracle's restatement of layers.py:52-74,178-219 and utils.py:43-51, through the public API
(predict_all_layers_full_cov / predict_f_full_cov, dgp.py:104-114; layer.conditional_ND(full_cov=True)).

Tolerance: the device arithmetic is float64 like the reference's; what remains is the float32 rounding of the inputs
(X, z, parameters -- removed by round_f32) and of the returned arrays: 5e-6 of the layer's scale, and 2e-4 on deeper
layers where float32-rounded parameters held on the device (q_sqrt etc. are stored in fp32) compound."""
import numpy as np
import pytest
from numpy.testing import assert_allclose

from tests.synth import build_oracle, make_problem, round_f32

pytestmark = pytest.mark.gpu

CASES = [
    dict(dims=[3, 1], N=17, M=6, S=2),
    dict(dims=[3, 3, 2], N=40, M=12, S=3),
    dict(dims=[4, 2, 3, 1], N=33, M=20, S=2, ard=True, kern='matern52'),
    dict(dims=[1, 1, 1], N=300, M=30, S=10),           # the plotting demos' shape (demos/using_natural_gradients.ipynb)
]


@pytest.mark.parametrize("white", [False, True])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_full_cov_propagate_matches_oracle(case, white):
    from tests.gpu_common import build_model, record, rel_err
    prob = round_f32(make_problem(seed=900 + case, white=white, inner_q_scale=0.3, **CASES[case]))
    m = build_model(prob)
    S, X = prob['S'], prob['X']
    # injected z through the C-ABI (propagate(zs=...)); the public predict_* entry points draw with Philox (below)
    Fs, Fm, Fv = m.propagate(X, full_cov=True, S=S, zs=prob['zs'])
    o = build_oracle(prob)
    oFs, oFm, oFv = o.propagate(X, full_cov=True, S=S, zs=prob['zs'])
    errs = {}
    for l in range(len(Fs)):
        N, D = prob['N'], prob['dims'][l + 1]
        assert Fv[l].shape == (S, N, N, D) and Fm[l].shape == (S, N, D) and Fs[l].shape == (S, N, D)
        sc = max(1.0, float(np.abs(oFm[l].numpy()).max()))
        tol = (5e-6 if l == 0 else 2e-4) * sc
        errs[f"mean{l}"] = rel_err(Fm[l], oFm[l].numpy()); errs[f"var{l}"] = rel_err(Fv[l], oFv[l].numpy())
        assert_allclose(Fm[l], oFm[l].numpy(), atol=tol, rtol=0, err_msg=f"mean l={l}")
        assert_allclose(Fv[l], oFv[l].numpy(), atol=tol, rtol=0, err_msg=f"var l={l}")
        assert_allclose(Fs[l], oFs[l].numpy(), atol=4 * tol, rtol=0, err_msg=f"F l={l}")
    record("full_cov", case=case, white=float(white), **errs)


def test_full_cov_public_entry_points_and_diag_consistency():
    """predict_f_full_cov / predict_all_layers_full_cov (dgp.py:104-114) with Philox draws: shapes, symmetry, positive
    diagonal; for a single layer the diagonal of the full covariance equals the full_cov=False variance of the fp32 path."""
    from tests.gpu_common import build_model
    prob = round_f32(make_problem(seed=950, dims=[2, 2, 1], N=50, M=10, S=4, inner_q_scale=0.3))
    m = build_model(prob)
    mean, var = m.predict_f_full_cov(prob['X'], 4)
    assert mean.shape == (4, 50, 1) and var.shape == (4, 50, 50, 1)
    assert_allclose(var, np.swapaxes(var, 1, 2), atol=1e-6)
    assert np.all(np.einsum('snnd->snd', var) > 0)
    Fs, Fm, Fv = m.predict_all_layers_full_cov(prob['X'], 4)
    assert [f.shape for f in Fv] == [(4, 50, 50, 2), (4, 50, 50, 1)]
    assert np.all(np.isfinite(Fs[-1]))
    # layer 1 does not depend on the draws: full-cov diagonal == diagonal path
    dFs, dFm, dFv = m.predict_all_layers(prob['X'], 4)
    # (fp64 full_cov pipeline vs fp32/TF32 row path on a cond ~ 3e4 problem: 1e-4 of the layer's scale)
    assert_allclose(np.einsum('snnd->snd', Fv[0]), dFv[0], atol=1e-4 * max(1.0, np.abs(dFv[0]).max()), rtol=0)
    assert_allclose(Fm[0], dFm[0], atol=1e-4 * max(1.0, np.abs(dFm[0]).max()), rtol=0)


def test_standalone_layer_conditional_and_sample_full_cov():
    """layer.conditional_ND(X, full_cov=True) -> (N,D), (N,N,D); layer.sample_from_conditional(X, z, full_cov=True)
    (layers.py:76-119) on a layer used outside a model."""
    from doubly_stochastic_dgp import settings
    from doubly_stochastic_dgp.kernels import RBF
    from doubly_stochastic_dgp.layers import SVGP_Layer
    from doubly_stochastic_dgp.mean_functions import Zero
    from oracle import reference_dgp as R
    import torch
    rng = np.random.default_rng(3)
    N, M, Din, D, S = 12, 5, 2, 2, 3
    Z = np.float32(rng.normal(size=(M, Din))).astype(np.float64)
    X = np.float32(rng.normal(size=(S, N, Din))).astype(np.float64)
    z = np.float32(rng.normal(size=(S, N, D))).astype(np.float64)
    q_mu = np.float32(rng.normal(size=(M, D))).astype(np.float64)
    settings.jitter = 1e-6
    R.settings.jitter = 1e-6
    lay = SVGP_Layer(RBF(Din, lengthscales=1.5), Z, D, Zero())
    lay.q_mu = q_mu
    olay = R.SVGP_Layer(R.RBF(Din, lengthscales=1.5), Z, D, R.Zero())
    olay.q_mu = torch.as_tensor(q_mu)
    olay.q_sqrt = torch.as_tensor(np.float32(lay.q_sqrt.value).astype(np.float64))
    mean, var = lay.conditional_ND(X[0], full_cov=True)
    omean, ovar = olay.conditional_ND(torch.as_tensor(X[0]), full_cov=True)
    assert var.shape == (N, N, D)
    assert_allclose(mean, omean.numpy(), atol=1e-5)
    assert_allclose(var, ovar.numpy(), atol=1e-5)
    f, fm, fv = lay.sample_from_conditional(X, z, full_cov=True)
    of, ofm, ofv = olay.sample_from_conditional(torch.as_tensor(X), z=torch.as_tensor(z), full_cov=True)
    assert f.shape == (S, N, D) and fv.shape == (S, N, N, D)
    assert_allclose(f, of.numpy(), atol=1e-4)
    # and the diagonal flavour of the same call
    f2, fm2, fv2 = lay.sample_from_conditional(X, z)
    of2, ofm2, ofv2 = olay.sample_from_conditional(torch.as_tensor(X), z=torch.as_tensor(z))
    assert_allclose(f2, of2.numpy(), atol=2e-3)
    assert_allclose(fv2, ofv2.numpy(), atol=2e-3)
