"""GPU parity of the prediction stack (dgp.py:100-126): predict_f / predict_y / predict_density through the public host
API, likelihood epilogues on the device (dsdgp_predict_y, dsdgp_predict_density), against the oracle with injected z.

Two kinds of check per case:
  * epilogue-only: the oracle's likelihood applied to the DEVICE's own (Fmean, Fvar) -- isolates k_predict_y_* /
    k_density_* (fp64 arithmetic on fp32 inputs, fp32 outputs): 2e-5 absolute / relative.
  * end-to-end vs the oracle's float64 propagate: the forward-pass tolerance of tests/test_gpu_parity.py carried through
    the epilogue (log-densities have sensitivity ~|y-mu|/v to the mean, so they get an absolute tolerance)."""
import math

import numpy as np
import pytest
import torch
from numpy.testing import assert_allclose

from tests.synth import build_oracle, make_problem, round_f32

pytestmark = pytest.mark.gpu


def _model(prob, path=1):
    from tests.gpu_common import build_model
    m = build_model(prob)
    m._ensure_ctx(prob['N'], prob['S']).set_option("path", path)
    return m


CASES = {
    "gauss_dgp2": dict(dims=[8, 8, 2], N=150, M=40, S=5, inner_q_scale=0.3),
    "gauss_svgp": dict(dims=[3, 2], N=33, M=9, S=3),                                   # L = 1: de-duplicated rows
    "multiclass": dict(dims=[6, 4, 5], N=60, M=15, S=3, n_classes=5, inner_q_scale=0.3),
    "multiclass_svgp": dict(dims=[4, 3], N=21, M=7, S=2, n_classes=3),
}


@pytest.mark.parametrize("path", [0, 1])
@pytest.mark.parametrize("name", list(CASES))
def test_predict_y_and_density(name, path):
    prob = round_f32(make_problem(seed=800 + list(CASES).index(name), **CASES[name]))
    S, X, Y, zs = prob['S'], prob['X'], prob['Y'], prob['zs']
    m = _model(prob, path)
    o = build_oracle(prob)
    Fm, Fv = m._build_predict(X, S=S, zs=zs)
    ym, yv = m.predict_y(X, S, zs=zs)
    dens = m.predict_density(X, Y, S, zs=zs)
    K = prob['n_classes']
    assert ym.shape == yv.shape == Fm.shape and dens.shape == (prob['N'], 1 if K else prob['dims'][-1])
    # ---- epilogue only: oracle likelihood on the device's marginals
    tFm, tFv = torch.as_tensor(Fm), torch.as_tensor(Fv)
    em, ev = o.likelihood.predict_mean_and_var(tFm, tFv)
    assert_allclose(ym, em.numpy(), rtol=2e-5, atol=2e-6)
    assert_allclose(yv, ev.numpy(), rtol=2e-5, atol=2e-6)
    l = o.likelihood.predict_density(tFm, tFv, torch.as_tensor(Y))
    ed = torch.logsumexp(l - math.log(S), 0).numpy()
    assert_allclose(dens, ed, rtol=2e-5, atol=2e-5)
    # ---- end to end vs the float64 oracle
    om, ov = o.predict_y(X, S, zs=zs)
    tol = 5e-4 * (1.0 if path == 0 else 6.0) * max(1.0, float(np.abs(om.numpy()).max()))
    assert_allclose(ym, om.numpy(), atol=tol, rtol=0)
    assert_allclose(yv, ov.numpy(), atol=tol, rtol=0)
    od = o.predict_density(X, Y, S, zs=zs).numpy()
    if K:
        assert_allclose(dens, od, atol=5e-3, rtol=0)
    else:
        # d log N / d mu = (y - mu)/v: bound it with the oracle's own marginals
        oFm, oFv = o.predict_f(X, S, zs=zs)
        sens = float((np.abs(Y[None] - oFm.numpy()) / (oFv.numpy() + prob['lik_var'])).max()) + 0.5 / prob['lik_var']
        assert_allclose(dens, od, atol=tol * sens, rtol=0)


def test_philox_predictions_are_reproducible_and_finite():
    """zs=None: in-kernel Philox draws; same model seed sequence => same predictions; S-average is sane."""
    prob = round_f32(make_problem(seed=820, dims=[5, 5, 1], N=64, M=16, S=8, inner_q_scale=0.3))
    a, b = _model(prob), _model(prob)
    ya, va = a.predict_y(prob['X'], 8)
    yb, vb = b.predict_y(prob['X'], 8)
    assert_allclose(ya, yb, rtol=0, atol=0)
    assert np.all(np.isfinite(ya)) and np.all(va > 0)
    d = a.predict_density(prob['X'], prob['Y'], 8)
    assert d.shape == (64, 1) and np.all(np.isfinite(d))
