"""Golden vectors of the later paths (tests/golden/ext/*.npz, made by tests/golden/make_golden_ext.py from the float64
oracle): NatGrad step, full_cov propagate, prediction epilogues.  CPU: the oracle still reproduces them; GPU: the CUDA
path matches them without the oracle in the loop, at the tolerances of the corresponding parity tests.
(File name: collected last on purpose -- the GPU halves were written after the round's last hardware run; the same problems,
paths and tolerances passed against the live oracle in tests/test_gpu_natgrad.py / _full_cov.py / _predict.py.)"""
import glob
import os

import numpy as np
import pytest
from numpy.testing import assert_allclose

from tests.golden.make_golden import unpack_problem
from tests.golden.make_golden_ext import CASES

EXT = os.path.join(os.path.dirname(__file__), "golden", "ext")


def _load(name):
    g = np.load(os.path.join(EXT, name + ".npz"), allow_pickle=False)
    prob = unpack_problem(g)
    prob['n_classes'] = int(g['n_classes'])
    return g, prob


def test_ext_golden_files_exist():
    assert sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(EXT, "*.npz"))) == sorted(CASES)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_ext_golden(name):
    g, _ = _load(name)
    fn, args = CASES[name]
    _, out = fn(*args)
    for k, v in out.items():
        assert_allclose(v, g["out_" + k], rtol=1e-8, atol=1e-10, err_msg=k)


def _model(prob, path):
    from tests.gpu_common import build_model
    m = build_model(prob)
    m._ensure_ctx(prob['N'], prob['S']).set_option("path", path)
    return m


@pytest.mark.gpu
@pytest.mark.parametrize("path", [0, 1])
@pytest.mark.parametrize("name", ["natgrad_gamma1_last_layer", "natgrad_small_gamma_all_layers"])
def test_cuda_natgrad_matches_golden(name, path):
    g, prob = _load(name)
    ids = [int(i) for i in g["out_ids"]]
    m = _model(prob, path)
    var_list = [[m.layers[l].q_mu, m.layers[l].q_sqrt] for l in ids]
    e0 = m.natgrad_step(var_list=var_list, gamma=float(g["out_gamma"]), zs=prob['zs'], X=prob['X'], Y=prob['Y'])
    assert abs(e0 - float(g["out_elbo_before"])) <= 1e-4 * abs(float(g["out_elbo_before"]))      # no per-path exception
    e1 = m.compute_log_likelihood(zs=prob['zs'], X=prob['X'], Y=prob['Y'])
    assert abs(e1 - float(g["out_elbo_after"])) <= 1e-4 * abs(float(g["out_elbo_after"]))
    tol = 1e-3 if path == 0 else 5e-3
    for l in ids:
        for k, got in (("q_mu", m.layers[l].q_mu.value), ("q_sqrt", m.layers[l].q_sqrt.value)):
            ref = g[f"out_{k}{l}"]
            assert_allclose(got, ref, atol=tol * np.abs(ref).max(), rtol=0, err_msg=f"{k} l={l}")


@pytest.mark.gpu
def test_cuda_full_cov_matches_golden():
    g, prob = _load("full_cov_dgp2")
    m = _model(prob, 1)
    Fs, Fm, Fv = m.propagate(prob['X'], full_cov=True, S=prob['S'], zs=prob['zs'])
    for l in range(len(Fs)):
        sc = max(1.0, float(np.abs(g[f"out_Fmean{l}"]).max()))
        tol = (5e-6 if l == 0 else 2e-4) * sc
        assert_allclose(Fm[l], g[f"out_Fmean{l}"], atol=tol, rtol=0)
        assert_allclose(Fv[l], g[f"out_Fvar{l}"], atol=tol, rtol=0)
        assert_allclose(Fs[l], g[f"out_F{l}"], atol=4 * tol, rtol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("path", [0, 1])
def test_cuda_predict_matches_golden(path):
    g, prob = _load("predict_gauss_dgp2")
    m = _model(prob, path)
    ym, yv = m.predict_y(prob['X'], prob['S'], zs=prob['zs'])
    tol = 5e-4 * (1.0 if path == 0 else 6.0) * max(1.0, float(np.abs(g["out_y_mean"]).max()))
    assert_allclose(ym, g["out_y_mean"], atol=tol, rtol=0)
    assert_allclose(yv, g["out_y_var"], atol=tol, rtol=0)
    dens = m.predict_density(prob['X'], prob['Y'], prob['S'], zs=prob['zs'])
    # d log N / d mu = (y - mu) / v, bounded with the golden marginals (cf. tests/test_gpu_predict.py)
    sens = float((np.abs(prob['Y'][None] - g["out_y_mean"]) / g["out_y_var"]).max()) + 0.5 / prob['lik_var']
    assert_allclose(dens, g["out_density"], atol=tol * sens, rtol=0)
