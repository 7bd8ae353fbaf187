"""Golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the float64 oracle):
CPU: the oracle still reproduces them (guards the checker); GPU: the CUDA path matches them without the oracle."""
import glob
import os

import numpy as np
import pytest
from numpy.testing import assert_allclose

from tests.golden.make_golden import oracle_outputs, unpack_problem

FILES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def test_golden_files_exist():
    assert len(FILES) >= 5


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_oracle_reproduces_golden(path):
    g = np.load(path, allow_pickle=False)
    out = oracle_outputs(unpack_problem(g))
    for k, v in out.items():
        assert_allclose(v, g["out_" + k], rtol=1e-9, atol=1e-11, err_msg=k)


@pytest.mark.gpu
@pytest.mark.parametrize("tc", [0, 1])
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_cuda_matches_golden(path, tc):
    from tests.gpu_common import build_model
    g = np.load(path, allow_pickle=False)
    prob = unpack_problem(g)
    m = build_model(prob)
    m._ensure_ctx(prob['N'], prob['S']).set_option("path", tc)
    Fs, Fm, Fv = m.propagate(prob['X'], S=prob['S'], zs=prob['zs'])
    tol = 5e-4 if tc == 0 else 3e-3
    for l in range(len(Fs)):
        sc = max(1.0, float(np.abs(g[f"out_Fmean{l}"]).max()))
        assert_allclose(Fm[l], g[f"out_Fmean{l}"], atol=tol * sc, rtol=0)
        assert_allclose(Fv[l], g[f"out_Fvar{l}"], atol=tol * sc, rtol=0)
        assert_allclose(Fs[l], g[f"out_F{l}"], atol=2 * tol * sc, rtol=0)
    e, grads, glik = m.compute_log_likelihood_and_grad(zs=prob['zs'])
    assert abs(e - float(g["out_elbo"])) <= 1e-4 * abs(float(g["out_elbo"]))
    gt = 2e-3 if tc == 0 else 1e-2
    for l, gr in enumerate(grads):
        for k in ("Z", "q_mu", "q_sqrt", "variance", "lengthscales"):
            ref = g[f"out_g{l}_{k}"]
            t_ = 5e-2 if (tc == 1 and k in ("variance", "lengthscales")) else gt
            assert_allclose(gr[k], ref, atol=t_ * (np.abs(ref).max() + 1e-12), rtol=0, err_msg=f"{k} l={l}")
