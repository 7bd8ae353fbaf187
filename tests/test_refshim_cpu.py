"""The oracle against the reference's OWN SOURCE (as far as this container allows).

tests/golden/refshim_*.npz were produced by tests/golden/make_from_reference_shim.py: the unmodified files
/root/reference/doubly_stochastic_dgp/{dgp,layers,utils,layer_initializations}.py executed over a float64 torch stand-in for the
TensorFlow ops / GPflow classes they touch (oracle/tf_gpflow_shim.py).  Each file holds the model the reference's constructors
built (init_layers_linear / init_layers_input_prop / DGP / DGP_Base / DGP_Quad), the draws the reference consumed, and what
it computed: per-layer propagate outputs, per-layer KL, the ELBO, predict_f / predict_y / predict_density, full-covariance
propagation.  The oracle (oracle/reference_dgp.py) must reproduce all of it to rounding.  GPflow's own pieces are the shim's
restatement, so this pins the oracle's transcription of the reference-owned code, not GPflow ("parity unpinned" stays)."""
import glob
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
from numpy.testing import assert_allclose

from oracle import reference_dgp as R

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "refshim_*.npz")))
RTOL, ATOL = 1e-9, 1e-10


def _kernel(g, l, lm):
    cls = {"RBF": R.RBF, "Matern52": R.Matern52}[lm["kern"]]
    ls = g[f"in_lengthscales{l}"]
    k = cls(lm["input_dim"], variance=float(g[f"in_variance{l}"]), lengthscales=ls if lm["ard"] else float(ls[0]), ARD=lm["ard"])
    if lm["has_white"]:
        k = R.Sum([k, R.White(lm["input_dim"], variance=float(g[f"in_white_variance{l}"]))])
    return k


def oracle_from_fixture(g):
    meta = json.loads(str(g["meta"]))
    R.settings.jitter = meta["jitter"]
    layers = []
    for l, lm in enumerate(meta["layers"]):
        mf = {"Zero": lambda: R.Zero(), "Identity": lambda: R.Identity(),
              "Linear": lambda: R.Linear(g[f"in_A{l}"], g[f"in_b{l}"])}[lm["mean"]]()
        layer = R.SVGP_Layer(_kernel(g, l, lm), g[f"in_Z{l}"], lm["num_outputs"], mf, white=lm["white"],
                             input_prop_dim=lm["input_prop_dim"] or None)
        layer.q_mu = torch.as_tensor(g[f"in_q_mu{l}"]).clone()
        layer.q_sqrt = torch.as_tensor(g[f"in_q_sqrt{l}"]).clone()
        layers.append(layer)
    spec = meta["spec"]
    lik = R.Bernoulli() if spec["lik"] == "bernoulli" else R.Gaussian(float(g["in_lik_variance"]))
    kw = dict(num_samples=meta["S"], num_data=meta["num_data"])
    if spec.get("quad_H"):
        return R.DGP_Quad(g["in_X"], g["in_Y"], lik, layers, H=spec["quad_H"], **kw), meta
    return R.DGP_Base(g["in_X"], g["in_Y"], lik, layers, **kw), meta


def _draws(g, tag, L):
    return [torch.as_tensor(g[f"in_draw_{tag}{i}"]) for i in range(L)] if f"in_draw_{tag}0" in g.files else None


def test_fixtures_exist():
    assert len(FILES) >= 8


@pytest.fixture(autouse=True)
def _restore_jitter():
    old = R.settings.jitter
    yield
    R.settings.jitter = old


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[8:-4] for f in FILES])
def test_oracle_reproduces_the_reference_source(path):
    g = np.load(path, allow_pickle=False)
    o, meta = oracle_from_fixture(g)
    L, S = len(o.layers), meta["S"]
    # propagate with the reference's zs (dgp.py:61-76 over layers.py:81-119,178-219) and KL (layers.py:221-246)
    with torch.no_grad():
        Fs, Fm, Fv = o.propagate(g["in_X"], S=S, zs=[g[f"in_z{l}"] for l in range(L)])
    for l in range(L):
        assert_allclose(Fm[l].numpy(), g[f"out_Fmean{l}"], rtol=RTOL, atol=ATOL, err_msg=f"Fmean {l}")
        assert_allclose(Fv[l].numpy(), g[f"out_Fvar{l}"], rtol=RTOL, atol=ATOL, err_msg=f"Fvar {l}")
        assert_allclose(Fs[l].numpy(), g[f"out_F{l}"], rtol=RTOL, atol=ATOL, err_msg=f"F {l}")
        assert_allclose(float(o.layers[l].KL()), float(g[f"out_KL{l}"]), rtol=RTOL, err_msg=f"KL {l}")
    # ELBO (dgp.py:83-98; DGP_Quad dgp.py:129-166) with the draws the reference made
    assert_allclose(o.compute_log_likelihood(zs=_draws(g, "elbo", L)), float(g["out_elbo"]), rtol=RTOL)
    # prediction stack (dgp.py:100-126, utils.py:53-121)
    Xs, Ys, Sp = g["in_Xs"], g["in_Ys"], 3
    m_, v_ = o.predict_f(Xs, Sp, zs=_draws(g, "pf", L))
    assert_allclose(m_.numpy(), g["out_predict_f_mean"], rtol=RTOL, atol=ATOL)
    assert_allclose(v_.numpy(), g["out_predict_f_var"], rtol=RTOL, atol=ATOL)
    m_, v_ = o.predict_y(Xs, Sp, zs=_draws(g, "py", L))
    assert_allclose(np.asarray(m_), g["out_predict_y_mean"], rtol=RTOL, atol=ATOL)
    assert_allclose(np.asarray(v_), g["out_predict_y_var"], rtol=RTOL, atol=ATOL)
    d_ = o.predict_density(Xs, Ys, Sp, zs=_draws(g, "pd", L))
    assert_allclose(np.asarray(d_), g["out_predict_density"], rtol=1e-8, atol=1e-9)
    # full covariance (layers.py:66-69,206-217; utils.py:43-51)
    if "out_fc_F0" in g.files:
        with torch.no_grad():
            Fs, Fm, Fv = o.propagate(Xs[:6], full_cov=True, S=2, zs=_draws(g, "fc", L))
        for l in range(L):
            assert_allclose(Fm[l].numpy(), g[f"out_fc_Fmean{l}"], rtol=RTOL, atol=ATOL, err_msg=f"full_cov Fmean {l}")
            assert_allclose(Fv[l].numpy(), g[f"out_fc_Fvar{l}"], rtol=RTOL, atol=ATOL, err_msg=f"full_cov Fvar {l}")
            assert_allclose(Fs[l].numpy(), g[f"out_fc_F{l}"], rtol=1e-8, atol=1e-9, err_msg=f"full_cov F {l}")


def test_oracle_constructor_matches_the_reference_constructor():
    """init_layers_linear (layer_initializations.py:16-52) incl. the PCA / padding mean functions: the oracle's DGP(...) builds
    the layers the reference built (fixture of the 5-3-4-2 case: one step down, one step up)."""
    g = np.load(os.path.join(HERE, "golden", "refshim_dgp3_linear_means_ard.npz"), allow_pickle=False)
    meta = json.loads(str(g["meta"]))
    R.settings.jitter = meta["jitter"]
    kernels = [_kernel(g, l, lm) for l, lm in enumerate(meta["layers"])]
    m = R.DGP(g["in_X"], g["in_Y"], g["in_Z"], kernels, R.Gaussian(0.05), num_outputs=meta["spec"]["dims"][-1], white=False)
    for l, lm in enumerate(meta["layers"]):
        assert type(m.layers[l].mean_function).__name__ == lm["mean"]
        assert_allclose(m.layers[l].Z.numpy(), g[f"in_Z{l}"], rtol=1e-12, atol=1e-13)
        if lm["mean"] == "Linear":
            assert_allclose(m.layers[l].mean_function.A.numpy(), g[f"in_A{l}"], rtol=1e-12, atol=1e-13)


@pytest.mark.skipif(not os.path.isdir("/root/reference/doubly_stochastic_dgp"), reason="the reference checkout is not here")
def test_committed_fixture_is_what_the_generator_produces(tmp_path):
    """re-runs the reference source for one case in a clean interpreter (the reference package and this repo's host package
    share the name doubly_stochastic_dgp) and compares with the committed file"""
    out = str(tmp_path / "regen.npz")
    subprocess.run([sys.executable, os.path.join(HERE, "golden", "make_from_reference_shim.py"), "--case", "dgp2_rbf", "--out", out],
                   check=True, cwd=str(tmp_path), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL,
                   env={k: v for k, v in os.environ.items() if k != "PYTHONPATH"})
    a = np.load(out, allow_pickle=False)
    b = np.load(os.path.join(HERE, "golden", "refshim_dgp2_rbf.npz"), allow_pickle=False)
    assert sorted(a.files) == sorted(b.files)
    for k in a.files:
        if k != "meta":
            assert_allclose(a[k], b[k], rtol=1e-12, atol=1e-13, err_msg=k)


def _host_kernel(g, l, lm):
    from doubly_stochastic_dgp import kernels as K
    ls = g[f"in_lengthscales{l}"]
    k = getattr(K, lm["kern"])(lm["input_dim"], variance=float(g[f"in_variance{l}"]),
                               lengthscales=ls if lm["ard"] else float(ls[0]), ARD=lm["ard"])
    if lm["has_white"]:
        k = k + K.White(lm["input_dim"], variance=float(g[f"in_white_variance{l}"]))
    return k


def test_host_package_constructors_build_what_the_reference_built():
    """The product's host mirror (doubly_stochastic_dgp.DGP / init_layers_linear / init_layers_input_prop: construction-time
    NumPy, no device) against the layers the reference's constructors produced: inducing inputs, mean-function matrices,
    q_sqrt initialisation, input_prop_dim."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "doubly-stochastic-dgp_b200"))
    from doubly_stochastic_dgp.dgp import DGP
    from doubly_stochastic_dgp.layer_initializations import init_layers_input_prop
    from doubly_stochastic_dgp.likelihoods import Gaussian
    # init_layers_linear through DGP(...): 5 -> 3 (PCA) -> 4 (padding) -> 2
    g = np.load(os.path.join(HERE, "golden", "refshim_dgp3_linear_means_ard.npz"), allow_pickle=False)
    meta = json.loads(str(g["meta"]))
    kernels = [_host_kernel(g, l, lm) for l, lm in enumerate(meta["layers"])]
    m = DGP(g["in_X"], g["in_Y"], g["in_Z"], kernels, Gaussian(), num_outputs=meta["spec"]["dims"][-1], white=False)
    for l, lm in enumerate(meta["layers"]):
        assert type(m.layers[l].mean_function).__name__ == lm["mean"]
        assert m.layers[l].num_outputs == lm["num_outputs"]
        assert_allclose(np.asarray(m.layers[l].feature.Z.value), g[f"in_Z{l}"], rtol=1e-12, atol=1e-13)
        if lm["mean"] == "Linear":
            assert_allclose(np.asarray(m.layers[l].mean_function.A.value), g[f"in_A{l}"], rtol=1e-12, atol=1e-13)
    # q_sqrt = chol(Kuu + jitter I) of a non-white layer (layers.py:160-163): same prior as the reference's kernel gives
    q0 = np.asarray(m.layers[0].q_sqrt.value)[0]
    from oracle import reference_dgp as R
    R.settings.jitter = meta["jitter"]
    K = _kernel(g, 0, meta["layers"][0]).K(torch.as_tensor(g["in_Z0"])).numpy() + meta["jitter"] * np.eye(g["in_Z0"].shape[0])
    assert_allclose(q0 @ q0.T, K, rtol=1e-9, atol=1e-11)
    # init_layers_input_prop: Z padded with np.random.randn columns (same global seed as the generator), input_prop_dim = D
    g = np.load(os.path.join(HERE, "golden", "refshim_dgp2_input_prop.npz"), allow_pickle=False)
    meta = json.loads(str(g["meta"]))
    kernels = [_host_kernel(g, l, lm) for l, lm in enumerate(meta["layers"])]
    seed = 1 + sorted(f[8:-4] for f in map(os.path.basename, FILES)).index("dgp2_input_prop")
    state = np.random.get_state()
    try:
        np.random.seed(seed)
        layers = init_layers_input_prop(g["in_X"], g["in_Y"], g["in_Z"], kernels, num_outputs=1, white=True)
    finally:
        np.random.set_state(state)
    for l, lm in enumerate(meta["layers"]):
        assert (layers[l].input_prop_dim or 0) == lm["input_prop_dim"]
        assert_allclose(np.asarray(layers[l].feature.Z.value), g[f"in_Z{l}"], rtol=1e-12, atol=1e-13)
