"""CPU-side checks of the boundary: libdsdgp.so loads, exports every symbol include/dsdgp.h declares, the ctypes
structs match the header, and the host-side mirror of the reference API behaves (no compute calls without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ensure_built():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "doubly-stochastic-dgp_b200"))
    import build
    return build.build()


def test_library_exports_every_declared_symbol():
    path = _ensure_built()
    lib = ctypes.CDLL(path)
    hdr = open(os.path.join(ROOT, "include", "dsdgp.h")).read()
    declared = sorted(set(re.findall(r"DSDGP_API[^;(]*?\b(dsdgp_\w+)\s*\(", hdr)))
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), name
    from doubly_stochastic_dgp import _lib
    assert sorted(_lib.SYMBOLS) == declared


def test_ctypes_struct_matches_header_layout():
    from doubly_stochastic_dgp import _lib
    assert ctypes.sizeof(_lib.LayerDesc) == 9 * 4
    # int L; LayerDesc[16]; int lik, K, D_y; (pad) double jitter; int N_max, S_max, device; (pad)
    assert _lib.Desc.jitter.offset % 8 == 0
    assert _lib.Desc.layers.offset == 4
    assert ctypes.sizeof(_lib.Desc) == 4 + 16 * 36 + 12 + 8 + 12 + 4


def test_create_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from doubly_stochastic_dgp import _lib
    _ensure_built()
    with pytest.raises(_lib.DsdgpError):
        _lib.Context([(4, 2, 1, 0, 0, 0, 0)], 0, 0, 1, 1e-6, 8, 1)


def test_invalid_descriptors_are_rejected_before_touching_the_device():
    from doubly_stochastic_dgp import _lib
    _ensure_built()
    for layers, kw in [([(4, 2, 3, 0, 0, 0, 1)], {}),                       # Identity mean with D_in != D_out
                       ([(4, 2, 2, 0, 0, 0, 0), (4, 3, 1, 0, 0, 0, 0)], {}),  # width mismatch
                       ([(4, 2, 1, 7, 0, 0, 0)], {})]:                        # unknown kernel
        with pytest.raises(_lib.DsdgpError) as ei:
            _lib.Context(layers, 0, 0, 1, 1e-6, 8, 1)
        assert ei.value.code in (-1, -5)


def test_host_api_mirrors_reference_constructor():
    """DGP(X, Y, Z, kernels, likelihood, ...) builds the same layer structure as layer_initializations.py:16-52."""
    from doubly_stochastic_dgp.dgp import DGP
    from doubly_stochastic_dgp.kernels import RBF
    from doubly_stochastic_dgp.likelihoods import Gaussian
    from doubly_stochastic_dgp.mean_functions import Identity, Linear, Zero
    rng = np.random.default_rng(0)
    X = rng.normal(size=(50, 5)); Y = rng.normal(size=(50, 2)); Z = rng.normal(size=(7, 5))
    m = DGP(X, Y, Z, [RBF(5), RBF(5), RBF(3), RBF(4)], Gaussian(), num_samples=3)
    assert [type(l.mean_function) for l in m.layers] == [Identity, Linear, Linear, Zero]
    assert [l.num_outputs for l in m.layers] == [5, 3, 4, 2]
    assert m.layers[2].feature.Z.shape == (7, 3) and m.layers[3].feature.Z.shape == (7, 4)
    W = m.layers[1].mean_function.A.value
    _, _, V = np.linalg.svd(X, full_matrices=False)
    np.testing.assert_allclose(W, V[:3].T)
    np.testing.assert_allclose(m.layers[2].mean_function.A.value, np.eye(3, 4))
    # non-white: q_sqrt initialised to chol(Kuu + jitter I) (layers.py:160-163); q_mu zeros
    q = m.layers[0].q_sqrt.value
    assert q.shape == (5, 7, 7) and np.allclose(q[0], q[4]) and np.all(np.triu(q[0], 1) == 0)
    K = np.exp(-0.5 * ((Z[:, None] - Z[None]) ** 2).sum(-1)) + 1e-6 * np.eye(7)
    np.testing.assert_allclose(q[0] @ q[0].T, K, atol=1e-10)
    # assignment semantics (tests/test_dgp.py:91-92, demos/run_regression.py:72-74)
    m.layers[-1].q_mu = np.ones((7, 2))
    assert np.all(m.layers[-1].q_mu.value == 1)
    m.layers[0].q_sqrt = m.layers[0].q_sqrt.value * 1e-5
    m.likelihood.likelihood.variance = 0.05
    assert float(m.likelihood.variance.value) == 0.05
    m.likelihood.variance = 0.07                      # the demo's spelling
    assert float(m.likelihood.likelihood.variance.value) == 0.07
    white = DGP(X, Y, Z, [RBF(5)], Gaussian(), white=True)
    np.testing.assert_allclose(white.layers[0].q_sqrt.value[0], np.eye(7))


def test_minibatch_is_aligned_and_covers_data():
    from doubly_stochastic_dgp.dgp import DGP
    from doubly_stochastic_dgp.kernels import RBF
    from doubly_stochastic_dgp.likelihoods import Gaussian
    X = np.arange(40, dtype=float).reshape(20, 2); Y = X[:, :1] * 10
    m = DGP(X, Y, X[:3], [RBF(2)], Gaussian(), minibatch_size=8)
    seen = []
    for _ in range(5):
        xb, yb = m._minibatch()
        assert xb.shape == (8, 2) and np.all(yb[:, 0] == xb[:, 0] * 10)
        seen.extend(xb[:, 0].tolist())
    assert set(seen) == set(X[:, 0].tolist())


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The boundary is a C ABI: include/dsdgp.h parses as C99 and a C program links against libdsdgp.so and gets the
    documented error behaviour (negative code + message, no exception) for an invalid descriptor."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    lib = _ensure_built()
    inc = os.path.join(ROOT, "include")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c", os.path.join(inc, "dsdgp.h")])
    src = tmp_path / "abi.c"
    src.write_text('#include "dsdgp.h"\n#include <stdio.h>\n#include <string.h>\n'
                   'int main(void) { dsdgp_ctx* c = 0; dsdgp_desc d; memset(&d, 0, sizeof d);\n'
                   '  int rc = dsdgp_create(&c, &d); printf("%s|%d|%s\\n", dsdgp_version(), rc, dsdgp_last_error());\n'
                   '  return (rc == DSDGP_ERR_INVALID && c == 0) ? 0 : 1; }\n')
    exe = tmp_path / "abi"
    libdir = os.path.dirname(lib)
    subprocess.check_call(["gcc", "-std=c99", "-I", inc, str(src), "-o", str(exe), "-L", libdir, "-ldsdgp",
                           "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    ver, rc, msg = out.stdout.strip().split("|")
    assert ver.startswith("dsdgp") and int(rc) == -1 and "out of range" in msg
