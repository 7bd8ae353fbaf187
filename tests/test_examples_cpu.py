"""The example scripts (the reference's demo workflows on this engine) run end to end against the oracle-backed fake
context: checks that they only use API that exists, with shapes that fit -- not numerics (those are the GPU tests')."""
import importlib.util
import os

import numpy as np
import pytest

from tests.fake_ctx import FakeContext

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(autouse=True)
def fake_device(monkeypatch):
    from doubly_stochastic_dgp import _lib
    monkeypatch.setattr(_lib, "Context", FakeContext)


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "examples", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_run_regression_example():
    rmse, ll = _load("run_regression").main(["--layers", "2", "--iterations", "4", "--log-every", "2", "--n", "300", "--d", "3",
                                             "--inducing", "8", "--minibatch", "64", "--test-samples", "12"])
    assert np.isfinite(rmse) and np.isfinite(ll)


def test_natural_gradients_example():
    res = _load("natural_gradients").main(["--iterations", "2", "--grid", "25", "--samples", "6"])
    assert set(res) == {"adam", "nat grads with adam"} and all(np.isfinite(v[0]) for v in res.values())
