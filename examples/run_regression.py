#!/usr/bin/env python
"""The regression workflow of the reference's demos/run_regression.py on this engine: same model construction, same
initialisation tweaks, Adam training on minibatches, batched S=100 test-set predictions, test RMSE and log-likelihood.

    python examples/run_regression.py [--layers 2] [--iterations 2000] [--n 8192] [--d 8]

Differences from the reference script: the data are synthetic (kin8nm shape; no dataset download here), the inducing inputs
are a random subset of X instead of scipy k-means centres, and GPflow's objects are this package's descriptors:

    from gpflow.likelihoods import Gaussian      ->  from doubly_stochastic_dgp.likelihoods import Gaussian
    from gpflow.kernels import RBF               ->  from doubly_stochastic_dgp.kernels import RBF
    from gpflow.training import AdamOptimizer    ->  from doubly_stochastic_dgp.training import AdamOptimizer
    from doubly_stochastic_dgp.dgp import DGP        (unchanged)
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "doubly-stochastic-dgp_b200"))

from doubly_stochastic_dgp.dgp import DGP  # noqa: E402
from doubly_stochastic_dgp.kernels import RBF  # noqa: E402
from doubly_stochastic_dgp.likelihoods import Gaussian  # noqa: E402
from doubly_stochastic_dgp.training import AdamOptimizer  # noqa: E402


def make_data(n, d, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, d))
    f = np.sin(X[:, :1] * 2.0) * np.cos(X[:, 1:2]) + 0.3 * X[:, 2:3] ** 2
    Y = f + 0.1 * rng.normal(size=(n, 1))
    Y = (Y - Y.mean()) / Y.std()
    nt = n // 10
    return X[nt:], Y[nt:], X[:nt], Y[:nt]


def assess(model, Xs, Ys, S=100, batch=1000):
    """demos/run_regression.py:108-128: S=100 samples, batches of 1000 test rows."""
    means, dens = [], []
    for i in range(0, len(Xs), batch):
        m, v = model.predict_y(Xs[i:i + batch], S)                         # (S, n, 1) each
        means.append(m.mean(0))
        dens.append(model.predict_density(Xs[i:i + batch], Ys[i:i + batch], S))
    mean = np.concatenate(means, 0)
    return float(np.sqrt(np.mean((Ys - mean) ** 2))), float(np.mean(np.concatenate(dens, 0)))


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=2)
    ap.add_argument("--iterations", type=int, default=2000)
    ap.add_argument("--log-every", type=int, default=500)
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--d", type=int, default=8)
    ap.add_argument("--inducing", type=int, default=100)
    ap.add_argument("--minibatch", type=int, default=1000)
    ap.add_argument("--test-samples", type=int, default=100)
    a = ap.parse_args(argv)

    X, Y, Xs, Ys = make_data(a.n, a.d)
    rng = np.random.default_rng(1)
    Z = X[rng.choice(len(X), a.inducing, replace=False)]
    kernels = [RBF(a.d, lengthscales=float(np.sqrt(a.d))) for _ in range(a.layers)]
    mb = a.minibatch if len(X) > a.minibatch else None
    model = DGP(X, Y, Z, kernels, Gaussian(), num_samples=1, minibatch_size=mb)
    for layer in model.layers[:-1]:                     # start the inner layers almost deterministically
        layer.q_sqrt = layer.q_sqrt.value * 1e-5
    model.likelihood.variance = 0.05

    opt = AdamOptimizer(0.01)
    done = 0
    out = None
    while done < a.iterations:
        k = min(a.log_every, a.iterations - done)
        elbo = opt.minimize(model, maxiter=k)
        done += k
        out = assess(model, Xs, Ys, S=a.test_samples)
        print(f"iteration {done:6d}  minibatch ELBO {elbo:12.3f}  test RMSE {out[0]:.4f}  test log-lik {out[1]:.4f}", flush=True)
    return out


if __name__ == "__main__":
    main()
