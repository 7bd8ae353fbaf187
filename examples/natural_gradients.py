#!/usr/bin/env python
"""The workflow of the reference's demos/using_natural_gradients.ipynb on this engine: a 2-layer DGP on a 1-D toy function,
trained (a) with Adam only and (b) with a natural-gradient step (gamma = 1) on the final layer's q(u) alternating with Adam
on everything else, then posterior percentiles from joint (full-covariance) samples.

    python examples/natural_gradients.py [--iterations 2000]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "doubly-stochastic-dgp_b200"))

from doubly_stochastic_dgp.dgp import DGP  # noqa: E402
from doubly_stochastic_dgp.kernels import RBF  # noqa: E402
from doubly_stochastic_dgp.likelihoods import Gaussian as Gaussian_lik  # noqa: E402
from doubly_stochastic_dgp.training import AdamOptimizer, Loop, NatGradOptimizer  # noqa: E402


def f(X):
    return -(np.sin(40 * (X - 0.85) ** 4) * np.cos(2.5 * (X - 0.95)) + (X - 0.9) / 2 + 1) / 2


def make_dgp2(X, Y):
    kernels = [RBF(1, lengthscales=0.1), RBF(1, lengthscales=0.1)]
    model = DGP(X, Y, X, kernels, Gaussian_lik(), num_samples=10)
    model.likelihood.likelihood.variance = 1e-4
    for layer in model.layers[:-1]:
        layer.q_sqrt = layer.q_sqrt.value * 1e-5
    return model


def percentiles(model, Xs, S=100):
    Fs, ms, vs = model.predict_all_layers_full_cov(Xs, S)
    return [np.percentile(Fs[-1][:, :, -1], q, axis=0) for q in (10., 50., 90.)]


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--iterations", type=int, default=2000)
    ap.add_argument("--grid", type=int, default=300)
    ap.add_argument("--samples", type=int, default=100)
    a = ap.parse_args(argv)

    X = np.linspace(0, 1, 30).reshape(-1, 1)
    Y = f(X)
    model_adam = make_dgp2(X, Y)
    model_nat_grads = make_dgp2(X, Y)

    AdamOptimizer(0.001).minimize(model_adam, maxiter=a.iterations)

    ng_vars = [[model_nat_grads.layers[-1].q_mu, model_nat_grads.layers[-1].q_sqrt]]
    for v in ng_vars[0]:
        v.set_trainable(False)
    ng_action = NatGradOptimizer(gamma=1.).make_optimize_action(model_nat_grads, var_list=ng_vars)
    adam_action = AdamOptimizer(0.001).make_optimize_action(model_nat_grads)
    Loop([ng_action, adam_action], stop=a.iterations)()

    Xs = np.linspace(-0.1, 1.1, a.grid).reshape(-1, 1)
    res = {}
    for name, model in (("adam", model_adam), ("nat grads with adam", model_nat_grads)):
        lo, med, hi = percentiles(model, Xs, a.samples)
        elbo = np.mean([model.compute_log_likelihood() for _ in range(5)])
        fit = float(np.sqrt(np.mean((np.interp(X[:, 0], Xs[:, 0], med) - Y[:, 0]) ** 2)))
        print(f"{name:22s} ELBO {elbo:10.3f}   RMSE of the median at the data {fit:.4f}   "
              f"mean 10-90% band {float(np.mean(hi - lo)):.4f}")
        res[name] = (elbo, fit)
    return res


if __name__ == "__main__":
    main()
