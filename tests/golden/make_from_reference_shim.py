#!/usr/bin/env python
"""Golden vectors from the reference's OWN SOURCE, executed unmodified in this container.

    python tests/golden/make_from_reference_shim.py          # writes tests/golden/refshim_*.npz (needs /root/reference)

/root/reference/doubly_stochastic_dgp/{dgp,layers,utils,layer_initializations}.py are imported as they are; `tensorflow` and
`gpflow` (neither installable here) are replaced by the float64 torch stand-in oracle/tf_gpflow_shim.py, which implements the ~35
TF ops and the GPflow classes those files touch.  So every reference-owned line on the hot path runs as written; GPflow's own
numerics are the shim's (see its docstring) -- the oracle therefore stays "parity unpinned" at the GPflow/TF boundary, but a
transcription error in oracle/reference_dgp.py's restatement of the reference's files would show up here.
tests/test_refshim_cpu.py compares the oracle with every refshim_*.npz (and, when /root/reference is present, regenerates one
case live and checks that the committed file is what this script produces)."""
import contextlib
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = os.environ.get("DSDGP_REFERENCE", "/root/reference")

# name -> spec.  dims = [D_X, inner..., D_Y]; kernels are built per layer input dimension.
CASES = {
    "svgp_rbf": dict(dims=[3, 1], N=30, M=8, S=1, kern="RBF", white=False, lik="gaussian"),
    "dgp2_rbf": dict(dims=[3, 3, 1], N=40, M=10, S=3, kern="RBF", white=False, lik="gaussian", full_cov=True),
    "dgp3_matern_white": dict(dims=[4, 4, 4, 2], N=50, M=12, S=2, kern="Matern52", white=True, lik="gaussian"),
    "dgp3_linear_means_ard": dict(dims=[5, 3, 4, 2], N=36, M=9, S=2, kern="RBF", ard=True, white=False, lik="gaussian",
                                  full_cov=True),
    "dgp2_sum_white": dict(dims=[3, 3, 1], N=32, M=8, S=2, kern="RBF", white=False, lik="gaussian", white_var=0.02),
    "dgp2_input_prop": dict(dims=[2, 3, 1], N=28, M=7, S=2, kern="RBF", white=True, lik="gaussian", input_prop=True),
    "dgp2_bernoulli_white": dict(dims=[2, 2, 3], N=19, M=19, S=2, kern="Matern52", white=True, lik="bernoulli"),
    "quad_dgp2": dict(dims=[2, 1, 1], N=15, M=6, S=1, kern="RBF", white=True, lik="gaussian", quad_H=7),
}


def load_reference():
    """install the stand-ins, import the reference package from where it lies, return the namespace used below"""
    sys.path.insert(0, ROOT)
    from oracle import tf_gpflow_shim as shim
    shim.install()
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    import warnings
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from doubly_stochastic_dgp import dgp as ref_dgp
        from doubly_stochastic_dgp import layer_initializations as ref_init
    return shim, ref_dgp, ref_init


def build_reference_model(spec, shim, ref_dgp, ref_init, seed):
    """the reference's constructors on the case's data; returns (model, data dict)"""
    from gpflow import kernels as gk
    from gpflow import likelihoods as gl
    rng = np.random.RandomState(seed)
    dims, N, M, S = spec["dims"], spec["N"], spec["M"], spec["S"]
    X = rng.randn(N, dims[0])
    if spec["lik"] == "bernoulli":
        Y = rng.choice([-1., 1.], N * dims[-1]).reshape(N, dims[-1])          # tests/test_dgp.py:51
    else:
        Y = np.tile(np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.randn(N, 1), (1, dims[-1]))
    Z = X[:M] + (0.3 * rng.randn(M, dims[0]) if M < N else 0.0)
    D = dims[0]
    kdims = [D] + [D + d for d in dims[1:-1]] if spec.get("input_prop") else dims[:-1]

    def kern(d):
        ls = (np.sqrt(d) * (1.0 + 0.1 * np.arange(d))) if spec.get("ard") else float(np.sqrt(d))
        k = getattr(gk, spec["kern"])(d, lengthscales=ls, variance=0.5, ARD=bool(spec.get("ard")))
        if spec.get("white_var"):
            k = k + gk.White(d, variance=spec["white_var"])
        return k
    kernels = [kern(d) for d in kdims]
    if spec["lik"] == "bernoulli":
        lik = gl.Bernoulli()
    else:
        lik = gl.Gaussian()
        lik.variance = 0.05
    shim.settings.jitter = 1e-6
    with contextlib.redirect_stdout(io.StringIO()):          # init_layers_linear prints the layer dimensions
        if spec.get("input_prop"):
            np.random.seed(seed)                                 # init_layers_input_prop pads Z with np.random.randn
            layers = ref_init.init_layers_input_prop(X, Y, Z, kernels, num_outputs=dims[-1], white=spec["white"])
            m = ref_dgp.DGP_Base(X, Y, lik, layers, num_samples=S, num_data=3 * N)
        elif spec.get("quad_H"):
            layers = ref_init.init_layers_linear(X, Y, Z, kernels, num_outputs=dims[-1], white=spec["white"])
            m = ref_dgp.DGP_Quad(X, Y, lik, layers, H=spec["quad_H"], num_samples=S, num_data=3 * N)
        else:
            m = ref_dgp.DGP(X, Y, Z, kernels, lik, num_outputs=dims[-1], white=spec["white"], num_samples=S, num_data=3 * N)
    for layer in m.layers:
        Dl = layer.q_mu.shape[1]
        layer.q_mu = 0.3 * rng.randn(M, Dl)
        layer.q_sqrt = np.tril(0.1 * rng.randn(Dl, M, M) / np.sqrt(M)) + 0.3 * np.eye(M)[None]
    return m, dict(X=X, Y=Y, Z=Z, rng=rng)


def describe(m, spec, data):
    """portable description of the model the reference built (what tests/test_refshim_cpu.py rebuilds the oracle from)"""
    out = dict(in_X=data["X"], in_Y=data["Y"], in_Z=data["Z"])
    meta = dict(spec=spec, jitter=1e-6, num_data=int(m.num_data), S=int(m.num_samples), layers=[])
    for l, layer in enumerate(m.layers):
        k = layer.kern
        parts = getattr(k, "kern_list", [k])
        lm = dict(num_outputs=int(layer.num_outputs), white=bool(layer.white), input_prop_dim=int(layer.input_prop_dim or 0),
                  kern=type(parts[0]).__name__, input_dim=int(k.input_dim), ard=bool(parts[0].ARD),
                  mean=type(layer.mean_function).__name__, has_white=len(parts) > 1)
        out[f"in_Z{l}"] = layer.feature.Z.read_value()
        out[f"in_q_mu{l}"] = layer.q_mu.read_value()
        out[f"in_q_sqrt{l}"] = layer.q_sqrt.read_value()
        out[f"in_variance{l}"] = parts[0].variance.read_value()
        out[f"in_lengthscales{l}"] = np.atleast_1d(parts[0].lengthscales.read_value())
        if len(parts) > 1:
            out[f"in_white_variance{l}"] = parts[1].variance.read_value()
        if lm["mean"] == "Linear":
            out[f"in_A{l}"] = layer.mean_function.A.read_value()
            out[f"in_b{l}"] = layer.mean_function.b.read_value()
        meta["layers"].append(lm)
    if spec["lik"] == "gaussian":
        out["in_lik_variance"] = m.likelihood.likelihood.variance.read_value()
    out["meta"] = np.array(json.dumps(meta))
    return out


def evaluate(m, spec, data, shim):
    """everything the reference computes on the path, with the draws it used"""
    rng = data["rng"]
    X, N, S = data["X"], spec["N"], spec["S"]
    L = len(m.layers)
    out = {}
    zs = [rng.randn(S, N, int(layer.num_outputs)) for layer in m.layers]
    Fs, Fm, Fv = shim.run_as_tensors(m, lambda: m.propagate(shim._t(X), S=S, zs=[shim._t(z) for z in zs]))
    for l in range(L):
        out[f"in_z{l}"] = zs[l]
        out[f"out_F{l}"], out[f"out_Fmean{l}"], out[f"out_Fvar{l}"] = Fs[l], Fm[l], Fv[l]
        out[f"out_KL{l}"] = float(shim.run_as_tensors(m, m.layers[l].KL))

    def with_draws(tag, fn):
        shim.reset_draws(seed=11)
        res = fn()
        for i, z in enumerate(shim.DRAWS):
            out[f"in_draw_{tag}{i}"] = z.numpy().copy()
        return res
    out["out_elbo"] = with_draws("elbo", m.compute_log_likelihood)
    Ns = 11
    Xs = rng.randn(Ns, spec["dims"][0])
    Ys = (rng.choice([-1., 1.], Ns * spec["dims"][-1]).reshape(Ns, -1) if spec["lik"] == "bernoulli"
          else rng.randn(Ns, spec["dims"][-1]))
    out["in_Xs"], out["in_Ys"] = Xs, Ys
    Sp = 3
    out["out_predict_f_mean"], out["out_predict_f_var"] = with_draws("pf", lambda: m.predict_f(Xs, Sp))
    out["out_predict_y_mean"], out["out_predict_y_var"] = with_draws("py", lambda: m.predict_y(Xs, Sp))
    out["out_predict_density"] = with_draws("pd", lambda: m.predict_density(Xs, Ys, Sp))
    if spec.get("full_cov"):
        Fs, Fm, Fv = with_draws("fc", lambda: m.predict_all_layers_full_cov(Xs[:6], 2))
        for l in range(L):
            out[f"out_fc_F{l}"], out[f"out_fc_Fmean{l}"], out[f"out_fc_Fvar{l}"] = Fs[l], Fm[l], Fv[l]
    return out


def generate(name, shim=None, ref=None):
    if shim is None:
        shim, ref_dgp, ref_init = load_reference()
    else:
        ref_dgp, ref_init = ref
    spec = CASES[name]
    seed = 1 + sorted(CASES).index(name)
    m, data = build_reference_model(spec, shim, ref_dgp, ref_init, seed)
    out = describe(m, spec, data)
    out.update(evaluate(m, spec, data, shim))
    return out


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default=None, help="one case only")
    ap.add_argument("--out", default=None, help="with --case: write here instead of tests/golden/")
    a = ap.parse_args(argv)
    shim, ref_dgp, ref_init = load_reference()
    for name in ([a.case] if a.case else CASES):
        out = generate(name, shim, (ref_dgp, ref_init))
        path = a.out if (a.case and a.out) else os.path.join(HERE, f"refshim_{name}.npz")
        np.savez_compressed(path, **out)
        print("wrote %s  elbo=%.10g" % (os.path.basename(path), out["out_elbo"]))


if __name__ == "__main__":
    sys.exit(main())
