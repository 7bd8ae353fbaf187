#!/usr/bin/env python
"""Golden vectors from the REAL reference (UCL-SML/Doubly-Stochastic-DGP on GPflow 1.1.1 / TensorFlow 1.8).

Neither package can be installed in the build container (no network, Python 3.12), so this script has NOT been run here and
the oracle (oracle/reference_dgp.py) stays "parity unpinned" at the GPflow/TF boundary.  It is committed so that anyone with a
Python <= 3.6 environment holding `gpflow==1.1.1`, `tensorflow==1.8` and the reference checkout can produce fixtures the tests
pick up automatically:

    PYTHONPATH=/path/to/Doubly-Stochastic-DGP python tests/golden/make_from_reference.py      # writes tests/golden/ref_*.npz

tests/test_golden.py::test_reference_generated_fixtures then compares the oracle (CPU) and the CUDA path (GPU) with every
ref_*.npz it finds (and is skipped while there are none).

Only quantities the reference computes deterministically are stored: per-layer `propagate(X, zs=...)` outputs (dgp.py:61-76 takes
explicit zs; `_build_likelihood` does not), per-layer KL (layers.py:221-246), and the ELBO / predict_f / predict_y /
predict_density of single-layer models, which do not depend on the draws for a Gaussian likelihood."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    import gpflow
    import tensorflow as tf
    from gpflow.kernels import RBF, Matern52
    from gpflow.likelihoods import Gaussian
    from doubly_stochastic_dgp.dgp import DGP

    assert gpflow.__version__.startswith("1.1"), gpflow.__version__
    rng = np.random.RandomState(0)
    cases = {
        "ref_svgp_rbf": dict(dims=[3, 1], N=30, M=8, S=1, kern=RBF, white=False),
        "ref_dgp2_rbf": dict(dims=[3, 3, 1], N=40, M=10, S=3, kern=RBF, white=False),
        "ref_dgp3_matern_white": dict(dims=[4, 4, 4, 2], N=50, M=12, S=2, kern=Matern52, white=True),
    }
    for name, c in cases.items():
        dims, N, M, S = c["dims"], c["N"], c["M"], c["S"]
        X = rng.randn(N, dims[0])
        Y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.randn(N, 1)
        Y = np.tile(Y, (1, dims[-1]))
        Z = X[:M] + 0.3 * rng.randn(M, dims[0])
        kernels = [c["kern"](d, lengthscales=float(np.sqrt(d)), variance=0.5) for d in dims[:-1]]
        lik = Gaussian()
        lik.variance = 0.05
        with gpflow.defer_build():
            m = DGP(X, Y, Z, kernels, lik, num_outputs=dims[-1], white=c["white"], num_samples=S)
        out = dict(in_X=X, in_Y=Y, in_Z=Z, in_dims=np.array(dims), in_S=S, in_white=c["white"],
                   in_kernel=c["kern"].__name__, in_lik_var=0.05, in_jitter=gpflow.settings.numerics.jitter_level)
        for l, layer in enumerate(m.layers):
            q_mu = 0.3 * rng.randn(*layer.q_mu.shape)
            q_sqrt = np.tril(0.1 * rng.randn(*layer.q_sqrt.shape)) + 0.3 * np.eye(M)[None]
            layer.q_mu = q_mu
            layer.q_sqrt = q_sqrt
            out["in_q_mu%d" % l], out["in_q_sqrt%d" % l] = q_mu, q_sqrt
            out["in_Z%d" % l] = layer.feature.Z.read_value() if hasattr(layer.feature.Z, "read_value") else layer.feature.Z.value
            out["in_mean%d" % l] = type(layer.mean_function).__name__
            if hasattr(layer.mean_function, "A"):
                out["in_W%d" % l] = layer.mean_function.A.read_value()
        m.compile()
        sess = m.enquire_session()
        zs = [rng.randn(S, N, layer.num_outputs) for layer in m.layers]
        with tf.name_scope("golden"):
            Xph = tf.constant(X)
            Fs, Fmeans, Fvars = m.propagate(Xph, S=S, zs=[tf.constant(z) for z in zs])
            kls = [layer.KL() for layer in m.layers]
        with gpflow.params_as_tensors_for(m):
            pass
        vals = sess.run([Fs, Fmeans, Fvars, kls])
        for l in range(len(m.layers)):
            out["in_z%d" % l] = zs[l]
            out["out_F%d" % l], out["out_Fmean%d" % l], out["out_Fvar%d" % l] = vals[0][l], vals[1][l], vals[2][l]
            out["out_KL%d" % l] = vals[3][l]
        if len(m.layers) == 1:
            out["out_elbo"] = m.compute_log_likelihood()
            Xs = rng.randn(11, dims[0])
            out["in_Xs"] = Xs
            out["out_predict_f_mean"], out["out_predict_f_var"] = m.predict_f(Xs, 1)
            out["out_predict_y_mean"], out["out_predict_y_var"] = m.predict_y(Xs, 1)
            Ys = rng.randn(11, dims[-1])
            out["in_Ys"] = Ys
            out["out_predict_density"] = m.predict_density(Xs, Ys, 1)
        np.savez(os.path.join(HERE, name + ".npz"), **out)
        print("wrote", name)


if __name__ == "__main__":
    sys.exit(main())
