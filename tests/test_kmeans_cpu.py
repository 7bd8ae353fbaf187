"""k-means inducing-point initialisation (demos/run_regression.py:57) and the kernel-sum descriptor: host-side pieces."""
import numpy as np
import pytest


def test_kmeans_inducing_points_are_cluster_centres():
    from doubly_stochastic_dgp.layer_initializations import kmeans_inducing_points
    rng = np.random.default_rng(0)
    centres = np.array([[0.0, 0.0], [10.0, 0.0], [0.0, 10.0], [10.0, 10.0]])
    X = np.concatenate([c + 0.1 * rng.normal(size=(50, 2)) for c in centres])
    Z = kmeans_inducing_points(X, 8, seed=1)
    assert Z.shape == (8, 2)
    assert np.allclose(Z, kmeans_inducing_points(X, 8, seed=1))     # seeded
    cost = lambda C: np.sum(np.min(np.linalg.norm(X[:, None] - C[None], axis=-1), 1) ** 2)
    assert cost(Z) <= cost(X[rng.choice(len(X), 8, replace=False)])   # Lloyd iterations only lower the quantisation error
    assert np.all(Z.min(0) >= X.min(0) - 1e-9) and np.all(Z.max(0) <= X.max(0) + 1e-9)
    assert kmeans_inducing_points(X[:3], 10).shape == (3, 2)      # M capped at the number of points


def test_sum_kernel_descriptor():
    from doubly_stochastic_dgp.kernels import RBF, Matern52, White
    k = RBF(3, variance=2.0, lengthscales=1.5) + White(3, variance=0.01)
    assert k.code == 0 and k.input_dim == 3 and float(k.white_variance.value) == 0.01 and float(k.variance.value) == 2.0
    k2 = k + White(3, variance=0.5)
    assert abs(float(k2.white_variance.value) - 0.51) < 1e-12
    assert RBF(3).white_variance is None
    with pytest.raises(NotImplementedError):
        RBF(3) + Matern52(3)
