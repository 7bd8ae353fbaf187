"""Host mirror of doubly_stochastic_dgp/layer_initializations.py:16-52 (`init_layers_linear`): Identity mean
when widths match, PCA projection when stepping down, identity padding when stepping up; Z pushed through W."""
import numpy as np

from .layers import SVGP_Layer
from .mean_functions import Identity, Linear, Zero


def init_layers_linear(X, Y, Z, kernels, num_outputs=None, mean_function=None, Layer=SVGP_Layer, white=False):
    num_outputs = num_outputs or Y.shape[1]
    mean_function = Zero() if mean_function is None else mean_function
    layers = []
    X_running, Z_running = np.array(X, dtype=np.float64), np.array(Z, dtype=np.float64)
    for kern_in, kern_out in zip(kernels[:-1], kernels[1:]):
        dim_in, dim_out = kern_in.input_dim, kern_out.input_dim
        if dim_in == dim_out:
            mf = Identity()
        else:
            if dim_in > dim_out:      # stepping down: PCA projection (layer_initializations.py:34-36)
                _, _, V = np.linalg.svd(X_running, full_matrices=False)
                W = V[:dim_out, :].T
            else:                     # stepping up: identity + zero padding (:38-39)
                W = np.concatenate([np.eye(dim_in), np.zeros((dim_in, dim_out - dim_in))], 1)
            mf = Linear(W)
        layers.append(Layer(kern_in, Z_running, dim_out, mf, white=white))
        if dim_in != dim_out:
            Z_running = Z_running.dot(W)
            X_running = X_running.dot(W)
    layers.append(Layer(kernels[-1], Z_running, num_outputs, mean_function, white=white))
    return layers


def init_layers_input_prop(X, Y, Z, kernels, num_outputs=None, mean_function=None, Layer=SVGP_Layer, white=False):
    """layer_initializations.py:55-81: every non-final layer propagates the D input columns in front of its outputs (kernel l
    sees D + D_out(l-1) inputs); inducing inputs are Z padded with random columns of the previous kernel's scale."""
    num_outputs = num_outputs or Y.shape[1]
    mean_function = Zero() if mean_function is None else mean_function
    X, Z = np.asarray(X, dtype=np.float64), np.asarray(Z, dtype=np.float64)
    D, M = X.shape[1], Z.shape[0]
    layers = []
    for kern_in, kern_out in zip(kernels[:-1], kernels[1:]):
        dim_in = kern_in.input_dim
        dim_out = kern_out.input_dim - D
        std_in = float(kern_in.variance.read_value()) ** 0.5
        pad = np.random.randn(M, dim_in - D) * 2. * std_in
        layers.append(Layer(kern_in, np.concatenate([Z, pad], 1), dim_out, Zero(), white=white, input_prop_dim=D))
    dim_in = kernels[-1].input_dim
    std_in = float(kernels[-2].variance.read_value()) ** 0.5 if dim_in > D else 1.
    pad = np.random.randn(M, dim_in - D) * 2. * std_in
    layers.append(Layer(kernels[-1], np.concatenate([Z, pad], 1), num_outputs, mean_function, white=white))
    return layers


def kmeans_inducing_points(X, M, seed=0):
    """Z = kmeans2(X, M, minit='points')[0] (demos/run_regression.py:57, demos/demo_regression_UCI.ipynb): M cluster centres of
    the training inputs as the first layer's inducing inputs.  Construction-time host work (SciPy), like the reference."""
    from scipy.cluster.vq import kmeans2
    X = np.asarray(X, dtype=np.float64)
    M = min(int(M), X.shape[0])
    try:
        return kmeans2(X, M, minit='points', seed=seed)[0]
    except TypeError:            # SciPy < 1.7: no seed argument
        state = np.random.get_state()
        np.random.seed(seed)
        try:
            return kmeans2(X, M, minit='points')[0]
        finally:
            np.random.set_state(state)
