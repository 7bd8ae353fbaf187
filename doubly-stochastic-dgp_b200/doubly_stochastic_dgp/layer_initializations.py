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
