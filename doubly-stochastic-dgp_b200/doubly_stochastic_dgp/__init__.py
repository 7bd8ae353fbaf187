"""B200-native drop-in for the hot path of UCL-SML/Doubly-Stochastic-DGP.

Same module names and constructor / method signatures as the reference package
(`doubly_stochastic_dgp.dgp.DGP`, `.layers.SVGP_Layer`, `.utils.reparameterize`, ...); the
arithmetic runs in hand-written sm_100a CUDA behind the C-ABI of include/dsdgp.h (ctypes, `_lib.py`).
The GPflow objects the reference takes as arguments (kernels, likelihoods, mean functions,
settings) are replaced by the light descriptors in `.kernels`, `.likelihoods`, `.mean_functions`,
`.settings`.  There is no CPU fallback: without libdsdgp.so or a CUDA device every compute call raises.
"""
from . import settings  # noqa: F401
