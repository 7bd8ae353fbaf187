"""gpflow.settings stand-in (reference: `settings.jitter` layers.py:162,171, utils.py:41,47;
`settings.float_type` dgp.py:26).  `jitter` is read when a model first touches the device."""
import contextlib

import numpy as np

jitter = 1e-6          # gpflow default numerics.jitter_level
float_type = np.float64


@contextlib.contextmanager
def temp_settings(jitter_level=None):
    """tests/test_dgp.py:7-11 `settings.temp_settings(custom_config)` equivalent."""
    global jitter
    old = jitter
    if jitter_level is not None:
        jitter = jitter_level
    try:
        yield
    finally:
        jitter = old
