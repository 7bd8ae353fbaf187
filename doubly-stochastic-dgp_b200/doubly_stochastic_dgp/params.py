"""Minimal stand-in for gpflow.params.Parameter: a named float64 array with `.value` /
`.read_value()` (demos/run_regression.py:73 `layer.q_sqrt.value`), plus the owner-side
`obj.name = ndarray` assignment semantics (tests/test_dgp.py:91-92)."""
import numpy as np


class Parameter:
    def __init__(self, value, trainable=True):
        self._value = np.array(value, dtype=np.float64)
        self.trainable = trainable
        self._owner = None        # object notified when the value changes / is needed
        self._version = 0

    # --- gpflow-like accessors
    @property
    def value(self):
        return self.read_value()

    def read_value(self):
        if self._owner is not None:
            self._owner._refresh_from_device()
        return self._value.copy()

    @property
    def shape(self):
        return self._value.shape

    def assign(self, value):
        value = np.asarray(value, dtype=np.float64)
        if value.shape != self._value.shape:
            value = np.broadcast_to(value, self._value.shape)
        if self._owner is not None:
            self._owner._refresh_from_device()
        self._value = np.array(value, dtype=np.float64)
        self._version += 1
        if self._owner is not None:
            self._owner._mark_host_dirty()

    def set_trainable(self, flag):
        """gpflow Parameter.set_trainable: an untrainable parameter is left alone by AdamOptimizer
        (demos/using_natural_gradients.ipynb takes the NatGrad-managed q_mu, q_sqrt away from Adam this way)."""
        self.trainable = bool(flag)
        if self._owner is not None:
            self._owner._mark_trainable_dirty()

    def __array__(self, dtype=None, copy=None):
        v = self.read_value()
        return v.astype(dtype) if dtype is not None else v

    def __repr__(self):
        return f"Parameter(shape={self._value.shape})"


class Parameterized:
    """Attribute assignment onto an existing Parameter assigns its value (gpflow semantics)."""
    def __setattr__(self, name, value):
        cur = self.__dict__.get(name)
        if isinstance(cur, Parameter) and not isinstance(value, Parameter):
            cur.assign(value)
        else:
            object.__setattr__(self, name, value)

    def parameters(self):
        out = []
        for v in self.__dict__.values():
            if isinstance(v, Parameter):
                out.append(v)
            elif isinstance(v, Parameterized):
                out.extend(v.parameters())
            elif isinstance(v, (list, tuple)):
                for x in v:
                    if isinstance(x, Parameterized):
                        out.extend(x.parameters())
        return out
