"""ctypes binding of libdsdgp.so (include/dsdgp.h).  This is the only place the Python host touches
the device; it fails loudly if the library is missing -- there is no CPU fallback."""
import ctypes as C
import os

import numpy as np

MAX_LAYERS = 16
F_Z, F_Q_MU, F_Q_SQRT, F_LENGTHSCALES, F_VARIANCE, F_MEAN_W, F_MEAN_B, F_LIK_VARIANCE, F_WHITE_VARIANCE = range(9)
FLAG_DEVICE_PTRS, FLAG_NO_SYNC = 1, 2
ERR_NOT_PD = -3


class LayerDesc(C.Structure):
    _fields_ = [("M", C.c_int), ("D_in", C.c_int), ("D_out", C.c_int), ("kernel", C.c_int),
                ("ard", C.c_int), ("white", C.c_int), ("mean", C.c_int), ("kernel_white", C.c_int),
                ("input_prop_dim", C.c_int)]


class Desc(C.Structure):
    _fields_ = [("L", C.c_int), ("layers", LayerDesc * MAX_LAYERS), ("likelihood", C.c_int),
                ("num_classes", C.c_int), ("D_y", C.c_int), ("jitter", C.c_double), ("N_max", C.c_int),
                ("S_max", C.c_int), ("device", C.c_int)]


class DsdgpError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"dsdgp error {code}: {msg}")
        self.code = code


_lib = None
FP = C.POINTER(C.c_float)
DP = C.POINTER(C.c_double)
FPP = C.POINTER(FP)

SYMBOLS = ["dsdgp_last_error", "dsdgp_version", "dsdgp_create", "dsdgp_destroy", "dsdgp_set_param",
           "dsdgp_get_param", "dsdgp_get_grad", "dsdgp_propagate", "dsdgp_elbo", "dsdgp_elbo_grad",
           "dsdgp_adam_init", "dsdgp_train_step", "dsdgp_kl", "dsdgp_comm_unique_id", "dsdgp_comm_init", "dsdgp_sync",
           "dsdgp_launch_count", "dsdgp_last_step_ms", "dsdgp_set_option", "dsdgp_timer_start",
           "dsdgp_timer_stop", "dsdgp_profile", "dsdgp_set_trainable", "dsdgp_natgrad_step", "dsdgp_predict_y",
           "dsdgp_predict_density", "dsdgp_propagate_full_cov", "dsdgp_set_sample_weights", "dsdgp_likelihood_apply",
           "dsdgp_set_stream", "dsdgp_device_buffers", "dsdgp_param_offset"]


def lib_path():
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lib", "libdsdgp.so")


def load():
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise DsdgpError(-100, f"{path} not found: build it with `python doubly-stochastic-dgp_b200/build.py` "
                               "(there is no CPU fallback)")
    lib = C.CDLL(path)
    lib.dsdgp_last_error.restype = C.c_char_p
    lib.dsdgp_version.restype = C.c_char_p
    lib.dsdgp_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(Desc)]
    lib.dsdgp_destroy.argtypes = [C.c_void_p]
    for f in (lib.dsdgp_set_param, lib.dsdgp_get_param, lib.dsdgp_get_grad):
        f.argtypes = [C.c_void_p, C.c_int, C.c_int, DP, C.c_size_t]
    lib.dsdgp_propagate.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint64,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint]
    for f in (lib.dsdgp_elbo, lib.dsdgp_elbo_grad):
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_uint64,
                      C.c_uint, DP]
    lib.dsdgp_adam_init.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double]
    lib.dsdgp_train_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p,
                                     C.c_uint64, C.c_uint, DP]
    lib.dsdgp_propagate_full_cov.argtypes = lib.dsdgp_propagate.argtypes
    lib.dsdgp_predict_y.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p,
                                    C.c_void_p, C.c_uint]
    lib.dsdgp_predict_density.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint64,
                                          C.c_void_p, C.c_uint]
    lib.dsdgp_set_sample_weights.argtypes = [C.c_void_p, DP, C.c_int]
    lib.dsdgp_likelihood_apply.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_uint]
    lib.dsdgp_set_trainable.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.dsdgp_natgrad_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p,
                                       C.c_uint64, C.c_uint, C.POINTER(C.c_int), C.c_int, C.c_double, DP]
    lib.dsdgp_timer_start.argtypes = [C.c_void_p]
    lib.dsdgp_timer_stop.argtypes = [C.c_void_p, FP]
    lib.dsdgp_profile.argtypes = [C.c_void_p, FP, C.c_int]
    lib.dsdgp_kl.argtypes = [C.c_void_p, DP]
    lib.dsdgp_comm_unique_id.argtypes = [C.c_void_p]
    lib.dsdgp_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    lib.dsdgp_sync.argtypes = [C.c_void_p]
    lib.dsdgp_launch_count.argtypes = [C.c_void_p]
    lib.dsdgp_launch_count.restype = C.c_longlong
    lib.dsdgp_last_step_ms.argtypes = [C.c_void_p, FP]
    lib.dsdgp_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
    lib.dsdgp_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    lib.dsdgp_device_buffers.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    lib.dsdgp_param_offset.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.dsdgp_param_offset.restype = C.c_longlong
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise DsdgpError(rc, load().dsdgp_last_error().decode())


def _ptr(a):
    """host ndarray (float32, C-contiguous) or an int device pointer -> void*"""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


def _ptr_array(arrs, L):
    if arrs is None:
        return None, None
    arr = (C.c_void_p * L)()
    for i in range(L):
        a = arrs[i] if i < len(arrs) else None
        arr[i] = None if a is None else (a if isinstance(a, int) else a.ctypes.data)
    return arr, arrs


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Context:
    """Owns one dsdgp_ctx."""
    def __init__(self, layer_descs, likelihood, num_classes, D_y, jitter, N_max, S_max, device=0):
        lib = load()
        d = Desc()
        d.L = len(layer_descs)
        for i, ld in enumerate(layer_descs):
            d.layers[i] = LayerDesc(*ld)
        d.likelihood, d.num_classes, d.D_y = likelihood, num_classes, D_y
        d.jitter, d.N_max, d.S_max, d.device = jitter, N_max, S_max, device
        self.desc = d
        self.L = d.L
        self.h = C.c_void_p()
        check(lib.dsdgp_create(C.byref(self.h), C.byref(d)))
        self.lib = lib
        self.N_max, self.S_max = N_max, S_max

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.dsdgp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_param(self, layer, field, value):
        v = np.ascontiguousarray(value, dtype=np.float64).reshape(-1)
        check(self.lib.dsdgp_set_param(self.h, layer, field, v.ctypes.data_as(DP), v.size))

    def _get(self, fn, layer, field, shape):
        out = np.empty(int(np.prod(shape)) if len(shape) else 1, dtype=np.float64)
        check(fn(self.h, layer, field, out.ctypes.data_as(DP), out.size))
        return out.reshape(shape)

    def get_param(self, layer, field, shape):
        return self._get(self.lib.dsdgp_get_param, layer, field, shape)

    def get_grad(self, layer, field, shape):
        return self._get(self.lib.dsdgp_get_grad, layer, field, shape)

    def propagate(self, X, S, zs=None, seed=0, want=(True, True, True), flags=0):
        X, _ = self._xy(X)
        N = X.shape[0]
        L = self.L
        douts = [self.desc.layers[l].D_out for l in range(L)]
        ipds = [self.desc.layers[l].input_prop_dim for l in range(L)]
        zarr, keep = self._zs(zs, S, N)
        outs = []
        arrs = []
        for k, w in enumerate(want):
            if w:     # Fs carry the propagated inputs in front of the samples (layers.py:105-117)
                o = [np.empty((S, N, douts[l] + (ipds[l] if k == 0 else 0)), dtype=np.float32) for l in range(L)]
                a, _ = _ptr_array(o, L)
            else:
                o, a = None, None
            outs.append(o); arrs.append(a)
        check(self.lib.dsdgp_propagate(self.h, _ptr(X), N, S, zarr, seed, arrs[0], arrs[1], arrs[2], flags))
        return outs

    def propagate_full_cov(self, X, S, zs=None, seed=0):
        """full_cov=True propagate: Fs, Fmeans lists of (S,N,D_l); Fvars list of (S,N,N,D_l)."""
        X, _ = self._xy(X)
        N, L = X.shape[0], self.L
        douts = [self.desc.layers[l].D_out for l in range(L)]
        zarr, keep = self._zs(zs, S, N)
        Fs = [np.empty((S, N, d), dtype=np.float32) for d in douts]
        Fm = [np.empty((S, N, d), dtype=np.float32) for d in douts]
        Fv = [np.empty((S, N, N, d), dtype=np.float32) for d in douts]
        a0, a1, a2 = _ptr_array(Fs, L)[0], _ptr_array(Fm, L)[0], _ptr_array(Fv, L)[0]
        check(self.lib.dsdgp_propagate_full_cov(self.h, _ptr(X), N, S, zarr, seed, a0, a1, a2, 0))
        return Fs, Fm, Fv

    def _zs(self, zs, S=None, N=None):
        """Per-layer draws -> (void*[L], keep-alive list).  Host arrays are broadcast to (S, N, D_out_l) like the reference's
        `mean + z * sqrt(var)` (utils.py:41; DGP_Quad passes (S,1,D) nodes, dgp.py:147-148) or rejected: the C side copies
        exactly S*N*D_out_l floats per layer, so a short buffer must never reach it.  ints are device pointers."""
        if zs is None:
            return None, None
        if len(zs) > self.L:
            raise ValueError(f"zs has {len(zs)} entries for {self.L} layers")
        zs32 = []
        for l, z in enumerate(zs):
            if z is None or isinstance(z, int):
                zs32.append(z)
                continue
            z = np.asarray(z)
            if S is not None:
                want = (S, N, self.desc.layers[l].D_out)
                if z.shape != want:
                    try:
                        z = np.broadcast_to(z, want)
                    except ValueError:
                        raise ValueError(f"zs[{l}] has shape {z.shape}, not broadcastable to (S, N, D_out) = {want}") from None
            zs32.append(f32(z))
        zarr, _ = _ptr_array(zs32, self.L)
        return zarr, zs32

    def _xy(self, X, Y=None):
        """float32 C-contiguous copies of host X (N, D_in) and Y (N, D_y) with the shapes the C side will read."""
        X = f32(X)
        if X.ndim != 2 or X.shape[1] != self.desc.layers[0].D_in:
            raise ValueError(f"X has shape {X.shape}, expected (N, {self.desc.layers[0].D_in})")
        if Y is None:
            return X, None
        Y = f32(Y)
        if Y.ndim == 1:
            Y = Y.reshape(-1, 1)
        if Y.shape != (X.shape[0], self.desc.D_y):
            raise ValueError(f"Y has shape {Y.shape}, expected ({X.shape[0]}, {self.desc.D_y})")
        return X, Y

    def predict_y(self, X, S, zs=None, seed=0):
        """likelihood.predict_mean_and_var of the last layer, per sample: two (S,N,D_last) arrays."""
        X, _ = self._xy(X)
        N, D = X.shape[0], self.desc.layers[self.L - 1].D_out
        zarr, keep = self._zs(zs, S, N)
        mean, var = np.empty((S, N, D), dtype=np.float32), np.empty((S, N, D), dtype=np.float32)
        check(self.lib.dsdgp_predict_y(self.h, _ptr(X), N, S, zarr, seed, _ptr(mean), _ptr(var), 0))
        return mean, var

    def predict_density(self, X, Y, S, zs=None, seed=0):
        """logsumexp_S(likelihood.predict_density - log S): (N,D_y) Gaussian, (N,1) MultiClass."""
        X, Y = self._xy(X, Y)
        N = X.shape[0]
        Do = 1 if self.desc.likelihood == 1 else self.desc.D_y      # MultiClass: one density per row
        zarr, keep = self._zs(zs, S, N)
        out = np.empty((N, Do), dtype=np.float32)
        check(self.lib.dsdgp_predict_density(self.h, _ptr(X), _ptr(Y), N, S, zarr, seed, _ptr(out), 0))
        return out

    def likelihood_apply(self, what, Fmu, Fvar, Y=None):
        """BroadcastingLikelihood.{variational_expectations (0), predict_mean_and_var (1), predict_density (2)} on (S,N,D)
        marginals (utils.py:88-121), evaluated by the device epilogues.  Returns float32 arrays of shape (S,N,-1)."""
        Fmu, Fvar = f32(Fmu), f32(Fvar)
        D = self.desc.layers[self.L - 1].D_out
        if Fmu.ndim != 3 or Fmu.shape[2] != D or Fvar.shape != Fmu.shape:
            raise ValueError(f"Fmu / Fvar must be (S, N, {D}) arrays of the same shape, got {Fmu.shape} and {Fvar.shape}")
        S, N = Fmu.shape[:2]
        Do = 1 if self.desc.likelihood == 1 else D
        if what == 1:
            mean, var = np.empty((S, N, D), dtype=np.float32), np.empty((S, N, D), dtype=np.float32)
            check(self.lib.dsdgp_likelihood_apply(self.h, what, _ptr(Fmu), _ptr(Fvar), None, S, N, _ptr(mean), _ptr(var), 0))
            return mean, var
        Y = f32(Y)
        if Y.ndim == 1:
            Y = Y.reshape(-1, 1)
        if Y.shape != (N, self.desc.D_y):
            raise ValueError(f"Y has shape {Y.shape}, expected ({N}, {self.desc.D_y})")
        out = np.empty((S, N, Do), dtype=np.float32)
        check(self.lib.dsdgp_likelihood_apply(self.h, what, _ptr(Fmu), _ptr(Fvar), _ptr(Y), S, N, _ptr(out), None, 0))
        return out

    def _elbo(self, fn, X, Y, S, num_data, zs, seed, flags):
        if isinstance(X, int):
            raise TypeError("device-pointer calls need explicit N: use elbo_dev")
        X, Y = self._xy(X, Y)
        zarr, keep = self._zs(zs, S, X.shape[0])
        e = C.c_double()
        check(fn(self.h, _ptr(X), _ptr(Y), X.shape[0], S, float(num_data), zarr, seed, flags, C.byref(e)))
        return e.value

    def elbo(self, X, Y, S, num_data, zs=None, seed=0):
        return self._elbo(self.lib.dsdgp_elbo, X, Y, S, num_data, zs, seed, 0)

    def elbo_grad(self, X, Y, S, num_data, zs=None, seed=0):
        return self._elbo(self.lib.dsdgp_elbo_grad, X, Y, S, num_data, zs, seed, 0)

    def adam_init(self, lr=0.01, beta1=0.9, beta2=0.999, eps=1e-8):
        check(self.lib.dsdgp_adam_init(self.h, lr, beta1, beta2, eps))

    def train_step(self, X, Y, N, S, num_data, seed, flags=0, want_elbo=True, zs=None):
        """X, Y: float32 host arrays (pinned for the e2e path) or int device pointers (with FLAG_DEVICE_PTRS)."""
        e = C.c_double()
        if not isinstance(X, int):
            X, Y = self._xy(X, Y)
            if X.shape[0] != N:
                raise ValueError(f"X has {X.shape[0]} rows, N={N}")
        zarr, keep = self._zs(zs, S, N)
        check(self.lib.dsdgp_train_step(self.h, _ptr(X), _ptr(Y), N, S, float(num_data), zarr, seed, flags,
                                        C.byref(e) if want_elbo else None))
        return e.value if want_elbo else None

    def set_sample_weights(self, w):
        """w: S weights summing to 1 (DGP_Quad), or None for the Monte-Carlo mean."""
        if w is None:
            check(self.lib.dsdgp_set_sample_weights(self.h, None, 0))
        else:
            w = np.ascontiguousarray(w, dtype=np.float64).reshape(-1)
            check(self.lib.dsdgp_set_sample_weights(self.h, w.ctypes.data_as(DP), w.size))

    def set_trainable(self, layer, field, flag):
        check(self.lib.dsdgp_set_trainable(self.h, layer, field, int(bool(flag))))

    def natgrad_step(self, X, Y, N, S, num_data, seed, layers, gamma, flags=0, zs=None):
        """ELBO + gradient pass, then the natural-gradient update of (q_mu, q_sqrt) of `layers`; returns the ELBO."""
        e = C.c_double()
        if not isinstance(X, int):
            X, Y = self._xy(X, Y)
            if X.shape[0] != N:
                raise ValueError(f"X has {X.shape[0]} rows, N={N}")
        zarr, keep = self._zs(zs, S, N)
        ids = (C.c_int * len(layers))(*[int(l) for l in layers])
        check(self.lib.dsdgp_natgrad_step(self.h, _ptr(X), _ptr(Y), N, S, float(num_data), zarr, seed, flags, ids,
                                          len(layers), float(gamma), C.byref(e)))
        return e.value

    def timer_start(self):
        check(self.lib.dsdgp_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        check(self.lib.dsdgp_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def profile(self):
        n = 5 + 3 * self.L
        buf = (C.c_float * n)()
        got = self.lib.dsdgp_profile(self.h, buf, n)
        if got < 0:
            check(got)
        return list(buf)

    def kl(self):
        out = np.empty(self.L, dtype=np.float64)
        check(self.lib.dsdgp_kl(self.h, out.ctypes.data_as(DP)))
        return out

    def comm_init(self, id_bytes, rank, world):
        buf = C.create_string_buffer(bytes(id_bytes), 128)
        check(self.lib.dsdgp_comm_init(self.h, buf, rank, world))

    def sync(self):
        check(self.lib.dsdgp_sync(self.h))

    def launch_count(self):
        return int(self.lib.dsdgp_launch_count(self.h))

    def last_step_ms(self):
        ms = C.c_float()
        check(self.lib.dsdgp_last_step_ms(self.h, C.byref(ms)))
        return ms.value

    def set_stream(self, stream):
        """Enqueue all following work on a caller-owned CUDA stream (int handle, e.g. torch.cuda.Stream().cuda_stream); None:
        back to the context's own stream."""
        check(self.lib.dsdgp_set_stream(self.h, C.c_void_p(int(stream)) if stream else None))

    def device_buffers(self):
        """(params_ptr, grads_ptr, n): device addresses of the flat fp32 parameter / gradient buffers."""
        p, g, n = C.c_void_p(), C.c_void_p(), C.c_size_t()
        check(self.lib.dsdgp_device_buffers(self.h, C.byref(p), C.byref(g), C.byref(n)))
        return p.value, g.value, n.value

    def param_offset(self, layer, field):
        return int(self.lib.dsdgp_param_offset(self.h, layer, field))

    def set_option(self, name, value):
        check(self.lib.dsdgp_set_option(self.h, name.encode(), float(value)))


def comm_unique_id():
    buf = C.create_string_buffer(128)
    check(load().dsdgp_comm_unique_id(buf))
    return buf.raw
