"""TEST INFRASTRUCTURE -- a minimal torch-backed stand-in for the parts of TensorFlow 1.x and GPflow 1.1 that the reference's
own source files touch, so that /root/reference/doubly_stochastic_dgp/{dgp,layers,utils,layer_initializations}.py can be
IMPORTED AND EXECUTED UNMODIFIED in a container that can install neither package (tests/golden/make_from_reference_shim.py).

What this pins and what it does not: every line of the reference-owned code on the hot path (propagate, conditional_ND /
conditional_SND, sample_from_conditional incl. input propagation, reparameterize diagonal and full-covariance, KL, E_log_p_Y,
_build_likelihood, predict_*, BroadcastingLikelihood's tiling, DGP_Quad, init_layers_linear / init_layers_input_prop) runs as
written, eagerly, in float64.  The GPflow pieces underneath (kernels, Gaussian / Bernoulli likelihoods, InducingPoints, mean
functions, mvhermgauss, the LowerTriangular transform, params_as_tensors / autoflow) are re-implemented HERE from GPflow 1.1's
documented behaviour -- they are not GPflow, so the oracle stays "parity unpinned" at that boundary.

Semantics kept from TF 1.x graph mode where eager torch would differ:
  * `a += b` on a tensor rebinds instead of writing in place (layers.py:201 adds a (D,M,M) product to a (1,M,M) `-I`);
  * SVGP_Layer caches Ku / Lu behind `needs_build_cholesky` (layers.py:167-176) -- in a TF graph those are symbolic and follow
    the parameters; here the cache flag is reset before every top-level evaluation (`_fresh`);
  * `tf.random_normal` draws are recorded (`DRAWS`) so that the same z can be fed to the oracle.
Only tests/ and tests/golden/ may import this module."""
import itertools
import sys
import types

import numpy as np
import torch

F64 = torch.float64
DRAWS = []                 # every tf.random_normal draw since the last reset_draws(), in call order
_GEN = torch.Generator().manual_seed(0)
_TENSOR_MODE = [0]         # depth of active params_as_tensors / autoflow scopes (GPflow: inherited through the parent chain)


class T(torch.Tensor):
    """torch.Tensor whose augmented assignments rebind (TF graph semantics) instead of mutating"""
    __array_ufunc__ = None          # numpy operands on the LEFT (DGP_Quad's gh_x * tensor) defer to the reflected methods below

    def __iadd__(self, o): return self + o
    def __isub__(self, o): return self - o
    def __imul__(self, o): return self * o
    def __itruediv__(self, o): return self / o
    def __add__(self, o): return torch.Tensor.__add__(self, _nd(o))
    def __sub__(self, o): return torch.Tensor.__sub__(self, _nd(o))
    def __mul__(self, o): return torch.Tensor.__mul__(self, _nd(o))
    def __truediv__(self, o): return torch.Tensor.__truediv__(self, _nd(o))
    def __radd__(self, o): return torch.Tensor.__add__(_nd(o, self), self)
    def __rsub__(self, o): return torch.Tensor.__sub__(_nd(o, self), self)
    def __rmul__(self, o): return torch.Tensor.__mul__(_nd(o, self), self)
    def __rtruediv__(self, o): return torch.Tensor.__truediv__(_nd(o, self), self)


def _nd(o, like=None):
    """numpy operand -> tensor; python scalars on the reflected side -> 0-dim tensor"""
    if isinstance(o, np.ndarray):
        return torch.as_tensor(o, dtype=F64).as_subclass(T)
    if like is not None and not isinstance(o, torch.Tensor):
        return torch.as_tensor(o, dtype=like.dtype).as_subclass(T)
    return o


def _t(x, dtype=F64):
    if isinstance(x, Parameter):
        x = x._value
    if isinstance(x, torch.Tensor):
        return x.to(dtype).as_subclass(T) if x.dtype != dtype else x.as_subclass(T)
    return torch.as_tensor(np.asarray(x), dtype=dtype).as_subclass(T)


def reset_draws(seed=0):
    DRAWS.clear()
    _GEN.manual_seed(seed)


# ------------------------------------------------------------------------------------------------------------ tensorflow
def _shape_list(shape):
    return [int(s) for s in shape]


def _build_tf():
    tf = types.ModuleType("tensorflow")
    tf.float64, tf.float32, tf.int32, tf.int64 = torch.float64, torch.float32, torch.int32, torch.int64
    tf.shape = lambda x: tuple(_t(x).shape)
    tf.size = lambda x: int(_t(x).numel())
    tf.tile = lambda x, m: _t(x).repeat(*_shape_list(m))
    tf.expand_dims = lambda x, axis: _t(x).unsqueeze(axis)
    tf.reshape = lambda x, shape: _t(x).reshape(*_shape_list(shape))
    tf.eye = lambda n, dtype=F64: torch.eye(int(n), dtype=dtype).as_subclass(T)
    tf.zeros = lambda shape, dtype=F64: torch.zeros(*_shape_list(shape), dtype=dtype).as_subclass(T)
    tf.zeros_like = lambda x: torch.zeros_like(_t(x))
    tf.constant = lambda v, dtype=F64: _t(v, dtype)
    tf.identity = lambda x: _t(x).clone()
    tf.square = lambda x: _t(x) ** 2
    tf.sqrt = lambda x: torch.sqrt(_t(x))
    tf.log = lambda x: torch.log(_t(x))
    tf.exp = lambda x: torch.exp(_t(x))
    tf.negative = lambda x: -_t(x)
    tf.cholesky = lambda x: torch.linalg.cholesky(_t(x))
    tf.matrix_diag_part = lambda x: torch.diagonal(_t(x), dim1=-2, dim2=-1)
    tf.matrix_diag = lambda d: torch.diag_embed(_t(d))
    tf.concat = lambda xs, axis: torch.cat([_t(x) for x in xs], dim=axis)

    def cast(x, dtype):
        return _t(x, dtype)
    tf.cast = cast

    def stack(xs, axis=0):
        if isinstance(xs, torch.Tensor):
            return xs
        if all(isinstance(x, (int, np.integer)) for x in xs):
            return tuple(int(x) for x in xs)          # a shape
        return torch.stack([_t(x) for x in xs], dim=axis)
    tf.stack = stack

    def transpose(x, perm=None):
        x = _t(x)
        return x.permute(*(perm if perm is not None else reversed(range(x.dim()))))
    tf.transpose = transpose

    def matmul(a, b, transpose_a=False, transpose_b=False):
        a, b = _t(a), _t(b)
        if transpose_a:
            a = a.transpose(-1, -2)
        if transpose_b:
            b = b.transpose(-1, -2)
        return torch.matmul(a, b)
    tf.matmul = matmul

    def matrix_triangular_solve(matrix, rhs, lower=True):
        return torch.linalg.solve_triangular(_t(matrix), _t(rhs), upper=not lower)
    tf.matrix_triangular_solve = matrix_triangular_solve
    tf.cholesky_solve = lambda chol, rhs: torch.cholesky_solve(_t(rhs), _t(chol))

    def _reduce(fn):
        def red(x, axis=None, keepdims=False):
            if isinstance(x, (list, tuple)):
                x = torch.stack([_t(v) for v in x])
            x = _t(x)
            return fn(x) if axis is None else fn(x, dim=axis, keepdim=keepdims)
        return red
    tf.reduce_sum = _reduce(torch.sum)
    tf.reduce_mean = _reduce(torch.mean)
    tf.reduce_logsumexp = lambda x, axis=None: torch.logsumexp(_t(x), dim=axis)

    def random_normal(shape, dtype=F64, **kw):
        z = torch.randn(*_shape_list(shape), dtype=dtype, generator=_GEN).as_subclass(T)
        DRAWS.append(z.clone())
        return z
    tf.random_normal = random_normal

    def map_fn(fn, elems, dtype=None, **kw):
        outs = [fn(e) for e in _t(elems)]
        if isinstance(outs[0], (tuple, list)):
            return tuple(torch.stack([_t(o[i]) for o in outs]) for i in range(len(outs[0])))
        return torch.stack([_t(o) for o in outs])
    tf.map_fn = map_fn
    tf.fill = lambda shape, v: torch.full(_shape_list(shape), 1.0, dtype=F64).as_subclass(T) * _t(v)
    tf.squeeze = lambda x: _t(x).squeeze()
    return tf


# ---------------------------------------------------------------------------------------------------------------- gpflow
class _Settings:
    float_type = F64
    int_type = torch.int32
    jitter = 1e-6

    class numerics:
        jitter_level = 1e-6


settings = _Settings()


class Transform:
    def forward_value(self, v):
        return v


class LowerTriangular(Transform):
    def __init__(self, N, num_matrices=1, squeeze=False):
        self.N, self.num_matrices = N, num_matrices

    def forward_value(self, v):                 # GPflow packs the lower triangle: whatever sits above the diagonal is dropped
        return np.tril(v)


class Parameter:
    def __init__(self, value, transform=None, prior=None, trainable=True, dtype=None, name=None):
        self.transform = transform or Transform()
        self.trainable = trainable
        self.assign(value)

    def assign(self, value):
        if isinstance(value, Parameter):
            value = value.read_value()
        if isinstance(value, torch.Tensor):
            value = value.detach().numpy()
        self._value = torch.as_tensor(self.transform.forward_value(np.array(value, dtype=np.float64)), dtype=F64).as_subclass(T)

    def read_value(self, session=None):
        return self._value.numpy().copy()

    value = property(read_value)
    shape = property(lambda self: tuple(self._value.shape))

    def set_trainable(self, flag):
        self.trainable = flag


class DataHolder(Parameter):
    def __init__(self, value, **kw):
        Parameter.__init__(self, value)


class Minibatch(DataHolder):
    def __init__(self, value, batch_size=None, shuffle=True, seed=None, **kw):
        DataHolder.__init__(self, np.asarray(value)[:batch_size])


class Parameterized:
    def __init__(self, name=None, **kw):
        pass

    def __getattribute__(self, name):
        attr = object.__getattribute__(self, name)
        if _TENSOR_MODE[0] and isinstance(attr, Parameter):
            return attr._value
        return attr

    def __setattr__(self, name, value):
        cur = self.__dict__.get(name)
        if isinstance(cur, Parameter) and not isinstance(value, Parameter):
            cur.assign(value)                   # `layer.q_sqrt = ndarray` assigns to the existing Parameter (layers.py:164)
        else:
            object.__setattr__(self, name, value)

    def set_trainable(self, flag):
        for v in self.__dict__.values():
            if isinstance(v, (Parameter, Parameterized)):
                v.set_trainable(flag)


class ParamList(Parameterized):
    def __init__(self, items, **kw):
        object.__setattr__(self, "_items", list(items))

    def __iter__(self): return iter(self._items)
    def __len__(self): return len(self._items)
    def __getitem__(self, i): return self._items[i]


def _walk(obj, seen=None):
    seen = seen if seen is not None else set()
    if id(obj) in seen:
        return
    seen.add(id(obj))
    yield obj
    children = obj._items if isinstance(obj, ParamList) else list(vars(obj).values())
    for c in children:
        if isinstance(c, Parameterized):
            yield from _walk(c, seen)
        elif isinstance(c, (list, tuple)):
            for cc in c:
                if isinstance(cc, Parameterized):
                    yield from _walk(cc, seen)


def _fresh(root):
    """graph semantics for the Ku / Lu cache of SVGP_Layer (see the module docstring)"""
    for o in _walk(root):
        if "needs_build_cholesky" in vars(o):
            object.__setattr__(o, "needs_build_cholesky", True)


def _to_numpy(x):
    if isinstance(x, torch.Tensor):
        return x.detach().numpy().copy()
    if isinstance(x, (list, tuple)):
        return type(x)(_to_numpy(v) for v in x)
    return x


def params_as_tensors(method):
    def wrapper(obj, *a, **kw):
        _TENSOR_MODE[0] += 1
        try:
            return method(obj, *a, **kw)
        finally:
            _TENSOR_MODE[0] -= 1
    wrapper.__name__ = getattr(method, "__name__", "wrapped")
    return wrapper


def run_as_tensors(root, fn):
    """evaluate fn() the way a session.run of a freshly built graph would: tensor mode on, caches reset, numpy out"""
    _fresh(root)
    _TENSOR_MODE[0] += 1
    try:
        return _to_numpy(fn())
    finally:
        _TENSOR_MODE[0] -= 1


def autoflow(*specs):
    def deco(method):
        def wrapper(obj, *args):
            conv = [_t(a) if isinstance(a, np.ndarray) else a for a in args]
            return run_as_tensors(obj, lambda: method(obj, *conv))
        wrapper.__name__ = method.__name__
        return wrapper
    return deco


class Model(Parameterized):
    def __init__(self, name=None, **kw):
        Parameterized.__init__(self)

    def compute_log_likelihood(self):
        return float(run_as_tensors(self, self._build_likelihood))

    def compile(self, session=None):
        pass


# -- kernels (gpflow/kernels.py, 1.1: Stationary.square_dist / euclid_dist, RBF, Matern52, White, Combination/Sum)
class Kern(Parameterized):
    def __init__(self, input_dim, active_dims=None, name=None):
        Parameterized.__init__(self)
        self.input_dim = int(input_dim)

    def __add__(self, other):
        return Sum([self, other])

    def compute_K_symm(self, X):
        return run_as_tensors(self, lambda: self.K(_t(X)))

    def compute_K(self, X, Z):
        return run_as_tensors(self, lambda: self.K(_t(X), _t(Z)))


class Stationary(Kern):
    def __init__(self, input_dim, variance=1.0, lengthscales=None, active_dims=None, ARD=False, name=None):
        Kern.__init__(self, input_dim)
        self.variance = Parameter(variance)
        if ARD:
            lengthscales = np.ones(input_dim) if lengthscales is None else lengthscales * np.ones(input_dim)
        elif lengthscales is None:
            lengthscales = 1.0
        self.lengthscales = Parameter(lengthscales)
        self.ARD = ARD

    @params_as_tensors
    def square_dist(self, X, X2):
        X = _t(X) / self.lengthscales
        Xs = torch.sum(X ** 2, dim=1)
        if X2 is None:
            dist = -2 * torch.matmul(X, X.t())
            dist = dist + Xs.reshape(-1, 1) + Xs.reshape(1, -1)
            return dist
        X2 = _t(X2) / self.lengthscales
        X2s = torch.sum(X2 ** 2, dim=1)
        dist = -2 * torch.matmul(X, X2.t())
        dist = dist + Xs.reshape(-1, 1) + X2s.reshape(1, -1)
        return dist

    def euclid_dist(self, X, X2):
        return torch.sqrt(self.square_dist(X, X2) + 1e-12)

    @params_as_tensors
    def Kdiag(self, X):
        return torch.ones(_t(X).shape[0], dtype=F64).as_subclass(T) * self.variance.squeeze()


class RBF(Stationary):
    @params_as_tensors
    def K(self, X, X2=None):
        return self.variance * torch.exp(-self.square_dist(X, X2) / 2)


class Matern52(Stationary):
    @params_as_tensors
    def K(self, X, X2=None):
        r = self.euclid_dist(X, X2)
        return self.variance * (1.0 + np.sqrt(5.0) * r + 5.0 / 3.0 * r ** 2) * torch.exp(-np.sqrt(5.0) * r)


class White(Kern):
    def __init__(self, input_dim, variance=1.0, active_dims=None, name=None):
        Kern.__init__(self, input_dim)
        self.variance = Parameter(variance)

    @params_as_tensors
    def K(self, X, X2=None):
        X = _t(X)
        if X2 is None:
            return torch.diag_embed(torch.ones(X.shape[0], dtype=F64) * self.variance.squeeze()).as_subclass(T)
        return torch.zeros(X.shape[0], _t(X2).shape[0], dtype=F64).as_subclass(T)

    @params_as_tensors
    def Kdiag(self, X):
        return torch.ones(_t(X).shape[0], dtype=F64).as_subclass(T) * self.variance.squeeze()


class Sum(Kern):
    def __init__(self, kern_list, name=None):
        Kern.__init__(self, max(k.input_dim for k in kern_list))
        self.kern_list = list(kern_list)

    @property
    def variance(self):            # (init_layers_input_prop reads kern.variance of a Sum's first part in the demos)
        return self.kern_list[0].variance

    def K(self, X, X2=None):
        return sum(k.K(X, X2) for k in self.kern_list)

    def Kdiag(self, X):
        return sum(k.Kdiag(X) for k in self.kern_list)


# -- features (gpflow/features.py: InducingPoints.Kuu adds jitter * I, Kuf = K(Z, Xnew))
class InducingPoints(Parameterized):
    def __init__(self, Z, **kw):
        Parameterized.__init__(self)
        self.Z = Parameter(Z)

    def __len__(self):
        return self.Z.shape[0]

    @params_as_tensors
    def Kuu(self, kern, jitter=0.0):
        Kzz = kern.K(self.Z)
        return Kzz + jitter * torch.eye(Kzz.shape[0], dtype=F64)

    @params_as_tensors
    def Kuf(self, kern, Xnew):
        return kern.K(self.Z, Xnew)


# -- mean functions (gpflow/mean_functions.py)
class MeanFunction(Parameterized):
    pass


class Zero(MeanFunction):
    def __init__(self, output_dim=1):
        MeanFunction.__init__(self)
        self.output_dim = output_dim

    def __call__(self, X):
        return torch.zeros(_t(X).shape[0], self.output_dim, dtype=F64).as_subclass(T)


class Identity(MeanFunction):
    def __init__(self, input_dim=None):
        MeanFunction.__init__(self)

    def __call__(self, X):
        return _t(X)


class Linear(MeanFunction):
    def __init__(self, A=None, b=None):
        MeanFunction.__init__(self)
        A = np.ones((1, 1)) if A is None else A
        b = np.zeros(1) if b is None else b
        self.A, self.b = Parameter(np.atleast_2d(A)), Parameter(b)

    @params_as_tensors
    def __call__(self, X):
        return torch.matmul(_t(X), self.A) + self.b


# -- likelihoods (gpflow/likelihoods.py: Gaussian closed forms; Bernoulli with the probit link and 20-point Gauss-Hermite)
def _gauss_logdensity(x, mu, var):
    return -0.5 * (np.log(2 * np.pi) + torch.log(var) + (mu - x) ** 2 / var)


class Likelihood(Parameterized):
    num_gauss_hermite_points = 20

    def _gh(self):
        x, w = np.polynomial.hermite.hermgauss(self.num_gauss_hermite_points)
        return _t(x), _t(w / np.sqrt(np.pi))

    def predict_mean_and_var(self, Fmu, Fvar):
        x, w = self._gh()
        X = _t(Fmu)[..., None] + torch.sqrt(2.0 * _t(Fvar))[..., None] * x
        m = torch.sum(self.conditional_mean(X) * w, -1)
        e2 = torch.sum((self.conditional_variance(X) + self.conditional_mean(X) ** 2) * w, -1)
        return m, e2 - m ** 2

    def predict_density(self, Fmu, Fvar, Y):
        x, w = self._gh()
        X = _t(Fmu)[..., None] + torch.sqrt(2.0 * _t(Fvar))[..., None] * x
        return torch.log(torch.sum(torch.exp(self.logp(X, _t(Y)[..., None])) * w, -1))

    def variational_expectations(self, Fmu, Fvar, Y):
        x, w = self._gh()
        X = _t(Fmu)[..., None] + torch.sqrt(2.0 * _t(Fvar))[..., None] * x
        return torch.sum(self.logp(X, _t(Y)[..., None]) * w, -1)


class Gaussian(Likelihood):
    def __init__(self, variance=1.0, **kw):
        Likelihood.__init__(self)
        self.variance = Parameter(variance)

    @params_as_tensors
    def logp(self, F, Y):
        return _gauss_logdensity(_t(F), _t(Y), self.variance)

    @params_as_tensors
    def conditional_mean(self, F):
        return _t(F).clone()

    @params_as_tensors
    def conditional_variance(self, F):
        return torch.ones_like(_t(F)) * self.variance.squeeze()

    @params_as_tensors
    def predict_mean_and_var(self, Fmu, Fvar):
        return _t(Fmu).clone(), _t(Fvar) + self.variance

    @params_as_tensors
    def predict_density(self, Fmu, Fvar, Y):
        return _gauss_logdensity(_t(Fmu), _t(Y), _t(Fvar) + self.variance)

    @params_as_tensors
    def variational_expectations(self, Fmu, Fvar, Y):
        return -0.5 * np.log(2 * np.pi) - 0.5 * torch.log(self.variance) \
               - 0.5 * ((_t(Y) - _t(Fmu)) ** 2 + _t(Fvar)) / self.variance


def _probit(x):
    return 0.5 * (1.0 + torch.erf(x / np.sqrt(2.0))) * (1 - 2e-3) + 1e-3


class Bernoulli(Likelihood):
    def __init__(self, invlink=None, **kw):
        Likelihood.__init__(self)

    def logp(self, F, Y):
        p = _probit(_t(F))
        return torch.log(torch.where(_t(Y) == 1, p, 1 - p))

    def conditional_mean(self, F):
        return _probit(_t(F))

    def conditional_variance(self, F):
        p = _probit(_t(F))
        return p - p ** 2

    def predict_mean_and_var(self, Fmu, Fvar):      # probit link: closed form (gpflow Bernoulli.predict_mean_and_var)
        p = _probit(_t(Fmu) / torch.sqrt(1 + _t(Fvar)))
        return p, p - p ** 2

    def predict_density(self, Fmu, Fvar, Y):
        p = self.predict_mean_and_var(Fmu, Fvar)[0]
        return torch.log(torch.where(_t(Y) == 1, p, 1 - p))


def mvhermgauss(H, D):
    gh_x, gh_w = np.polynomial.hermite.hermgauss(H)
    x = np.array(list(itertools.product(*(gh_x,) * D)))
    w = np.prod(np.array(list(itertools.product(*(gh_w,) * D))), 1)
    return x, w


def _unavailable(name):
    def f(*a, **kw):
        raise NotImplementedError(f"gpflow.{name} is outside the shim (not on the SVGP_Layer / DGP path)")
    f.__name__ = name
    return f


def install():
    """register the stand-ins as `tensorflow` and `gpflow` (+ the submodules the reference imports) in sys.modules"""
    tf = _build_tf()
    g = types.ModuleType("gpflow")
    g.__version__ = "1.1-shim"
    g.settings, g.params_as_tensors, g.autoflow = settings, params_as_tensors, autoflow
    g.Parameterized, g.ParamList, g.Param, g.Parameter = Parameterized, ParamList, Parameter, Parameter

    def sub(name, **attrs):
        m = types.ModuleType("gpflow." + name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules["gpflow." + name] = m
        parent, _, leaf = name.rpartition(".")
        setattr(sys.modules["gpflow." + parent] if parent else g, leaf, m)
        return m

    sys.modules["tensorflow"], sys.modules["gpflow"] = tf, g
    sub("params", Parameter=Parameter, Parameterized=Parameterized, ParamList=ParamList, DataHolder=DataHolder, Minibatch=Minibatch)
    sub("models", Model=Model)
    sub("models.model", Model=Model)
    sub("models.gplvm", BayesianGPLVM=type("BayesianGPLVM", (Model,), {}))
    sub("mean_functions", Zero=Zero, Identity=Identity, Linear=Linear, MeanFunction=MeanFunction)
    sub("quadrature", mvhermgauss=mvhermgauss)
    sub("likelihoods", Gaussian=Gaussian, Bernoulli=Bernoulli, Likelihood=Likelihood)
    sub("kernels", RBF=RBF, Matern52=Matern52, White=White, Sum=Sum, Stationary=Stationary)
    sub("features", InducingPoints=InducingPoints)
    sub("transforms", LowerTriangular=LowerTriangular, Transform=Transform, positive=Transform())
    sub("conditionals", conditional=_unavailable("conditionals.conditional"))
    sub("kullback_leiblers", gauss_kl=_unavailable("kullback_leiblers.gauss_kl"))
    sub("priors", Gaussian=type("GaussianPrior", (), {"__init__": lambda self, *a, **k: None}))
    sub("expectations", expectation=_unavailable("expectations.expectation"))
    sub("probability_distributions", DiagonalGaussian=type("DiagonalGaussian", (), {"__init__": lambda self, *a, **k: None}))
    sub("logdensities", multivariate_normal=_unavailable("logdensities.multivariate_normal"))
    sys.modules["gpflow.settings"] = settings
    return tf, g
