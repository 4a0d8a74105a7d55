"""Eager numpy stand-in for the TensorFlow 1.3 ops the reference hot path calls.

TEST INFRASTRUCTURE ONLY (part of `oracle/`).  Purpose: the reference
(emtiyaz/vmp-for-svae) is pure Python over `tensorflow==1.3.0`
(environment.yml:33), which cannot be installed in this image (Python 3.12, no
network).  The *algorithm text* of the hot path lives in the reference's own
files (models/svae.py, models/gmm.py, models/smm.py, distributions/*.py,
helpers/tf_utils.py); TensorFlow only supplies the primitive array ops.  This
module restates those primitives (each is a published, unambiguous array
operation: batched LU solve, lower Cholesky, einsum, reductions, digamma, ...)
over numpy/LAPACK so that the UNMODIFIED reference source can be imported from
/root/reference and executed to generate golden vectors
(tests/golden/make_golden.py).  Nothing here is imported by the product path.

Semantics restated (TF 1.3 API name -> numpy):
  matrix_solve -> LAPACK gesv (LU, partial pivoting)   cholesky -> LAPACK potrf (lower)
  matrix_inverse / matrix_determinant -> LAPACK getri / getrf
  einsum -> numpy.einsum (same subscript notation)
  multinomial -> the CPU kernel's algorithm (core/kernels/multinomial_op.cc):
      cdf = cumsum(exp(logits - max)) in double, draw u~U[0,1),
      sample = upper_bound(cdf, u * cdf[-1])
  random_normal / random_uniform -> numpy RandomState(seed) (TF's Philox stream
      is not reproducible; every draw is appended to `rng_log` so the generator
      script can store the injected noise next to the outputs).
Graph-mode features (sessions, queues, summaries, devices, scopes) are no-ops.
All float tensors are computed in `FLOAT` (float64 by default so that golden
vectors are fp64 truth; set_float(np.float32) reproduces the reference's fp32).
"""
import builtins as _builtins
import contextlib

import numpy as np
from scipy import special as _sp

FLOAT = np.float64
rng_log = []          # [(kind, array)] in call order; cleared by the caller

float32 = 'float32'
float64 = 'float64'
int32 = 'int32'
int64 = 'int64'
uint8 = 'uint8'
string = 'string'
bool_ = 'bool'


def set_float(dt):
    global FLOAT
    FLOAT = dt


def _np_dtype(dtype):
    if dtype is None:
        return None
    if dtype in (float32, float64) or dtype in (np.float32, np.float64, float):
        return FLOAT
    if dtype in (int32, np.int32):
        return np.int32
    if dtype in (int64, np.int64, int):
        return np.int64
    if dtype in (bool_, bool, np.bool_):
        return np.bool_
    if dtype == uint8:
        return np.uint8
    return dtype


class Dimension(int):
    @property
    def value(self):
        return int(self)


class TensorShape(object):
    def __init__(self, dims):
        self._dims = _builtins.tuple(Dimension(d) for d in dims)

    def as_list(self):
        return [int(d) for d in self._dims]

    def __iter__(self):
        return iter(self._dims)

    def __len__(self):
        return len(self._dims)

    def __getitem__(self, i):
        return self._dims[i]

    def __eq__(self, other):
        if isinstance(other, TensorShape):
            return self._dims == other._dims
        if isinstance(other, (int, np.integer)):
            return self._dims == (int(other),)
        try:
            return _builtins.tuple(int(d) for d in self._dims) == _builtins.tuple(int(d) for d in other)
        except TypeError:
            return False

    def __ne__(self, other):
        return not self.__eq__(other)

    def __repr__(self):
        return 'TensorShape(%s)' % (list(self._dims),)

    __str__ = __repr__


_scope_stack = []


class Tensor(object):
    __array_priority__ = 1000

    def __init__(self, value, name=None):
        if isinstance(value, Tensor):
            value = value.a
        a = np.asarray(value)
        if a.dtype.kind == 'f' and a.dtype != FLOAT:
            a = a.astype(FLOAT)
        self.a = a
        self.name = name if name is not None else 'Tensor:0'

    # --- shape protocol used by the reference -------------------------------
    def get_shape(self):
        return TensorShape(self.a.shape)

    @property
    def shape(self):
        return TensorShape(self.a.shape)

    @property
    def dtype(self):
        return self.a.dtype

    def __array__(self, dtype=None, copy=None):
        return self.a if dtype is None else self.a.astype(dtype)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        # the reference applies np.multiply / np.divide to tensors (gmm.py:66,76)
        args = [x.a if isinstance(x, Tensor) else x for x in inputs]
        return Tensor(getattr(ufunc, method)(*args, **kwargs))

    def __getitem__(self, idx):
        return Tensor(self.a[idx])

    def __len__(self):
        return len(self.a)

    def assign(self, value, **k):
        self.a[...] = _a(value)
        return self

    def __repr__(self):
        return 'shim.Tensor(%r)' % (self.a,)

    # --- arithmetic ----------------------------------------------------------
    def _b(op):
        def f(self, other):
            return Tensor(op(self.a, _a(other)))
        return f

    def _r(op):
        def f(self, other):
            return Tensor(op(_a(other), self.a))
        return f

    __add__ = _b(np.add); __radd__ = _r(np.add)
    __sub__ = _b(np.subtract); __rsub__ = _r(np.subtract)
    __mul__ = _b(np.multiply); __rmul__ = _r(np.multiply)
    __truediv__ = _b(np.true_divide); __rtruediv__ = _r(np.true_divide)
    __div__ = __truediv__; __rdiv__ = __rtruediv__
    __pow__ = _b(np.power)
    __gt__ = _b(np.greater); __lt__ = _b(np.less)
    __ge__ = _b(np.greater_equal); __le__ = _b(np.less_equal)

    def __neg__(self):
        return Tensor(-self.a)

    __hash__ = object.__hash__


def _a(x):
    """to ndarray (python floats become FLOAT scalars)."""
    if isinstance(x, Tensor):
        return x.a
    a = np.asarray(x)
    if a.dtype.kind == 'f' and a.dtype != FLOAT:
        a = a.astype(FLOAT)
    return a


def _t(x, name=None):
    return Tensor(x, name=name)


# ----------------------------------------------------------------------------------------------
# scopes / graph plumbing (no-ops)
@contextlib.contextmanager
def name_scope(name=None, *a, **k):
    yield name


@contextlib.contextmanager
def device(name=None):
    yield


class _VarScope(object):
    def __init__(self, name):
        self.name = name

    def reuse_variables(self):
        pass


_variables = {}


@contextlib.contextmanager
def variable_scope(name_or_scope=None, *a, **k):
    nm = name_or_scope.name if isinstance(name_or_scope, _VarScope) else (name_or_scope or '')
    _scope_stack.append(nm)
    try:
        yield _VarScope(nm)
    finally:
        _scope_stack.pop()


def get_variable_scope():
    return _VarScope('/'.join(s for s in _scope_stack if s))


def reset_default_graph():
    _variables.clear()
    del _scope_stack[:]


def get_variable(name, shape=None, initializer=None, trainable=True, dtype=None, **k):
    full = '/'.join([s for s in _scope_stack if s] + [name])
    if full in _variables:
        return _variables[full]
    if callable(initializer) and not isinstance(initializer, Tensor):
        init = initializer(shape)
    else:
        init = initializer
    v = Tensor(np.array(_a(init), copy=True), name=full + ':0')
    v.trainable = trainable
    _variables[full] = v
    return v


def Variable(initial_value, dtype=None, name=None, trainable=True, **k):
    v = Tensor(np.array(_a(initial_value), copy=True), name=(name or 'Variable') + ':0')
    v.trainable = trainable
    return v


def assign(ref, value, name=None, **k):
    ref.a[...] = _a(value)
    return ref


def group(*a, **k):
    return None


def tuple(tensors, name=None, **k):   # noqa: A001  (mirrors tf.tuple)
    return [t if isinstance(t, Tensor) else _t(t) for t in tensors]


def identity(x, name=None):
    return _t(_a(x))


def stop_gradient(x, name=None):
    return _t(_a(x))


def set_random_seed(seed):
    pass


# ----------------------------------------------------------------------------------------------
# constructors
def constant(value, dtype=None, shape=None, name=None):
    a = np.asarray(_a(value), dtype=_np_dtype(dtype)) if dtype is not None else _a(value)
    if shape is not None:
        a = np.broadcast_to(a, shape).copy()
    return _t(a)


def convert_to_tensor(value, dtype=None, name=None):
    return constant(value, dtype=dtype)


def ones(shape, dtype=float32, name=None):
    return _t(np.ones(_shape(shape), dtype=_np_dtype(dtype)))


def zeros(shape, dtype=float32, name=None):
    return _t(np.zeros(_shape(shape), dtype=_np_dtype(dtype)))


def ones_like(x, dtype=None, name=None):
    return _t(np.ones_like(_a(x), dtype=_np_dtype(dtype)))


def zeros_like(x, dtype=None, name=None):
    return _t(np.zeros_like(_a(x), dtype=_np_dtype(dtype)))


def eye(num_rows, num_columns=None, batch_shape=None, dtype=float32, name=None):
    return _t(np.eye(int(num_rows), None if num_columns is None else int(num_columns), dtype=_np_dtype(dtype)))


def range(start, limit=None, delta=1, dtype=None, name=None):   # noqa: A001
    if limit is None:
        start, limit = 0, start
    a = np.arange(float(_a(start)), float(_a(limit)), delta)
    if dtype is None:
        dtype = int32 if all(float(v).is_integer() for v in (float(_a(start)), float(_a(limit)), delta)) else float32
    return _t(a.astype(_np_dtype(dtype)))


def _shape(shape):
    if isinstance(shape, TensorShape):
        return builtins_tuple(shape.as_list())
    if isinstance(shape, Tensor):
        return builtins_tuple(int(v) for v in shape.a)
    if isinstance(shape, (int, np.integer)):
        return (int(shape),)
    return builtins_tuple(int(s) for s in shape)


builtins_tuple = _builtins.tuple


def random_normal(shape, mean=0.0, stddev=1.0, dtype=float32, seed=None, name=None):
    rs = np.random.RandomState(seed)
    z = rs.standard_normal(_shape(shape)).astype(FLOAT)
    rng_log.append(('random_normal', z.copy()))
    return _t(mean + stddev * z)


def random_uniform(shape, minval=0, maxval=None, dtype=float32, seed=None, name=None):
    if maxval is None:
        maxval = 1
    rs = np.random.RandomState(seed)
    u = rs.random_sample(_shape(shape)).astype(FLOAT)
    rng_log.append(('random_uniform', u.copy()))
    return _t(minval + (maxval - minval) * u)


def random_normal_initializer(mean=0.0, stddev=1.0, seed=None, dtype=float32):
    def init(shape):
        return random_normal(shape, mean, stddev, seed=seed)
    return init


def constant_initializer(value=0, dtype=float32):
    def init(shape):
        return _t(np.full(_shape(shape if shape is not None else ()), value, dtype=FLOAT))
    return init


def multinomial(logits, num_samples, seed=None, name=None):
    """CPU-kernel algorithm of tf.multinomial (multinomial_op.cc): unnormalised
    cdf of exp(logit - max) accumulated in double; sample = upper_bound(cdf, u*total)."""
    lg = np.asarray(_a(logits), dtype=np.float64)
    n, k = lg.shape
    rs = np.random.RandomState(seed)
    u = rs.random_sample((n, int(num_samples)))
    rng_log.append(('multinomial_uniform', u.copy()))
    mx = np.max(np.where(np.isfinite(lg), lg, -np.inf), axis=1, keepdims=True)
    cdf = np.cumsum(np.where(np.isfinite(lg), np.exp(lg - mx), 0.0), axis=1)
    out = np.empty((n, int(num_samples)), dtype=np.int64)
    for i in _builtins.range(n):
        out[i] = np.searchsorted(cdf[i], u[i] * cdf[i, -1], side='right')
    return _t(np.minimum(out, k - 1))


# ----------------------------------------------------------------------------------------------
# elementwise
def add(x, y, name=None):
    return _t(_a(x) + _a(y))


def subtract(x, y, name=None):
    return _t(_a(x) - _a(y))


def multiply(x, y, name=None):
    return _t(_a(x) * _a(y))


def divide(x, y, name=None):
    return _t(_a(x) / _a(y))


def pow(x, y, name=None):   # noqa: A001
    return _t(np.power(_a(x), _a(y)))


def square(x, name=None):
    return _t(np.square(_a(x)))


def sqrt(x, name=None):
    return _t(np.sqrt(_a(x)))


def exp(x, name=None):
    return _t(np.exp(_a(x)))


def log(x, name=None):
    with np.errstate(divide='ignore', invalid='ignore'):
        return _t(np.log(_a(x)))


def log1p(x, name=None):
    return _t(np.log1p(_a(x)))


def tanh(x, name=None):
    return _t(np.tanh(_a(x)))


def digamma(x, name=None):
    return _t(_sp.digamma(_a(x)))


def lgamma(x, name=None):
    return _t(_sp.gammaln(_a(x)))


def is_nan(x, name=None):
    return _t(np.isnan(_a(x)))


def equal(x, y, name=None):
    return _t(np.equal(_a(x), _a(y)))


def logical_not(x, name=None):
    return _t(np.logical_not(_a(x)))


def where(condition, x=None, y=None, name=None):
    return _t(np.where(_a(condition), _a(x), _a(y)))


def cast(x, dtype, name=None):
    return _t(_a(x).astype(_np_dtype(dtype)))


def to_float(x, name=None):
    return _t(_a(x).astype(FLOAT))


def to_int32(x, name=None):
    return _t(_a(x).astype(np.int32))


def argmax(x, axis=None, name=None, **k):
    return _t(np.argmax(_a(x), axis=axis))


def one_hot(indices, depth, dtype=float32, **k):
    return _t(np.eye(int(depth), dtype=_np_dtype(dtype))[_a(indices)])


# ----------------------------------------------------------------------------------------------
# reductions (TF 1.3 spelling: keep_dims)
def _red(fn):
    def f(x, axis=None, keep_dims=False, name=None, keepdims=None, reduction_indices=None):
        if keepdims is not None:
            keep_dims = keepdims
        if axis is None:
            axis = reduction_indices
        return _t(fn(_a(x), axis=axis if axis is None or isinstance(axis, int) else builtins_tuple(axis),
                     keepdims=keep_dims))
    return f


reduce_sum = _red(np.sum)
reduce_mean = _red(np.mean)
reduce_max = _red(np.max)
reduce_min = _red(np.min)


def reduce_logsumexp(x, axis=None, keep_dims=False, name=None):
    return _t(_sp.logsumexp(_a(x), axis=axis, keepdims=keep_dims))


# ----------------------------------------------------------------------------------------------
# shape ops
def expand_dims(x, axis=None, name=None, dim=None):
    return _t(np.expand_dims(_a(x), axis if axis is not None else dim))


def reshape(x, shape, name=None):
    return _t(np.reshape(_a(x), _shape(shape)))


def tile(x, multiples, name=None):
    return _t(np.tile(_a(x), _shape(multiples)))


def transpose(x, perm=None, name=None):
    return _t(np.transpose(_a(x), perm))


def concat(values, axis, name=None):
    return _t(np.concatenate([_a(v) for v in values], axis=axis))


def split(value, num_or_size_splits, axis=0, name=None):
    return [_t(p) for p in np.split(_a(value), num_or_size_splits, axis=axis)]


def gather_nd(params, indices, name=None):
    p, idx = _a(params), _a(indices)
    return _t(p[builtins_tuple(idx[..., i] for i in _builtins.range(idx.shape[-1]))])


# ----------------------------------------------------------------------------------------------
# linear algebra
def matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    a, b = _a(a), _a(b)
    if transpose_a:
        a = np.swapaxes(a, -1, -2)
    if transpose_b:
        b = np.swapaxes(b, -1, -2)
    return _t(np.matmul(a, b))


def einsum(equation, *inputs):
    return _t(np.einsum(equation, *[_a(x) for x in inputs]))


def matrix_transpose(x, name=None):
    return _t(np.swapaxes(_a(x), -1, -2))


def matrix_diag(diagonal, name=None):
    d = _a(diagonal)
    out = np.zeros(d.shape + (d.shape[-1],), dtype=d.dtype)
    i = np.arange(d.shape[-1])
    out[..., i, i] = d
    return _t(out)


def matrix_diag_part(x, name=None):
    return _t(np.diagonal(_a(x), axis1=-2, axis2=-1).copy())


def matrix_set_diag(x, diagonal, name=None):
    out = np.array(_a(x), copy=True)
    i = np.arange(out.shape[-1])
    out[..., i, i] = _a(diagonal)
    return _t(out)


def matrix_solve(matrix, rhs, adjoint=False, name=None):
    m = _a(matrix)
    if adjoint:
        m = np.swapaxes(m, -1, -2)
    return _t(np.linalg.solve(m, _a(rhs)))


def matrix_inverse(x, adjoint=False, name=None):
    return _t(np.linalg.inv(_a(x)))


def matrix_determinant(x, name=None):
    return _t(np.linalg.det(_a(x)))


def cholesky(x, name=None):
    return _t(np.linalg.cholesky(_a(x)))


# ----------------------------------------------------------------------------------------------
# sub-namespaces
class _NS(object):
    pass


nn = _NS()
nn.softplus = lambda x, name=None: _t(np.logaddexp(0.0, _a(x)))
nn.softmax = lambda x, dim=-1, name=None: _t(_sp.softmax(_a(x), axis=dim))
nn.sigmoid = lambda x, name=None: _t(_sp.expit(_a(x)))
nn.tanh = tanh


class _TriL(object):
    def __init__(self, tril, name=None):
        self._m = _a(tril)

    def to_dense(self):
        return _t(np.tril(self._m))


class _Dirichlet(object):
    def __init__(self, concentration):
        self._c = _a(concentration)

    def sample(self, n, seed=None):
        rs = np.random.RandomState(seed)
        s = rs.dirichlet(self._c, size=int(n)).astype(FLOAT)
        rng_log.append(('dirichlet', s.copy()))
        return _t(s)


contrib = _NS()
contrib.linalg = _NS()
contrib.linalg.LinearOperatorTriL = _TriL
contrib.distributions = _NS()
contrib.distributions.Dirichlet = _Dirichlet


class _Layers(object):
    @staticmethod
    def dense(inputs, units, activation=None, kernel_initializer=None, bias_initializer=None, name=None, **k):
        x = _a(inputs)
        with variable_scope(name or 'dense'):
            w = get_variable('kernel', (x.shape[-1], int(units)), initializer=kernel_initializer)
            b = get_variable('bias', (int(units),), initializer=bias_initializer)
        y = _t(np.matmul(x, w.a) + b.a)
        return activation(y) if activation is not None else y


layers = _Layers()


class _Train(object):
    @staticmethod
    def exponential_decay(learning_rate, global_step, decay_steps, decay_rate, staircase=False, name=None):
        p = float(_a(global_step)) / float(decay_steps)
        if staircase:
            p = np.floor(p)
        return _t(np.asarray(learning_rate * decay_rate ** p, dtype=FLOAT))


train = _Train()


class _Summary(object):
    def __getattr__(self, item):
        return lambda *a, **k: None


summary = _Summary()


def get_default_graph():
    return None
