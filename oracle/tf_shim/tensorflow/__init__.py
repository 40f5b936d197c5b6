"""numpy emulation of the TensorFlow / Keras calls made by the reference's SRF model files.
TEST INFRASTRUCTURE ONLY (oracle/): it lets tests/golden/make_golden.py execute the
reference's own, unmodified tfsr/model/sequence_router_naive.py (and _einsum.py) in a
container where TensorFlow cannot be installed.

Only what those files (plus tfsr/model/sequence_router.py and tfsr/helper/model_helper.py)
touch is implemented, with TF semantics:
  tf.matmul (batched, transpose_a), tf.tile, tf.reshape, tf.concat, tf.transpose, tf.squeeze,
  tf.expand_dims, tf.shape, tf.zeros/ones/constant, tf.less/add/multiply/square/sqrt,
  tf.reduce_sum, tf.nn.softmax, tf.einsum, tf.sequence_mask, tf.while_loop, tf.TensorArray,
  tf.random.normal, tf.Variable, tf.function, tf.cast, tf.math.{maximum,ceil,reduce_max},
  keras: Model, layers.{Layer, Dense, Conv2D (SAME padding, strides), Dropout,
  BatchNormalization (inference), LayerNormalization (eps 1e-3), Lambda, Masking,
  ZeroPadding2D}, initializers.{VarianceScaling, RandomUniform}.

Tensors are numpy arrays.  The float type is switchable (set_float) so that the same
reference code can be run in float64 (structural pin) and float32 (TF's dtype).
"""
import math as _math
import types as _types

import numpy as _np

_FLOAT = _np.float32
_RNG = _np.random.default_rng(0)


def set_float(dtype):
  global _FLOAT, float32
  _FLOAT = _np.dtype(dtype).type
  float32 = _FLOAT


def set_seed(seed):
  global _RNG
  _RNG = _np.random.default_rng(seed)


float32 = _FLOAT
int32 = _np.int32
bool = _np.bool_  # pylint: disable=redefined-builtin
newaxis = None
dtypes = _types.SimpleNamespace(int32=_np.int32, float32=_np.float32)


def _dt(dtype):
  if dtype is None or dtype in (_np.float32, _np.float64):
    return _FLOAT
  return dtype


def _arr(x):
  return x.value if isinstance(x, _Variable) else _np.asarray(x)


class _Variable:
  """tf.Variable: the reference only reads variables (tf.shape, tf.tile, tf.matmul, +)."""

  def __init__(self, initial_value, trainable=True, name=None, dtype=None):
    self.value = _np.array(_arr(initial_value), dtype=_dt(dtype) if dtype else None)
    self.trainable, self.name = trainable, name

  @property
  def shape(self):
    return self.value.shape

  def numpy(self):
    return self.value

  def assign(self, v):
    self.value = _np.array(v, dtype=self.value.dtype)

  def __array__(self, dtype=None, copy=None):
    return self.value if dtype is None else self.value.astype(dtype)

  def __add__(self, o):
    return self.value + _arr(o)

  __radd__ = __add__

  def __mul__(self, o):
    return self.value * _arr(o)

  __rmul__ = __mul__

  def __getitem__(self, k):
    return self.value[k]


Variable = _Variable


def function(fn=None, **_):
  if fn is None:
    return lambda f: f
  return fn


def shape(x):
  return _np.array(_arr(x).shape, dtype=_np.int64)


def reshape(x, shp, name=None):
  return _np.reshape(_arr(x), [int(s) for s in shp])


def expand_dims(x, axis):
  return _np.expand_dims(_arr(x), axis)


def squeeze(x, axis=None):
  return _np.squeeze(_arr(x), axis=tuple(axis) if isinstance(axis, (list, tuple)) else axis)


def concat(values, axis):
  return _np.concatenate([_arr(v) for v in values], axis=axis)


def tile(x, multiples):
  return _np.tile(_arr(x), [int(m) for m in multiples])


def transpose(x, perm=None):
  return _np.transpose(_arr(x), perm)


def zeros(shp, dtype=None):
  return _np.zeros([int(s) for s in shp], dtype=_dt(dtype))


def ones(shp, dtype=None):
  return _np.ones([int(s) for s in shp], dtype=_dt(dtype))


def constant(v, dtype=None):
  return _np.array(v) if dtype is None else _np.array(v, dtype=_dt(dtype))


def cast(x, dtype):
  return _arr(x).astype(_dt(dtype))


def less(a, b):
  return _np.less(a, b)


def add(a, b):
  return _np.add(a, b)


def multiply(a, b):
  return _np.multiply(_arr(a), _arr(b))


def square(x):
  return _np.square(_arr(x))


def sqrt(x):
  return _np.sqrt(_arr(x))


def exp(x):
  return _np.exp(_arr(x))


def sin(x):
  return _np.sin(_arr(x))


def cos(x):
  return _np.cos(_arr(x))


def range(*a):  # pylint: disable=redefined-builtin
  return _np.arange(*a)


def reduce_sum(x, axis=None, keepdims=False):
  return _np.sum(_arr(x), axis=axis, keepdims=keepdims)


def matmul(a, b, transpose_a=False, transpose_b=False):
  a, b = _arr(a), _arr(b)
  if transpose_a:
    a = _np.swapaxes(a, -1, -2)
  if transpose_b:
    b = _np.swapaxes(b, -1, -2)
  return _np.matmul(a, b)


def einsum(eq, *ops):
  return _np.einsum(eq.replace(" ", ""), *[_arr(o) for o in ops])


def sequence_mask(lengths, maxlen=None, dtype=_np.bool_):
  lengths = _np.asarray(lengths).astype(_np.int64)
  if maxlen is None:
    maxlen = int(lengths.max())
  return (_np.arange(maxlen)[None, :] < lengths[..., None]).astype(_dt(dtype))


def while_loop(cond, body, loop_vars, **_):
  loop_vars = list(loop_vars)
  while cond(*loop_vars):
    loop_vars = list(body(*loop_vars))
  return loop_vars


class TensorArray:
  """tf.TensorArray(dtype, size, dynamic_size): write(idx, v) returns the array;
  concat() concatenates the elements along axis 0."""

  def __init__(self, dtype=None, size=0, dynamic_size=False, infer_shape=True, **_):
    self._items = {}

  def write(self, index, value):
    self._items[int(index)] = _arr(value)
    return self

  def concat(self):
    return _np.concatenate([self._items[k] for k in sorted(self._items)], axis=0)

  def stack(self):
    return _np.stack([self._items[k] for k in sorted(self._items)], axis=0)


def _softmax(x, axis=-1):
  x = _arr(x)
  m = _np.max(x, axis=axis, keepdims=True)
  e = _np.exp(x - m)
  return e / _np.sum(e, axis=axis, keepdims=True)


nn = _types.SimpleNamespace(softmax=_softmax)
math = _types.SimpleNamespace(
    maximum=lambda a, b: _np.maximum(_arr(a), _arr(b)),
    ceil=lambda x: _np.ceil(_arr(x)),
    reduce_max=lambda x, axis=None, keepdims=False: _np.max(_arr(x), axis=axis, keepdims=keepdims),
    equal=lambda a, b: _np.equal(a, b), sqrt=sqrt)
maximum = math.maximum
random = _types.SimpleNamespace(
    normal=lambda shape, mean=0.0, stddev=1.0, dtype=None, seed=None:
    (_RNG.standard_normal([int(s) for s in shape]) * stddev + mean).astype(_dt(dtype)))
linalg = _types.SimpleNamespace(band_part=None)


# --------------------------------------------------------------------------------------
# keras
# --------------------------------------------------------------------------------------
class _VarianceScaling:
  def __init__(self, scale=1.0, mode="fan_in", distribution="truncated_normal", seed=None):
    self.scale, self.mode, self.distribution = scale, mode, distribution

  def __call__(self, shp):
    if len(shp) == 2:
      fan_in, fan_out = shp
    else:
      rf = int(_np.prod(shp[:-2]))
      fan_in, fan_out = shp[-2] * rf, shp[-1] * rf
    n = {"fan_in": fan_in, "fan_out": fan_out, "fan_avg": (fan_in + fan_out) / 2.0}[self.mode]
    limit = _math.sqrt(3.0 * self.scale / n)
    return _RNG.uniform(-limit, limit, size=shp).astype(_FLOAT)


class _RandomUniform:
  def __init__(self, minval=-0.05, maxval=0.05, seed=None):
    self.minval, self.maxval = minval, maxval

  def __call__(self, shp):
    return _RNG.uniform(self.minval, self.maxval, size=shp).astype(_FLOAT)


def _get_init(init):
  if callable(init):
    return init
  return _VarianceScaling(scale=1.0, mode="fan_avg", distribution="uniform")  # glorot_uniform


class _Layer:
  def __init__(self, name=None, **_):
    self.name = name
    self.built = False
    self.last_input = None
    self.last_output = None

  def build(self, input_shape):
    pass

  def __call__(self, inputs, *args, **kwargs):
    if not self.built:
      shp = [_np.shape(i) for i in inputs] if isinstance(inputs, (list, tuple)) else _arr(inputs).shape
      self.build(shp)
      self.built = True
    out = self.call(inputs, *args, **kwargs)
    self.last_input, self.last_output = inputs, out
    return out


class _Model(_Layer):
  def summary(self):
    pass


class _Dense(_Layer):
  def __init__(self, units, activation=None, kernel_initializer=None, name=None, **_):
    super().__init__(name=name)
    self.units, self.init = units, _get_init(kernel_initializer)

  def build(self, input_shape):
    self.kernel = self.init((int(input_shape[-1]), self.units))
    self.bias = _np.zeros(self.units, dtype=_FLOAT)

  def call(self, x, **_):
    return _np.matmul(_arr(x), self.kernel) + self.bias


class _Conv2D(_Layer):
  """NHWC conv, padding='same' with TF's rule: out = ceil(in/stride),
  pad_total = max((out-1)*stride + k - in, 0), pad_before = pad_total // 2."""

  def __init__(self, filters, kernel_size, activation=None, padding="valid", strides=1,
               kernel_initializer=None, name=None, **_):
    super().__init__(name=name)
    self.filters, self.k, self.s = filters, int(kernel_size), int(strides)
    assert padding == "same"
    self.init = _get_init(kernel_initializer)

  def build(self, input_shape):
    self.kernel = self.init((self.k, self.k, int(input_shape[-1]), self.filters))
    self.bias = _np.zeros(self.filters, dtype=_FLOAT)

  def call(self, x, **_):
    x = _arr(x)
    B, H, W, _C = x.shape
    k, s = self.k, self.s
    oh, ow = -(-H // s), -(-W // s)
    ph, pw = max((oh - 1) * s + k - H, 0), max((ow - 1) * s + k - W, 0)
    xp = _np.pad(x, ((0, 0), (ph // 2, ph - ph // 2), (pw // 2, pw - pw // 2), (0, 0)))
    out = _np.zeros((B, oh, ow, self.filters), dtype=x.dtype)
    for di in _np.arange(k):
      for dj in _np.arange(k):
        patch = xp[:, di:di + (oh - 1) * s + 1:s, dj:dj + (ow - 1) * s + 1:s, :]
        out += _np.einsum("bhwc,cf->bhwf", patch, self.kernel[di, dj])
    return out + self.bias


class _Dropout(_Layer):
  def __init__(self, rate, name=None, **_):
    super().__init__(name=name)
    self.rate = rate
    self.mask = None   # optional injected scaled keep mask (training-mode goldens)

  def call(self, x, training=None, **_):
    x = _arr(x)
    if training and self.mask is not None:
      return x * self.mask
    return x


class _BatchNormalization(_Layer):
  def __init__(self, axis=-1, momentum=0.99, epsilon=1e-3, name=None, **_):
    super().__init__(name=name)
    self.eps = epsilon

  def build(self, input_shape):
    c = int(input_shape[-1])
    self.gamma, self.beta = _np.ones(c, _FLOAT), _np.zeros(c, _FLOAT)
    self.moving_mean, self.moving_variance = _np.zeros(c, _FLOAT), _np.ones(c, _FLOAT)

  def call(self, x, **_):
    return (_arr(x) - self.moving_mean) / _np.sqrt(self.moving_variance + self.eps) * self.gamma + self.beta


class _LayerNormalization(_Layer):
  def __init__(self, axis=-1, epsilon=1e-3, name=None, **_):
    super().__init__(name=name)
    self.eps = epsilon

  def build(self, input_shape):
    c = int(input_shape[-1])
    self.gamma, self.beta = _np.ones(c, _FLOAT), _np.zeros(c, _FLOAT)

  def call(self, x, **_):
    x = _arr(x)
    mean = x.mean(axis=-1, keepdims=True)
    var = ((x - mean) ** 2).mean(axis=-1, keepdims=True)
    return (x - mean) / _np.sqrt(var + self.eps) * self.gamma + self.beta


class _Lambda(_Layer):
  def __init__(self, fn, name=None, **_):
    super().__init__(name=name)
    self.fn = fn

  def call(self, x, **_):
    return self.fn(x)


class _Masking(_Layer):
  def __init__(self, mask_value=0.0, name=None, **_):
    super().__init__(name=name)
    self.mask_value = mask_value

  def call(self, x, **_):
    x = _arr(x)
    keep = _np.any(x != self.mask_value, axis=-1, keepdims=True)
    return x * keep.astype(x.dtype)


class _ZeroPadding2D(_Layer):
  def __init__(self, padding=(1, 1), **_):
    super().__init__()
    self.padding = padding

  def call(self, x, **_):
    (t, b), (l, r) = self.padding
    return _np.pad(_arr(x), ((0, 0), (int(t), int(b)), (int(l), int(r)), (0, 0)))


keras = _types.SimpleNamespace(
    Model=_Model,
    layers=_types.SimpleNamespace(
        Layer=_Layer, Dense=_Dense, Conv2D=_Conv2D, Dropout=_Dropout,
        BatchNormalization=_BatchNormalization, LayerNormalization=_LayerNormalization,
        Lambda=_Lambda, Masking=_Masking, ZeroPadding2D=_ZeroPadding2D),
    initializers=_types.SimpleNamespace(VarianceScaling=_VarianceScaling,
                                        RandomUniform=_RandomUniform))
