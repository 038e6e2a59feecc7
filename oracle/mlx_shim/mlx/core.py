"""NumPy-backed stand-in for `mlx.core` — oracle/test infrastructure only.

Implements, with MLX's dtype rules (float32 / complex64 / int32 defaults, weak Python
scalars == NumPy >= 2 promotion), exactly the calls made by the reference files on the
STFT / log-mel / iSTFT hot path.  CPU FFT = numpy pocketfft in single precision (MLX's CPU
FFT is pocketfft as well).
"""
import numpy as _np

float32 = _np.float32
float16 = _np.float16
int32 = _np.int32
int64 = _np.int64
complex64 = _np.complex64
bool_ = _np.bool_
pi = _np.pi


class _At:
    def __init__(self, arr):
        self._arr = arr

    def __getitem__(self, idx):
        arr = self._arr

        class _Upd:
            def add(self, values):
                out = _np.array(arr, copy=True)
                _np.add.at(out, _np.asarray(idx), _np.asarray(values, dtype=out.dtype))
                return out.view(array)

        return _Upd()


class array(_np.ndarray):
    def __new__(cls, value, dtype=None):
        a = _np.asarray(value)
        if dtype is None:
            if a.dtype == _np.float64:
                dtype = _np.float32
            elif a.dtype == _np.int64:
                dtype = _np.int32
            elif a.dtype == _np.complex128:
                dtype = _np.complex64
            else:
                dtype = a.dtype
        return _np.array(a, dtype=dtype).view(cls)

    # MLX promotes `integer array (op) python float` to float32; NumPy would go to float64.  (Needed by the Kaldi
    # mel banks, dsp.py:558-560: `mel_low_freq + bin_idx * mel_freq_delta` with an int32 bin_idx.)
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method == "__call__" and any(isinstance(v, float) for v in inputs):
            inputs = tuple(_np.asarray(v, dtype=_np.float32) if isinstance(v, _np.ndarray) and v.dtype.kind in "iu" else v
                           for v in inputs)
        inputs = tuple(_np.asarray(v) if isinstance(v, array) else v for v in inputs)
        if "out" in kwargs:
            kwargs["out"] = tuple(_np.asarray(v) if isinstance(v, array) else v for v in kwargs["out"])
        r = getattr(ufunc, method)(*inputs, **kwargs)
        return r.view(array) if isinstance(r, _np.ndarray) else r

    # mx.array method surface used by the reference
    def abs(self):
        return _np.abs(self).view(array)

    def square(self):
        return _np.square(self).view(array)

    def log(self):
        return _np.log(self).view(array)

    def log10(self):
        return _np.log10(self).view(array)

    def exp(self):
        return _np.exp(self).view(array)

    def sqrt(self):
        return _np.sqrt(self).view(array)

    def moveaxis(self, s, d):
        return _np.moveaxis(self, s, d).view(array)

    def split(self, n, axis=0):
        return [p.view(array) for p in _np.split(self, n, axis=axis)]

    @property
    def at(self):
        return _At(self)


def _w(x):
    return _np.asarray(x).view(array) if isinstance(x, _np.ndarray) else x


def zeros(shape, dtype=float32):
    return _np.zeros(shape, dtype=dtype).view(array)


def ones(shape, dtype=float32):
    return _np.ones(shape, dtype=dtype).view(array)


def zeros_like(x):
    return _np.zeros_like(x).view(array)


def arange(*args, dtype=None):
    a = _np.arange(*args)
    if dtype is None:
        dtype = _np.int32 if a.dtype.kind in "iu" else _np.float32
    return a.astype(dtype).view(array)


def linspace(start, stop, num=50, dtype=float32):
    if num == 1:
        return _np.array([start], dtype=dtype).view(array)
    t = _np.arange(num, dtype=_np.float32) / _np.float32(num - 1)
    r = (_np.float32(1) - t) * _np.float32(start) + t * _np.float32(stop)
    return r.astype(dtype).view(array)


def pad(x, pad_width, mode="constant", constant_values=0):
    x = _np.asarray(x)
    if isinstance(pad_width, int):
        pad_width = [(pad_width, pad_width)] * x.ndim
    elif isinstance(pad_width, tuple) and len(pad_width) == 2 and all(isinstance(p, int) for p in pad_width):
        pad_width = [tuple(pad_width)] * x.ndim
    return _np.pad(x, pad_width, mode="constant", constant_values=constant_values).astype(x.dtype).view(array)


def concatenate(arrs, axis=0):
    return _np.concatenate([_np.asarray(a) for a in arrs], axis=axis).view(array)


concat = concatenate


def stack(arrs, axis=0):
    return _np.stack([_np.asarray(a) for a in arrs], axis=axis).view(array)


def as_strided(x, shape, strides, offset=0):
    x = _np.ascontiguousarray(x)
    item = x.itemsize
    flat = x.reshape(-1)[offset:]
    return _np.lib.stride_tricks.as_strided(flat, shape=shape, strides=tuple(s * item for s in strides)).copy().view(array)


def where(c, a, b):
    with _np.errstate(all="ignore"):
        return _w(_np.where(c, a, b))


def _unary(fn):
    def f(x, *a, **k):
        with _np.errstate(all="ignore"):
            return _w(fn(x, *a, **k))

    return f


exp = _unary(_np.exp)
log = _unary(_np.log)
log10 = _unary(_np.log10)
sqrt = _unary(_np.sqrt)
cos = _unary(_np.cos)
sin = _unary(_np.sin)
abs = _unary(_np.abs)
square = _unary(_np.square)
floor = _unary(_np.floor)
imag = _unary(_np.imag)
real = _unary(_np.real)
maximum = _unary(_np.maximum)
minimum = _unary(_np.minimum)
arctan2 = _unary(_np.arctan2)
power = _unary(_np.power)
matmul = _unary(_np.matmul)
tile = _unary(_np.tile)
expand_dims = _unary(_np.expand_dims)
squeeze = _unary(_np.squeeze)
swapaxes = _unary(_np.swapaxes)
broadcast_to = _unary(_np.broadcast_to)


def take(x, indices, axis=None):
    return _w(_np.take(_np.asarray(x), _np.asarray(indices), axis=axis))


def reshape(x, shape):
    return _w(_np.reshape(x, shape))


def repeat(x, repeats, axis=None):
    return _w(_np.repeat(_np.asarray(x), repeats, axis=axis))


def transpose(x, axes=None):
    return _w(_np.transpose(x, axes))


def clip(x, a_min, a_max):
    return _w(_np.clip(x, a_min, a_max))


def mean(x, axis=None, keepdims=False):
    return _w(_np.mean(x, axis=axis, keepdims=keepdims, dtype=_np.asarray(x).dtype))


def sum(x, axis=None, keepdims=False):
    return _w(_np.sum(x, axis=axis, keepdims=keepdims, dtype=_np.asarray(x).dtype))


def std(x, axis=None, keepdims=False, ddof=0):
    return _w(_np.std(x, axis=axis, keepdims=keepdims, ddof=ddof, dtype=_np.asarray(x).dtype))


def max(x, axis=None, keepdims=False):
    return _w(_np.max(x, axis=axis, keepdims=keepdims))


def cumsum(x, axis=None):
    return _w(_np.cumsum(x, axis=axis, dtype=_np.asarray(x).dtype))


def eval(*a, **k):
    return None


class fft:
    @staticmethod
    def rfft(x, n=None, axis=-1):
        return _np.fft.rfft(_np.asarray(x), n=n, axis=axis).astype(_np.complex64).view(array)

    @staticmethod
    def fft(x, n=None, axis=-1):
        return _np.fft.fft(_np.asarray(x).astype(_np.complex64), n=n, axis=axis).astype(_np.complex64).view(array)

    @staticmethod
    def ifft(x, n=None, axis=-1):
        return _np.fft.ifft(_np.asarray(x).astype(_np.complex64), n=n, axis=axis).astype(_np.complex64).view(array)

    @staticmethod
    def irfft(x, n=None, axis=-1):
        return _np.fft.irfft(_np.asarray(x, dtype=_np.complex64), n=n, axis=axis).astype(_np.float32).view(array)

Dtype = type(_np.dtype("float32"))
bfloat16 = _np.float32  # no bf16 in NumPy; goldens only use float32 input
