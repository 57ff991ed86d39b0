"""TEST INFRASTRUCTURE ONLY -- a minimal eager stand-in for the TensorFlow 1.x ops
that the reference's hot-path modules use (dpc/util/point_cloud.py, drc.py,
gauss_kernel.py, quaternion.py, camera.py, point_cloud_distance.py).

Why it exists: TensorFlow cannot be installed in this image (py3.12, no wheel, no
network), so the reference's Python source cannot run on its own runtime.  With this
module first on sys.path, `import tensorflow as tf` inside the UNMODIFIED reference
files resolves here and every `tf.*` call is executed eagerly, one op at a time, in
fp32 on torch-CPU.  That pins what the oracle restatement must reproduce -- the
reference's op order, axis conventions, flips and quirks -- to the reference's own
source text.  What it cannot pin is TensorFlow's kernel-internal rounding (Eigen
reduction order, conv accumulation order); see DESIGN.md "parity status".

Semantics implemented from the public TF 1.x op documentation:
  * tensors are immutable values: `x += y` rebinds (no __iadd__ defined);
  * every op rounds to fp32 on its own (no cross-op fusion);
  * tf.clip_by_value / tf.reduce_max gradients follow TF (inclusive clip mask, ties
    split evenly) because torch.clamp / torch.amax have the same conventions;
  * tf.nn.conv3d is NDHWC cross-correlation, "SAME" = zero pad (k-1)//2 low, rest high;
  * tf.scatter_nd sums duplicate indices;
  * tf.norm over the last axis sums squares left to right (documented choice).

Nothing in the product path may import this package (only tests/ and the golden
generator do).
"""
import builtins

import numpy as np
import torch
import torch.nn.functional as F

float32 = torch.float32
float64 = torch.float64
int32 = torch.int32
int64 = torch.int64
bool = torch.bool  # noqa: A001  (tf.bool)


class TensorShape(tuple):
    def as_list(self):
        return list(self)


class Tensor:
    """Immutable-value wrapper around a torch tensor (`.t`)."""

    __array_priority__ = 1000

    def __init__(self, t):
        assert isinstance(t, torch.Tensor)
        self.t = t

    # -- TF-like attributes
    @property
    def shape(self):
        return TensorShape(int(s) for s in self.t.shape)

    @property
    def dtype(self):
        return self.t.dtype

    def get_shape(self):
        return self.shape

    def numpy(self):
        return self.t.detach().numpy()

    def __len__(self):
        return self.t.shape[0]

    def __iter__(self):
        for i in builtins.range(self.t.shape[0]):
            yield Tensor(self.t[i])

    def __getitem__(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        key = tuple(_raw(k) if isinstance(k, Tensor) else k for k in key)
        return Tensor(self.t[key])

    # -- arithmetic: out of place only (no __i*__ => `a += b` rebinds like TF)
    def _bin(self, other, fn, rev=False):
        a, b = self.t, _raw(other, like=self.t)
        return Tensor(fn(b, a) if rev else fn(a, b))

    def __add__(self, o): return self._bin(o, torch.add)
    def __radd__(self, o): return self._bin(o, torch.add, True)
    def __sub__(self, o): return self._bin(o, torch.sub)
    def __rsub__(self, o): return self._bin(o, torch.sub, True)
    def __mul__(self, o): return self._bin(o, torch.mul)
    def __rmul__(self, o): return self._bin(o, torch.mul, True)
    def __truediv__(self, o): return self._bin(o, torch.div)
    def __rtruediv__(self, o): return self._bin(o, torch.div, True)
    def __pow__(self, o): return self._bin(o, torch.pow)
    def __neg__(self): return Tensor(torch.neg(self.t))
    def __ge__(self, o): return self._bin(o, torch.ge)
    def __le__(self, o): return self._bin(o, torch.le)
    def __gt__(self, o): return self._bin(o, torch.gt)
    def __lt__(self, o): return self._bin(o, torch.lt)

    def __repr__(self):
        return "tf1_shim.Tensor(%r)" % (self.t,)


def _raw(x, like=None, dtype=None):
    """torch tensor (or python scalar) from anything tensor-like."""
    if isinstance(x, Tensor):
        return x.t
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, (builtins.bool, int, float)):
        if like is not None and dtype is None:
            return x  # torch handles python scalars with the tensor's dtype
        return torch.tensor(x, dtype=dtype if dtype is not None else _infer(x))
    arr = np.asarray(x)
    if dtype is None:
        if arr.dtype == np.float64:
            dtype = torch.float32  # TF converts python floats to float32
        elif arr.dtype == np.int64:
            dtype = torch.int32  # and python ints to int32
    t = torch.from_numpy(np.ascontiguousarray(arr))
    return t.to(dtype) if dtype is not None else t


def _infer(x):
    if isinstance(x, builtins.bool):
        return torch.bool
    if isinstance(x, int):
        return torch.int32
    return torch.float32


def _t(x, dtype=None):
    r = _raw(x, dtype=dtype)
    if not isinstance(r, torch.Tensor):
        r = torch.tensor(r, dtype=dtype if dtype is not None else _infer(r))
    return r


def _int(x):
    if isinstance(x, Tensor):
        return int(x.t.item())
    if isinstance(x, torch.Tensor):
        return int(x.item())
    return int(x)


def _axes(axis):
    if axis is None:
        return None
    if isinstance(axis, (list, tuple)):
        return [int(a) for a in axis]
    return int(axis)


# ---------------------------------------------------------------- creation
def constant(value, dtype=None, shape=None):
    t = _t(value, dtype=dtype)
    if shape is not None:
        shape = [int(s) for s in shape]
        t = t.reshape(-1)
        n = int(np.prod(shape)) if len(shape) else 1
        if t.numel() == 1 and n != 1:
            t = t.repeat(n)
        t = t.reshape(shape)
    return Tensor(t)


def convert_to_tensor(value, dtype=None):
    if isinstance(value, Tensor) and dtype is None:
        return value
    return Tensor(_t(value, dtype=dtype))


def ones(shape, dtype=float32):
    return Tensor(torch.ones([_int(s) for s in shape], dtype=dtype))


def zeros(shape, dtype=float32):
    return Tensor(torch.zeros([_int(s) for s in shape], dtype=dtype))


def ones_like(x, dtype=None):
    return Tensor(torch.ones_like(_t(x), dtype=dtype))


def zeros_like(x, dtype=None):
    return Tensor(torch.zeros_like(_t(x), dtype=dtype))


def range(start, limit=None, delta=1, dtype=None):  # noqa: A001
    def py(v):
        if isinstance(v, (Tensor, torch.Tensor)):
            return _t(v).item()
        return v
    start, limit, delta = py(start), py(limit), py(delta)
    if limit is None:
        start, limit = 0, start
    if dtype is None:
        dtype = float32 if any(isinstance(v, float) for v in (start, limit, delta)) else int32
    return Tensor(torch.arange(start, limit, delta, dtype=dtype))


def linspace(start, stop, num):
    return Tensor(torch.linspace(float(start), float(stop), int(num), dtype=float32))


def meshgrid(*args, **kw):
    indexing = kw.get("indexing", "xy")
    outs = torch.meshgrid(*[_t(a) for a in args], indexing=indexing)
    return [Tensor(o) for o in outs]


# ---------------------------------------------------------------- shape ops
def shape(x):
    return TensorShape(int(s) for s in _t(x).shape)


def expand_dims(x, axis):
    return Tensor(torch.unsqueeze(_t(x), int(axis)))


def squeeze(x, axis=None):
    t = _t(x)
    if axis is None:
        return Tensor(t.squeeze())
    for a in sorted([a % t.dim() for a in (axis if isinstance(axis, (list, tuple)) else [axis])], reverse=True):
        t = t.squeeze(a)
    return Tensor(t)


def reshape(x, shape):  # noqa: A002
    return Tensor(_t(x).reshape([_int(s) for s in shape]))


def tile(x, multiples):
    return Tensor(_t(x).repeat([_int(m) for m in multiples]))


def concat(values, axis):
    return Tensor(torch.cat([_t(v) for v in values], dim=int(axis)))


def stack(values, axis=0):
    return Tensor(torch.stack([_t(v) for v in values], dim=int(axis)))


def unstack(x, axis=0):
    return [Tensor(t) for t in torch.unbind(_t(x), dim=int(axis))]


def transpose(x, perm):
    return Tensor(_t(x).permute([int(p) for p in perm]).contiguous())


def reverse(x, axis):
    return Tensor(torch.flip(_t(x), dims=[int(a) for a in axis]))


def slice(x, begin, size):  # noqa: A001
    t = _t(x)
    idx = []
    for d, (b, s) in enumerate(zip(begin, size)):
        b = _int(b)
        s = _int(s)
        e = t.shape[d] if s == -1 else b + s
        idx.append(builtins.slice(b, e))
    return Tensor(t[tuple(idx)].clone())


def pad(x, paddings, mode="CONSTANT", constant_values=0):
    assert mode == "CONSTANT"
    t = _t(x)
    p = _t(paddings).tolist() if isinstance(paddings, (Tensor, torch.Tensor)) else [list(q) for q in paddings]
    flat = []
    for lo, hi in reversed(p):
        flat += [int(lo), int(hi)]
    return Tensor(F.pad(t, flat, mode="constant", value=float(constant_values)))


def cast(x, dtype):
    return Tensor(_t(x).to(dtype))


def to_int32(x):
    return cast(x, int32)


def stop_gradient(x):
    return Tensor(_t(x).detach())


# ---------------------------------------------------------------- math
def _un(fn):
    def f(x, name=None):
        return Tensor(fn(_t(x)))
    return f


def _bi(fn):
    def f(x, y, name=None):
        a = _t(x)
        return Tensor(fn(a, _raw(y, like=a)))
    return f


def _sqrt_rn(t):
    # torch's CPU float32 sqrt is not correctly rounded (off by one ulp for ~0.7% of inputs);
    # TF/Eigen use sqrtps (IEEE).  float64 sqrt + one rounding is exact for float32.
    if t.dtype == torch.float32:
        return torch.sqrt(t.double()).float()
    if t.dtype == torch.float64:     # numpy's sqrt is the hardware (correctly rounded) one
        import numpy as _np
        return torch.from_numpy(_np.sqrt(t.detach().numpy())) if not t.requires_grad else torch.sqrt(t)
    return torch.sqrt(t)


exp = _un(torch.exp)
log = _un(torch.log)
sqrt = _un(_sqrt_rn)
square = _un(torch.square)
floor = _un(torch.floor)
sigmoid = _un(torch.sigmoid)
tanh = _un(torch.tanh)
add = _bi(torch.add)
subtract = _bi(torch.sub)
multiply = _bi(torch.mul)
divide = _bi(torch.div)
logical_and = _bi(torch.logical_and)
maximum = _bi(torch.maximum)
minimum = _bi(torch.minimum)


def pow(x, y):  # noqa: A001
    return Tensor(torch.pow(_t(x), _raw(y, like=_t(x))))


def add_n(inputs):
    out = _t(inputs[0])
    for v in inputs[1:]:
        out = out + _t(v)
    return Tensor(out)


def clip_by_value(x, lo, hi):
    return Tensor(torch.clamp(_t(x), min=float(lo), max=float(hi)))


def matmul(a, b):
    return Tensor(torch.matmul(_t(a), _t(b)))


def _reduce(fn):
    def f(x, axis=None, keepdims=False, keep_dims=None, reduction_indices=None):
        if keep_dims is not None:
            keepdims = keep_dims
        if reduction_indices is not None:
            axis = reduction_indices
        t = _t(x)
        ax = _axes(axis)
        if ax is None:
            ax = list(builtins.range(t.dim()))
        return Tensor(fn(t, ax, keepdims))
    return f


reduce_sum = _reduce(lambda t, ax, k: torch.sum(t, dim=ax, keepdim=k))
reduce_mean = _reduce(lambda t, ax, k: torch.mean(t, dim=ax, keepdim=k))
reduce_max = _reduce(lambda t, ax, k: torch.amax(t, dim=ax, keepdim=k))
reduce_all = _reduce(lambda t, ax, k: torch.all(t, dim=ax[0] if isinstance(ax, list) and len(ax) == 1 else ax, keepdim=k))


def norm(x, axis=None, keepdims=False):
    """sqrt(sum x^2) along `axis`; squares are added left to right (see module doc)."""
    t = _t(x)
    assert axis is not None
    parts = torch.unbind(t * t, dim=int(axis))
    s = parts[0]
    for p in parts[1:]:
        s = s + p
    r = _sqrt_rn(s)
    if keepdims:
        r = r.unsqueeze(int(axis))
    return Tensor(r)


def cumsum(x, axis=0):
    return Tensor(torch.cumsum(_t(x), dim=int(axis)))


def cumprod(x, axis=0):
    return Tensor(torch.cumprod(_t(x), dim=int(axis)))


def argmin(x, axis=None):
    return Tensor(torch.argmin(_t(x), dim=int(axis)))


# ---------------------------------------------------------------- gather / scatter
def boolean_mask(tensor, mask):
    return Tensor(_t(tensor)[_t(mask)])


def scatter_nd(indices, updates, shape):  # noqa: A002
    idx = _t(indices).long()
    upd = _t(updates)
    shape = [_int(s) for s in shape]
    out = torch.zeros(shape, dtype=upd.dtype)
    k = idx.shape[-1]
    idx = idx.reshape(-1, k)
    upd = upd.reshape([idx.shape[0]] + shape[k:])
    out = out.index_put(tuple(idx[:, d] for d in builtins.range(k)), upd, accumulate=True)
    return Tensor(out)


def gather_nd(params, indices):
    p = _t(params)
    idx = _t(indices).long()
    k = idx.shape[-1]
    res = p[tuple(idx[..., d] for d in builtins.range(k))]
    return Tensor(res)


def py_func(func, inp, Tout):
    args = [(_t(a).numpy() if isinstance(a, (Tensor, torch.Tensor)) else np.asarray(a)) for a in inp]
    out = func(*args)
    return Tensor(torch.from_numpy(np.ascontiguousarray(out)).to(Tout))


# ---------------------------------------------------------------- tf.nn
class _NN:
    @staticmethod
    def _same_pad(k):
        total = k - 1
        lo = total // 2
        return lo, total - lo

    @staticmethod
    def conv3d(input, filter, strides, padding):  # noqa: A002
        assert padding == "SAME" and list(strides) == [1, 1, 1, 1, 1]
        x = _t(input).permute(0, 4, 1, 2, 3)  # NDHWC -> NCDHW
        w = _t(filter).permute(4, 3, 0, 1, 2)  # DHWIO -> OIDHW
        kd, kh, kw = w.shape[2:]
        pd, ph, pw = _NN._same_pad(kd), _NN._same_pad(kh), _NN._same_pad(kw)
        x = F.pad(x, [pw[0], pw[1], ph[0], ph[1], pd[0], pd[1]])
        y = F.conv3d(x, w.contiguous())
        return Tensor(y.permute(0, 2, 3, 4, 1).contiguous())

    @staticmethod
    def depthwise_conv2d(input, filter, strides, padding):  # noqa: A002
        assert padding == "SAME" and list(strides) == [1, 1, 1, 1]
        x = _t(input).permute(0, 3, 1, 2)  # NHWC -> NCHW
        w = _t(filter)  # [kh,kw,in,mult]
        kh, kw, cin, mult = w.shape
        w = w.permute(2, 3, 0, 1).reshape(cin * mult, 1, kh, kw)
        ph, pw = _NN._same_pad(kh), _NN._same_pad(kw)
        x = F.pad(x, [pw[0], pw[1], ph[0], ph[1]])
        y = F.conv2d(x, w.contiguous(), groups=cin)
        return Tensor(y.permute(0, 2, 3, 1).contiguous())

    @staticmethod
    def l2_loss(x):
        t = _t(x)
        return Tensor(torch.sum(t * t) / 2)


nn = _NN()


# ---------------------------------------------------------------- ops used by the loss surface (dpc/util/losses.py,
# dpc/models/model_pc.py:308-445); added in round 2 so that the reference's loss code runs unmodified as well
less = _bi(torch.lt)
equal = _bi(torch.eq)
not_equal = _bi(torch.ne)


def to_float(x):
    return cast(x, float32)


def where(condition, x=None, y=None):
    c = _t(condition)
    a = _t(x)
    return Tensor(torch.where(c, a, _raw(y, like=a)))


def one_hot(indices, depth, dtype=float32):
    return Tensor(F.one_hot(_t(indices).long(), int(depth)).to(dtype))


def Variable(initial_value, name=None, dtype=None, trainable=True):  # noqa: N802
    t = _raw(initial_value)
    if dtype is not None:
        t = t.to(dtype)
    return Tensor(t.clone())


class _Image:
    """tf.image.resize_images of TF 1.x (align_corners=False, no half-pixel centres): source coordinate = i * in/out."""

    class ResizeMethod:
        BILINEAR, NEAREST_NEIGHBOR, BICUBIC, AREA = 0, 1, 2, 3

    @staticmethod
    def resize_images(images, size, method=0, align_corners=False):
        assert not align_corners
        x = _t(images)                      # [B,H,W,C]
        oh, ow = int(size[0]), int(size[1])
        ih, iw = x.shape[1], x.shape[2]
        if method == _Image.ResizeMethod.NEAREST_NEIGHBOR:
            ri = torch.clamp(torch.floor(torch.arange(oh, dtype=torch.float32) * (ih / oh)).long(), max=ih - 1)
            ci = torch.clamp(torch.floor(torch.arange(ow, dtype=torch.float32) * (iw / ow)).long(), max=iw - 1)
            return Tensor(x[:, ri][:, :, ci])
        if method != _Image.ResizeMethod.BILINEAR:
            raise NotImplementedError("resize method %r" % (method,))

        def axis_weights(n_in, n_out):
            src = torch.arange(n_out, dtype=torch.float32) * (n_in / n_out)
            lo = torch.floor(src)
            frac = src - lo
            lo = lo.long()
            hi = torch.clamp(lo + 1, max=n_in - 1)
            return lo, hi, frac

        rl, rh, rf = axis_weights(ih, oh)
        cl, ch, cf = axis_weights(iw, ow)
        rf = rf.reshape(1, -1, 1, 1)
        cf = cf.reshape(1, 1, -1, 1)
        top = x[:, rl][:, :, cl] + (x[:, rl][:, :, ch] - x[:, rl][:, :, cl]) * cf
        bot = x[:, rh][:, :, cl] + (x[:, rh][:, :, ch] - x[:, rh][:, :, cl]) * cf
        return Tensor(top + (bot - top) * rf)


image = _Image()

from . import contrib  # noqa: E402,F401  (tf.contrib.summary / tf.contrib.slim stubs)
