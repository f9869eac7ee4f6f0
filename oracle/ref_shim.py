"""A minimal stand-in for the `tensorflow` module, backed by torch CPU tensors (TEST INFRASTRUCTURE).

TensorFlow (the reference pins 2.9.1) cannot be installed in this image, so the reference's own source files
cannot be imported as they are.  This module implements exactly the TensorFlow symbols that
/root/reference/histogram.py, io_utils.py and dataset_utils.py touch on the colour path, each with the semantics
TensorFlow documents for it, so that `oracle/run_reference.py` can execute the reference's *unmodified source*
(its control flow, axis conventions, transposes, broadcasting, dtype promotion) and, because the backing tensors
are torch tensors, differentiate it with autograd the way `tape.gradient` differentiates the TF graph.

What is and is not pinned by this: every line of the reference's Python runs as written; what is assumed is the
per-op behaviour listed below (documented TensorFlow semantics, unverifiable here — SURVEY.md §8c):

  linspace        `start + delta * [1 .. num-2]` framed by the exact end points, delta = (stop-start)/(num-1), float32
  pow/sqrt/log/exp/abs and the arithmetic operators: float32 element-wise; python scalars are weakly typed
  matmul          float32; the (n,4)x(4,1) grayness product is evaluated left to right without fused multiply-add
  UniqueWithCountsV2(axis=[0])  unique rows in order of first occurrence (+ inverse index, counts)
  argsort(stable=True)          stable ascending sort
  gather          params[indices]
  scatter_nd      zeros(shape) with updates ADDED at the indices (duplicates accumulate)
  where(cond)     coordinates of the true elements in row-major order, int64
  repeat(x, [n], axis=0)        n copies; a negative n is an error
  one_hot         1.0 at the index, an index outside [0, depth) gives an all-zero row
  reduce_sum / reduce_mean, reshape, transpose(perm), stack, concat, expand_dims, squeeze, cast, shape, zeros,
  constant, equal, tf.function (identity decorator), tf.newaxis (None)
  io.read_file + image.decode_png(channels=4): the PNG decoded to RGBA uint8 (done with PIL here)
  image.stateless_random_hue, keras.layers.RandomTranslation, random.uniform, split (augmentation,
  dataset_utils.py:80-120): adjust_hue / nearest constant-fill translation as restated in oracle/augment_oracle.py;
  the random draws come from a numpy generator (TensorFlow's Philox streams are not restated) and are recorded in
  `DRAWS` so that the fixture stores them next to the outputs

Nothing in the product imports this file.
"""
from __future__ import annotations

import sys
import types

import numpy as np
import torch

_DTYPES = {"float32": torch.float32, "int32": torch.int32, "int64": torch.int64, "bool": torch.bool,
           "float64": torch.float64, "uint8": torch.uint8}


class RefTensor(torch.Tensor):
    """torch.Tensor whose `x[::-1]` reverses the first axis like a TensorFlow / numpy slice (io_utils.py:48)."""

    def __getitem__(self, idx):
        if isinstance(idx, slice) and idx.step == -1 and idx.start is None and idx.stop is None:
            return torch.flip(self, [0])
        return super().__getitem__(idx)

    def numpy(self):  # noqa: D102
        return torch.Tensor.numpy(self.detach().as_subclass(torch.Tensor))


def _dt(d):
    return _DTYPES[d] if isinstance(d, str) else d


def T(x, dtype=None):
    """python scalar / list / ndarray / tensor -> RefTensor (float -> float32, int -> int32 like tf.constant)."""
    if isinstance(x, torch.Tensor):
        t = x
    else:
        a = np.asarray(x)
        if dtype is None:
            if a.dtype.kind == "f":
                a = a.astype(np.float32)
            elif a.dtype.kind in "iu" and a.dtype != np.uint8:
                a = a.astype(np.int32)
        t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(_dt(dtype))
    return t.as_subclass(RefTensor)


def _axes(axis):
    if axis is None:
        return None
    return tuple(int(a) for a in axis) if isinstance(axis, (list, tuple)) else int(axis)


def linspace(start, stop, num):
    num = int(num)
    start_t, stop_t = torch.tensor(float(start), dtype=torch.float32), torch.tensor(float(stop), dtype=torch.float32)
    if num == 1:
        return T(start_t.reshape(1))
    delta = (stop_t - start_t) / torch.tensor(float(num - 1), dtype=torch.float32)
    middle = start_t + delta * torch.arange(1, num - 1, dtype=torch.float32)
    return T(torch.cat([start_t.reshape(1), middle, stop_t.reshape(1)]))


def matmul(a, b):
    a, b = T(a), T(b)
    if a.dim() == 2 and b.dim() == 2 and b.shape[0] <= 8:
        # small inner dimension (the grayness key): left-to-right float32, no fused multiply-add
        acc = a[:, 0:1] * b[0:1, :]
        for k in range(1, b.shape[0]):
            acc = acc + a[:, k:k + 1] * b[k:k + 1, :]
        return acc
    return torch.matmul(a, b)


def unique_with_counts_v2(x, axis):
    """Unique rows in order of FIRST OCCURRENCE, the index of each input row among them, and the counts."""
    assert list(axis) == [0]
    rows = T(x)
    a = rows.numpy().reshape(rows.shape[0], -1)
    _, first, inverse, counts = np.unique(a, axis=0, return_index=True, return_inverse=True, return_counts=True)
    order = np.argsort(first, kind="stable")          # sorted-unique position -> rank by first occurrence
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    uniq = rows[torch.from_numpy(first[order].astype(np.int64))]
    return uniq, T(rank[np.asarray(inverse).reshape(-1)].astype(np.int32)), T(counts[order].astype(np.int32))


def scatter_nd(indices, updates, shape):
    out = torch.zeros([int(s) for s in shape], dtype=updates.dtype).as_subclass(RefTensor)
    assert indices.dim() == 2 and indices.shape[1] == 1 and len(shape) == 1
    out.index_add_(0, indices[:, 0].long(), updates)
    return out


def repeat(x, repeats, axis=None):
    t = T(x)
    r = torch.as_tensor(repeats).reshape(-1)
    if (r < 0).any():
        raise ValueError("repeats must be non-negative")  # io_utils.py:62 with more than 256 colours
    if axis is None:
        return torch.repeat_interleave(t.reshape(-1), r if r.numel() > 1 else int(r))
    return torch.repeat_interleave(t, r if r.numel() > 1 else int(r), dim=int(axis))


def where(condition, x=None, y=None):
    if x is None:
        return torch.nonzero(T(condition)).as_subclass(RefTensor)
    xt = x if isinstance(x, torch.Tensor) else torch.tensor(x, dtype=(y.dtype if isinstance(y, torch.Tensor) else torch.float32))
    yt = y if isinstance(y, torch.Tensor) else torch.tensor(y, dtype=xt.dtype)
    return torch.where(T(condition), xt, yt).as_subclass(RefTensor)


def one_hot(indices, depth, axis=-1):
    idx = T(indices).long()
    ok = (idx >= 0) & (idx < depth)
    out = torch.zeros(tuple(idx.shape) + (int(depth),), dtype=torch.float32)
    out.scatter_(-1, idx.clamp(0, depth - 1).unsqueeze(-1), ok.to(torch.float32).unsqueeze(-1))
    assert axis == -1
    return out.as_subclass(RefTensor)


def _pow(x, y):
    return torch.pow(T(x), y)


def _read_file(path):
    return str(path)


def _decode_png(path, channels=4):
    from PIL import Image

    assert channels == 4
    return T(np.asarray(Image.open(path).convert("RGBA"), dtype=np.uint8))


DRAWS = []  # (kind, values) of every random draw the augmentation stand-ins made, in call order
_RNG = np.random.default_rng(47)


def _stateless_random_hue(image, max_delta, seed):
    from oracle import augment_oracle as ao

    seed = [int(v) for v in np.asarray(T(seed).numpy()).reshape(-1)]
    delta = np.float32(np.random.default_rng(seed).uniform(-max_delta, max_delta))
    DRAWS.append(("hue_delta", float(delta)))
    return T(ao.adjust_hue_f32(T(image).numpy(), delta))


class _RandomTranslation:
    """keras.layers.RandomTranslation(height_factor, width_factor, fill_mode="constant", interpolation="nearest")
    applied to an unbatched (H,W,C) image (training=True is the default of TF 2.9's layer)."""

    def __init__(self, height_factor, width_factor, fill_mode="reflect", interpolation="bilinear"):
        assert fill_mode == "constant" and interpolation == "nearest", "only the reference's configuration"
        as_pair = lambda f: (float(f[0]), float(f[1])) if isinstance(f, (tuple, list)) else (-float(f), float(f))
        self.height, self.width = as_pair(height_factor), as_pair(width_factor)

    def __call__(self, image):
        from oracle import augment_oracle as ao

        img = T(image).numpy()
        h, w = img.shape[:2]
        dy = np.float32(np.float32(_RNG.uniform(*self.height)) * np.float32(h))
        dx = np.float32(np.float32(_RNG.uniform(*self.width)) * np.float32(w))
        DRAWS.append(("translation", (float(dx), float(dy))))
        return T(ao.translate_nearest(img, dx, dy))


def _random_uniform(shape, minval=0, maxval=None, dtype="float32"):
    if _dt(dtype) in (torch.int32, torch.int64):
        return T(_RNG.integers(minval, maxval, size=[int(s) for s in shape]).astype(np.int32))
    hi = 1.0 if maxval is None else maxval
    return T(np.asarray(_RNG.uniform(minval, hi, size=[int(s) for s in shape]), np.float32))


def build():
    """-> a module object to install as sys.modules['tensorflow']."""
    tf = types.ModuleType("tensorflow")
    tf.newaxis = None
    tf.function = lambda f=None, **kw: f if f is not None else (lambda g: g)
    tf.constant = lambda v, dtype=None: T(v, dtype)
    tf.cast = lambda x, dtype: T(x).to(_dt(dtype)).as_subclass(RefTensor)
    tf.shape = lambda x: tuple(int(s) for s in T(x).shape)
    tf.reshape = lambda x, shape: T(x).reshape([int(s) for s in shape])
    tf.transpose = lambda x, perm: T(x).permute(*[int(p) for p in perm])
    tf.expand_dims = lambda x, axis: T(x).unsqueeze(int(axis))
    tf.squeeze = lambda x: T(x).squeeze()
    tf.stack = lambda xs, axis=0: torch.stack([T(x) for x in xs], dim=int(axis)).as_subclass(RefTensor)
    tf.concat = lambda xs, axis: torch.cat([T(x) for x in xs], dim=int(axis)).as_subclass(RefTensor)
    tf.zeros = lambda shape, dtype="float32": torch.zeros([int(s) for s in shape], dtype=_dt(dtype)).as_subclass(RefTensor)
    tf.pow = _pow
    tf.sqrt = lambda x: torch.sqrt(T(x))
    tf.exp = lambda x: torch.exp(T(x))
    tf.abs = lambda x: torch.abs(T(x))
    tf.linspace = linspace
    tf.matmul = matmul
    tf.reduce_sum = lambda x, axis=None, keepdims=False: (T(x).sum() if axis is None
                                                          else T(x).sum(dim=_axes(axis), keepdim=keepdims))
    tf.reduce_mean = lambda x, axis=None: T(x).mean() if axis is None else T(x).mean(dim=_axes(axis))
    tf.argsort = lambda v, direction="ASCENDING", stable=False: torch.argsort(T(v), stable=True, descending=direction != "ASCENDING")
    tf.gather = lambda params, indices: T(params)[T(indices).long()]
    tf.scatter_nd = scatter_nd
    tf.repeat = repeat
    tf.where = where
    tf.one_hot = one_hot
    tf.tuple = lambda xs: tuple(xs)
    tf.math = types.SimpleNamespace(log=lambda x: torch.log(T(x)), equal=lambda a, b: T(a) == b)
    tf.equal = lambda a, b: T(a) == b
    tf.raw_ops = types.SimpleNamespace(UniqueWithCountsV2=lambda x, axis: unique_with_counts_v2(x, axis))
    tf.strings = types.SimpleNamespace(join=lambda parts, sep="": sep.join(str(p) for p in parts))
    tf.io = types.SimpleNamespace(read_file=_read_file)
    tf.image = types.SimpleNamespace(decode_png=_decode_png, stateless_random_hue=_stateless_random_hue)
    tf.random = types.SimpleNamespace(uniform=_random_uniform)
    tf.keras = types.SimpleNamespace(layers=types.SimpleNamespace(RandomTranslation=_RandomTranslation))
    tf.split = lambda x, n, axis=0: [t.as_subclass(RefTensor) for t in torch.chunk(T(x), int(n), dim=int(axis))]
    tf.data = types.SimpleNamespace(AUTOTUNE=-1)
    return tf


def install(reference_dir="/root/reference"):
    """Put the shim (and an empty matplotlib) into sys.modules and the reference directory on sys.path."""
    sys.modules["tensorflow"] = build()
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = types.ModuleType("matplotlib.pyplot")
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", mpl.pyplot)
    if reference_dir not in sys.path:
        sys.path.insert(0, reference_dir)
