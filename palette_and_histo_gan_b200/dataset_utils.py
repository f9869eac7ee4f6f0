"""Drop-in for the palette call sites and pixel helpers of the reference's `dataset_utils.py`.

    blacken_transparent_pixels   dataset_utils.py:11-20
    normalize / denormalize      dataset_utils.py:39-60
    load_indexed_images          dataset_utils.py:138-151 (the body of create_indexed_image_loader,
                                 after PNG decode): shared palette of source||target + two index images
    augment_hue_rotation / augment_translation / augment_two / normalize_two /
    create_augmentation_with_prob   dataset_utils.py:80-120, one fused kernel

PNG decoding, file naming and the `tf.data` plumbing around these are out of scope; the functions here take
already-decoded pixel tensors, single or batched.
"""
from __future__ import annotations

import torch

from . import _lib, io_utils
from ._tensor import from_any, ptr, require_cuda, stream_ptr, to_caller_framework
from .configuration import MAX_PALETTE_SIZE


def _pixel_map(image, op, name):
    """One launch of `ph_pixel_map` over a float32 CUDA tensor (integer tensors are cast first, as the reference's
    callers do with `tf.cast(image, "float32")`, dataset_utils.py:72)."""
    img = from_any(image, name=name)
    if not img.is_cuda:
        raise ValueError(f"{name} must live on a CUDA device (got {img.device}); there is no CPU fallback")
    src = require_cuda(img if img.dtype == torch.float32 else img.to(torch.float32), torch.float32, name=name)
    if op == "blacken" and (src.dim() == 0 or src.shape[-1] != 4):
        raise ValueError(f"{name} must be RGBA (last dimension 4), got {tuple(src.shape)}")
    out = torch.empty_like(src)
    if src.numel():
        with torch.cuda.device(src.device):
            _lib.call("ph_pixel_map", ptr(src), src.numel(), _lib.MAP_OPS[op], ptr(out), stream_ptr(src.device))
    if op == "blacken" and img.dtype != torch.float32:
        out = out.to(img.dtype)  # blacken keeps the dtype of its argument (tf.where)
    return to_caller_framework(out, image)


def blacken_transparent_pixels(image):
    """dataset_utils.py:11-20: pixels whose alpha is 0 become (0,0,0,0).  (…,4) CUDA tensor."""
    return _pixel_map(image, "blacken", "image")


def normalize(image):
    """dataset_utils.py:39-48: [0,255] -> [-1,1] as a true division by 127.5 and a subtraction (TensorFlow's two
    roundings; a multiplication by the reciprocal is 1 ulp away)."""
    return _pixel_map(image, "normalize", "image")


def denormalize(image):
    """dataset_utils.py:51-60: [-1,1] -> [0,255]."""
    return _pixel_map(image, "denormalize", "image")


def load_image(image_u8, should_normalize=True):
    """dataset_utils.py:66-77 after `decode_png`: uint8 RGBA (…,4) CUDA tensor -> float32 with
    `blacken_transparent_pixels` and (optionally) `normalize` fused in one pass on the device."""
    img = require_cuda(from_any(image_u8, name="image"), torch.uint8, name="image")
    if img.shape[-1] != 4:
        raise ValueError("load_image expects RGBA uint8 pixels (last dimension 4)")
    out = torch.empty(img.shape, dtype=torch.float32, device=img.device)
    if img.numel():
        with torch.cuda.device(img.device):
            _lib.call("ph_u8_to_float_image", ptr(img), img.numel() // 4, 1, 1 if should_normalize else 0, ptr(out),
                      stream_ptr(img.device))
    return to_caller_framework(out, image_u8)


# ---------------------------------------------------------------------------------------------------
# Augmentation (dataset_utils.py:80-120).  The pixel work — tf.image.adjust_hue's algorithm and the nearest,
# constant-fill translation of keras RandomTranslation — runs in ONE kernel for both images of a pair
# (`ph_augment_pair`); the random draws are made on the host with a torch generator: TensorFlow's Philox streams
# are not reproduced (the reference seeds them from `tf.random.uniform`, so only the distributions are defined).
# ---------------------------------------------------------------------------------------------------
MAX_HUE_DELTA = 0.5               # dataset_utils.py:82
HEIGHT_FACTOR = (-0.15, 0.075)    # dataset_utils.py:89
WIDTH_FACTOR = (-0.125, 0.125)    # dataset_utils.py:89


def _as_batch(img, name):
    img = require_cuda(from_any(img, name=name), torch.float32, name=name)
    batched = img.dim() == 4
    if not batched:
        img = img.unsqueeze(0)
    if img.dim() != 4 or img.shape[-1] != 4:
        raise ValueError(f"{name} must be RGBA float32 (H,W,4) or (B,H,W,4), got {tuple(img.shape)}")
    return img, batched


def _per_image(values, b, width, device, name, limit=None):
    """python scalar / sequence / tensor -> float32 device tensor (b,) or (b,width); scalars are shared."""
    if values is None:
        return None
    t = values.detach().to(torch.float32) if isinstance(values, torch.Tensor) else torch.tensor(values, dtype=torch.float32)
    shape = (b,) if width == 1 else (b, width)
    if t.dim() < len(shape):
        t = t.expand(shape)
    if tuple(t.shape) != shape:
        raise ValueError(f"{name} must have shape {shape} (or be shared by the batch), got {tuple(t.shape)}")
    if limit is not None and not t.is_cuda and t.numel() and not bool((t.abs() <= limit).all()):
        raise ValueError(f"{name} must lie in [-{limit}, {limit}]")  # device-resident draws are not read back
    return t.to(device).contiguous()


def _augment(first, second, hue_delta, translations, apply, should_normalize):
    a, batched = _as_batch(first, "first")
    b2 = None
    if second is not None:
        b2, batched2 = _as_batch(second, "second")
        if b2.shape != a.shape or batched2 != batched or b2.device != a.device:
            raise ValueError("both images must have the same shape and device")
    n, h, w, _ = a.shape
    hue = _per_image(hue_delta, n, 1, a.device, "hue delta", limit=1.0)  # tf.image.adjust_hue: delta in [-1, 1]
    tr = _per_image(translations, n, 2, a.device, "translations")
    ap = None
    if apply is not None:
        ap = torch.as_tensor(apply)
        if ap.dtype == torch.bool and ap.is_cuda and tuple(ap.shape) == (n,) and ap.is_contiguous():
            ap = ap.view(torch.uint8)  # device-resident gate: no conversion kernels
        else:
            ap = ap.to(torch.bool).expand(n).to(torch.uint8).to(a.device).contiguous()
    out_a = torch.empty_like(a)
    out_b = torch.empty_like(b2) if b2 is not None else None
    if a.numel():
        with torch.cuda.device(a.device):
            _lib.call("ph_augment_pair", ptr(a), ptr(b2), n, h, w, ptr(hue), ptr(tr), ptr(ap),
                      1 if should_normalize else 0, ptr(out_a), ptr(out_b), stream_ptr(a.device))
    if not batched:
        out_a = out_a[0]
        out_b = out_b[0] if out_b is not None else None
    return to_caller_framework(out_a, first), (to_caller_framework(out_b, first) if out_b is not None else None)


def _draw_hue_delta(n, seed, generator):
    if seed is not None:
        s = [int(v) for v in torch.as_tensor(seed).reshape(-1).tolist()]
        generator = torch.Generator().manual_seed((s[0] << 20) ^ s[-1])
        return (torch.rand((), generator=generator) * 2 - 1) * MAX_HUE_DELTA  # one seed -> one delta, shared
    return (torch.rand(n, generator=generator) * 2 - 1) * MAX_HUE_DELTA


def _draw_translations(n, h, w, generator):
    u = torch.rand(n, 2, generator=generator)
    dx = (WIDTH_FACTOR[0] + u[:, 0] * (WIDTH_FACTOR[1] - WIDTH_FACTOR[0])) * w
    dy = (HEIGHT_FACTOR[0] + u[:, 1] * (HEIGHT_FACTOR[1] - HEIGHT_FACTOR[0])) * h
    return torch.stack([dx, dy], 1)


def adjust_hue(image, delta):
    """`tf.image.adjust_hue(image, delta)` for RGBA float32 pixels (alpha untouched): the op behind
    dataset_utils.py:82.  `delta` in [-1, 1], a scalar or one value per image."""
    return _augment(image, None, delta, None, None, False)[0]


def augment_hue_rotation(image, seed=None, *, delta=None, generator=None):
    """dataset_utils.py:80-84.  `seed` (two integers, as the reference passes) fixes the draw of delta in
    [-0.5, 0.5) — equal seeds give equal rotations, which is what `augment_two` relies on; `delta=` bypasses
    the draw."""
    if delta is None:
        n = 1 if from_any(image).dim() == 3 else from_any(image).shape[0]
        delta = _draw_hue_delta(n, seed, generator)
    return adjust_hue(image, delta)


def augment_translation(images, *, translations=None, generator=None):
    """dataset_utils.py:87-92: the images (a pair) are moved by ONE shared translation per sample: (dx, dy) pixels
    drawn from width·U(-0.125, 0.125), height·U(-0.15, 0.075) unless `translations` ((2,) or (B,2)) is given;
    nearest source pixel, zeros where it falls outside."""
    first, second = images
    t0 = from_any(first)
    if translations is None:
        n = 1 if t0.dim() == 3 else t0.shape[0]
        translations = _draw_translations(n, t0.shape[-3], t0.shape[-2], generator)
        if t0.dim() == 3:
            translations = translations[0]
    return _augment(first, second, None, translations, None, False)


def augment_two(first, second, *, hue_delta=None, translations=None, generator=None, apply=None,
                should_normalize=False):
    """dataset_utils.py:95-102: one hue rotation shared by both images, then one shared translation — a single
    kernel launch.  Keyword arguments give the draws explicitly (tests, reproducible pipelines); `apply` is the
    per-sample gate of `create_augmentation_with_prob`; `should_normalize` fuses the `normalize_two` that follows
    in `load_rgba_ds` (dataset_utils.py:220-225)."""
    t0 = from_any(first)
    n = 1 if t0.dim() == 3 else t0.shape[0]
    if hue_delta is None:
        hue_delta = _draw_hue_delta(n, None, generator)
        hue_delta = hue_delta[0] if t0.dim() == 3 else hue_delta
    if translations is None:
        translations = _draw_translations(n, t0.shape[-3], t0.shape[-2], generator)
        translations = translations[0] if t0.dim() == 3 else translations
    return _augment(first, second, hue_delta, translations, apply, should_normalize)


def normalize_two(first, second):
    """dataset_utils.py:105-106."""
    return normalize(first), normalize(second)


def create_augmentation_with_prob(prob=0.8, *, generator=None, should_normalize=False):
    """dataset_utils.py:109-120: augment a sample when uniform() < prob; batched inputs draw one choice per
    sample (the reference maps over single samples)."""

    def augmentation_wrapper(first, second):
        t0 = from_any(first)
        n = 1 if t0.dim() == 3 else t0.shape[0]
        choice = torch.rand(n, generator=generator) < prob
        return augment_two(first, second, generator=generator, apply=choice, should_normalize=should_normalize)

    return augmentation_wrapper


def load_indexed_images(source_image, target_image, palette_ordering="grayness", *, check=True, generator=None):
    """dataset_utils.py:138-151 for decoded images: `concat([source, target], -1)` -> `extract_palette`
    -> `rgba_to_indexed` twice with the shared palette, in ONE launch.  (H,W,4) or (B,H,W,4) CUDA tensors, int32
    (the reference's `tf.cast(image, "int32")`) or uint8 (the decoded PNG as it is: a quarter of the bytes read).
    `palette_ordering="shuffled"` (io_utils.py:56-58) draws its permutation from `generator`.
    Returns (source_indexed (…,H,W,1), target_indexed (…,H,W,1), palette (…,256,4)), all int32."""
    src0 = from_any(source_image, name="source_image")
    dt = torch.uint8 if src0.dtype == torch.uint8 else torch.int32
    src = require_cuda(src0, dt, name="source_image")
    tgt = require_cuda(from_any(target_image, name="target_image"), dt, name="target_image")
    if src.shape != tgt.shape:
        raise ValueError("source and target images must have the same shape")
    batched = src.dim() == 4
    if not batched:
        src, tgt = src.unsqueeze(0), tgt.unsqueeze(0)
    if src.dim() != 4 or src.shape[-1] != 4:
        raise ValueError(f"images must be (H,W,4) or (B,H,W,4), got {tuple(src.shape)}")
    order = io_utils._ordering_id(palette_ordering)
    b, h, w, _ = src.shape
    keys = io_utils._shuffle_keys(b, src.device, generator) if palette_ordering == "shuffled" else None
    s_idx = torch.empty((b, h, w, 1), dtype=torch.int32, device=src.device)
    t_idx = torch.empty((b, h, w, 1), dtype=torch.int32, device=src.device)
    palette = torch.empty((b, MAX_PALETTE_SIZE, 4), dtype=torch.int32, device=src.device)
    ncolors = torch.empty((b,), dtype=torch.int32, device=src.device)
    if b:
        with torch.cuda.device(src.device):
            _lib.call("ph_load_indexed_images_u8" if dt == torch.uint8 else "ph_load_indexed_images", ptr(src), ptr(tgt),
                      b, h * w, order, ptr(keys), ptr(s_idx), ptr(t_idx), ptr(palette), ptr(ncolors),
                      stream_ptr(src.device))
    if check:
        io_utils._check_ncolors(ncolors)
    if not batched:
        s_idx, t_idx, palette = s_idx[0], t_idx[0], palette[0]
    return (to_caller_framework(s_idx, source_image), to_caller_framework(t_idx, source_image),
            to_caller_framework(palette, source_image))


def create_indexed_image_loader(palette_ordering):
    """Shape of the reference's factory (dataset_utils.py:123-129) minus the file-system arguments:
    returns `load_indexed_images(source, target)` bound to an ordering."""

    def loader(source_image, target_image):
        return load_indexed_images(source_image, target_image, palette_ordering)

    return loader
