"""Drop-in for the palette half of the reference's `io_utils.py` — same callables, CUDA underneath.

    extract_palette        io_utils.py:25-65
    rgba_to_single_int     io_utils.py:68-75   (dead code in the reference; kept for API parity)
    rgba_to_indexed        io_utils.py:78-93
    indexed_to_rgba        io_utils.py:96-103
    one_hot                pix2pix_model.py:300-301 (`tf.one_hot(idx, MAX_PALETTE_SIZE)` + reshape)

Every function also accepts a leading batch dimension (the reference maps them over a `tf.data`
pipeline one sample at a time; here a whole batch is one launch).  Inputs are int32 CUDA tensors.
"""
from __future__ import annotations

import torch

from . import _lib
from ._tensor import from_any, ptr, require_cuda, stream_ptr, to_caller_framework
from .configuration import INVALID_INDEX_COLOR, MAX_PALETTE_SIZE, OUTPUT_CHANNELS


class PaletteOverflowError(ValueError):
    """More than MAX_PALETTE_SIZE unique colours — the reference's `tf.repeat` receives a negative
    count there and raises (io_utils.py:61-62)."""


def _ordering_id(palette_ordering) -> int:
    try:
        return _lib.ORDERINGS[palette_ordering]
    except KeyError:
        # the reference treats every unknown string as "shuffled" (io_utils.py:56-58); be explicit instead
        raise ValueError("palette_ordering must be 'grayness', 'top2bottom', 'bottom2top' or 'shuffled', "
                         f"got {palette_ordering!r}") from None


def _check_ncolors(ncolors: torch.Tensor):
    """One device->host read of the per-image status (the only synchronisation of the palette path)."""
    nc = ncolors.cpu()
    if bool((nc == _lib.PALETTE_BAD_VALUE).any()):
        raise ValueError("extract_palette: colour values must lie in [0, 255]")
    if bool((nc > MAX_PALETTE_SIZE).any()):
        worst = int(nc.max())
        raise PaletteOverflowError(f"image has more than MAX_PALETTE_SIZE={MAX_PALETTE_SIZE} unique colours "
                                   f"(hash table saw {worst})")
    return nc


def _shuffle_keys(batch: int, device, generator=None) -> torch.Tensor:
    """`tf.random.shuffle(colors)` (io_utils.py:58): independent uniform keys, one per palette row; the kernel ranks
    the first n of them, which permutes the n colours uniformly at random.  Drawn on the host from `generator`
    (reproducible), 1 KiB per image uploaded."""
    keys = torch.rand((batch, MAX_PALETTE_SIZE), generator=generator, dtype=torch.float32)
    return keys.to(device, non_blocking=True)


def extract_palette(image, palette_ordering, channels=OUTPUT_CHANNELS, *, batched=None, return_counts=False,
                    check=True, generator=None):
    """io_utils.py:25-65.  `image` (H,W,channels·k) int32 — or (B,H,W,channels·k) — is flattened with
    reshape(-1, channels) per image; returns the (256, channels) int32 palette padded with
    INVALID_INDEX_COLOR (batched input -> (B,256,channels)).

    `batched`: None = infer (4-D input is a batch).  `check=False` skips the status read-back (no host
    synchronisation; overflow then goes unnoticed, pass `return_counts=True` to inspect it later)."""
    if channels != 4:
        raise ValueError("the CUDA palette path packs RGBA pixels: channels must be 4")
    img = require_cuda(from_any(image, name="image"), torch.int32, name="image")
    if batched is None:
        batched = img.dim() == 4
    if not batched:
        img = img.unsqueeze(0)
    b = img.shape[0]
    per_image = img[0].numel() if b else 0
    if per_image % 4 != 0:
        raise ValueError(f"image size {per_image} is not a multiple of channels=4")
    rows = per_image // 4
    palette = torch.empty((b, MAX_PALETTE_SIZE, 4), dtype=torch.int32, device=img.device)
    ncolors = torch.empty((b,), dtype=torch.int32, device=img.device)
    order = _ordering_id(palette_ordering)
    keys = _shuffle_keys(b, img.device, generator) if palette_ordering == "shuffled" else None
    if b:
        with torch.cuda.device(img.device):
            _lib.call("ph_extract_palette", ptr(img), b, rows, order, ptr(keys), ptr(palette), ptr(ncolors),
                      stream_ptr(img.device))
    if check:
        _check_ncolors(ncolors)
    out = palette if batched else palette[0]
    out = to_caller_framework(out, image)
    if return_counts:
        return out, (ncolors if batched else ncolors[0])
    return out


def rgba_to_single_int(values_in_rgba):
    """io_utils.py:68-75, verbatim semantics including the reference's last multiplier of 0
    (alpha is dropped).  Dead code in the reference; plain tensor arithmetic here."""
    v = from_any(values_in_rgba)
    converted = torch.zeros(v.shape[:-1], dtype=torch.int32, device=v.device)
    for i, multiplier in enumerate([16777216, 65536, 256, 0]):
        converted = converted + v[..., i].to(torch.int32) * multiplier
    return converted


def rgba_to_indexed(image, palette, *, mode="exact", with_one_hot=False, depth=MAX_PALETTE_SIZE):
    """io_utils.py:78-93.  image (H,W,4) int32 + palette (256,4) -> (H,W,1) int32, or batched
    (B,H,W,4) + (B,256,4)|(256,4) -> (B,H,W,1).

    mode="exact" is the reference: index = sum of all palette rows equal to the pixel (scatter_nd adds
    duplicates), 0 when none.  mode="nearest" is the arg-min of squared RGBA distance.
    with_one_hot=True also returns the float32 (…,H,W,depth) one-hot of pix2pix_model.py:300-301,
    written by the same kernel."""
    img = require_cuda(from_any(image, name="image"), torch.int32, name="image")
    pal = require_cuda(from_any(palette, name="palette"), torch.int32, name="palette")
    batched = img.dim() == 4
    if not batched:
        img = img.unsqueeze(0)
    if img.dim() != 4 or img.shape[-1] != 4:
        raise ValueError(f"image must be (H,W,4) or (B,H,W,4), got {tuple(img.shape)}")
    if pal.dim() == 2:
        pal = pal.unsqueeze(0)
    if pal.shape[1:] != (MAX_PALETTE_SIZE, 4) or pal.shape[0] not in (1, img.shape[0]):
        raise ValueError(f"palette must be ({MAX_PALETTE_SIZE},4) or (B,{MAX_PALETTE_SIZE},4), got {tuple(pal.shape)}")
    b, h, w, _ = img.shape
    idx = torch.empty((b, h, w, 1), dtype=torch.int32, device=img.device)
    oh = torch.empty((b, h, w, depth), dtype=torch.float32, device=img.device) if with_one_hot else None
    if b and h * w:
        with torch.cuda.device(img.device):
            _lib.call("ph_rgba_to_indexed", ptr(img), b, h * w, ptr(pal), pal.shape[0], _lib.INDEX_MODES[mode],
                      ptr(idx), ptr(oh), depth, stream_ptr(img.device))
    if not batched:
        idx = idx[0]
        oh = oh[0] if oh is not None else None
    idx = to_caller_framework(idx, image)
    if with_one_hot:
        return idx, to_caller_framework(oh, image)
    return idx


def indexed_to_rgba(indexed_image, palette):
    """io_utils.py:96-103.  (H,W,1) int32 + (256,C) -> (H,W,C); batched (B,H,W,1) + (B,256,C)|(256,C)."""
    idx = require_cuda(from_any(indexed_image, name="indexed_image"), torch.int32, name="indexed_image")
    pal = require_cuda(from_any(palette, name="palette"), torch.int32, name="palette")
    batched = idx.dim() == 4
    if not batched:
        idx = idx.unsqueeze(0)
    if idx.dim() != 4:
        raise ValueError(f"indexed_image must be (H,W,1) or (B,H,W,1), got {tuple(idx.shape)}")
    if pal.dim() == 2:
        pal = pal.unsqueeze(0)
    if pal.shape[0] not in (1, idx.shape[0]):
        raise ValueError("palette batch must be 1 or match the image batch")
    b, h, w = idx.shape[:3]
    npix = h * w * idx.shape[3]
    rows, ch = pal.shape[1], pal.shape[2]
    out = torch.empty((b, h, w, idx.shape[3] * ch), dtype=torch.int32, device=idx.device)
    if b and npix:
        with torch.cuda.device(idx.device):
            _lib.call("ph_indexed_to_rgba", ptr(idx), b, npix, ptr(pal), pal.shape[0], rows, ch, ptr(out),
                      stream_ptr(idx.device))
    return to_caller_framework(out if batched else out[0], indexed_image)


def probabilities_to_indexed(probabilities, palette=None):
    """Inference ops of the indexed model in one pass: `tf.expand_dims(tf.argmax(probs, -1, "int32"), -1)`
    (pix2pix_model.py:283-287) and, when `palette` is given, `indexed_to_rgba` of the result (:356, :446-447).
    probabilities (H,W,D) or (B,H,W,D) float32; palette (256,4) or (B,256,4) int32.
    Returns indexed (…,1) int32, or (indexed, rgba (…,4) int32) with a palette."""
    pr = require_cuda(from_any(probabilities, name="probabilities"), torch.float32, name="probabilities")
    batched = pr.dim() == 4
    if not batched:
        pr = pr.unsqueeze(0)
    if pr.dim() != 4:
        raise ValueError(f"probabilities must be (H,W,D) or (B,H,W,D), got {tuple(pr.shape)}")
    b, h, w, depth = pr.shape
    idx = torch.empty((b, h, w, 1), dtype=torch.int32, device=pr.device)
    pal = rgba = None
    pal_b = rows = 0
    if palette is not None:
        pal = require_cuda(from_any(palette, name="palette"), torch.int32, name="palette")
        if pal.dim() == 2:
            pal = pal.unsqueeze(0)
        if pal.shape[0] not in (1, b) or pal.shape[2] != 4:
            raise ValueError("palette must be (256,4) or (B,256,4) with B matching the image batch")
        pal_b, rows = pal.shape[0], pal.shape[1]
        rgba = torch.empty((b, h, w, 4), dtype=torch.int32, device=pr.device)
    if b and h * w:
        with torch.cuda.device(pr.device):
            _lib.call("ph_argmax_indexed", ptr(pr), b, h * w, depth, ptr(pal) if pal is not None else None, pal_b, rows,
                      ptr(idx), ptr(rgba) if rgba is not None else None, stream_ptr(pr.device))
    idx = to_caller_framework(idx if batched else idx[0], probabilities)
    if rgba is None:
        return idx
    return idx, to_caller_framework(rgba if batched else rgba[0], probabilities)


def one_hot(indices, depth=MAX_PALETTE_SIZE):
    """`tf.reshape(tf.one_hot(idx, depth, axis=-1), [B,H,W,-1])` of pix2pix_model.py:300-301:
    (…,1) or (…) int32 -> (…,depth) float32; an index outside [0,depth) gives an all-zero row."""
    idx = require_cuda(from_any(indices, name="indices"), torch.int32, name="indices")
    shape = tuple(idx.shape[:-1]) if idx.dim() and idx.shape[-1] == 1 else tuple(idx.shape)
    out = torch.empty(shape + (depth,), dtype=torch.float32, device=idx.device)
    if idx.numel():
        with torch.cuda.device(idx.device):
            _lib.call("ph_one_hot", ptr(idx), idx.numel(), depth, ptr(out), stream_ptr(idx.device))
    return to_caller_framework(out, indices)


__all__ = ["extract_palette", "rgba_to_single_int", "rgba_to_indexed", "indexed_to_rgba", "one_hot",
           "probabilities_to_indexed",
           "PaletteOverflowError", "MAX_PALETTE_SIZE", "INVALID_INDEX_COLOR"]
