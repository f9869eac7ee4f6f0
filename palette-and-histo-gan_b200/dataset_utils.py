"""Drop-in for the palette call sites and pixel helpers of the reference's `dataset_utils.py`.

    blacken_transparent_pixels   dataset_utils.py:11-20
    normalize / denormalize      dataset_utils.py:39-60
    load_indexed_images          dataset_utils.py:138-151 (the body of create_indexed_image_loader,
                                 after PNG decode): shared palette of source||target + two index images

PNG decoding, file naming, augmentation and the `tf.data` plumbing around these are out of scope
(SURVEY.md §8f f2/f4); the functions here take already-decoded pixel tensors, batched.
"""
from __future__ import annotations

import torch

from . import _lib, io_utils
from ._tensor import from_any, ptr, require_cuda, stream_ptr, to_caller_framework
from .configuration import MAX_PALETTE_SIZE


def blacken_transparent_pixels(image):
    """dataset_utils.py:11-20: pixels whose alpha is 0 become (0,0,0,0). (…,4) any dtype."""
    img = from_any(image)
    return to_caller_framework(torch.where(img[..., 3:4] == 0, torch.zeros_like(img), img), image)


def normalize(image):
    """dataset_utils.py:39-48: [0,255] -> [-1,1]."""
    return to_caller_framework((from_any(image) / 127.5) - 1, image)


def denormalize(image):
    """dataset_utils.py:51-60: [-1,1] -> [0,255]."""
    return to_caller_framework((from_any(image) + 1) * 127.5, image)


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


def load_indexed_images(source_image, target_image, palette_ordering="grayness", *, check=True):
    """dataset_utils.py:138-151 for decoded images: `concat([source, target], -1)` -> `extract_palette`
    -> `rgba_to_indexed` twice with the shared palette.  (H,W,4) or (B,H,W,4) int32 CUDA tensors.
    Returns (source_indexed (…,H,W,1), target_indexed (…,H,W,1), palette (…,256,4))."""
    src = require_cuda(from_any(source_image, name="source_image"), torch.int32, name="source_image")
    tgt = require_cuda(from_any(target_image, name="target_image"), torch.int32, name="target_image")
    if src.shape != tgt.shape:
        raise ValueError("source and target images must have the same shape")
    batched = src.dim() == 4
    if not batched:
        src, tgt = src.unsqueeze(0), tgt.unsqueeze(0)
    if src.dim() != 4 or src.shape[-1] != 4:
        raise ValueError(f"images must be (H,W,4) or (B,H,W,4), got {tuple(src.shape)}")
    if palette_ordering == "shuffled":
        # nondeterministic in the reference as well (io_utils.py:56-58): extract, permute, then index
        cat = torch.cat([src, tgt], dim=-1)
        palette = io_utils.extract_palette(cat, "shuffled", batched=True, check=True)
        s_idx = io_utils.rgba_to_indexed(src, palette)
        t_idx = io_utils.rgba_to_indexed(tgt, palette)
    else:
        b, h, w, _ = src.shape
        s_idx = torch.empty((b, h, w, 1), dtype=torch.int32, device=src.device)
        t_idx = torch.empty((b, h, w, 1), dtype=torch.int32, device=src.device)
        palette = torch.empty((b, MAX_PALETTE_SIZE, 4), dtype=torch.int32, device=src.device)
        ncolors = torch.empty((b,), dtype=torch.int32, device=src.device)
        if b:
            with torch.cuda.device(src.device):
                _lib.call("ph_load_indexed_images", ptr(src), ptr(tgt), b, h * w,
                          io_utils._ordering_id(palette_ordering), ptr(s_idx), ptr(t_idx), ptr(palette),
                          ptr(ncolors), stream_ptr(src.device))
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
