"""ctypes binding of libpalhist.so (C ABI declared in include/palhist.h).

The product path has no CPU or framework fallback: if the shared library is missing, or a compute
entry point is called without a CUDA device, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PALHIST_LIB: another build of the same library (kernel experiments: tools/gpu_r2_variants.sh times several
# builds on one box); it must export the same ABI, which _load() checks
LIB_PATH = os.environ.get("PALHIST_LIB") or os.path.join(_HERE, "libpalhist.so")

PH_OK = 0
PH_ERR_INVALID, PH_ERR_CUDA, PH_ERR_UNSUPPORTED = -1, -2, -3
METHODS = {"inverse-quadratic": 0, "RBF": 1}
ORDERINGS = {"top2bottom": 0, "bottom2top": 1, "grayness": 2, "shuffled": 3}
MAP_OPS = {"blacken": 0, "normalize": 1, "denormalize": 2}
ASYNC_RANGE = 1
ASYNC_MIRROR = 2  # a forward launched with PH_IMPL_MIRROR found bin centres that are not antisymmetric
ABI_VERSION = 3
COMM_HANDLE_BYTES = 64
INDEX_MODES = {"exact": 0, "nearest": 1}
IMPLS = {"auto": 0, "simt": 1, "tc": 2}
PALETTE_BAD_VALUE = -1


class PalHistError(RuntimeError):
    """A libpalhist entry point returned a non-zero status."""

    def __init__(self, name, code, message):
        super().__init__(f"{name} failed with status {code}: {message}")
        self.code = code


_p = C.c_void_p
_i64 = C.c_int64
_int = C.c_int
_f = C.c_float
_sz = C.c_size_t

# name -> (restype, argtypes); mirrors include/palhist.h one to one (tests/test_abi.py checks it)
PROTOTYPES = {
    "ph_abi_version": (_int, []),
    "ph_last_error": (C.c_char_p, []),
    "ph_launch_count": (_i64, []),
    "ph_reset_launch_count": (None, []),
    "ph_async_status": (_int, [_int, _int]),
    "ph_device_info": (_int, [_int, C.POINTER(_int), C.POINTER(_int), C.POINTER(_int)]),
    "ph_hist256_plan": (_int, [_i64, _i64, C.POINTER(_i64)]),
    "ph_hist_workspace_bytes": (_sz, [_i64, _i64, _int, _int]),
    "ph_hist_forward": (_int, [_p, _i64, _i64, _int, _p, _int, _int, _f, _f, _p, _p, _p, _sz, _int, _p]),
    "ph_hist_forward_ssum": (_int, [_p, _i64, _i64, _int, _p, _int, _int, _f, _f, _p, _p, _p, _p, _int, _p, _sz, _int,
                                    _p]),
    "ph_component_histogram": (_int, [_p, _p, _p, _p, _i64, _i64, _p, _int, _int, _f, _f, _p, _p]),
    "ph_hist_backward": (_int, [_p, _i64, _i64, _int, _p, _int, _int, _f, _f, _p, _p, _p, _p, _p, _i64, _p,
                                _p, _p, _sz, _int, _p]),
    "ph_hellinger_ssum": (_int, [_p, _p, _i64, _p, _p]),
    "ph_hellinger_finish": (_int, [_p, _i64, _p, _p]),
    "ph_hellinger_backward": (_int, [_p, _p, _i64, _p, _i64, _p, _p, _p, _p]),
    "ph_mean_abs_or_sq_diff": (_int, [_p, _p, _i64, _int, _p, _p]),
    "ph_extract_palette": (_int, [_p, _i64, _i64, _int, _p, _p, _p, _p]),
    "ph_rgba_to_indexed": (_int, [_p, _i64, _i64, _p, _i64, _int, _p, _p, _int, _p]),
    "ph_one_hot": (_int, [_p, _i64, _int, _p, _p]),
    "ph_u8_to_float_image": (_int, [_p, _i64, _int, _int, _p, _p]),
    "ph_augment_pair": (_int, [_p, _p, _i64, _int, _int, _p, _p, _p, _int, _p, _p, _p]),
    "ph_indexed_to_rgba": (_int, [_p, _i64, _i64, _p, _i64, _int, _int, _p, _p]),
    "ph_argmax_indexed": (_int, [_p, _i64, _i64, _int, _p, _i64, _int, _p, _p, _p]),
    "ph_load_indexed_images": (_int, [_p, _p, _i64, _i64, _int, _p, _p, _p, _p, _p, _p]),
    "ph_load_indexed_images_u8": (_int, [_p, _p, _i64, _i64, _int, _p, _p, _p, _p, _p, _p]),
    "ph_pixel_map": (_int, [_p, _i64, _int, _p, _p]),
    "ph_comm_create": (_int, [_int, _int, _int, C.POINTER(_p)]),
    "ph_comm_export": (_int, [_p, _p]),
    "ph_comm_connect": (_int, [_p, _p]),
    "ph_comm_allreduce_sum_f64": (_int, [_p, _p, _int, _p]),
    "ph_comm_destroy": (None, [_p]),
    "ph_host_ctx_create": (_int, [_int, C.POINTER(_p)]),
    "ph_host_ctx_destroy": (None, [_p]),
    "ph_host_hist_loss": (_int, [_p, _p, _p, _i64, _i64, _int, _p, _int, _int, _f, _f, _int, _p, _p]),
    "ph_host_hist_begin": (_int, [_p, _p, _p, _i64, _i64, _int, _p, _int, _int, _f, _f, _int, _p]),
    "ph_host_hist_begin_u8real": (_int, [_p, _p, _p, _i64, _i64, _p, _int, _int, _f, _f, _int, _p]),
    "ph_host_hist_finish": (_int, [_p, C.c_double, _i64, _p, _p, _p]),
    "ph_host_hist_finish_comm": (_int, [_p, _p, _i64, _p, _p, _p]),
    "ph_host_hist_loss_sharded": (_int, [_p, _p, _p, _int, _p, _i64, _i64, _int, _p, _int, _int, _f, _f, _int, _i64, _p, _p,
                                         _p]),
    "ph_host_load_indexed_images": (_int, [_p, _p, _p, _i64, _i64, _int, _p, _p, _p, _p, _p, _p]),
    "ph_host_load_indexed_images_u8": (_int, [_p, _p, _p, _i64, _i64, _int, _p, _p, _p, _p, _p, _p]),
}

_lib = None


def load():
    """Load libpalhist.so (once). Raises ImportError with build instructions if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"or `make -C {os.path.join(_HERE, 'csrc')}`. There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.ph_abi_version() != ABI_VERSION:
        raise ImportError(f"libpalhist ABI version {lib.ph_abi_version()} != {ABI_VERSION}; rebuild the library")
    _lib = lib
    return lib


def last_error() -> str:
    return load().ph_last_error().decode("utf-8", "replace")


def call(name, *args):
    """Call a status-returning entry point; raise PalHistError (or ValueError for bad arguments)."""
    rc = getattr(load(), name)(*args)
    if rc != PH_OK:
        msg = last_error()
        if rc == PH_ERR_INVALID:
            raise ValueError(f"{name}: {msg}")
        raise PalHistError(name, rc, msg)


def async_status(device: int = -1, clear: bool = False) -> int:
    """Sticky asynchronous status word of `device` (-1 = current): bit ASYNC_RANGE = a tensor-core histogram
    forward met pixels outside its operand range.  Never synchronises; as current as the last finished kernel."""
    return int(load().ph_async_status(int(device), 1 if clear else 0))


def launch_count() -> int:
    return int(load().ph_launch_count())


def reset_launch_count() -> None:
    load().ph_reset_launch_count()
