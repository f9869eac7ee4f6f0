"""B200-native colour kernels behind the call signatures of fegemo/palette-and-histo-gan.

    from palette_and_histo_gan_b200 import histogram, io_utils, dataset_utils

`histogram`, `io_utils` and `dataset_utils` mirror the reference modules of the same names for the
per-pixel colour path (RGB-uv histogram + Hellinger loss forward/backward, palette extraction,
colour indexing, one-hot).  Everything computes in hand-written sm_100a CUDA inside
`libpalhist.so` (C ABI: include/palhist.h); importing this package without the built library
raises ImportError — there is no CPU or framework fallback.
"""
from . import _lib

_lib.load()  # fail loudly at import time if the CUDA library is missing

from . import configuration, dataset_utils, histogram, hostapi, io_utils  # noqa: E402

__all__ = ["configuration", "dataset_utils", "histogram", "hostapi", "io_utils"]
__version__ = "0.1.0"
