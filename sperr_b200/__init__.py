"""sperr_b200 -- B200-native SPERR hot path. The product is sperr_b200/libsperr_b200.so (C ABI in
include/sperr_b200.h, hand-written sm_100a CUDA); this package is only a thin ctypes face over it
for Python callers, the tests and bench.py. There is no CPU execution path: loading fails loudly
when the library has not been built, and every compute call returns -1 without a CUDA device."""
from .api import (Library, load, compress_3d, decompress_3d, parse_header,  # noqa: F401
                  MODE_BPP, MODE_PSNR, MODE_PWE)
