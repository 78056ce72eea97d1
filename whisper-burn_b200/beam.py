"""Mirror of the reference's ``beam`` module (src/beam.rs): the search itself runs in the library's
C++ host code (host/beam.hpp); this exposes its selection primitive for tie-break tests."""
from __future__ import annotations

import numpy as np

from . import ffi


def get_top_elements(scores, num: int) -> list[int]:
    """beam::get_top_elements (beam.rs:81-110) over f64 scores: kept indices, ascending score."""
    s = np.ascontiguousarray(scores, dtype=np.float64)
    out = np.empty(max(num, 1), dtype=np.int64)
    import ctypes as C
    n = ffi.lib().wb_beam_get_top_elements(s.ctypes.data_as(C.POINTER(C.c_double)), s.shape[0], num, ffi.i64ptr(out))
    if n < 0:
        raise ffi.WbError(ffi.WB_ERR_INVALID_ARG, "get_top_elements: bad arguments")
    return [int(v) for v in out[:n]]
