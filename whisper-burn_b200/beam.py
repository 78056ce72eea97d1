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


def beam_search_table(table, first_token: int, eot: int, beam_size: int, max_depth: int) -> list[int]:
    """beam::beam_search (beam.rs:9-37) in the library's C++ host code (host/beam.hpp) over a table-driven `next`:
    log-prob of token v after a beam ending in t with length n = table[(t * 131 + n) % n_ctx][v]."""
    import ctypes as C
    t = np.ascontiguousarray(table, dtype=np.float64)
    out = np.zeros(max_depth + 2, dtype=np.int64)
    n = ffi.lib().wb_beam_search_table(t.ctypes.data_as(C.POINTER(C.c_double)), t.shape[0], t.shape[1], first_token, eot, beam_size,
                                      max_depth, ffi.i64ptr(out), out.shape[0])
    if n < 0:
        raise ffi.WbError(ffi.WB_ERR_INVALID_ARG, "beam_search_table: bad arguments")
    return [int(v) for v in out[:n]]
