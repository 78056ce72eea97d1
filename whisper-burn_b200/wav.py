"""load_audio_waveform of the reference's transcribe binary (src/bin/transcribe/main.rs:31-55) over wb_load_wav."""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import numpy as np

from . import ffi


def load_audio_waveform(filename: str, strict: bool = True) -> Tuple[np.ndarray, int]:
    """Returns (samples f32, sample_rate).  strict=True keeps the reference's asserts: 16 kHz, single channel
    (main.rs:43-44) -> WbError(WB_ERR_INVALID_ARG).  Integer PCM is scaled by 1 / (2^(bits-1) - 1) (main.rs:46-53)."""
    n = C.c_int64(0)
    sr = C.c_int64(0)
    ch = C.c_int(0)
    path = str(filename).encode()
    ffi.check(ffi.lib().wb_load_wav(path, 1 if strict else 0, None, 0, C.byref(n), C.byref(sr), C.byref(ch)))
    out = np.zeros(n.value, dtype=np.float32)
    ffi.check(ffi.lib().wb_load_wav(path, 1 if strict else 0, ffi.fptr(out), n.value, C.byref(n), C.byref(sr), C.byref(ch)))
    if ch.value > 1:
        out = out.reshape(-1, ch.value)
    return out, int(sr.value)
