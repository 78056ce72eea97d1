"""Mirror of the reference's ``audio`` module (src/audio.rs) over the C ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import ffi


def max_waveform_samples(n_frame_max: int) -> int:
    """audio::max_waveform_samples (audio.rs:12-17)."""
    return int(ffi.lib().wb_max_waveform_samples(n_frame_max))


def prep_audio(waveform: np.ndarray, sample_rate: float = 16000.0, device: int = 0) -> np.ndarray:
    """audio::prep_audio (audio.rs:34-56): [n_batch, n_samples] f32 -> [n_batch, 80, n_samples/160].
    Raises WbError(WB_ERR_INVALID_ARG) where the reference panics (n_samples < 400, audio.rs:292)."""
    if sample_rate != 16000.0:
        raise ffi.WbError(ffi.WB_ERR_UNSUPPORTED, "only 16 kHz input is supported (the reference asserts it, transcribe/main.rs:41)")
    w = np.ascontiguousarray(waveform, dtype=np.float32)
    if w.ndim != 2:
        raise ffi.WbError(ffi.WB_ERR_INVALID_ARG, "prep_audio expects [n_batch, n_samples]")
    n_batch, n = w.shape
    out = np.empty((n_batch, 80, max(n // 160, 0)), dtype=np.float32)
    nf = C.c_int64(0)
    ffi.check(ffi.lib().wb_prep_audio(device, ffi.fptr(w), n_batch, n, ffi.fptr(out), C.byref(nf)))
    return out
