"""Writer / prober for the reference's model-file format, the npy tree of python/dump.py:120-213 (read by
src/model/load.rs:19-310 and by wb_model_load_npy_tree): one f32 .npy per tensor with payload
[dims..., values...]; scalars as [1.0, value].  Used by the tests to round-trip synthetic weights."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import ffi
from .model import WhisperConfig


def _save_tensor(path: Path, arr: np.ndarray) -> None:          # dump.py:134-139 save_tensor
    path.parent.mkdir(parents=True, exist_ok=True)
    a = np.asarray(arr, dtype=np.float32)
    np.save(path.with_suffix(".npy"), np.concatenate([np.array(a.shape, dtype=np.float32), a.reshape(-1)]).astype(np.float32))


def _save_scalar(path: Path, v: float) -> None:                  # dump.py:130-132 save_scalar
    path.parent.mkdir(parents=True, exist_ok=True)
    np.save(path.with_suffix(".npy"), np.array([1.0, float(v)], dtype=np.float32))


def save_npy_tree(directory, dims, tensors: dict) -> None:
    """`tensors` uses the same path keys as Whisper(...): they ARE the tree paths."""
    root = Path(directory)
    for key, arr in tensors.items():
        if key.endswith("/eps"):
            _save_scalar(root / key, float(np.asarray(arr).reshape(-1)[-1]))
        else:
            _save_tensor(root / key, arr)
    _save_scalar(root / "encoder/n_layer", dims.n_audio_layer)        # dump.py:189-191
    _save_scalar(root / "encoder/n_mels", dims.n_mels)
    _save_scalar(root / "encoder/n_audio_state", dims.n_audio_state)
    _save_scalar(root / "decoder/n_layer", dims.n_text_layer)         # dump.py:199
    for i in range(dims.n_audio_layer):
        _save_scalar(root / f"encoder/block_{i}/attn/n_head", dims.n_audio_head)      # dump.py:168
    for i in range(dims.n_text_layer):
        _save_scalar(root / f"decoder/block_{i}/attn/n_head", dims.n_text_head)
        _save_scalar(root / f"decoder/block_{i}/cross_attn/n_head", dims.n_text_head)


def probe(directory) -> WhisperConfig:
    dims = ffi.Dims()
    ffi.check(ffi.lib().wb_npy_tree_probe(str(directory).encode(), C.byref(dims)))
    return WhisperConfig(*[getattr(dims, f) for f, _ in ffi.Dims._fields_])
