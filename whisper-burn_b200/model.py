"""Mirror of the reference's ``model`` module (src/model/mod.rs, src/model/load.rs) over the C ABI."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import ffi


@dataclass(frozen=True)
class WhisperConfig:
    """WhisperConfig (mod.rs:16-39) = AudioEncoderConfig (:164-171) + TextDecoderConfig (:73-80)."""
    n_mels: int
    n_audio_ctx: int
    n_audio_state: int
    n_audio_head: int
    n_audio_layer: int
    n_vocab: int
    n_text_ctx: int
    n_text_state: int
    n_text_head: int
    n_text_layer: int

    def to_c(self) -> ffi.Dims:
        return ffi.Dims(*[getattr(self, f) for f, _ in ffi.Dims._fields_])


class Whisper:
    """model::Whisper (mod.rs:41-71).  ``tensors`` uses the reference's npy-tree paths
    (load.rs / dump.py), Linear weights in burn layout [d_in, d_out]."""

    def __init__(self, config, tensors: dict, device: int = 0, ln_eps_outside: bool = True):
        self.config = config
        self._h = C.c_void_p()
        dims = config.to_c() if hasattr(config, "to_c") else ffi.Dims(*[getattr(config, f) for f, _ in ffi.Dims._fields_])
        L = ffi.lib()
        ffi.check(L.wb_model_create(C.byref(dims), device, C.byref(self._h)))
        try:
            for path, arr in tensors.items():
                a = np.ascontiguousarray(np.asarray(arr, dtype=np.float32))
                shape = np.asarray(a.shape if a.ndim else (1,), dtype=np.int64)
                ffi.check(L.wb_model_set_tensor(self._h, path.encode(), ffi.fptr(a.reshape(-1)), ffi.i64ptr(shape), len(shape)))
            ffi.check(L.wb_model_set_layernorm_eps_mode(self._h, 1 if ln_eps_outside else 0))
            ffi.check(L.wb_model_finalize(self._h))
        except Exception:
            L.wb_model_destroy(self._h)
            self._h = None
            raise

    @classmethod
    def from_npy_tree(cls, directory, device: int = 0, ln_eps_outside: bool = True) -> "Whisper":
        """model::load::load_whisper (load.rs:295-310): the npy tree python/dump.py writes."""
        dims = ffi.Dims()
        ffi.check(ffi.lib().wb_npy_tree_probe(str(directory).encode(), C.byref(dims)))
        self = cls.__new__(cls)
        self.config = WhisperConfig(*[getattr(dims, f) for f, _ in ffi.Dims._fields_])
        self._h = C.c_void_p()
        ffi.check(ffi.lib().wb_model_load_npy_tree(str(directory).encode(), device, 1 if ln_eps_outside else 0, C.byref(self._h)))
        return self

    def close(self):
        if getattr(self, "_h", None):
            ffi.lib().wb_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    @property
    def weights_fp16_exact(self) -> bool:
        return bool(ffi.lib().wb_model_weights_fp16_exact(self._h))

    def encoder_ctx_size(self) -> int:   # mod.rs:64-66
        return self.config.n_audio_ctx

    def decoder_ctx_size(self) -> int:   # mod.rs:68-70
        return self.config.n_text_ctx

    def forward_encoder(self, mel: np.ndarray) -> np.ndarray:
        """Whisper::forward_encoder (mod.rs:52-54): [B, 80, Tm] -> [B, (Tm-1)//2+1, d]."""
        mel = np.ascontiguousarray(mel, dtype=np.float32)
        if mel.ndim != 3:
            raise ffi.WbError(ffi.WB_ERR_INVALID_ARG, "forward_encoder expects [n_batch, n_mels, n_ctx]")
        b, n_mels, n_ctx = mel.shape
        out = np.empty((b, (max(n_ctx, 1) - 1) // 2 + 1, self.config.n_audio_state), dtype=np.float32)
        ffi.check(ffi.lib().wb_forward_encoder(self._h, ffi.fptr(mel), b, n_mels, n_ctx, ffi.fptr(out)))
        return out

    def forward_decoder(self, tokens: np.ndarray, encoder_output: np.ndarray) -> np.ndarray:
        """Whisper::forward_decoder (mod.rs:56-62), stateless: [nb, t] i64, [nb, T, d] -> [nb, t, V]."""
        tokens = np.ascontiguousarray(tokens, dtype=np.int64)
        enc = np.ascontiguousarray(encoder_output, dtype=np.float32)
        if tokens.ndim != 2 or enc.ndim != 3 or enc.shape[0] != tokens.shape[0]:
            raise ffi.WbError(ffi.WB_ERR_INVALID_ARG, "forward_decoder expects tokens [nb, t] and encoder_output [nb, T, d]")
        nb, t = tokens.shape
        out = np.empty((nb, t, self.config.n_vocab), dtype=np.float32)
        ffi.check(ffi.lib().wb_forward_decoder(self._h, ffi.i64ptr(tokens), nb, t, ffi.fptr(enc), enc.shape[1], ffi.fptr(out)))
        return out

    def forward(self, mel: np.ndarray, tokens: np.ndarray) -> np.ndarray:   # mod.rs:48-50
        return self.forward_decoder(tokens, self.forward_encoder(mel))
