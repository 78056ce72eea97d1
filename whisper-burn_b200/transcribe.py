"""Mirror of the reference's ``transcribe`` module (src/transcribe.rs), token side, over the C ABI.
Detokenisation (src/token.rs) stays with the caller: the library needs 5 ids and a bitmap."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import ffi
from .model import Whisper


class Session:
    """KV-cached decoding session (wb_session): encoder output, cross/self K/V, workspaces."""

    def __init__(self, whisper: Whisper, max_windows: int, max_beams: int = 5, max_text_len: int = 104,
                 kv_dtype: int = ffi.WB_KV_F32):
        self.whisper = whisper
        self.max_windows = max_windows
        self._h = C.c_void_p()
        ffi.check(ffi.lib().wb_session_create(whisper.handle, max_windows, max_beams, max_text_len, kv_dtype, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            ffi.lib().wb_session_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- encode
    def encode_waveforms(self, waves: Sequence[np.ndarray]) -> None:
        ws = [np.ascontiguousarray(w, dtype=np.float32) for w in waves]
        ptrs = (ffi._F * len(ws))(*[ffi.fptr(w) for w in ws])
        lens = np.asarray([w.shape[0] for w in ws], dtype=np.int64)
        ffi.check(ffi.lib().wb_session_encode_waveforms(self._h, ptrs, ffi.i64ptr(lens), len(ws)))

    def encode_mels(self, mels: np.ndarray) -> None:
        mels = np.ascontiguousarray(mels, dtype=np.float32)
        n, n_mels, n_ctx = mels.shape
        ffi.check(ffi.lib().wb_session_encode_mels(self._h, ffi.fptr(mels), n, n_mels, n_ctx))

    def get_mel(self, window: int) -> np.ndarray:
        cap = 80 * (self.whisper.config.n_audio_ctx + 2)
        buf = np.empty(cap, dtype=np.float32)
        n = C.c_int64(0)
        ffi.check(ffi.lib().wb_session_get_mel(self._h, window, ffi.fptr(buf), cap, C.byref(n)))
        return buf[:80 * n.value].reshape(80, n.value).copy()

    def get_encoder_output(self, window: int) -> np.ndarray:
        d = self.whisper.config.n_audio_state
        cap = d * self.whisper.config.n_audio_ctx
        buf = np.empty(cap, dtype=np.float32)
        n = C.c_int64(0)
        ffi.check(ffi.lib().wb_session_get_encoder_output(self._h, window, ffi.fptr(buf), cap, C.byref(n)))
        return buf[:d * n.value].reshape(n.value, d).copy()

    # ---- decode
    def begin(self, prompt: Sequence[int]) -> None:
        p = np.asarray(prompt, dtype=np.int64)
        ffi.check(ffi.lib().wb_session_begin(self._h, ffi.i64ptr(p), len(p)))

    def step(self, window_of_row, parent_row, token, apply_special_mask: bool, is_special: Optional[np.ndarray], k: int):
        w = np.asarray(window_of_row, dtype=np.int32)
        pr = np.asarray(parent_row, dtype=np.int32)
        tk = np.asarray(token, dtype=np.int64)
        n = len(w)
        ids = np.empty((n, k), dtype=np.int64)
        lps = np.empty((n, k), dtype=np.float32)
        sp = ffi.u8ptr(np.ascontiguousarray(is_special, dtype=np.uint8)) if is_special is not None else None
        ffi.check(ffi.lib().wb_session_step(self._h, n, ffi.i32ptr(w), ffi.i32ptr(pr), ffi.i64ptr(tk),
                                           1 if apply_special_mask else 0, sp, k, ffi.i64ptr(ids), ffi.fptr(lps)))
        return ids, lps

    # ---- pipelines
    def transcribe_windows(self, waves: Sequence[np.ndarray], special, is_special: np.ndarray, beam_size: int = 5,
                           max_depth: int = 100) -> List[List[int]]:
        """mels_to_text for a batch of windows (transcribe.rs:148-383), ids only."""
        ws = [np.ascontiguousarray(w, dtype=np.float32) for w in waves]
        ptrs = (ffi._F * len(ws))(*[ffi.fptr(w) for w in ws])
        lens = np.asarray([w.shape[0] for w in ws], dtype=np.int64)
        cap = 4 + max_depth + 1
        out = np.zeros((len(ws), cap), dtype=np.int64)
        out_len = np.zeros(len(ws), dtype=np.int64)
        ids = ffi.SpecialIds(special.sot, special.lang, special.transcribe, special.notimestamps, special.eot)
        sp = np.ascontiguousarray(is_special, dtype=np.uint8)
        ffi.check(ffi.lib().wb_transcribe_windows(self._h, ptrs, ffi.i64ptr(lens), len(ws), beam_size, max_depth,
                                                 C.byref(ids), ffi.u8ptr(sp), ffi.i64ptr(out), cap, ffi.i64ptr(out_len)))
        return [[int(t) for t in out[i, :out_len[i]]] for i in range(len(ws))]

    def transcribe_windows_dev(self, wave_dev_ptr: int, offsets, lens, special, is_special: np.ndarray,
                               beam_size: int = 5, max_depth: int = 100) -> List[List[int]]:
        """Same, windows already resident in HBM (device pointer + element offsets)."""
        offs = np.asarray(offsets, dtype=np.int64)
        ln = np.asarray(lens, dtype=np.int64)
        cap = 4 + max_depth + 1
        out = np.zeros((len(ln), cap), dtype=np.int64)
        out_len = np.zeros(len(ln), dtype=np.int64)
        ids = ffi.SpecialIds(special.sot, special.lang, special.transcribe, special.notimestamps, special.eot)
        sp = np.ascontiguousarray(is_special, dtype=np.uint8)
        ffi.check(ffi.lib().wb_transcribe_windows_dev(self._h, C.c_void_p(wave_dev_ptr), ffi.i64ptr(offs), ffi.i64ptr(ln),
                                                     len(ln), beam_size, max_depth, C.byref(ids), ffi.u8ptr(sp),
                                                     ffi.i64ptr(out), cap, ffi.i64ptr(out_len)))
        return [[int(t) for t in out[i, :out_len[i]]] for i in range(len(ln))]

    def waveform_to_tokens(self, waveform: np.ndarray, special, is_special: np.ndarray, sample_rate: int = 16000,
                           beam_size: int = 5, max_depth: int = 100) -> List[int]:
        w = np.ascontiguousarray(waveform, dtype=np.float32)
        cap = (len(w) // 1000 + 2) * (4 + max_depth + 1) + 16
        out = np.zeros(cap, dtype=np.int64)
        n = C.c_int64(0)
        ids = ffi.SpecialIds(special.sot, special.lang, special.transcribe, special.notimestamps, special.eot)
        sp = np.ascontiguousarray(is_special, dtype=np.uint8)
        ffi.check(ffi.lib().wb_waveform_to_tokens(self._h, ffi.fptr(w), len(w), sample_rate, beam_size, max_depth,
                                                 C.byref(ids), ffi.u8ptr(sp), ffi.i64ptr(out), cap, C.byref(n)))
        return [int(t) for t in out[:n.value]]

    def waveforms_to_tokens(self, waveforms: Sequence[np.ndarray], special, is_special: np.ndarray, sample_rate: int = 16000,
                            beam_size: int = 5, max_depth: int = 100) -> List[List[int]]:
        """Batched waveform_to_tokens: all windows of all waveforms decoded together (wb_waveforms_to_tokens)."""
        ws = [np.ascontiguousarray(w, dtype=np.float32) for w in waveforms]
        cap = (max(len(w) for w in ws) // 1000 + 2) * (4 + max_depth + 1) + 16
        out = np.zeros((len(ws), cap), dtype=np.int64)
        n = np.zeros(len(ws), dtype=np.int64)
        ptrs = (C.c_void_p * len(ws))(*[w.ctypes.data for w in ws])
        lens = np.array([len(w) for w in ws], dtype=np.int64)
        ids = ffi.SpecialIds(special.sot, special.lang, special.transcribe, special.notimestamps, special.eot)
        sp = np.ascontiguousarray(is_special, dtype=np.uint8)
        ffi.check(ffi.lib().wb_waveforms_to_tokens(self._h, ptrs, ffi.i64ptr(lens), len(ws), sample_rate, beam_size, max_depth,
                                                  C.byref(ids), ffi.u8ptr(sp), ffi.i64ptr(out), cap, ffi.i64ptr(n)))
        return [[int(t) for t in out[i, :n[i]]] for i in range(len(ws))]

    def last_decoder(self) -> int:
        return int(ffi.lib().wb_session_last_decoder(self._h))

    def last_timings_ms(self):
        buf = np.zeros(4, dtype=np.float32)
        ffi.check(ffi.lib().wb_session_last_timings(self._h, ffi.fptr(buf)))
        return {"logmel": float(buf[0]), "encoder": float(buf[1]), "decode": float(buf[2]), "total": float(buf[3])}

    def profile_decode(self, special, n_steps: int = 50):
        """(average logits-GEMV ms, average whole-step ms) over n_steps re-run greedy steps."""
        ids = ffi.SpecialIds(special.sot, special.lang, special.transcribe, special.notimestamps, special.eot)
        a, b = C.c_float(0), C.c_float(0)
        ffi.check(ffi.lib().wb_session_profile_decode(self._h, C.byref(ids), n_steps, C.byref(a), C.byref(b)))
        return float(a.value), float(b.value)

    def last_steps(self) -> int:
        n = C.c_int64(0)
        ffi.check(ffi.lib().wb_session_last_steps(self._h, C.byref(n)))
        return int(n.value)


def waveform_to_text(whisper: Whisper, special, is_special: np.ndarray, waveform: np.ndarray, sample_rate: int = 16000,
                     beam_size: int = 5, max_depth: int = 100, max_windows: int = 8) -> List[int]:
    """transcribe::waveform_to_text (transcribe.rs:23-74) without detokenisation: merged token ids.
    (`bpe` and `lang` of the reference signature collapse into `special` / `is_special`.)"""
    s = Session(whisper, max_windows=max_windows, max_beams=max(beam_size, 1), max_text_len=4 + max_depth + 1)
    try:
        return s.waveform_to_tokens(waveform, special, is_special, sample_rate, beam_size, max_depth)
    finally:
        s.close()


def window_bounds(n_samples: int, sample_rate: int, window_len: int):
    n = int(ffi.lib().wb_window_count(n_samples, sample_rate, window_len))
    st = np.zeros(n, dtype=np.int64)
    en = np.zeros(n, dtype=np.int64)
    ffi.check(ffi.lib().wb_window_bounds(n_samples, sample_rate, window_len, ffi.i64ptr(st), ffi.i64ptr(en)))
    return [(int(a), int(b)) for a, b in zip(st, en)]


def find_chunk_overlap(prev_tokens, curr_tokens, max_n_offsets: int, min_n_overlaps: int):
    p = np.asarray(prev_tokens, dtype=np.int64)
    c = np.asarray(curr_tokens, dtype=np.int64)
    pi, ci = C.c_int64(0), C.c_int64(0)
    found = ffi.lib().wb_find_chunk_overlap(ffi.i64ptr(p), len(p), ffi.i64ptr(c), len(c), max_n_offsets, min_n_overlaps,
                                            C.byref(pi), C.byref(ci))
    return (pi.value, ci.value) if found else None


def first_repetition_end(tokens, period: int) -> int:
    """transcribe.rs:385-393 (ValueError where the reference's usize arithmetic underflows)."""
    t = np.asarray(tokens, dtype=np.int64)
    r = int(ffi.lib().wb_first_repetition_end(ffi.i64ptr(t), len(t), period))
    if r < 0:
        raise ValueError("first_repetition_end: period exceeds the token count")
    return r


def repetition_period(tokens, min_repetitions: int):
    """transcribe.rs:395-419: the period, or None."""
    t = np.asarray(tokens, dtype=np.int64)
    r = int(ffi.lib().wb_repetition_period(ffi.i64ptr(t), len(t), min_repetitions))
    if r < 0:
        raise ValueError("repetition_period: invalid argument")
    return r if r > 0 else None


def find_repeated_tokens_index(tokens, window_size: int, min_repeat_count: int):
    """transcribe.rs:421-447: (index of the first repeat, index of the second) or None (ValueError where the reference panics)."""
    t = np.asarray(tokens, dtype=np.int64)
    a, b = C.c_int64(0), C.c_int64(0)
    r = int(ffi.lib().wb_find_repeated_tokens_index(ffi.i64ptr(t), len(t), window_size, min_repeat_count, C.byref(a), C.byref(b)))
    if r < 0:
        raise ValueError("find_repeated_tokens_index: the reference unwraps a repeat that does not exist")
    return (a.value, b.value) if r else None
