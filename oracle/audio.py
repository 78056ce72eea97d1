"""Oracle restatement of the log-mel frontend (reference: src/audio.rs, src/helper.rs).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every tensor op is PyTorch-CPU fp32 in the
same order the reference issues burn tensor ops, including the places where a "clean"
implementation would differ:

  * DFT angles are formed in f32 (audio.rs:349-356), so the twiddles are NOT exact.
  * max/min go through relu identities (helper.rs:8-22).
  * log10 = ln(x) / fl32(ln 10) (helper.rs:24-27).
  * the last STFT frame is dropped (audio.rs:42) and the max is global per call (audio.rs:50).
"""
from __future__ import annotations

import math

import numpy as np
import torch

N_FFT = 400        # audio.rs:5
HOP_LENGTH = 160   # audio.rs:6
N_MELS = 80        # audio.rs:7
WINDOW_LENGTH = N_FFT


def f32(x: float) -> float:
    """Round an f64 scalar to f32 (burn converts scalars to B::FloatElem = f32)."""
    return float(np.float32(x))


def max_waveform_samples(n_frame_max: int) -> int:
    """audio.rs:12-17."""
    n_samples_max = HOP_LENGTH * (n_frame_max + 1) + (N_FFT % 2)
    return n_samples_max - 1


# ---------------------------------------------------------------- helper.rs
def tensor_max_scalar(x: torch.Tensor, m: float) -> torch.Tensor:
    """helper.rs:8-10  relu(x - m) + m  (NOT bit-identical to max(x, m))."""
    m = f32(m)
    return torch.relu(x - m) + m


def tensor_max(x: torch.Tensor, m: torch.Tensor) -> torch.Tensor:
    """helper.rs:16-18."""
    return torch.relu(x - m) + m


def tensor_min(x: torch.Tensor, m: torch.Tensor) -> torch.Tensor:
    """helper.rs:20-22."""
    return -tensor_max(-x, -m)


def tensor_log10(x: torch.Tensor) -> torch.Tensor:
    """helper.rs:24-27: ln(x) / fl32(ln 10)."""
    return torch.log(x) / f32(math.log(10.0))


# ---------------------------------------------------------------- audio.rs
def hann_window() -> torch.Tensor:
    """audio.rs:272-278: sin^2(pi*i/N), periodic, f32."""
    i = torch.arange(0, WINDOW_LENGTH, dtype=torch.int64).to(torch.float32)
    return torch.pow(torch.sin(i * f32(math.pi / WINDOW_LENGTH)), 2.0)


def hz_to_mel(freq: float) -> float:
    """audio.rs:198-230 (Slaney, htk=false), f64 scalar."""
    f_min, f_sp = 0.0, 200.0 / 3.0
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = math.log(6.4) / 27.0
    if freq >= min_log_hz:
        return min_log_mel + math.log(freq / min_log_hz) / logstep
    return (freq - f_min) / f_sp


def mel_to_hz_tensor(mel: torch.Tensor) -> torch.Tensor:
    """audio.rs:232-266 (f32 tensor math, blend through a 0/1 mask)."""
    f_min, f_sp = 0.0, 200.0 / 3.0
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = math.log(6.4) / 27.0
    log_t = (mel >= f32(min_log_mel)).to(torch.float32)
    freq = log_t * (torch.exp((mel - f32(min_log_mel)) * f32(logstep)) * f32(min_log_hz)) \
        + (-log_t + 1.0) * (mel * f32(f_sp) + f32(f_min))
    return freq


def mel_frequencies(n_mels: int, fmin: float, fmax: float) -> torch.Tensor:
    """audio.rs:178-196."""
    min_mel, max_mel = hz_to_mel(fmin), hz_to_mel(fmax)
    mels = torch.arange(0, n_mels, dtype=torch.int64).to(torch.float32)
    mels = mels * f32((max_mel - min_mel) / (n_mels - 1)) + f32(min_mel)
    return mel_to_hz_tensor(mels)


def fft_frequencies(sample_rate: float, n_fft: int) -> torch.Tensor:
    """audio.rs:149-157."""
    return torch.arange(0, n_fft // 2 + 1, dtype=torch.int64).to(torch.float32) * f32(sample_rate / n_fft)


def get_mel_filters(sample_rate: float = 16000.0, n_fft: int = N_FFT, n_mels: int = N_MELS) -> torch.Tensor:
    """audio.rs:67-143 -> [n_mels, n_fft/2+1] f32 (Slaney norm)."""
    fmin, fmax = 0.0, sample_rate * 0.5
    fftfreqs = fft_frequencies(sample_rate, n_fft)
    n_fftfreqs = fftfreqs.shape[0]
    mel_f_size = n_mels + 2
    mel_f = mel_frequencies(mel_f_size, fmin, fmax)
    fdiff = mel_f[1:mel_f_size] - mel_f[0:mel_f_size - 1]
    ramps = mel_f.unsqueeze(1).repeat(1, n_fftfreqs) - fftfreqs.unsqueeze(0)
    lower = -ramps[0:n_mels] / fdiff[0:n_mels].unsqueeze(1)
    upper = ramps[2:2 + n_mels] / fdiff[1:1 + n_mels].unsqueeze(1)
    weights = torch.relu(tensor_min(lower, upper))
    enorm = torch.pow(mel_f[2:n_mels + 2] - mel_f[0:n_mels], -1.0) * 2.0
    return weights * enorm.unsqueeze(1)


def dft_basis() -> tuple[torch.Tensor, torch.Tensor]:
    """The two [201,400] matrices of audio.rs:349-364: cos(b)*w and sin(b)*(-w), where
    b[k][j] = fl32(fl32(k * fl32(2pi/400)) * j) -- angles in f32, up to ~1253 rad."""
    n_freq = N_FFT // 2 + 1
    window = hann_window()
    coe = f32(math.pi * 2.0 / N_FFT)
    kk = torch.arange(0, n_freq, dtype=torch.int64).to(torch.float32) * coe
    b = kk.unsqueeze(1).repeat(1, N_FFT) * torch.arange(0, N_FFT, dtype=torch.int64).to(torch.float32).unsqueeze(0)
    return torch.cos(b) * window.unsqueeze(0), torch.sin(b) * (-window).unsqueeze(0)


def stfft(inp: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """audio.rs:284-367.  inp [B, n] f32 -> (re, im) each [B, 201, n/160 + 1]."""
    n_batch, n = inp.shape
    assert n >= N_FFT, "audio.rs:292 assert!(orig_input_size >= n_fft)"
    pad = N_FFT // 2
    left = torch.flip(inp[:, 1:pad + 1], dims=[1])               # audio.rs:298
    right = torch.flip(inp[:, n - pad - 1:n - 1], dims=[1])      # audio.rs:299-305
    x = torch.cat([left, inp, right], dim=1)
    input_size = x.shape[1]
    n_frame = (input_size - N_FFT) // HOP_LENGTH + 1             # audio.rs:327
    # audio.rs:331-346 builds input_windows[b, j, t] = x[b, t*hop + j]
    idx = torch.arange(N_FFT).unsqueeze(1) + HOP_LENGTH * torch.arange(n_frame).unsqueeze(0)
    input_windows = x[:, idx]                                    # [B, 400, n_frame]
    cw, sw = dft_basis()
    real = torch.matmul(cw.unsqueeze(0), input_windows)
    imag = torch.matmul(sw.unsqueeze(0), input_windows)
    return real, imag


def prep_audio(waveform: torch.Tensor, sample_rate: float = 16000.0) -> torch.Tensor:
    """audio.rs:34-56.  [B, n] f32 -> [B, 80, n/160] f32."""
    waveform = waveform.to(torch.float32)
    re, im = stfft(waveform)
    magnitudes = torch.pow(re, 2.0) + torch.pow(im, 2.0)
    magnitudes = magnitudes[:, :, :magnitudes.shape[2] - 1]
    mel_spec = torch.matmul(get_mel_filters(sample_rate).unsqueeze(0), magnitudes)
    log_spec = tensor_log10(tensor_max_scalar(mel_spec, 1.0e-10))
    mx = float(log_spec.max())
    log_spec = tensor_max_scalar(log_spec, mx - 8.0)
    return (log_spec + 4.0) / 4.0


def prep_audio_f64(waveform: np.ndarray) -> np.ndarray:
    """Sanity reference: exact-twiddle f64 STFT (numpy rfft) + f64 Slaney filterbank.
    NOT the contract -- used to report how far the reference's f32-angle DFT is from an
    exact FFT (hard part 2)."""
    x = np.asarray(waveform, dtype=np.float64)
    n = x.shape[-1]
    xp = np.pad(x, [(0, 0)] * (x.ndim - 1) + [(200, 200)], mode="reflect")
    n_frame = n // 160 + 1
    idx = np.arange(400)[None, :] + 160 * np.arange(n_frame)[:, None]
    w = np.sin(np.pi * np.arange(400) / 400.0) ** 2
    fr = xp[..., idx] * w
    spec = np.fft.rfft(fr, axis=-1)
    power = (spec.real ** 2 + spec.imag ** 2)[..., :-1, :]
    power = np.swapaxes(power, -1, -2)
    filt = get_mel_filters().double().numpy()
    mel = filt @ power
    ls = np.log10(np.maximum(mel, 1e-10))
    ls = np.maximum(ls, ls.max() - 8.0)
    return (ls + 4.0) / 4.0
