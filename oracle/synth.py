"""Synthetic inputs for the oracle: re-exports the shared generator (whisper-burn_b200/synth.py,
pure numpy input generation) and adds the torch view of the weights the oracle computes with."""
from __future__ import annotations

import numpy as np
import torch

import wb200  # noqa: F401  (registers whisper_burn_b200)
from whisper_burn_b200.synth import (MODEL_DIMS, SpecialTokens, chunk_waveform, make_weights_np,  # noqa: F401
                                     special_tokens, waveform)


def to_torch(w_np: dict) -> dict:
    return {k: torch.from_numpy(np.asarray(v, dtype=np.float32)) for k, v in w_np.items()}


def make_weights(model_name_or_dims, seed: int = 0):
    dims = MODEL_DIMS[model_name_or_dims] if isinstance(model_name_or_dims, str) else model_name_or_dims
    w_np = make_weights_np(dims, seed)
    return dims, w_np, to_torch(w_np)
