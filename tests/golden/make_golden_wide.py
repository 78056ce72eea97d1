"""Golden token fixtures for the batched tensor-core decoder tests (models test-c / test-d / test-e), from the CPU oracle.
Same status as make_golden.py: oracle outputs (the reference ships no vectors: "parity unpinned"); they pin the oracle and
let the GPU tests compare against committed ids as well as against the live oracle.  Run: python tests/golden/make_golden_wide.py"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from oracle import audio, model, synth, transcribe  # noqa: E402

OUT = Path(__file__).resolve().parent


def tokens(w_t, dims, sp, wave, beam, depth, kv="f32"):
    mel = audio.prep_audio(torch.from_numpy(wave)[None])
    return transcribe.mels_to_tokens(w_t, dims, sp, mel, beam_size=beam, max_depth=depth, opts=model.OracleOptions(kv_dtype=kv))


def main():
    out = {}
    dims, _, w_t = synth.make_weights("test-c", seed=0)
    sp = synth.special_tokens(dims)
    waves = [synth.waveform(30000 + 7000 * i, seed=40 + i) for i in range(10)]
    out["test-c_greedy_depth14_f32"] = [tokens(w_t, dims, sp, w, 1, 14) for w in waves]
    out["test-c_greedy_depth14_f16"] = [tokens(w_t, dims, sp, w, 1, 14, "f16") for w in waves]
    waves = [synth.waveform(42000 + 9000 * i, seed=60 + i) for i in range(3)]
    out["test-c_beam5_depth8_f32"] = [tokens(w_t, dims, sp, w, 5, 8) for w in waves]
    dims, _, w_t = synth.make_weights("test-d", seed=0)
    sp = synth.special_tokens(dims)
    out["test-d_greedy_depth6_f32"] = {str(i): tokens(w_t, dims, sp, synth.waveform(24000 + 3000 * i, seed=80 + i), 1, 6) for i in (0, 7, 13, 19)}
    dims, _, w_t = synth.make_weights("test-e", seed=0)
    sp = synth.special_tokens(dims)
    out["test-e_greedy_depth6_f32"] = {str(i): tokens(w_t, dims, sp, synth.waveform(20000 + 2500 * i, seed=120 + i), 1, 6) for i in (0, 4, 8)}
    (OUT / "tokens_wide.json").write_text(json.dumps(out, indent=1))
    print({k: (len(v)) for k, v in out.items()})


if __name__ == "__main__":
    main()
