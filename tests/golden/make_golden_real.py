"""Golden token fixtures at the REAL shapes of BASELINE.json configs 3-5 (small.en / medium / large-v2), from the CPU oracle.

Same status as make_golden.py: oracle outputs (the reference ships no vectors and cannot be built here: "parity unpinned");
they pin the oracle and give the GPU tests committed ids at the shapes the metric is quoted on:

  tiny.en   seed 0, chunks 0..7 (24 windows), greedy, depth 100, fp32 and fp16 K/V cache (batched cluster decoder)
  small.en  seed 0, chunks 0..7 (24 reference windows), greedy, depth 100, fp32 cache and fp16 K/V cache
            + per-step top-5 ids / log-probs of every window (checked through wb_session_step with a tolerance: the synthetic
            deep models decode to 1-2 distinct tokens, so the continuous log-probs carry the parity evidence)
  medium    seed 0, chunk 0 (3 windows), greedy, depth 30
  large-v2  seed 0, chunk 0 window 2 (the short one), beam 5, depth 20, fp16 K/V cache (configs[4]) and fp32

Every case records the smallest top-1/top-2 log-prob margin met on the decoded path.
Run from the repo root (minutes of CPU):  python tests/golden/make_golden_real.py [small.en] [medium] [large-v2]
"""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from oracle import audio, model, synth, transcribe  # noqa: E402

OUT = Path(__file__).resolve().parent / "tokens_real.json"


def margins(trace):
    g = []
    for st in trace["log_probs"]:
        for row in st:
            if row is not None:
                s = np.partition(row, row.shape[0] - 2)[-2:]
                g.append(float(abs(s[1] - s[0])))
    return g


def windows_of(chunk_id, dims):
    chunk = synth.chunk_waveform(chunk_id)
    window_len = audio.max_waveform_samples(dims.n_audio_ctx - transcribe.PADDING)
    return chunk, transcribe.window_bounds(len(chunk), 16000, window_len)


def decode(w, dims, sp, wave, beam, depth, kv, want_top1=False, want_margins=False):
    mel = audio.prep_audio(torch.from_numpy(np.ascontiguousarray(wave))[None])
    tr = {}
    toks = transcribe.mels_to_tokens(w, dims, sp, mel, beam_size=beam, max_depth=depth, opts=model.OracleOptions(kv_dtype=kv), trace=tr)
    rec = {"tokens": toks, "min_margin": min(margins(tr))}
    if want_margins:   # top-1 / top-2 log-prob gap at every generated step (greedy): where it is tiny, fp32 rounding decides the id
        rec["margins"] = [round(g, 7) for g in margins(tr)]
    if want_top1 and beam == 1:   # per generated step: the 5 best ids (ties -> lower id) and their log-probs
        top = []
        for st in tr["log_probs"]:
            lp = st[0]
            order = np.lexsort((np.arange(lp.shape[0]), -lp))[:5]
            top.append([[int(i) for i in order], [float(lp[i]) for i in order]])
        rec["top5"] = top
    return rec


def small_en():
    dims, _, w = synth.make_weights("small.en", seed=0)
    sp = synth.special_tokens(dims)
    out = {"model": "small.en", "seed": 0, "depth": 100, "chunks": []}
    for c in range(8):
        chunk, bounds = windows_of(c, dims)
        rec = {"bounds": bounds, "f32": [], "f16": []}
        for i, (s, e) in enumerate(bounds):
            rec["f32"].append(decode(w, dims, sp, chunk[s:e], 1, 100, "f32", want_top1=True))
            rec["f16"].append(decode(w, dims, sp, chunk[s:e], 1, 100, "f16"))
        out["chunks"].append(rec)
        print("small.en chunk", c, "done", flush=True)
    out["min_margin_f32"] = min(r["min_margin"] for c in out["chunks"] for r in c["f32"])
    out["min_margin_f16"] = min(r["min_margin"] for c in out["chunks"] for r in c["f16"])
    return out


def tiny_en():
    """tiny.en, chunks 0..7 (24 windows): the batched case of the head-fused cluster decoder (decoder6.cu: 24 clusters, 3 n-tiles)."""
    dims, _, w = synth.make_weights("tiny.en", seed=0)
    sp = synth.special_tokens(dims)
    out = {"model": "tiny.en", "seed": 0, "depth": 100, "chunks": []}
    for c in range(8):
        chunk, bounds = windows_of(c, dims)
        rec = {"bounds": bounds, "f32": [], "f16": []}
        for (s, e) in bounds:
            rec["f32"].append(decode(w, dims, sp, chunk[s:e], 1, 100, "f32", want_margins=True))
            rec["f16"].append(decode(w, dims, sp, chunk[s:e], 1, 100, "f16", want_margins=True))
        out["chunks"].append(rec)
    out["min_margin_f32"] = min(r["min_margin"] for c in out["chunks"] for r in c["f32"])
    out["min_margin_f16"] = min(r["min_margin"] for c in out["chunks"] for r in c["f16"])
    return out


def medium():
    dims, _, w = synth.make_weights("medium", seed=0)
    sp = synth.special_tokens(dims)
    chunk, bounds = windows_of(0, dims)
    out = {"model": "medium", "seed": 0, "depth": 30, "bounds": bounds, "f32": []}
    for (s, e) in bounds:
        out["f32"].append(decode(w, dims, sp, chunk[s:e], 1, 30, "f32", want_top1=True))
    out["min_margin_f32"] = min(r["min_margin"] for r in out["f32"])
    return out


def large_v2():
    dims, _, w = synth.make_weights("large-v2", seed=0)
    sp = synth.special_tokens(dims)
    chunk, bounds = windows_of(0, dims)
    s, e = bounds[2]
    out = {"model": "large-v2", "seed": 0, "depth": 20, "beam": 5, "window": [s, e]}
    out["f16"] = decode(w, dims, sp, chunk[s:e], 5, 20, "f16")
    out["f32"] = decode(w, dims, sp, chunk[s:e], 5, 20, "f32")
    return out


def main():
    torch.manual_seed(0)
    which = sys.argv[1:] or ["tiny.en", "small.en", "medium", "large-v2"]
    data = json.loads(OUT.read_text()) if OUT.exists() else {}
    for name, fn in (("tiny.en", tiny_en), ("small.en", small_en), ("medium", medium), ("large-v2", large_v2)):
        if name in which:
            t0 = time.time()
            data[name] = fn()
            OUT.write_text(json.dumps(data))
            print(name, "written in", round(time.time() - t0), "s", flush=True)


if __name__ == "__main__":
    main()
