import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
os.environ["WB200_TRACE"] = "gpurun_out/d3_trace.txt"; os.makedirs("gpurun_out", exist_ok=True)
import numpy as np
import wb200  # noqa
from whisper_burn_b200 import model, synth, transcribe
dims, w_np = synth.make_weights(sys.argv[1] if len(sys.argv) > 1 else "tiny.en")
sp = synth.special_tokens(dims)
wh = model.Whisper(dims, w_np)
chunk = synth.chunk_waveform(0)
waves = [chunk[:238559], chunk[190559:429118], chunk[381118:]]
sess = transcribe.Session(wh, 3, 1, 105)
sess.transcribe_windows(waves, sp, sp.is_special_bitmap(), beam_size=1, max_depth=100)
t = np.loadtxt("gpurun_out/d3_trace.txt", dtype=np.uint64).astype(np.int64)
L = dims.n_text_layer
print("stamps", len(t), "total ms", (t[-1] - t[0]) / 1e6)
# layout: start; prefill steps: 8L stamps each; decode steps: 8L + 6 stamps (pre/post of 3 grid barriers)
pre = 1 + 3 * 8 * L
per = 8 * L + 8
body = t[pre:]
n = len(body) // per
body = body[:n * per].reshape(n, per)
prev = np.concatenate([[t[pre - 1]], body[:-1, -1]])
d = np.diff(np.concatenate([prev[:, None], body], axis=1), axis=1) / 1e3
names = [f"L{l}.{s}" for l in range(L) for s in ("qkv", "self", "out", "cq", "cross", "cout", "mlp1", "mlp2")] + ["publish", "G1", "lg_ln", "lg_loop", "lg_merge", "G2", "finish", "G3"]
print("steps", n, "mean step us", d.sum(1).mean())
m = d.mean(0)
kinds = ("qkv", "self", "out", "cq", "cross", "cout", "mlp1", "mlp2")
for k, nm in enumerate(kinds):
    print(f"{nm:10s} {np.mean([m[l * 8 + k] for l in range(L)]):7.2f} us (mean over layers, stage + cluster barrier)")
for nm, v in zip(names[8 * L:], m[8 * L:]):
    print(f"{nm:10s} {v:7.2f} us")
