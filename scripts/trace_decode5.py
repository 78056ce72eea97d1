"""Stage trace of the batched tensor-core decoder (decoder5.cu): WB200_TRACE stamps of CTA 0.
usage: python scripts/trace_decode5.py [model] [n_chunks] [kv]"""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
os.environ["WB200_TRACE"] = "gpurun_out/d3_trace.txt"; os.makedirs("gpurun_out", exist_ok=True)
import numpy as np
import wb200  # noqa
from whisper_burn_b200 import ffi, model, synth, transcribe
name = sys.argv[1] if len(sys.argv) > 1 else "small.en"
n_chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 8
kv = sys.argv[3] if len(sys.argv) > 3 else "f32"
dims, w_np = synth.make_weights(name)
sp = synth.special_tokens(dims)
wh = model.Whisper(dims, w_np)
waves = []
for c in range(n_chunks):
    chunk = synth.chunk_waveform(c)
    waves += [chunk[:238559], chunk[190559:429118], chunk[381118:]]
depth = 40
sess = transcribe.Session(wh, len(waves), 1, 4 + depth + 1, kv_dtype=ffi.WB_KV_F16 if kv == "f16" else ffi.WB_KV_F32)
sess.transcribe_windows(waves, sp, sp.is_special_bitmap(), beam_size=1, max_depth=depth)
print("decoder", sess.last_decoder(), "rows", len(waves))
t = np.loadtxt("gpurun_out/d3_trace.txt", dtype=np.uint64).astype(np.int64)
print("stamps", len(t), "total ms", (t[-1] - t[0]) / 1e6)
L = dims.n_text_layer
kinds = ["ln1", "qkv", "self", "out", "ln2", "cq", "cross", "cout", "ln3", "mlp1", "mlp2"]
tail = ["lnf", "logits", "softmax", "finish"]
NK = len(kinds)
per_step = 2 * (NK * L + len(tail))
body = t[1:]
prefill = 3 * 2 * (NK * L)
steps = body[prefill:]
n = len(steps) // per_step
steps = steps[:n * per_step].reshape(n, per_step // 2, 2)
prev_end = np.concatenate([[body[prefill - 1]], steps[:-1, -1, 1]])
stage = steps[:, :, 0] - np.concatenate([prev_end[:, None], steps[:, :-1, 1]], axis=1)
barr = steps[:, :, 1] - steps[:, :, 0]
print(f"steps {n}; mean step us {(steps[:, -1, 1] - prev_end).mean() / 1e3:.1f}")
for k, nm in enumerate(kinds):
    idx = [l * NK + k for l in range(L)]
    print(f"{nm:8s} stage {stage[:, idx].mean() / 1e3:7.2f} us   barrier(wait) {barr[:, idx].mean() / 1e3:7.2f} us   (mean over layers)")
for i, nm in enumerate(tail):
    print(f"{nm:8s} stage {stage[:, NK * L + i].mean() / 1e3:7.2f} us   barrier(wait) {barr[:, NK * L + i].mean() / 1e3:7.2f} us")
print("sum stage", stage.mean(0).sum() / 1e3, "sum barrier", barr.mean(0).sum() / 1e3)
