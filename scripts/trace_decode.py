import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
os.environ["WB200_TRACE"] = "1"
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
print("stamps", len(t), "total ms", (t[-1] - t[0]) / 1e6)
# stamps: start, then per barrier (stage_end, barrier_end)
L = dims.n_text_layer
per_step = 2 * (8 * L + 2)
names = [f"L{l}.{n}" for l in range(L) for n in ("qkv", "self", "out", "cq", "cross", "cout", "mlp1", "mlp2")] + ["logits", "finish"]
body = t[1:]
prefill = 3 * 2 * (8 * L)
steps = body[prefill:]
n = len(steps) // per_step
steps = steps[:n * per_step].reshape(n, per_step // 2, 2)
prev_end = np.concatenate([[body[prefill - 1]], steps[:-1, -1, 1]])
stage = steps[:, :, 0] - np.concatenate([prev_end[:, None], steps[:, :-1, 1]], axis=1)
barr = steps[:, :, 1] - steps[:, :, 0]
print(f"steps {n}; mean step us {(steps[:, -1, 1] - prev_end).mean() / 1e3:.1f}")
for i, nm in enumerate(names):
    print(f"{nm:10s} stage {stage[:, i].mean() / 1e3:6.2f} us   barrier {barr[:, i].mean() / 1e3:6.2f} us")
print("sum stage", stage.mean(0).sum() / 1e3, "sum barrier", barr.mean(0).sum() / 1e3)
