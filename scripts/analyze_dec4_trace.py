"""Stage table of the cluster decoder (decoder4.cu) from a WB200_TRACE stamp file written by scripts/trace_decode.py.
usage: python scripts/analyze_dec4_trace.py <trace.npy> <n_layers>
Stamps of CTA 0 per logits position: 8 per layer (after each cluster barrier) + 8 around the vocabulary stage."""
import sys
import numpy as np

t = np.load(sys.argv[1]).astype(np.int64)
L = int(sys.argv[2])
per = 8 * L + 8
names = ["qkv", "self", "out", "cq", "cross", "cout", "mlp1", "mlp2"]
tail = ["publish", "G1", "lg_ln", "lg_loop", "lg_merge", "G2", "finish", "G3"]
d = np.diff(t)
# logits positions are the last ones: align on the end of the trace
n = (len(d) // per) - 4          # skip the prefill positions (fewer stamps) at the start
d = d[len(d) - n * per:].reshape(n, per)
print(f"positions {n}; mean position {d.sum(1).mean() / 1e3:.2f} us")
lay = d[:, :8 * L].reshape(n, L, 8).mean((0, 1)) / 1e3
for k, v in zip(names, lay):
    print(f"{k:8s} {v:6.2f} us (mean over layers, stage + cluster barrier)")
print(f"layer    {lay.sum():6.2f} us")
for k, v in zip(tail, d[:, 8 * L:].mean(0) / 1e3):
    print(f"{k:8s} {v:6.2f} us")
