"""Sub-stage table of the cluster decoder from a trace of a -DD4_FINE=1 build (see D4_STAGE_END / WB_FINE in decoder4.cu).
usage: python scripts/analyze_dec4_fine.py <trace.npy> <n_layers>"""
import sys
import numpy as np

t = np.load(sys.argv[1]).astype(np.int64)
L = int(sys.argv[2])
end = ["sync", "send", "prefetch", "WAIT"]


def stage(name, pre):
    return [f"{name} {x}" for x in pre + end]


labels = (stage("S1", ["ln+dot+stg"]) + stage("S2", ["self attention"]) + stage("S3", ["dot", "stg"]) + stage("S4", ["ln", "dot", "stg"]) +
          stage("S5", ["cross attention"]) + stage("S6", ["merge+dot+stg"]) + stage("S7", ["ln", "dot+gelu", "stg"]) + stage("S8", ["dot+stg"]))
PL = len(labels)
per = PL * L + 8
d = np.diff(t)
n = (len(d) // per) - 4
d = d[len(d) - n * per:].reshape(n, per)
print(f"positions {n}; mean position {d.sum(1).mean() / 1e3:.2f} us (stamps add overhead)")
lay = d[:, :PL * L].reshape(n, L, PL).mean((0, 1)) / 1e3
for k, v in zip(labels, lay):
    print(f"{k:24s} {v:6.2f} us")
print(f"layer {lay.sum():.2f} us; tail {d[:, PL * L:].mean(0).sum() / 1e3:.2f} us")
