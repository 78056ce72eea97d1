"""Per-stage means of a decoder6.cu stage trace (scripts/trace_decode.py with WB200_TRACE): python scripts/analyze_dec6_trace.py file.npy [L] [depth]"""
import sys
import numpy as np
names = ['comb1', 'qkv', 'self', 'oproj', 'send1', 'comb2', 'cq', 'cross', 'coproj', 'send2', 'comb3', 'mlp1', 'mlp2', 'send3']
t = np.load(sys.argv[1])
L = int(sys.argv[2]) if len(sys.argv) > 2 else 4
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 20
pre, per = 1 + 14 * L, 1 + 14 * L + 6
idx = 3 * pre
lg = np.array([t[idx + per * p: idx + per * (p + 1)] for p in range(depth)])
m = (np.diff(lg, axis=1) / 1e3)[2:].mean(axis=0)
for l in range(L):
    print('  L%d ' % l + ' '.join(f'{n}={v:.2f}' for n, v in zip(names, m[l * 14:(l + 1) * 14])), ' sum=%.2f' % m[l * 14:(l + 1) * 14].sum())
print('  final-combine, grid barrier, planes, stream, merge, flag:', np.round(m[14 * L:], 2), ' to next position', round(float(((lg[1:, 0] - lg[:-1, -1]) / 1e3).mean()), 2))
print('  per position us', round(float(m.sum() + ((lg[1:, 0] - lg[:-1, -1]) / 1e3).mean()), 2))
