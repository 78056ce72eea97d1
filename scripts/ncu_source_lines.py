"""Aggregates the per-instruction warp-stall samples of an `ncu --set full --import-source on` capture by CUDA source line.
  ncu -i X.ncu-rep --page source --csv > x_sass.csv                       (SASS view: one row per instruction, in address order)
  nvdisasm -g -c <cubin of the same build> > x.sass                       (line table of the same instructions)
  python scripts/ncu_source_lines.py x_sass.csv x.sass <kernel substring, e.g. dec4_kernelILi384ELi4EfE> [top]
The two listings are matched by instruction order inside the kernel's .text section."""
import csv, re, sys, collections

sass_csv, disasm, ksub = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]


def f(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


# line table from nvdisasm: instructions of the kernel in order, each with the last "//## File ..., line N" seen
lines, on, cur = [], False, ("?", 0)
for l in open(disasm):
    if l.startswith(".text."):
        on = ksub in l
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append((cur, l.split("*/", 1)[1].strip()[:60]))
print(f"{len(data)} profiled instructions, {len(lines)} disassembled")
n = min(len(data), len(lines))
agg = collections.defaultdict(lambda: [0.0, collections.Counter()])
tot = 0.0
for i in range(n):
    s = f(data[i][idx["# Samples"]])
    tot += s
    a = agg[lines[i][0]]
    a[0] += s
    for h in stalls:
        a[1][h[6:]] += f(data[i][idx[h]])
print("total samples", tot)
src_cache = {}
for (fn, ln), (s, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if fn not in src_cache:
        try:
            src_cache[fn] = open(f"whisper-burn_b200/csrc/{fn}").read().split("\n")
        except OSError:
            src_cache[fn] = []
    text = src_cache[fn][ln - 1].strip()[:100] if 0 < ln <= len(src_cache[fn]) else ""
    print(f"{100 * s / tot:5.1f}%  {fn}:{ln:<5d} {dict(st.most_common(3))}  | {text}")
