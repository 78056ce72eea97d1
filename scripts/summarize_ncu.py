"""Summarises gpurun_out ncu outputs into profiles/ (launch-share table + per-kernel full-set metrics)."""
import csv, collections, io, re, subprocess, sys

tag = sys.argv[1]          # e.g. r01_v4
launch_csv = sys.argv[2]
rep = sys.argv[3]

lines = [l for l in open(launch_csv) if not l.startswith('==')]
r = csv.reader(lines); hdr = next(r); idx = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict(); tot = 0.0
for row in r:
    if len(row) < len(hdr): continue
    name = re.sub(r'\(.*', '', row[idx['Kernel Name']]).replace('void wb::<unnamed>::', '').replace('wb::<unnamed>::', '')
    v = float(row[idx['Metric Value']].replace(',', '')) / 1000
    a = agg.setdefault((name[:70], row[idx['Grid Size']]), [0, 0.0]); a[0] += 1; a[1] += v; tot += v
out = [f"# ncu launch list ({tag}): gpu__time_duration.sum, --clock-control none; python bench.py --steps 1 --warmup 0",
       f"# total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches; cold-cache and serialised: compare SHARES, not absolutes",
       "share%  count     avg_us  kernel  grid"]
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{100 * t / tot:6.2f} {n:5d} {t / n:10.2f}  {k[0]}  {k[1]}")
open(f'profiles/{tag}_launches.txt', 'w').write("\n".join(out) + "\n")
print("\n".join(out[:14]))

raw = subprocess.run(f"ncu -i {rep} --page raw --csv", shell=True, capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw))); h = rr[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__cluster_size',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.max']
seen = collections.OrderedDict()
for row in rr[2:]:
    name = re.sub(r'\(.*', '', row[h.index('Kernel Name')]).replace('void unnamed>::', '')
    key = (name, row[h.index('Grid Size')])
    rec = seen.setdefault(key, {'n': 0})
    rec['n'] += 1
    for w in want:
        if w in h:
            try: rec[w] = rec.get(w, 0.0) + float(row[h.index(w)].replace(',', ''))
            except ValueError: pass
unit = {w: rr[1][h.index(w)] for w in want if w in h}
lines = [f"# ncu --set full --clock-control none --import-source on ({tag}); averages per launch; raw report not committed (size)"]
for (name, grid), rec in seen.items():
    lines.append(f"--- {name} grid={grid} launches={rec['n']}")
    for w in want:
        if w in rec: lines.append(f"    {w}: {rec[w] / rec['n']:.4g} {unit.get(w, '')}")
open(f'profiles/{tag}_ncu_full_summary.txt', 'w').write("\n".join(lines) + "\n")
print("\n".join(lines[:70]))
