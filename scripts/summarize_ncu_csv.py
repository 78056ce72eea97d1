"""Summarises ncu outputs pulled back as CSV (the .ncu-rep files stay on the GPU box: size) into profiles/.
  python scripts/summarize_ncu_csv.py launches <tag> <launch_list.csv> "<command that was profiled>"
  python scripts/summarize_ncu_csv.py full <tag> <raw_page.csv> [<raw_page.csv> ...]
`launches`: gpu__time_duration.sum per launch (ncu --metrics ... --csv --log-file)  -> profiles/<tag>_launches.txt
`full`:     ncu -i X.ncu-rep --page raw --csv of --set full captures                 -> profiles/<tag>_ncu_full_summary.txt"""
import collections, csv, io, re, sys


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("void wb::<unnamed>::", "").replace("wb::<unnamed>::", "").replace("void wb::", "").strip()


def launches(tag, path, cmd):
    lines = [l for l in open(path) if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    idx = {h: i for i, h in enumerate(hdr)}
    agg, tot = collections.OrderedDict(), 0.0
    for row in r:
        if len(row) < len(hdr):
            continue
        v = float(row[idx["Metric Value"]].replace(",", ""))
        unit = row[idx["Metric Unit"]]
        v = v / 1000 if unit in ("ns", "nsecond") else v * 1000 if unit in ("ms", "msecond") else v     # -> us
        a = agg.setdefault((short(row[idx["Kernel Name"]])[:72], row[idx["Grid Size"]]), [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    out = [f"# ncu launch list ({tag}): gpu__time_duration.sum, --clock-control none; {cmd}",
           f"# total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches; cold-cache and serialised under the profiler: compare SHARES, not absolutes",
           "share%  count     avg_us  kernel  grid"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{100 * t / tot:6.2f} {n:5d} {t / n:10.2f}  {k[0]}  {k[1]}")
    open(f"profiles/{tag}_launches.txt", "w").write("\n".join(out) + "\n")
    print("\n".join(out[:16]))


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
        "l1tex__t_bytes.sum", "sm__icc_request_hit_rate.pct"]


def full(tag, paths):
    lines = [f"# ncu --set full --clock-control none ({tag}); per-launch averages; the .ncu-rep files are not committed (size)"]
    for path in paths:
        rr = list(csv.reader(io.StringIO("".join(l for l in open(path) if not l.startswith("==")))))
        if len(rr) < 3:
            lines.append(f"=== {path}: empty")
            continue
        h, units = rr[0], rr[1]
        seen = collections.OrderedDict()
        for row in rr[2:]:
            if len(row) < len(h):
                continue
            key = (short(row[h.index("Kernel Name")]), row[h.index("Grid Size")])
            rec = seen.setdefault(key, {"n": 0})
            rec["n"] += 1
            for w in WANT:
                if w in h:
                    try:
                        rec[w] = rec.get(w, 0.0) + float(row[h.index(w)].replace(",", ""))
                    except ValueError:
                        pass
        lines.append(f"=== {path.split('/')[-1]}")
        for (name, grid), rec in seen.items():
            lines.append(f"--- {name} grid={grid} launches={rec['n']}")
            for w in WANT:
                if w in rec:
                    lines.append(f"    {w}: {rec[w] / rec['n']:.6g} {units[h.index(w)]}")
    open(f"profiles/{tag}_ncu_full_summary.txt", "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:80]))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        full(sys.argv[2], sys.argv[3:])
