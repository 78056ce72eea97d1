#!/bin/bash
# Round-end evidence on one B200: the whole GPU suite, smoke, the default bench line (both arms), ncu launch lists and
# --set full summaries of the decoders of the final build.  Outputs under gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
(timeout 560 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.txt 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_final.txt)
(timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo smoke rc=$?; tail -1 gpurun_out/smoke.txt)
(timeout 240 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo bench rc=$?)
(timeout 200 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo ref rc=$?; cut -c1-200 gpurun_out/bench_reference.json)
(WB200_NO_COOP=1 timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tiny.csv python bench.py --configs tiny.en:1:f32 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/launches_tiny.log 2>&1; echo ncu-list-tiny rc=$?)
(timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_small.csv python bench.py --configs small.en:8:f16 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/launches_small.log 2>&1; echo ncu-list-small rc=$?)
(timeout 200 ncu --set full --clock-control none -k regex:dec5_kernel -s 1 -c 1 -f -o gpurun_out/full_small_dec5_f16 python scripts/run_config.py small.en 8 1 f16 100 > gpurun_out/ncu_full5.log 2>&1; echo full5 rc=$?
 ncu -i gpurun_out/full_small_dec5_f16.ncu-rep --page raw --csv > gpurun_out/r02f_full_small_dec5_f16.csv 2>/dev/null)
(WB200_NO_COOP=1 timeout 200 ncu --set full --clock-control none -k regex:dec4_kernel -s 1 -c 1 -f -o gpurun_out/full_tiny_dec4 python scripts/run_config.py tiny.en 1 1 f32 100 > gpurun_out/ncu_full4.log 2>&1; echo full4 rc=$?
 ncu -i gpurun_out/full_tiny_dec4.ncu-rep --page raw --csv > gpurun_out/r02f_full_tiny_dec4.csv 2>/dev/null)
rm -f gpurun_out/full_*.ncu-rep
python - <<'P'
import json
d = json.loads(open("gpurun_out/bench_final.json").read())
print("clocks", d["clocks"], "cpu", d["cpu_baseline"])
for c in d["configs"]:
    print(c["model"], c["kv_cache"], round(c["value"], 1), round(c["ms_per_step"], 2), round(c["e2e"]["value"], 1), c["phase_ms"],
          round(c["roofline"]["us_per_position"], 1), round(c["roofline"]["frac"], 3), c["tokens_checksum"], c["gpu_launches"])
P
