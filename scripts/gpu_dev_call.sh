#!/bin/bash
# One development call on the GPU box: targeted parity tests, stage traces of the persistent decoders, a bench line,
# and source-level ncu captures of the two persistent decoders (read on the CPU box with ncu -i ... --page source --csv).
mkdir -p gpurun_out
(timeout 200 python -m pytest tests/test_parity_gpu.py tests/test_real_shapes_gpu.py -m gpu -x -q -k "tiny_en or fp16_kv or tokens_small_model or eot" --durations=5 > gpurun_out/pytest_c.txt 2>&1; echo pytest rc=$?; tail -10 gpurun_out/pytest_c.txt)
(timeout 60 python scripts/trace_decode4.py > gpurun_out/dec4_trace.txt 2>&1; echo t4 rc=$?; tail -18 gpurun_out/dec4_trace.txt)
(timeout 120 python bench.py --no-cpu-baseline --configs tiny.en:1:f32,tiny.en:1:f16 > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; echo bench rc=$?)
python - <<'P'
import json
d = json.loads(open("gpurun_out/bench_c.json").read())
for c in d["configs"]:
    print(c["model"], c["kv_cache"], round(c["value"], 1), round(c["ms_per_step"], 2), round(c["e2e"]["value"], 1), c["phase_ms"],
          round(c["roofline"]["us_per_position"], 1), round(c["roofline"]["frac"], 3), c["tokens_checksum"])
P
(timeout 240 ncu --set full --clock-control none --import-source on -k regex:dec5_kernel -s 1 -c 1 -f -o gpurun_out/prof_dec5_f16 python scripts/run_config.py small.en 8 1 f16 20 > gpurun_out/ncu_dec5.log 2>&1; echo ncu5 rc=$?; tail -3 gpurun_out/ncu_dec5.log)
(WB200_NO_COOP=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:dec4_kernel -s 1 -c 1 -f -o gpurun_out/prof_dec4 python scripts/run_config.py tiny.en 1 1 f32 100 > gpurun_out/ncu_dec4.log 2>&1; echo ncu4 rc=$?; tail -3 gpurun_out/ncu_dec4.log)
ls -la gpurun_out/*.ncu-rep
