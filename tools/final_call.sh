#!/bin/bash
# dev tool (one gpurun call): GPU parity tests, smoke, default bench lines
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/fin_tests.txt
tail -3 gpurun_out/fin_tests.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 200 python tools/perf_quick.py 10 2>&1 | tee gpurun_out/fin_perf10.txt
timeout 200 python tools/perf_quick.py 20 2>&1 | tee gpurun_out/fin_perf20.txt
timeout 400 python bench.py 2>gpurun_out/fin_bench.err | tail -1 > gpurun_out/fin_bench_B1024.json
timeout 300 python bench.py --batch 32768 --steps 40 --warmup 5 --no-cpu-baseline 2>>gpurun_out/fin_bench.err | tail -1 > gpurun_out/fin_bench_B32768.json
python - <<'PY'
import json
for f in ("gpurun_out/fin_bench_B1024.json", "gpurun_out/fin_bench_B32768.json"):
    try:
        d = json.load(open(f)); print(f, "value %.0f" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "p50", d.get("p50_solve_us"), d.get("roofline_fp64", {}).get("frac"), d.get("clocks"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/fin_bench.err
