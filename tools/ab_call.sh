#!/bin/bash
# dev tool (one gpurun call): GPU parity tests of the current build, then A/B timing against ab/liba1mpc_old.so, one ncu capture
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/ab_tests.txt
cat gpurun_out/ab_tests.txt | tail -3
echo "== new"; timeout 300 python tools/perf_quick.py 10 2>&1 | tee gpurun_out/ab_new.txt
echo "== old"; A1MPC_LIB=$PWD/ab/liba1mpc_old.so timeout 300 python tools/perf_quick.py 10 2>&1 | tee gpurun_out/ab_old.txt
timeout 300 python bench.py --steps 500 --warmup 10 --no-cpu-baseline 2>gpurun_out/ab_bench.err | tail -1 > gpurun_out/ab_bench_B1024.json
timeout 300 python bench.py --batch 32768 --steps 30 --warmup 5 --no-cpu-baseline 2>>gpurun_out/ab_bench.err | tail -1 > gpurun_out/ab_bench_B32768.json
python - <<'PY'
import json
for f in ("gpurun_out/ab_bench_B1024.json", "gpurun_out/ab_bench_B32768.json"):
    try:
        d = json.load(open(f)); print(f, "value %.0f" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], d.get("roofline_fp64", {}).get("frac"))
    except Exception as e:
        print(f, "ERR", e)
PY
timeout 400 ncu --set full --import-source on --clock-control none -k regex:solve_kernel -s 10 -c 1 -o gpurun_out/r01b_trot_mix16k python tools/prof_target.py 16384 0 > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out/*.ncu-rep
