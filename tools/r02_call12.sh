#!/bin/bash
# round 2, call 12: warp teams (two warps per QP) for the wrench classes
mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee $O/r02k_tests.txt
echo "== team2"; timeout 200 python tools/perf_quick.py 10 | tee $O/r02k_base10.txt; timeout 300 python tools/perf_quick.py 20 | tee $O/r02k_base20.txt
timeout 300 python bench.py --steps 1000 --no-cpu-baseline --no-subrecords 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('bench B=1024: %.3f M  %.4f ms  p50 %.1f  classes %s'%(d['value']/1e6,d['ms_per_step'],d['p50_solve_us'],d['class_kernel_ms_per_step']))" | tee $O/r02k_bench.txt
timeout 300 python bench.py --batch 32768 --steps 100 --no-cpu-baseline --no-subrecords 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('bench B=32768: %.3f M  %.4f ms  classes %s'%(d['value']/1e6,d['ms_per_step'],d['class_kernel_ms_per_step']))" | tee -a $O/r02k_bench.txt
