#!/bin/bash
# round 2, call 17: class kernels draw QPs from a device-wide queue after their first one (default) against the static split (q0)
mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee $O/r02r_tests.txt
for v in new q0; do
  echo "== $v"
  if [ $v = new ]; then unset A1MPC_LIB; else export A1MPC_LIB=$PWD/ab/liba1mpc_$v.so; fi
  timeout 300 python tools/perf_quick.py 10 2>&1 | tee $O/r02r_n10_$v.txt
  timeout 300 python tools/perf_quick.py 20 2>&1 | grep mix | tee $O/r02r_n20_$v.txt
  timeout 300 python bench.py --steps 1000 --no-cpu-baseline 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('bench B=1024: %.3f M  %.4f ms  p50 %.1f  e2e %.3f M classes %s'%(d['value']/1e6,d['ms_per_step'],d['p50_solve_us'],d['e2e']['value']/1e6,d['class_kernel_ms_per_step']));print('config3 %.3f M  config4 %.3f M  plugin %.3f M'%(d['config3']['value']/1e6,d['config4']['value']/1e6,d['e2e_plugin_pageable']['value']/1e6))" | tee $O/r02r_bench_$v.txt
  timeout 300 python bench.py --batch 32768 --steps 100 --no-cpu-baseline --no-subrecords 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('bench B=32768: %.3f M  %.4f ms  e2e %.3f M classes %s'%(d['value']/1e6,d['ms_per_step'],d['e2e']['value']/1e6,d['class_kernel_ms_per_step']))" | tee -a $O/r02r_bench_$v.txt
done
unset A1MPC_LIB
for t in racecheck synccheck memcheck; do echo "== $t"; timeout 600 compute-sanitizer --tool $t python tools/prof_target2.py 10 600 2>&1 | tail -1; done
timeout 900 python tools/robust_sweep.py 2>&1 | tail -12 | tee $O/r02r_robust.txt
