#!/bin/bash
# round 2, call 16: team Cholesky with the barriers in common code (mode 2, default) against mode 0 (specialised loops) and mode 1 (barrier behind a call)
mkdir -p gpurun_out; O=gpurun_out
for n in 10 20; do for t in synccheck racecheck memcheck; do echo "== $t N=$n (default, B=96 mix)"; timeout 600 compute-sanitizer --tool $t --print-limit 50 python tools/prof_target2.py $n 96 > $O/r02q_${t}_n$n.txt 2>&1; tail -1 $O/r02q_${t}_n$n.txt; done; done
for v in new m0 m1; do
  echo "== $v"
  if [ $v = new ]; then unset A1MPC_LIB; else export A1MPC_LIB=$PWD/ab/liba1mpc_$v.so; fi
  timeout 300 python tools/perf_quick.py 10 2>&1 | grep four | tee $O/r02q_n10_$v.txt
  timeout 300 python tools/perf_quick.py 20 2>&1 | tee $O/r02q_n20_$v.txt
  timeout 300 python bench.py --steps 1000 --no-cpu-baseline --no-subrecords 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('bench B=1024: %.3f M  %.4f ms  p50 %.1f  classes %s'%(d['value']/1e6,d['ms_per_step'],d['p50_solve_us'],d['class_kernel_ms_per_step']))" | tee $O/r02q_bench_$v.txt
  timeout 300 python bench.py --batch 32768 --steps 100 --no-cpu-baseline --no-subrecords 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('bench B=32768: %.3f M  %.4f ms  classes %s'%(d['value']/1e6,d['ms_per_step'],d['class_kernel_ms_per_step']))" | tee -a $O/r02q_bench_$v.txt
done
unset A1MPC_LIB
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee $O/r02q_tests.txt
