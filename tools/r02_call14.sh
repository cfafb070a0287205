#!/bin/bash
# round 2, call 14: synccheck of the N=20 team kernels after "teams leave together"; wrench team of three at N=20; wrench team of two at N=10 with the low-spill Cholesky
mkdir -p gpurun_out; O=gpurun_out
echo "== synccheck N=20 (B=96 mix)"; timeout 600 compute-sanitizer --tool synccheck python tools/prof_target2.py 20 96 > $O/r02n_sanitizer_n20_synccheck.txt 2>&1; tail -2 $O/r02n_sanitizer_n20_synccheck.txt
echo "== racecheck N=20"; timeout 600 compute-sanitizer --tool racecheck python tools/prof_target2.py 20 96 > $O/r02n_sanitizer_n20_racecheck.txt 2>&1; tail -2 $O/r02n_sanitizer_n20_racecheck.txt
for v in new w3; do
  echo "== N=20 $v"
  if [ $v = new ]; then unset A1MPC_LIB; else export A1MPC_LIB=$PWD/ab/liba1mpc_$v.so; fi
  timeout 300 python tools/perf_quick.py 20 2>&1 | tee $O/r02n_n20_$v.txt
done
for v in new t2n10; do
  echo "== N=10 $v"
  if [ $v = new ]; then unset A1MPC_LIB; else export A1MPC_LIB=$PWD/ab/liba1mpc_$v.so; fi
  timeout 300 python tools/perf_quick.py 10 2>&1 | tee $O/r02n_n10_$v.txt
  timeout 300 python bench.py --steps 1000 --no-cpu-baseline --no-subrecords 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('bench B=1024: %.3f M  %.4f ms  p50 %.1f  classes %s'%(d['value']/1e6,d['ms_per_step'],d['p50_solve_us'],d['class_kernel_ms_per_step']))" | tee $O/r02n_bench_$v.txt
  timeout 300 python bench.py --batch 32768 --steps 100 --no-cpu-baseline --no-subrecords 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('bench B=32768: %.3f M  %.4f ms  classes %s'%(d['value']/1e6,d['ms_per_step'],d['class_kernel_ms_per_step']))" | tee -a $O/r02n_bench_$v.txt
done
unset A1MPC_LIB
echo "== synccheck N=10 teams"; A1MPC_LIB=$PWD/ab/liba1mpc_t2n10.so timeout 600 compute-sanitizer --tool synccheck python tools/prof_target2.py 10 96 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee $O/r02n_tests.txt
