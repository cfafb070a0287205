#!/bin/bash
# round 2, call 15: wrench teams of two at N=10 by default; synccheck experiments (which build trips the report); team of three
mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee $O/r02p_tests.txt
for v in new prev olbar; do
  if [ $v = new ]; then unset A1MPC_LIB; else export A1MPC_LIB=$PWD/ab/liba1mpc_$v.so; fi
  echo "== synccheck N=20 $v"; timeout 600 compute-sanitizer --tool synccheck --print-limit 3000 python tools/prof_target2.py 20 96 > $O/r02p_synccheck_n20_$v.txt 2>&1; tail -1 $O/r02p_synccheck_n20_$v.txt
done
unset A1MPC_LIB
echo "== synccheck N=10 new"; timeout 600 compute-sanitizer --tool synccheck --print-limit 3000 python tools/prof_target2.py 10 96 > $O/r02p_synccheck_n10_new.txt 2>&1; tail -1 $O/r02p_synccheck_n10_new.txt
echo "== racecheck N=10 new"; timeout 600 compute-sanitizer --tool racecheck python tools/prof_target2.py 10 96 2>&1 | tail -1
echo "== memcheck N=10 new"; timeout 600 compute-sanitizer --tool memcheck python tools/prof_target2.py 10 96 2>&1 | tail -1
for v in new w3 olbar; do
  echo "== N=10 $v"
  if [ $v = new ]; then unset A1MPC_LIB; else export A1MPC_LIB=$PWD/ab/liba1mpc_$v.so; fi
  timeout 300 python tools/perf_quick.py 10 2>&1 | tee $O/r02p_n10_$v.txt
  timeout 300 python bench.py --steps 1000 --no-cpu-baseline --no-subrecords 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('bench B=1024: %.3f M  %.4f ms  p50 %.1f  classes %s'%(d['value']/1e6,d['ms_per_step'],d['p50_solve_us'],d['class_kernel_ms_per_step']))" | tee $O/r02p_bench_$v.txt
  timeout 300 python bench.py --batch 32768 --steps 100 --no-cpu-baseline --no-subrecords 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('bench B=32768: %.3f M  %.4f ms  classes %s'%(d['value']/1e6,d['ms_per_step'],d['class_kernel_ms_per_step']))" | tee -a $O/r02p_bench_$v.txt
done
unset A1MPC_LIB
timeout 300 python tools/perf_quick.py 20 2>&1 | tee $O/r02p_n20_new.txt
timeout 900 python tools/robust_sweep.py 2>&1 | tail -12 | tee $O/r02p_robust.txt
