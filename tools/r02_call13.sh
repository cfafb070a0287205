#!/bin/bash
# round 2, call 13: warp teams for the direct (one / two stance feet) classes at N = 20 too; team Cholesky out of line with a compile-time warp index
mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $O/r02m_tests.txt
for v in prev new t2w2 t3 t4; do
  echo "== $v"
  if [ $v = new ]; then unset A1MPC_LIB; else export A1MPC_LIB=$PWD/ab/liba1mpc_$v.so; fi
  timeout 300 python tools/perf_quick.py 20 2>&1 | tee $O/r02m_n20_$v.txt
done
unset A1MPC_LIB
echo "== sanitizer on the N=20 team kernels (B=96 mix)"
for t in memcheck racecheck synccheck; do timeout 600 compute-sanitizer --tool $t python tools/prof_target2.py 20 96 > $O/r02m_sanitizer_n20_$t.txt 2>&1; tail -2 $O/r02m_sanitizer_n20_$t.txt; done
echo "== bench"; timeout 600 python bench.py --no-cpu-baseline > $O/r02m_bench_B1024.json 2> $O/r02m_bench.err; python -c "
import json;d=json.loads(open('$O/r02m_bench_B1024.json').read().strip().splitlines()[-1]);print('value %.3f M  %.4f ms'%(d['value']/1e6,d['ms_per_step']));print('config3',d['config3']['value'],d['config3']['ms_per_step']);print('config4',d['config4']['value'])"
timeout 600 python tools/robust_sweep.py 2>&1 | tail -5 | tee $O/r02m_robust.txt
