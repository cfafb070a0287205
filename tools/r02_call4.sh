#!/bin/bash
# round 2, fourth gpurun call: wrench classes without the stored B_k / B_k D^-1 (smaller shared memory), occupancy variants
mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee $O/r02d_tests.txt
echo "== base"; timeout 200 python tools/perf_quick.py 10 | tee $O/r02d_base10.txt; timeout 300 python tools/perf_quick.py 20 | tee $O/r02d_base20.txt
for v in w6 w6t9; do [ -f ab/liba1mpc_$v.so ] && { echo "== $v"; A1MPC_LIB=$PWD/ab/liba1mpc_$v.so timeout 200 python tools/perf_quick.py 10 | tee $O/r02d_$v.txt; }; done
[ -f ab/liba1mpc_w6t9.so ] && { echo "== w6t9 bench B=32768"; A1MPC_LIB=$PWD/ab/liba1mpc_w6t9.so timeout 300 python bench.py --batch 32768 --steps 100 --no-cpu-baseline --no-subrecords > $O/r02d_bench_B32768_w6t9.json 2> $O/r02d_bench.err; tail -c 400 $O/r02d_bench_B32768_w6t9.json; echo; echo "== w6t9 bench B=1024"; A1MPC_LIB=$PWD/ab/liba1mpc_w6t9.so timeout 300 python bench.py --steps 1000 --no-cpu-baseline --no-subrecords > $O/r02d_bench_B1024_w6t9.json 2>> $O/r02d_bench.err; python -c "import json;d=json.loads(open('$O/r02d_bench_B1024_w6t9.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['class_kernel_ms_per_step'])"; }
echo "== default bench B=32768"; timeout 300 python bench.py --batch 32768 --steps 100 --no-cpu-baseline --no-subrecords > $O/r02d_bench_B32768.json 2>> $O/r02d_bench.err; tail -c 400 $O/r02d_bench_B32768.json
echo "== robust sweep (every QP against the oracle)"; timeout 900 python tools/robust_sweep.py 2>&1 | tail -12 | tee $O/r02d_robust.txt
