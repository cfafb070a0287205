#!/bin/bash
# round 2, first gpurun call: HEAD's GPU suite, timings of the default and the A/B variants, bench lines, ncu of HEAD kernels, sanitizer
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tee $O/r02_gpu.txt; nproc | tee -a $O/r02_gpu.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee $O/r02_tests.txt
echo "== base"; timeout 200 python tools/perf_quick.py 10 | tee $O/r02_base10.txt; timeout 300 python tools/perf_quick.py 20 | tee $O/r02_base20.txt
for v in r1algo ffrag rsqrtlib fixedref; do [ -f ab/liba1mpc_$v.so ] && { echo "== $v"; A1MPC_LIB=$PWD/ab/liba1mpc_$v.so timeout 200 python tools/perf_quick.py 10 | tee $O/r02_$v.txt; }; done
[ -f ab/liba1mpc_sswitch.so ] && { echo "== sswitch (N=20)"; A1MPC_LIB=$PWD/ab/liba1mpc_sswitch.so timeout 300 python tools/perf_quick.py 20 | tee $O/r02_sswitch20.txt; }
echo "== config 4, general extended kernel"; timeout 200 python tools/ext_probe.py | tee $O/r02_ext.txt
echo "== config 4, compacted class (A1MPC_EXT_COMPACT=1)"; A1MPC_EXT_COMPACT=1 timeout 200 python tools/ext_probe.py | tee $O/r02_ext_compact.txt
echo "== bench"; timeout 300 python bench.py > $O/r02_bench_B1024.json 2> $O/r02_bench_B1024.err; tail -c 1500 $O/r02_bench_B1024.json
timeout 300 python bench.py --batch 32768 --steps 100 --no-cpu-baseline > $O/r02_bench_B32768.json 2>> $O/r02_bench_B1024.err; tail -c 600 $O/r02_bench_B32768.json
echo "== ncu"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -c 8 -f -o $O/r02_head_mix python tools/prof_target2.py 10 > $O/r02_ncu.log 2>&1; tail -3 $O/r02_ncu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_bench.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $O/r02_launches_bench.log 2>&1
echo "== sanitizer"
for t in memcheck racecheck synccheck; do timeout 400 compute-sanitizer --tool $t python __graft_entry__.py smoke > $O/r02_sanitizer_$t.txt 2>&1; tail -4 $O/r02_sanitizer_$t.txt; done
