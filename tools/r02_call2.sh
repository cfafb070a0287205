#!/bin/bash
# round 2, second gpurun call: new pin tests + full-size config tests, one-QP-per-CTA mapping, precision 32, bench with sub-records,
# ncu of the new default at both batch sizes
mkdir -p gpurun_out; O=gpurun_out
nproc | tee $O/r02b_gpu.txt; cat /sys/fs/cgroup/cpu.max 2>/dev/null | tee -a $O/r02b_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee $O/r02b_tests.txt
echo "== base"; timeout 200 python tools/perf_quick.py 10 | tee $O/r02b_base10.txt; timeout 300 python tools/perf_quick.py 20 | tee $O/r02b_base20.txt
[ -f ab/liba1mpc_n20w3.so ] && { echo "== n20w3 (N=20, 3 warps per CTA for the trot class)"; A1MPC_LIB=$PWD/ab/liba1mpc_n20w3.so timeout 300 python tools/perf_quick.py 20 | tee $O/r02b_n20w3.txt; }
echo "== bench"; timeout 600 python bench.py > $O/r02b_bench_B1024.json 2> $O/r02b_bench.err; tail -c 2500 $O/r02b_bench_B1024.json; tail -5 $O/r02b_bench.err
timeout 300 python bench.py --batch 32768 --steps 100 --no-cpu-baseline > $O/r02b_bench_B32768.json 2>> $O/r02b_bench.err; tail -c 900 $O/r02b_bench_B32768.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > $O/r02b_bench_reference.json 2>> $O/r02b_bench.err; cat $O/r02b_bench_reference.json
echo "== ncu"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -c 16 -f -o $O/r02b_mix python tools/prof_target2.py 10 > $O/r02b_ncu.log 2>&1; tail -3 $O/r02b_ncu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02b_launches_bench.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-subrecords --ring 8 > $O/r02b_launches_bench.log 2>&1
