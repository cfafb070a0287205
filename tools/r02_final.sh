#!/bin/bash
# round 2, final single-GPU call at HEAD: GPU suite, timings, every-QP sweep, bench lines, ncu evidence, sanitizer
mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $O/r02z_tests.txt
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2 | tee -a $O/r02z_tests.txt
echo "== perf_quick"; timeout 200 python tools/perf_quick.py 10 | tee $O/r02z_base10.txt; timeout 300 python tools/perf_quick.py 20 | tee $O/r02z_base20.txt
echo "== bench"; timeout 600 python bench.py > $O/r02z_bench_B1024.json 2> $O/r02z_bench.err; tail -c 1200 $O/r02z_bench_B1024.json; tail -3 $O/r02z_bench.err
timeout 300 python bench.py --batch 32768 --steps 100 --no-cpu-baseline > $O/r02z_bench_B32768.json 2>> $O/r02z_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > $O/r02z_bench_reference.json 2>> $O/r02z_bench.err
echo "== ncu"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_kernel --launch-skip 4 -c 4 -f -o $O/r02z_mix1024 python tools/prof_target2.py 10 > $O/r02z_ncu.log 2>&1; tail -1 $O/r02z_ncu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_kernel --launch-skip 12 -c 4 -f -o $O/r02z_mix32768 python tools/prof_target2.py 10 >> $O/r02z_ncu.log 2>&1; tail -1 $O/r02z_ncu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02z_launches_bench.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-subrecords --ring 8 > $O/r02z_launches_bench.log 2>&1
echo "== sanitizer"
for t in memcheck racecheck synccheck; do timeout 400 compute-sanitizer --tool $t python __graft_entry__.py smoke > $O/r02z_sanitizer_$t.txt 2>&1; tail -3 $O/r02z_sanitizer_$t.txt; done
# the team kernels and the queue with more QPs than slots: benchmark mix, N = 10 (B = 600) and N = 20 (B = 96)
for t in memcheck racecheck synccheck; do for n in 10 20; do b=$([ $n = 10 ] && echo 600 || echo 96); echo "== $t prof_target2.py $n $b" >> $O/r02z_sanitizer_mix.txt; timeout 600 compute-sanitizer --tool $t python tools/prof_target2.py $n $b 2>&1 | tail -2 >> $O/r02z_sanitizer_mix.txt; done; done; cat $O/r02z_sanitizer_mix.txt
echo "== robust sweep"; timeout 900 python tools/robust_sweep.py 2>&1 | tail -12 | tee $O/r02z_robust.txt
