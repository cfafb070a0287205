#!/bin/bash
# round 2, third gpurun call: default library after the revert of the QP spreading; bench line with sub-records and the plugin-level
# figure; ncu --set full of the second launch group at both batch sizes (two small reports: gpurun_out is capped at 64 MiB)
mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee $O/r02c_tests.txt
echo "== base"; timeout 200 python tools/perf_quick.py 10 | tee $O/r02c_base10.txt; timeout 300 python tools/perf_quick.py 20 | tee $O/r02c_base20.txt
echo "== bench"; timeout 600 python bench.py > $O/r02c_bench_B1024.json 2> $O/r02c_bench.err; tail -c 3000 $O/r02c_bench_B1024.json; tail -5 $O/r02c_bench.err
timeout 300 python bench.py --batch 32768 --steps 100 --no-cpu-baseline > $O/r02c_bench_B32768.json 2>> $O/r02c_bench.err; tail -c 600 $O/r02c_bench_B32768.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > $O/r02c_bench_reference.json 2>> $O/r02c_bench.err
echo "== ncu"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_kernel --launch-skip 4 -c 4 -f -o $O/r02c_mix1024 python tools/prof_target2.py 10 > $O/r02c_ncu.log 2>&1; tail -2 $O/r02c_ncu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_kernel --launch-skip 12 -c 4 -f -o $O/r02c_mix32768 python tools/prof_target2.py 10 >> $O/r02c_ncu.log 2>&1; tail -2 $O/r02c_ncu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02c_launches_bench.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-subrecords --ring 8 > $O/r02c_launches_bench.log 2>&1
ls -la $O | tail -20
