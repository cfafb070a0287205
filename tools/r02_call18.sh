#!/bin/bash
# round 2, call 18 (last GPU minutes): team Cholesky with one barrier per column (mode 3, default) against mode 2, then the evidence set for the default
mkdir -p gpurun_out; O=gpurun_out
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee $O/r02s_tests.txt
for v in new m2; do
  echo "== $v"
  if [ $v = new ]; then unset A1MPC_LIB; else export A1MPC_LIB=$PWD/ab/liba1mpc_$v.so; fi
  timeout 120 python tools/perf_quick.py 10 2>&1 | grep -E "four|mix" | tee $O/r02s_n10_$v.txt
  timeout 120 python tools/perf_quick.py 20 2>&1 | grep -E "B= 16384|B=     1" | tee $O/r02s_n20_$v.txt
  timeout 120 python bench.py --steps 1000 --no-cpu-baseline --no-subrecords 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('bench B=1024: %.3f M  %.4f ms  p50 %.1f  classes %s'%(d['value']/1e6,d['ms_per_step'],d['p50_solve_us'],d['class_kernel_ms_per_step']))" | tee $O/r02s_bench_$v.txt
  timeout 120 python bench.py --batch 32768 --steps 100 --no-cpu-baseline --no-subrecords 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('bench B=32768: %.3f M  %.4f ms  classes %s'%(d['value']/1e6,d['ms_per_step'],d['class_kernel_ms_per_step']))" | tee -a $O/r02s_bench_$v.txt
done
unset A1MPC_LIB
echo "== sanitizer (default)"
for t in memcheck racecheck synccheck; do for n in 10 20; do b=$([ $n = 10 ] && echo 600 || echo 96); echo "== $t prof_target2.py $n $b" >> $O/r02s_sanitizer_mix.txt; timeout 200 compute-sanitizer --tool $t python tools/prof_target2.py $n $b 2>&1 | tail -2 >> $O/r02s_sanitizer_mix.txt; done; done; cat $O/r02s_sanitizer_mix.txt
echo "== ncu"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:solve_kernel --launch-skip 4 -c 4 -f -o $O/r02s_mix1024 python tools/prof_target2.py 10 > $O/r02s_ncu.log 2>&1; tail -1 $O/r02s_ncu.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:solve_kernel --launch-skip 12 -c 4 -f -o $O/r02s_mix32768 python tools/prof_target2.py 10 >> $O/r02s_ncu.log 2>&1; tail -1 $O/r02s_ncu.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02s_launches_bench.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-subrecords --ring 8 > $O/r02s_launches_bench.log 2>&1
echo "== bench"; timeout 400 python bench.py > $O/r02s_bench_B1024.json 2> $O/r02s_bench.err; tail -c 600 $O/r02s_bench_B1024.json
timeout 200 python bench.py --batch 32768 --steps 100 --no-cpu-baseline > $O/r02s_bench_B32768.json 2>> $O/r02s_bench.err
echo "== robust sweep"; timeout 400 python tools/robust_sweep.py 2>&1 | tail -12 | tee $O/r02s_robust.txt
timeout 100 python tools/perf_quick.py 10 > $O/r02s_base10.txt 2>&1; timeout 100 python tools/perf_quick.py 20 > $O/r02s_base20.txt 2>&1
