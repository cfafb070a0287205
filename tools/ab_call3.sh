#!/bin/bash
# dev tool (one gpurun call): timing of the ab/ variants named on the command line, then one ncu capture of variant $NCU_V
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
for v in "$@"; do
  echo "== $v"; A1MPC_LIB=$PWD/ab/liba1mpc_$v.so timeout 200 python tools/perf_quick.py 10 2>&1 | tee gpurun_out/ab3_$v.txt
done
if [ -n "$NCU_V" ]; then
  A1MPC_LIB=$PWD/ab/liba1mpc_$NCU_V.so timeout 400 ncu --set full --import-source on --clock-control none -k regex:solve_kernel -s 10 -c 1 -o gpurun_out/r01c_trot_mix16k python tools/prof_target.py 16384 0 > gpurun_out/ncu_c.log 2>&1
  A1MPC_LIB=$PWD/ab/liba1mpc_$NCU_V.so timeout 400 ncu --set full --import-source on --clock-control none -k regex:solve_kernel -s 8 -c 1 -o gpurun_out/r01c_four_mix16k python tools/prof_target.py 16384 0 >> gpurun_out/ncu_c.log 2>&1
  ls -la gpurun_out/*.ncu-rep
fi
