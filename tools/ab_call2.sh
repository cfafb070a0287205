#!/bin/bash
# dev tool (one gpurun call): GPU parity tests of the default build, then timing of the ab/ variants
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/ab2_tests.txt
tail -3 gpurun_out/ab2_tests.txt
echo "== base"; timeout 200 python tools/perf_quick.py 10 2>&1 | tee gpurun_out/ab2_base.txt
for v in "$@"; do
  echo "== $v"; A1MPC_LIB=$PWD/ab/liba1mpc_$v.so timeout 200 python tools/perf_quick.py 10 2>&1 | tee gpurun_out/ab2_$v.txt
done
