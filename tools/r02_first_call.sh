#!/bin/bash
# dev tool for round 2, first gpurun call: (1) the GPU suite incl. the rows written after round 1's GPU minutes were spent
# (warm start, kinematics, EKF), (2) A/B of the macro-guarded variants prepared on the CPU emulator.
#   build the variants first (here, no GPU needed):
#     for v in r1algo:"-DA1MPC_GUESS_TAPIA=0 -DA1MPC_INIT_LAM=1.0 -DA1MPC_INIT_CENTRED=0 -DA1MPC_FIN_HYST=0" ffrag:-DA1MPC_FORM_FRAG=1 sswitch:-DA1MPC_SOLVE_SWITCH=1 rsqrtlib:-DA1MPC_RSQRT_NB=0 fixedref:"-DA1MPC_FIXED_REFINE=1 -DA1MPC_FIN_HYST=1"; do n=${v%%:*}; f=${v#*:}; \
#       make -j8 OBJ=build_$n LIB=ab/liba1mpc_$n.so EXTRA="$f" ab/liba1mpc_$n.so; done
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/r02_tests.txt
echo "== base"; timeout 200 python tools/perf_quick.py 10 | tee gpurun_out/r02_base10.txt; timeout 200 python tools/perf_quick.py 20 | tee gpurun_out/r02_base20.txt
for v in r1algo ffrag rsqrtlib fixedref; do [ -f ab/liba1mpc_$v.so ] && { echo "== $v"; A1MPC_LIB=$PWD/ab/liba1mpc_$v.so timeout 200 python tools/perf_quick.py 10 | tee gpurun_out/r02_$v.txt; }; done
[ -f ab/liba1mpc_sswitch.so ] && { echo "== sswitch (N=20)"; A1MPC_LIB=$PWD/ab/liba1mpc_sswitch.so timeout 200 python tools/perf_quick.py 20 | tee gpurun_out/r02_sswitch20.txt; }
timeout 100 python tools/hard_qp.py | tee gpurun_out/r02_hard_qp.txt
timeout 400 python tools/robust_sweep.py | tee gpurun_out/r02_robust.txt
for n in 0.0 0.1 0.3; do timeout 100 python tools/warm_bench.py 1024 $n; timeout 100 python tools/warm_bench.py 16384 $n; done | tee gpurun_out/r02_warm.txt
echo "== config 4, general extended kernel"; timeout 200 python tools/ext_probe.py | tee gpurun_out/r02_ext.txt
echo "== config 4, compacted class (A1MPC_EXT_COMPACT=1)"; A1MPC_EXT_COMPACT=1 timeout 200 python tools/ext_probe.py | tee gpurun_out/r02_ext_compact.txt
