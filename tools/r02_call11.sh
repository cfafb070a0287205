#!/bin/bash
# round 2, call 11: rendezvous A/B at the benchmark batch (does lock-step cost the slowest QP?), N = 20 with the 2-warp wrench CTAs
mkdir -p gpurun_out; O=gpurun_out
b() { name=$1; lib=$2; echo "== $name"; A1MPC_LIB=$lib timeout 200 python tools/perf_quick.py 10 | grep -E "B=  1024|B= 16384" | tee $O/r02j_$name.txt; A1MPC_LIB=$lib timeout 300 python bench.py --steps 1000 --no-cpu-baseline --no-subrecords 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('bench B=1024: %.3f M  %.4f ms  p50 %.1f  classes %s'%(d['value']/1e6,d['ms_per_step'],d['p50_solve_us'],d['class_kernel_ms_per_step']))" | tee -a $O/r02j_$name.txt; }
b default $PWD/a1-qp-mpc-controller_b200/liba1mpc.so
b norv_wrench $PWD/ab/liba1mpc_norv.so
b norv_all $PWD/ab/liba1mpc_norvall.so
echo "== N=20 default"; timeout 300 python tools/perf_quick.py 20 | tee $O/r02j_n20.txt
