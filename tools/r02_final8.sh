#!/bin/bash
# round 2, final multi-GPU call (8 x B200, one box): the driver's launch line at N = 8, 4, 2, 1 with the committed library; ncclAllGather at N = 8 for comparison
mkdir -p gpurun_out; O=gpurun_out
run() { n=$1; name=$2; shift; shift
  if [ $n = 1 ]; then timeout 300 python bench.py --gpus 1 --steps 1000 --warmup 10 --no-cpu-baseline "$@" > $O/r02y_${n}_$name.json 2> $O/r02y_${n}_$name.err
  else timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$n$n bench.py --gpus $n --steps 1000 --warmup 10 --no-cpu-baseline "$@" > $O/r02y_${n}_$name.json 2> $O/r02y_${n}_$name.err; fi
  echo "== $n GPUs $name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("$O/r02y_${n}_$name.json").read().strip().splitlines()[-1]); c=d.get("config5",{})
    print("value %.3f M ms %.4f p50 %.1f e2e %.3f M collect %s verified %s | config5 %.2f M %.3f ms verified %s timeouts %s" % (d["value"]/1e6,d["ms_per_step"],d["p50_solve_us"],d["e2e"]["value"]/1e6,str(d["config"].get("final_collect"))[:24],d["config"].get("final_collect_verified"),c.get("value",0)/1e6,c.get("ms_per_step",0),c.get("final_collect_verified"),c.get("peer_wait_timeouts")))
except Exception as e:
    print("no line:", e)
PY
tail -2 $O/r02y_${n}_$name.err; }
run 8 auto
run 8 nccl --collect nccl
run 4 auto
run 2 auto
run 1 auto
