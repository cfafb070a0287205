#!/bin/bash
# round 2, tenth gpurun call (8 GPUs): final version of the fused collect against ncclAllGather, config 5; plus the fixed API test
mkdir -p gpurun_out; O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_zy_round2_api.py -m gpu -q 2>&1 | tail -3
run() { n=$1; name=$2; shift; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 298$n$n bench.py --gpus $n --steps 400 --warmup 10 --no-cpu-baseline "$@" > $O/r02i_${n}_$name.json 2> $O/r02i_${n}_$name.err; echo "== $n GPUs $name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("$O/r02i_${n}_$name.json").read().strip().splitlines()[-1]); c=d.get("config5",{})
    print("value %.3f M ms %.4f p50 %.1f collect %s verified %s | config5 %.2f M %.3f ms verified %s timeouts %s" % (d["value"]/1e6,d["ms_per_step"],d["p50_solve_us"],d["config"]["final_collect"][:24],d["config"]["final_collect_verified"],c.get("value",0)/1e6,c.get("ms_per_step",0),c.get("final_collect_verified"),c.get("peer_wait_timeouts")))
except Exception as e:
    print("no line:", e)
PY
tail -2 $O/r02i_${n}_$name.err; }
run 8 peer --collect peer
run 8 nccl --collect nccl
