#!/bin/bash
# round 2, fifth gpurun call (--gpus 2): the fused final collect (peer stores from the solve epilogue) against ncclAllGather and no collect
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi topo -m 2>&1 | head -8 | tee $O/r02e_topo.txt
run() { name=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 600 --warmup 10 --no-cpu-baseline "$@" > $O/r02e_$name.json 2> $O/r02e_$name.err; echo "== $name rc=$?"; tail -c 1800 $O/r02e_$name.json; tail -3 $O/r02e_$name.err; }
run peer --collect peer
run nccl --collect nccl
run none --no-gather --no-subrecords
echo "== 1 GPU reference point"; timeout 300 python bench.py --steps 600 --warmup 10 --no-cpu-baseline --no-subrecords > $O/r02e_1gpu.json 2> $O/r02e_1gpu.err; python -c "import json;d=json.loads(open('$O/r02e_1gpu.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'])"
