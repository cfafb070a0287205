#!/bin/bash
# dev tool: config-5 shaped runs on 8 GPUs of one box
python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-8} --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus ${NG:-8} --steps 200 --warmup 10 --no-cpu-baseline 2>gpurun_out/run8.err | tail -1 > gpurun_out/bench_r1_${NG:-8}gpu_B1024.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-8} --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus ${NG:-8} --batch 32768 --steps 30 --warmup 5 --no-cpu-baseline 2>>gpurun_out/run8.err | tail -1 > gpurun_out/bench_r1_${NG:-8}gpu_B32768.json
python - <<'PY'
import json
for f in ("gpurun_out/bench_r1_${NG:-8}gpu_B1024.json", "gpurun_out/bench_r1_${NG:-8}gpu_B32768.json"):
    try:
        d = json.load(open(f)); print(f, "value %.0f" % d["value"], "ms/step %.3f" % d["ms_per_step"], d["n_gpus"], d["config"]["final_collect"], "e2e %.0f" % d["e2e"]["value"])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 gpurun_out/run8.err
