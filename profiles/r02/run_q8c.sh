#!/bin/bash
# round 2, call Q8c (8 GPUs): c4 strong scaling with the views dealt round-robin
set -u
mkdir -p gpurun_out/r02q8c
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 30 --warmup 5 --collective peer --no-e2e > gpurun_out/r02q8c/bench_n8_peer_graph.json 2> gpurun_out/r02q8c/bench_n8_peer_graph.err
echo "rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r02q8c/bench_n8_peer_graph.json"))
    sg=d.get("single_gpu",{}).get("ms_per_step")
    print(d["scaling"], "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "single_gpu", sg, "speedup", round(sg/d["ms_per_step"],3), "check", d.get("exchange_check"))
    print("   stages", {k:round(v,4) for k,v in d["roofline"]["stages_ms_per_step"].items() if v}, "launches", d["gpu_launches"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r02q8c/bench_n8_peer_graph.err").read()[-2500:])
PY
