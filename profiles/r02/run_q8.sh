#!/bin/bash
# round 2, call Q8 (8 GPUs): peer-exchange tests at 8 ranks, c4 strong-scaling bench at N = 8 and 4 (peer), N = 8 NCCL
set -u
mkdir -p gpurun_out/r02q8
nvidia-smi topo -m > gpurun_out/r02q8/topo.txt 2>&1
run() {  # n collective tag
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 --steps 30 --warmup 5 --collective $2 $4 > gpurun_out/r02q8/bench_$3.json 2> gpurun_out/r02q8/bench_$3.err
  echo "bench $3 rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02q8/bench_$3.json"))
    sg=d.get("single_gpu",{}).get("ms_per_step")
    print("$3", d["scaling"], "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "single_gpu", sg, "speedup", round(sg/d["ms_per_step"],3) if sg else None, "check", d.get("exchange_check"))
    if d.get("e2e"): print("   e2e", {k:d["e2e"][k] for k in ("value","ms_per_step","host_link_gbs_per_rank","host_link_ceiling_gbs_per_rank")})
    print("   stages", {k:round(v,4) for k,v in d["roofline"]["stages_ms_per_step"].items() if v}, "launches", d["gpu_launches"])
except Exception as e:
    print("$3 failed", e); print(open("gpurun_out/r02q8/bench_$3.err").read()[-1500:])
PY
}
run 8 peer n8_peer_graph "--no-e2e"
run 8 nccl n8_nccl_graph "--no-e2e"
run 4 peer n4_peer_graph "--no-e2e"
run 2 peer n2_peer_graph "--no-e2e"
