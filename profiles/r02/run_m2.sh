#!/bin/bash
# round 2, call M2 (2 GPUs): peer-exchange tests, c4 strong-scaling bench with the exchange check, both collectives
set -u
mkdir -p gpurun_out/r02m2
timeout 600 python -m pytest tests/test_gpu_peer_exchange.py -m gpu -q > gpurun_out/r02m2/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02m2/pytest.log
tail -4 gpurun_out/r02m2/pytest.log
for coll in peer nccl; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --collective $coll > gpurun_out/r02m2/bench_n2_$coll.json 2> gpurun_out/r02m2/bench_n2_$coll.err
  echo "bench n2 $coll rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02m2/bench_n2_$coll.json"))
    print("$coll", d["scaling"], "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "single_gpu", d.get("single_gpu",{}).get("ms_per_step"), "check", d.get("exchange_check"))
    print("   e2e", {k:d["e2e"][k] for k in ("value","ms_per_step","host_link_gbs_per_rank","host_link_ceiling_gbs_per_rank")})
    print("   stages", {k:round(v,4) for k,v in d["roofline"]["stages_ms_per_step"].items() if v})
except Exception as e:
    print("$coll failed", e); print(open("gpurun_out/r02m2/bench_n2_$coll.err").read()[-1500:])
PY
done
