#!/bin/bash
# round 2, call C: v4 backward kernel (single TMA box, opaque thread index): occupancy variants + parity field
set -u
mkdir -p gpurun_out/r02c
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_reference_dropin.py tests/test_gpu_render.py tests/test_gpu_shade.py -m gpu -x -q > gpurun_out/r02c/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c/pytest.log
tail -15 gpurun_out/r02c/pytest.log
for occ in 0 5 7 8; do
  PMR_BWD_OCC=$occ timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-parity --steps 30 --warmup 5 > gpurun_out/r02c/bench_c2_occ$occ.json 2> gpurun_out/r02c/bench_c2_occ$occ.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02c/bench_c2_occ$occ.json"))
    print("occ=$occ", "ms/step", round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["roofline"]["stages_ms_per_step"].items()})
except Exception as e:
    print("occ=$occ failed", e)
PY
done
timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 10 --warmup 3 > gpurun_out/r02c/bench_c2_parity.json 2> gpurun_out/r02c/bench_c2_parity.err
python -c "
import json
d=json.load(open('gpurun_out/r02c/bench_c2_parity.json')); print(json.dumps(d['parity'], indent=1))"
timeout 600 python bench.py --config c5 --batch 8 --no-cpu-baseline --no-e2e --steps 5 --warmup 3 > gpurun_out/r02c/bench_c5.json 2> gpurun_out/r02c/bench_c5.err
python -c "
import json
d=json.load(open('gpurun_out/r02c/bench_c5.json')); print('c5 b8', d['ms_per_step'], d['roofline']['stages_ms_per_step']); print(json.dumps(d['parity'], indent=1))"
