#!/bin/bash
# round 2, call D: backward kernel variants (strip loop at 6 / 5 CTAs per SM, one block per warp at 8 CTAs)
set -u
mkdir -p gpurun_out/r02e
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_dropin.py tests/test_gpu_full_size.py -m gpu -q > gpurun_out/r02e/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e/pytest.log
tail -8 gpurun_out/r02e/pytest.log
for var in 0 1 2 3; do
  PMR_BWD_VARIANT=$var timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-parity --steps 30 --warmup 5 > gpurun_out/r02e/bench_c2_var$var.json 2> gpurun_out/r02e/bench_c2_var$var.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02e/bench_c2_var$var.json"))
    print("variant=$var", "ms/step", round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["roofline"]["stages_ms_per_step"].items()})
except Exception as e:
    print("variant=$var failed", e)
PY
done
for var in 0 1; do
  PMR_BWD_VARIANT=$var PMR_BACKWARD_VARIANT=$var timeout 300 python bench.py --config c5 --batch 8 --no-cpu-baseline --no-e2e --no-parity --steps 5 --warmup 3 > gpurun_out/r02e/bench_c5_var$var.json 2> gpurun_out/r02e/bench_c5_var$var.err
  python -c "
import json
d=json.load(open('gpurun_out/r02e/bench_c5_var$var.json')); print('c5 b8 variant $var', d['ms_per_step'], d['roofline']['stages_ms_per_step'])"
done
timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 5 --warmup 3 > gpurun_out/r02e/bench_c2_parity.json 2> gpurun_out/r02e/bench_c2_parity.err
python -c "
import json
d=json.load(open('gpurun_out/r02e/bench_c2_parity.json')); print(json.dumps(d['parity']['d_vertices'])); print(json.dumps(d['parity']['d_attributes']))"
