#!/bin/bash
# round 2, call F: backward variants after the register-only predicated adds; resolve strip kernel A/B
set -u
mkdir -p gpurun_out/r02f
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -q -x > gpurun_out/r02f/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f/pytest.log
tail -5 gpurun_out/r02f/pytest.log
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2]))
    print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["roofline"]["stages_ms_per_step"].items() if v})
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
for var in 0 1 2 3; do
  PMR_BWD_VARIANT=$var timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-parity --steps 30 --warmup 5 > gpurun_out/r02f/bench_c2_var$var.json 2> gpurun_out/r02f/bench_c2_var$var.err
  show "bwd variant=$var" gpurun_out/r02f/bench_c2_var$var.json
done
PMR_RESOLVE_VARIANT=1 timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-parity --steps 30 --warmup 5 > gpurun_out/r02f/bench_c2_resolve_old.json 2> gpurun_out/r02f/bench_c2_resolve_old.err
show "resolve old" gpurun_out/r02f/bench_c2_resolve_old.json
for c in c3 c5; do
  timeout 600 python bench.py --config $c --no-cpu-baseline --no-e2e --no-parity --steps 10 --warmup 3 > gpurun_out/r02f/bench_$c.json 2> gpurun_out/r02f/bench_$c.err
  show "$c" gpurun_out/r02f/bench_$c.json
done
