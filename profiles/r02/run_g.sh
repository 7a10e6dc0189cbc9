#!/bin/bash
# round 2, call G: ordered (sort-based) backward tests + timing; ncu --set full of resolve strip + backward strip on c2
set -u
mkdir -p gpurun_out/r02g
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_reference_dropin.py tests/test_gpu_reference_jacobians.py -m gpu -q -x > gpurun_out/r02g/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02g/pytest.log
tail -12 gpurun_out/r02g/pytest.log
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2]))
    print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["roofline"]["stages_ms_per_step"].items() if v})
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
for c in c2 c3 c5; do
  timeout 600 python bench.py --config $c --mode ordered --no-cpu-baseline --no-e2e --steps 5 --warmup 3 > gpurun_out/r02g/bench_${c}_ordered.json 2> gpurun_out/r02g/bench_${c}_ordered.err
  show "$c ordered" gpurun_out/r02g/bench_${c}_ordered.json
done
CMD="python bench.py --no-cpu-baseline --no-e2e --no-parity --steps 2 --warmup 3"
$CMD > gpurun_out/r02g/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'backward_blocks_kernel|resolve_strip' -s 6 -c 2 -o gpurun_out/r02g/prof_c2 $CMD > gpurun_out/r02g/ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r02g/ncu.log
