#!/bin/bash
# round 2, call J: ncu --set full of the backward kernel with staged attributes (c2) + new fold kernel timing
set -u
mkdir -p gpurun_out/r02j
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -q -x > gpurun_out/r02j/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02j/pytest.log
tail -4 gpurun_out/r02j/pytest.log
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
  timeout 600 python bench.py --config $c --mode ordered --no-cpu-baseline --no-e2e --no-parity --no-configs --steps 5 --warmup 3 > gpurun_out/r02j/bench_${c}_ordered.json 2> gpurun_out/r02j/bench_${c}_ordered.err
  show "$c ordered" gpurun_out/r02j/bench_${c}_ordered.json
done
CMD="python bench.py --config c2 --no-cpu-baseline --no-e2e --no-parity --no-configs --steps 2 --warmup 3"
$CMD > gpurun_out/r02j/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'backward_blocks_kernel' -s 3 -c 1 -o gpurun_out/r02j/prof_bwd_staged $CMD > gpurun_out/r02j/ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r02j/ncu.log
