#!/bin/bash
# round 2, call X: resolve strip kernel at 6 (default) / 5 / 4 CTAs per SM, all configs at 6
set -u
mkdir -p gpurun_out/r02x
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2]))
    print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["roofline"]["stages_ms_per_step"].items() if v})
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
for t in 0 5 4; do
  PMR_TUNING=$t timeout 600 python bench.py --config c2 --no-cpu-baseline --no-e2e --no-parity --no-configs --steps 20 --warmup 5 > gpurun_out/r02x/bench_c2_t$t.json 2> gpurun_out/r02x/bench_c2_t$t.err
  show "c2 tuning $t" gpurun_out/r02x/bench_c2_t$t.json
done
for c in c3 c5 c4; do
  timeout 600 python bench.py --config $c --no-cpu-baseline --no-e2e --no-parity --no-configs --steps 20 --warmup 5 > gpurun_out/r02x/bench_$c.json 2> gpurun_out/r02x/bench_$c.err
  show "$c" gpurun_out/r02x/bench_$c.json
done
CMD="python bench.py --config c2 --no-cpu-baseline --no-e2e --no-parity --no-configs --no-graph --steps 2 --warmup 3"
$CMD > gpurun_out/r02x/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'resolve_kernel' -s 3 -c 1 -o gpurun_out/r02x/prof_resolve_strip6 $CMD > gpurun_out/r02x/ncu.log 2>&1
echo "ncu rc=$?"
