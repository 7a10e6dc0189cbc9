#!/bin/bash
# round 2, call W: resolve kernel as a strip loop (keys of 32 x 4 pixels by one tensor copy); 8 / 7 / 6 CTAs per SM
set -u
mkdir -p gpurun_out/r02w
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02w/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02w/pytest.log
tail -5 gpurun_out/r02w/pytest.log
grep -q "pytest rc=0" gpurun_out/r02w/pytest.log || { echo "tests failed: benches skipped"; exit 1; }
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2]))
    print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["roofline"]["stages_ms_per_step"].items() if v})
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
for t in 0 7 6; do
  PMR_TUNING=$t timeout 600 python bench.py --config c2 --no-cpu-baseline --no-e2e --no-parity --no-configs --steps 20 --warmup 5 > gpurun_out/r02w/bench_c2_t$t.json 2> gpurun_out/r02w/bench_c2_t$t.err
  show "c2 tuning $t" gpurun_out/r02w/bench_c2_t$t.json
done
for c in c3 c5 c4 c1; do
  timeout 600 python bench.py --config $c --no-cpu-baseline --no-e2e --no-parity --no-configs --steps 20 --warmup 5 > gpurun_out/r02w/bench_$c.json 2> gpurun_out/r02w/bench_$c.err
  show "$c" gpurun_out/r02w/bench_$c.json
done
PMR_NO_TMA=1 timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-configs --no-parity --steps 10 > gpurun_out/r02w/bench_c2_notma.json 2> gpurun_out/r02w/bench_c2_notma.err
show "c2 no-tma" gpurun_out/r02w/bench_c2_notma.json
CMD="python bench.py --config c2 --no-cpu-baseline --no-e2e --no-parity --no-configs --no-graph --steps 2 --warmup 3"
$CMD > gpurun_out/r02w/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'resolve_kernel' -s 3 -c 1 -o gpurun_out/r02w/prof_resolve_strip $CMD > gpurun_out/r02w/ncu.log 2>&1
echo "ncu rc=$?"
