#!/bin/bash
# round 2, call A: GPU tests of the new backward kernel + strip-length / TMA variants on c2
set -u
mkdir -p gpurun_out/r02a
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02a/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a/pytest.log
tail -5 gpurun_out/r02a/pytest.log
for nb in 0 1 2 4 8 16; do
  PMR_STRIP_BLOCKS=$nb timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 30 --warmup 5 > gpurun_out/r02a/bench_c2_nb$nb.json 2> gpurun_out/r02a/bench_c2_nb$nb.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02a/bench_c2_nb$nb.json"))
    print("nb=$nb", "ms/step", round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["roofline"]["stages_ms_per_step"].items()})
except Exception as e:
    print("nb=$nb failed", e)
PY
done
PMR_NO_TMA=1 timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 30 --warmup 5 > gpurun_out/r02a/bench_c2_notma.json 2> gpurun_out/r02a/bench_c2_notma.err
python -c "
import json
d=json.load(open('gpurun_out/r02a/bench_c2_notma.json')); print('no-tma', d['ms_per_step'], d['roofline']['stages_ms_per_step'])"
for c in c1 c3 c5; do
  timeout 600 python bench.py --config $c --no-cpu-baseline --no-e2e --steps 10 --warmup 3 > gpurun_out/r02a/bench_$c.json 2> gpurun_out/r02a/bench_$c.err
  python -c "
import json
d=json.load(open('gpurun_out/r02a/bench_$c.json')); print('$c', d['ms_per_step'], d['roofline']['stages_ms_per_step'])"
done
