#!/bin/bash
# round 2, call O: CUDA-graph replay of the step vs eager launches (c2, 1 GPU); default line with configs
set -u
mkdir -p gpurun_out/r02o
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2]))
    print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), "value", round(d["value"],1), d["config"].get("launches"), {k:round(v,4) for k,v in d["roofline"]["stages_ms_per_step"].items() if v})
except Exception as e:
    print(sys.argv[1], "failed", e); print(open(sys.argv[2].replace('.json','.err')).read()[-2000:])
PY
}
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-parity --no-configs --no-graph --steps 50 > gpurun_out/r02o/bench_c2_eager.json 2> gpurun_out/r02o/bench_c2_eager.err
show "c2 eager" gpurun_out/r02o/bench_c2_eager.json
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-parity --no-configs --steps 50 > gpurun_out/r02o/bench_c2_graph.json 2> gpurun_out/r02o/bench_c2_graph.err
show "c2 graph" gpurun_out/r02o/bench_c2_graph.json
timeout 600 python bench.py --mode ordered --no-cpu-baseline --no-e2e --no-configs --steps 10 > gpurun_out/r02o/bench_c2_ordered.json 2> gpurun_out/r02o/bench_c2_ordered.err
show "c2 ordered graph" gpurun_out/r02o/bench_c2_ordered.json
timeout 900 python bench.py > gpurun_out/r02o/bench_default.json 2> gpurun_out/r02o/bench_default.err; echo "default rc=$?"
show "default" gpurun_out/r02o/bench_default.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02o/bench_default.json'))
for c in d.get('configs') or []:
    print({k:(round(v,4) if isinstance(v,float) else v) for k,v in c.items() if k in ('workload','ms_per_step','launches','hbm_frac_step','frac_of_slower_roofline','error')})
PY
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02o/bench_reference.json 2> gpurun_out/r02o/bench_reference.err; echo "reference rc=$?"; cat gpurun_out/r02o/bench_reference.json | cut -c1-600
