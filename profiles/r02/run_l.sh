#!/bin/bash
# round 2, call L: full GPU test suite on the cleaned-up library; ordered-mode timing; default bench line
set -u
mkdir -p gpurun_out/r02l
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02l/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02l/pytest.log
tail -6 gpurun_out/r02l/pytest.log
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2]))
    print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["roofline"]["stages_ms_per_step"].items() if v})
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
for c in c2 c3; do
  timeout 600 python bench.py --config $c --mode ordered --no-cpu-baseline --no-e2e --no-parity --no-configs --steps 5 --warmup 3 > gpurun_out/r02l/bench_${c}_ordered.json 2> gpurun_out/r02l/bench_${c}_ordered.err
  show "$c ordered" gpurun_out/r02l/bench_${c}_ordered.json
done
timeout 900 python bench.py > gpurun_out/r02l/bench_default.json 2> gpurun_out/r02l/bench_default.err; echo "default bench rc=$?"
show "default" gpurun_out/r02l/bench_default.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02l/bench_default.json'))
print('e2e', d['e2e']); print('cpu', d['cpu_baseline']); print('step frac', d['roofline']['step']['frac'], 'bwd frac', d['roofline']['backward']['frac'], 'fwd', d['roofline']['forward']['frac'])
for c in d.get('configs') or []:
    print({k:(round(v,4) if isinstance(v,float) else v) for k,v in c.items() if k not in ('parity','stages_ms_per_step')})
    if 'parity' in c: print('   parity dv', c['parity'].get('d_vertices'), 'bits', c['parity'].get('ids_bit_exact'), c['parity'].get('image_bit_exact'))
PY
