#!/bin/bash
# round 2, call I: packed vertex records + quarter-aligned reducers; launch list of the ordered backward
set -u
mkdir -p gpurun_out/r02i
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_render.py tests/test_gpu_shade.py -m gpu -q -x > gpurun_out/r02i/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02i/pytest.log
tail -6 gpurun_out/r02i/pytest.log
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2]))
    print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["roofline"]["stages_ms_per_step"].items() if v})
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
for c in c2 c3 c5 c1; do
  timeout 600 python bench.py --config $c --no-cpu-baseline --no-e2e --no-parity --steps 20 --warmup 5 > gpurun_out/r02i/bench_$c.json 2> gpurun_out/r02i/bench_$c.err
  show "$c" gpurun_out/r02i/bench_$c.json
done
PMR_RESOLVE_VARIANT=1 timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-parity --steps 30 --warmup 5 > gpurun_out/r02i/bench_c2_resolve_old.json 2> gpurun_out/r02i/bench_c2_resolve_old.err
show "c2 resolve old" gpurun_out/r02i/bench_c2_resolve_old.json
CMD="python bench.py --mode ordered --no-cpu-baseline --no-e2e --no-parity --steps 2 --warmup 3"
$CMD > gpurun_out/r02i/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/r02i/launches_c2_ordered.csv $CMD > gpurun_out/r02i/ncu.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02i/launches_c2_ordered.csv')) if len(r)>10]
hdr=rows[0]; k=hdr.index('Kernel Name'); v=hdr.index('Metric Value')
from collections import OrderedDict
agg=OrderedDict()
for r in rows[1:]:
    name=r[k].split('(')[0][:50]; agg.setdefault(name,[]).append(float(r[v].replace(',','')))
for n,t in agg.items(): print('%-52s n=%2d mean %9.1f us'%(n,len(t),sum(t)/len(t)/1000 if max(t)>1e5 else sum(t)/len(t)))
PY
