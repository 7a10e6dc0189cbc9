#!/bin/bash
# round 2, call N: ordered mode through per-pixel rows (tests + timing + launch list)
set -u
mkdir -p gpurun_out/r02n
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_reference_dropin.py tests/test_gpu_reference_jacobians.py tests/test_gpu_render.py -m gpu -q -x > gpurun_out/r02n/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02n/pytest.log
tail -6 gpurun_out/r02n/pytest.log
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
  timeout 600 python bench.py --config $c --mode ordered --no-cpu-baseline --no-e2e --no-configs --steps 5 --warmup 3 > gpurun_out/r02n/bench_${c}_ordered.json 2> gpurun_out/r02n/bench_${c}_ordered.err
  show "$c ordered" gpurun_out/r02n/bench_${c}_ordered.json
done
CMD="python bench.py --mode ordered --no-cpu-baseline --no-e2e --no-parity --no-configs --steps 2 --warmup 3"
$CMD > gpurun_out/r02n/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 45 -c 30 --csv --log-file gpurun_out/r02n/launches_c2_ordered.csv $CMD > gpurun_out/r02n/ncu.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02n/launches_c2_ordered.csv')) if len(r)>10]
hdr=rows[0]; k=hdr.index('Kernel Name'); v=hdr.index('Metric Value'); u=hdr.index('Metric Unit')
from collections import OrderedDict
agg=OrderedDict()
for r in rows[1:]:
    val=float(r[v].replace(',','')); val = val/1000 if r[u]=='ns' else val
    agg.setdefault(r[k].split('(')[0][:50],[]).append(val)
for n,t in agg.items(): print('%-52s n=%2d mean %9.1f us'%(n,len(t),sum(t)/len(t)))
PY
