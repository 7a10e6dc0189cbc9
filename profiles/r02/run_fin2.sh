#!/bin/bash
# round 2, call FIN2 (final code): full GPU tests; default bench line; ncu launch list and --set full capture of the same command
set -u
mkdir -p gpurun_out/r02fin2
timeout 900 python bench.py > gpurun_out/r02fin2/bench_default.json 2> gpurun_out/r02fin2/bench_default.err; echo "default rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02fin2/bench_default.json'))
print('c2 ms/step', d['ms_per_step'], 'value', d['value'], d['config']['launches'])
print('stages', {k:round(v,4) for k,v in d['roofline']['stages_ms_per_step'].items() if v})
print('fracs step', d['roofline']['step']['frac'], 'fwd', d['roofline']['forward']['frac'], 'bwd', d['roofline']['backward']['frac'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['host_link_gbs_per_rank'], d['e2e']['host_link_ceiling_gbs_per_rank'])
for c in d.get('configs') or []:
    print({k:(round(v,4) if isinstance(v,float) else v) for k,v in c.items() if k in ('workload','ms_per_step','hbm_frac_step','frac_of_slower_roofline','error')})
PY
CMD="python bench.py --no-cpu-baseline --no-e2e --no-parity --no-configs --no-graph --steps 5 --warmup 3"
$CMD > gpurun_out/r02fin2/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 80 --csv --log-file gpurun_out/r02fin2/launches_c2_r2.csv $CMD > gpurun_out/r02fin2/ncu_list.log 2>&1
echo "ncu list rc=$?"
$CMD > gpurun_out/r02fin2/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'scatter_small_kernel|resolve_kernel|backward_blocks_kernel' -s 12 -c 3 -o gpurun_out/r02fin2/prof_c2_r2 $CMD > gpurun_out/r02fin2/ncu_full.log 2>&1
echo "ncu full rc=$?"
