#!/bin/bash
# round 2, call B: ncu --set full of the new backward kernel on c2 (after the same command ran clean without ncu)
set -u
mkdir -p gpurun_out/r02b
CMD="python bench.py --no-cpu-baseline --no-e2e --steps 2 --warmup 3"
$CMD > gpurun_out/r02b/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:backward_blocks -s 3 -c 1 -o gpurun_out/r02b/prof_bwd_v3 $CMD > gpurun_out/r02b/ncu.log 2>&1
echo "ncu rc=$?"
tail -3 gpurun_out/r02b/ncu.log
