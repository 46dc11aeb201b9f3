#!/usr/bin/env bash
# Round-1 profiling pass (run under gpurun): plain run first, then the ncu launch list and one
# --set full capture per hot kernel.  Outputs land in gpurun_out/.
set -uo pipefail
ARGS="--images 250 --chunk 250 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extras"
mkdir -p gpurun_out
python bench.py $ARGS > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py $ARGS > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'refine_kernel|center_kernel|existence_kernel|sat_kernel|score_kernel' -s 10 -c 5 -o gpurun_out/prof_r1 -f python bench.py $ARGS > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out
