#!/usr/bin/env bash
# Round-2 profiling pass (run under gpurun).  Order matters: every ncu pass comes AFTER the same command exited 0
# without ncu; numbers printed under ncu are never bench values.
#   1. the full default bench line                                   -> gpurun_out/r02_bench.json
#   2. plain + ncu launch list of the timed region (500 images)      -> r02_plain.json, r02_launches.csv
#   3. ncu dram bytes + issue-slot utilisation per hot kernel at the bench's launch shape (250 images x 4096
#      proposals; the SAT scan at the full 5000-image batch = 10 000 planes)   -> r02_traffic_*.csv
set -uo pipefail
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_bench.err
ARGS="--images 500 --chunk 250 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extras"
UNMORE_PROFILE_RANGE=1 python bench.py $ARGS > gpurun_out/r02_plain.json 2> gpurun_out/r02_plain.err || { echo "plain run failed"; tail -20 gpurun_out/r02_plain.err; exit 1; }
UNMORE_PROFILE_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python bench.py $ARGS > gpurun_out/r02_ncu_launches.log 2>&1
echo "launch list rc=$?"
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum"
python scripts/kbench.py 250 exist center refine > gpurun_out/r02_kbench250.log 2>&1; echo "kbench250 rc=$?"; cat gpurun_out/r02_kbench250.log
ncu --metrics $M --clock-control none -k regex:"refine_kernel|center_kernel|existence_kernel" -s 6 -c 3 --csv --log-file gpurun_out/r02_traffic_250.csv python scripts/kbench.py 250 exist center refine > gpurun_out/r02_ncu_traffic.log 2>&1; echo "traffic rc=$?"
python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-extras > /dev/null 2> gpurun_out/r02_sat_plain.err; echo "sat plain rc=$?"
ncu --metrics $M --clock-control none -k regex:sat_kernel -c 1 --csv --log-file gpurun_out/r02_traffic_sat.csv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_sat.log 2>&1; echo "sat traffic rc=$?"
ls -la gpurun_out | tail -15
