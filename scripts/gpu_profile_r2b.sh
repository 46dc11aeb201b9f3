#!/usr/bin/env bash
# Round-2 re-profiling after the center kernel's row skipping (run under gpurun): launch list of the timed region and the
# ncu counters of the center kernel (first pass) at the bench's launch shape; every ncu pass follows the same command run plain.
set -uo pipefail
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
ARGS="--images 500 --chunk 250 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extras"
UNMORE_PROFILE_RANGE=1 python bench.py $ARGS > gpurun_out/r02_plain.json 2> gpurun_out/r02_plain.err || { echo "plain run failed"; tail -20 gpurun_out/r02_plain.err; exit 1; }
UNMORE_PROFILE_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python bench.py $ARGS > gpurun_out/r02_ncu_launches.log 2>&1
echo "launch list rc=$?"
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,lts__t_sectors_op_read.sum,l1tex__t_sector_hit_rate.pct"
python scripts/kbench.py 250 center > gpurun_out/r02_kbench250_center.log 2>&1; echo "kbench250 rc=$?"; cat gpurun_out/r02_kbench250_center.log
ncu --metrics $M --clock-control none -k regex:"center_kernel" -s 2 -c 2 --csv --log-file gpurun_out/r02_traffic_center.csv python scripts/kbench.py 250 center > gpurun_out/r02_ncu_traffic_center.log 2>&1; echo "traffic rc=$?"
grep -v "^==" gpurun_out/r02_traffic_center.csv | cut -d, -f5,13- | head -30
