#!/usr/bin/env bash
# Round-1 profiling pass (run under gpurun): the plain run first, then the ncu launch list of the
# SAME command restricted to the timed region (cudaProfilerStart/Stop inside bench.py), then one
# --set full capture per hot kernel on the kernel micro-benchmark.  Outputs land in gpurun_out/.
set -uo pipefail
ARGS="--images 500 --chunk 250 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extras"
mkdir -p gpurun_out
UNMORE_PROFILE_RANGE=1 python bench.py $ARGS > gpurun_out/r1_plain.json 2> gpurun_out/r1_plain.err || { echo "plain run failed"; tail -20 gpurun_out/r1_plain.err; exit 1; }
UNMORE_PROFILE_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_launches.csv python bench.py $ARGS > gpurun_out/r1_ncu_launches.log 2>&1
echo "launch list rc=$?"
for k in refine center exist sat score; do
  case $k in refine) re=refine_kernel;; center) re=center_kernel;; exist) re=existence_kernel;; sat) re=sat_kernel;; score) re=score_kernel;; esac
  bash scripts/gpu_ncu_kernel.sh $re $k r1_$k 32 | tail -2
done
ls -la gpurun_out | tail -20
