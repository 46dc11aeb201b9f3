#!/usr/bin/env bash
# usage: gpu_ncu_kernel.sh <kernel-regex> <kbench kernel name> <out-name> [n_images]
set -uo pipefail
K="$1"; NAME="$2"; OUT="$3"; NI="${4:-32}"
mkdir -p gpurun_out
python scripts/kbench.py $NI $NAME > gpurun_out/${OUT}_plain.log 2>&1 || { tail -20 gpurun_out/${OUT}_plain.log; exit 1; }
cat gpurun_out/${OUT}_plain.log
ncu --set full --clock-control none --import-source on -k regex:"$K" -s 2 -c 1 -o gpurun_out/$OUT -f python scripts/kbench.py $NI $NAME > gpurun_out/${OUT}_ncu.log 2>&1
echo "ncu rc=$?"; grep -E "PROF|ERROR" gpurun_out/${OUT}_ncu.log | head
