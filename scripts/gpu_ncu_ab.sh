#!/usr/bin/env bash
# usage: gpu_ncu_ab.sh <kernel-regex> <kbench name> <libA> <libB>   -> gpurun_out/ab_A.ncu-rep, ab_B.ncu-rep
set -uo pipefail
K="$1"; NAME="$2"
for tag in A B; do
  if [ $tag = A ]; then LIB="$3"; else LIB="$4"; fi
  UNMORE_B200_LIB=$LIB python scripts/kbench.py 32 $NAME > gpurun_out/ab_${tag}_plain.log 2>&1 || { tail gpurun_out/ab_${tag}_plain.log; exit 1; }
  cat gpurun_out/ab_${tag}_plain.log
  UNMORE_B200_LIB=$LIB ncu --set full --clock-control none --import-source on -k regex:"$K" -s 2 -c 1 -o gpurun_out/ab_$tag -f python scripts/kbench.py 32 $NAME > gpurun_out/ab_${tag}_ncu.log 2>&1
  echo "ncu $tag rc=$?"
done
