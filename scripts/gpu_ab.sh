#!/usr/bin/env bash
# A/B of kernel builds on the kernel micro-benchmark (run under gpurun):
#   gpu_ab.sh <n_images> <kernels...> -- <libA> <libB> ...
# The first library is the reference: its outputs are saved and every other build is compared bit-for-bit.
set -uo pipefail
NI="$1"; shift
KS=()
while [ "$1" != "--" ]; do KS+=("$1"); shift; done
shift
mkdir -p gpurun_out
first=1
for LIB in "$@"; do
  tag=$(basename "$LIB" .so)
  if [ $first = 1 ]; then
    UNMORE_B200_LIB=$LIB KB_SAVE=gpurun_out/kb_ref.pt python scripts/kbench.py $NI "${KS[@]}" > gpurun_out/ab_$tag.log 2>&1
    first=0
  else
    UNMORE_B200_LIB=$LIB KB_CMP=gpurun_out/kb_ref.pt python scripts/kbench.py $NI "${KS[@]}" > gpurun_out/ab_$tag.log 2>&1
  fi
  echo "== $tag (rc=$?)"; grep -E "^(exist|center|center2|refine|score|sat|pack) +:|vs saved|Error|error" gpurun_out/ab_$tag.log
done
rm -f gpurun_out/kb_ref.pt*   # large; only needed between the builds of one call
