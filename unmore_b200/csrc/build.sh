#!/usr/bin/env bash
# Builds libunmore_b200.so (sm_100a only) next to the package so it travels with the repo snapshot.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../libunmore_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O3,-Wall
       --expt-relaxed-constexpr -Xptxas -v -shared -cudart static)
"${NVCC}" "${FLAGS[@]}" -o "${OUT}" "${HERE}"/*.cu 2> "${HERE}/../_build.log" || { cat "${HERE}/../_build.log" >&2; exit 1; }
echo "built ${OUT}"
