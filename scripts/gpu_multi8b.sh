#!/usr/bin/env bash
# N-GPU re-run of the strong / weak scaling lines only (after a kernel change): see gpu_multi8.sh for the full set
set -uo pipefail
N="${1:-8}"
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N "$@"; }
run --images 5000 --steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-e2e --scaling strong > gpurun_out/multi_strong_n$N.json 2> gpurun_out/multi_strong_n$N.err; echo "strong rc=$?"
run --images 1250 --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/multi_weak_n$N.json 2> gpurun_out/multi_weak_n$N.err; echo "weak rc=$?"
python - <<PY
import json
for f in ["multi_strong_n$N", "multi_weak_n$N"]:
    try:
        line = [l for l in open(f"gpurun_out/{f}.json").read().splitlines() if l.startswith("{")][-1]
        l = json.loads(line)
        print(f, {k: l.get(k) for k in ["n_gpus", "scaling", "value", "ms_per_step", "detections", "detections_sha256", "gather_equal"]}, "e2e", (l.get("e2e") or {}).get("value"))
    except Exception as e:
        print(f, "unreadable", e)
PY
