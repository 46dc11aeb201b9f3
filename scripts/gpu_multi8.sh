#!/usr/bin/env bash
# N-GPU runs of the north-star target: 5000 images over N GPUs (strong), weak scaling beside it, and configs[4].
set -uo pipefail
N="${1:-8}"
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
nvidia-smi -L | wc -l
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N "$@"; }
run --images 5000 --steps 3 --warmup 2 --no-extras --no-cpu-baseline --no-e2e --scaling strong > gpurun_out/multi_strong_n$N.json 2> gpurun_out/multi_strong_n$N.err; echo "strong rc=$?"
run --images 1250 --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/multi_weak_n$N.json 2> gpurun_out/multi_weak_n$N.err; echo "weak rc=$?"
run --config producer --images 8 --steps 2 --warmup 1 > gpurun_out/multi_producer_n$N.json 2> gpurun_out/multi_producer_n$N.err; echo "producer rc=$?"; tail -3 gpurun_out/multi_producer_n$N.err
python - <<PY
import json
for f in ["multi_strong_n$N", "multi_weak_n$N", "multi_producer_n$N"]:
    try:
        line = [l for l in open(f"gpurun_out/{f}.json").read().splitlines() if l.startswith("{")][-1]
        l = json.loads(line)
        print(f, {k: l.get(k) for k in ["n_gpus", "scaling", "value", "ms_per_step", "detections", "detections_sha256", "gather_equal"]}, "e2e", (l.get("e2e") or {}).get("value"))
        print("   stages", {k: round(v, 1) for k, v in l["stage_ms_per_step"].items() if v > 1})
    except Exception as e:
        print(f, "unreadable", e)
PY
