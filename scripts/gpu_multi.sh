#!/usr/bin/env bash
# Multi-GPU validation (run under gpurun --gpus N): the 2-GPU equality test, then weak and strong scaling
# bench lines at N ranks with the in-job hardware-equality check (gather_equal) and the detection digest.
set -uo pipefail
N="${1:-2}"; IMG="${2:-1000}"; STEPS="${3:-2}"
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -4
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 "${@:2}"; }
python bench.py --gpus 1 --images $IMG --steps $STEPS --warmup 1 --no-extras --no-cpu-baseline --no-e2e > gpurun_out/multi_n1.json 2> gpurun_out/multi_n1.err; echo "n1 rc=$?"
run $N --images $IMG --steps $STEPS --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/multi_weak_n$N.json 2> gpurun_out/multi_weak_n$N.err; echo "weak rc=$?"; tail -2 gpurun_out/multi_weak_n$N.err
run $N --images $IMG --steps $STEPS --warmup 1 --no-extras --no-cpu-baseline --no-e2e --scaling strong > gpurun_out/multi_strong_n$N.json 2> gpurun_out/multi_strong_n$N.err; echo "strong rc=$?"; tail -2 gpurun_out/multi_strong_n$N.err
python - <<PY
import json
for f in ["multi_n1", "multi_weak_n$N", "multi_strong_n$N"]:
    try:
        l = json.load(open(f"gpurun_out/{f}.json"))
        print(f, {k: l.get(k) for k in ["n_gpus", "scaling", "value", "ms_per_step", "detections", "detections_sha256", "gather_equal"]}, "e2e", (l.get("e2e") or {}).get("value"))
    except Exception as e:
        print(f, "unreadable", e)
PY
