set -uo pipefail
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/driver_like_n2.json 2> gpurun_out/driver_like_n2.err; echo "rc=$?"; tail -3 gpurun_out/driver_like_n2.err
python - <<'PY'
import json
line=[l for l in open('gpurun_out/driver_like_n2.json').read().splitlines() if l.startswith('{')][-1]
l=json.loads(line)
print({k:l.get(k) for k in ['n_gpus','scaling','value','ms_per_step','detections','gather_equal']}, 'e2e', l['e2e']['value'], 'extras' , bool(l['extras']))
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
