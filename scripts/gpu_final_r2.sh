#!/usr/bin/env bash
# Final verification pass of round 2 (run under gpurun): the GPU test-suite, smoke(), the default bench line, and the cost of
# the two tile-path modes on one benchmark image.
set -uo pipefail
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_bench_final.err
python - <<'PY'
import json, time, argparse, numpy as np, torch
l = json.load(open('gpurun_out/r02_bench_final.json'))
print({k: l[k] for k in ['value', 'ms_per_step', 'detections', 'detections_sha256', 'gpu_launches', 'parity_sample']})
print('e2e', {k: l['e2e'][k] for k in ['value', 'ms_per_step', 'd2h_bytes_per_step']}, 'rows_only', l['e2e']['rows_only']['value'])
print('pack', [(s['masks'], round(s['pack_frac_of_hbm_peak'], 3)) for s in l['extras']['nms_sweep_480x640']])
from unmore_b200 import synth
from unmore_b200.object_reasoning import Object_Discovery, default_args
from unmore_b200.object_scoring import Object_Scoring
dev = torch.device('cuda:0')
img = synth.make_fields(0).to(dev); props = synth.make_proposals(0, 4096)
for name, od, sc in [("fused (antialias=False)", Object_Discovery(device=dev), Object_Scoring(device=dev)),
                     ("tile path (antialias=True)", Object_Discovery(default_args(antialias=True), device=dev), Object_Scoring(argparse.Namespace(antialias=True), device=dev))]:
    od.discover_image(img, props); torch.cuda.synchronize()
    t0 = time.perf_counter(); det = od.discover_image(img, props); anns = sc.score_image(img, det.astype(np.float64).tolist()) if len(det) else []; torch.cuda.synchronize()
    print(f"one bench image, 4096 proposals, {name}: {time.perf_counter() - t0:.3f} s, {len(det)} boxes, {len(anns)} annotations")
PY
