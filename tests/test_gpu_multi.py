"""Multi-GPU equality on hardware: the gathered N-GPU result must equal the 1-GPU result, exactly
(the reference shards the same way by hand with --start_idx/--end_idx, datasets.py:432-435).
Skips when fewer than two devices are visible."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from unmore_b200 import ops, synth
from unmore_b200.pipeline import ReasoningPipeline
from unmore_b200.sharding import gather_rows, merge_rows, overflowed, rows_digest, shard_indices
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
n_images, n_prop, max_rows = 10, 384, 512
def run(ids, device):
    pipe = ReasoningPipeline(device)
    f = torch.stack([synth.make_fields(i) for i in ids]).to(device)
    p = torch.tensor(np.stack([synth.make_proposals(i, n_prop) for i in ids]), device=device)
    r = pipe.run_chunk(f, p)
    return ops.pack_detections(torch.tensor(ids, dtype=torch.int64, device=device), r["bbox"], r["out"], r["keep_counts"],
                               ops.detection_rows(max_rows, device))
for interleave in (True, False):
    mine = shard_indices(n_images, rank, world, interleave=interleave)
    g = gather_rows(run(mine, dev))
    assert not bool(overflowed(g))
    merged, total = merge_rows(g)
    got = merged[: int(total)]
    single = run(list(range(n_images)), dev)              # the same images in ONE process on this GPU
    ref = single[1:1 + int(single[0, 0])]
    assert got.shape == ref.shape and torch.equal(got, ref), (rank, interleave, got.shape, ref.shape)
    digests = [None] * world
    dist.all_gather_object(digests, rows_digest(got))
    assert len(set(digests)) == 1
# tensors on a device that is NOT current must still run on their own device (ADVICE r1: device guard)
other = torch.device("cuda", (rank + 1) % world)
with torch.cuda.device(dev):
    rows_other = run([3], other)
    rows_here = run([3], dev)
assert torch.equal(rows_other.cpu(), rows_here.cpu())
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_gather_equals_single_gpu(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


def test_device_guard_single_gpu():
    """The C ABI refuses a field tensor that lives on another device than the current one instead of launching
    on it (capi.cu check_fields); with one GPU this only checks the happy path through the Python guard."""
    from unmore_b200 import ops, synth
    dev = torch.device("cuda:0")
    f = synth.make_fields(1).to(dev)[None].contiguous()
    b = torch.tensor(synth.make_proposals(1, 16), device=dev)[None].contiguous()
    s = ops.existence_scores(f, b)
    assert s.shape == (1, 16) and bool(((s >= 0) & (s <= 1)).all())
