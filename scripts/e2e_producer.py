#!/usr/bin/env python
"""configs[4] / configs[0] of BASELINE.json: a random-init DPT-L objectness net (PyTorch, the producer)
feeding the CUDA reasoning path, images sharded across the GPUs of one node.

    python scripts/e2e_producer.py --size 1024 1024 --images 16 --batch 2          # one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/e2e_producer.py ...

Per step and rank: `images` synthetic RGB images (torch.rand, seed = global image index) -> FieldProducer
(ObjectnessNet + dense Binary_Classifier) -> [B,4,H,W] field stacks -> anchors of
generate_random_proposal (object_reasoning.py:110-137) -> discovery + scoring kernels -> one all-gather
of detections.  Prints one JSON line with the producer and reasoning shares (device time, CUDA events,
max over ranks).  The nets are random-init with the heads' last layer rescaled so the fields are not
degenerate (FieldProducer.calibrate_random_init)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, nargs=2, default=[1024, 1024])
    ap.add_argument("--images", type=int, default=16, help="images per rank and step")
    ap.add_argument("--batch", type=int, default=2, help="images per producer forward")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"], help="autocast dtype of the producer")
    a = ap.parse_args()
    from unmore_b200 import ops
    from unmore_b200.object_reasoning import Object_Discovery
    from unmore_b200.pipeline import ReasoningPipeline
    from unmore_b200.producer import FieldProducer
    from unmore_b200.sharding import gather_detections, pack_detections

    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    H, W = a.size
    torch.manual_seed(0)   # identical weights on every rank
    prod = FieldProducer(autocast_dtype=torch.bfloat16 if a.dtype == "bf16" else None).to(dev).eval()
    g = torch.Generator(device=dev)

    def images_of(i0, n):
        out = torch.empty((n, 3, H, W), device=dev)
        for k in range(n):
            g.manual_seed(rank * a.images + i0 + k)
            out[k] = torch.rand((3, H, W), generator=g, device=dev)
        return out

    prod.calibrate_random_init(images_of(0, 1))
    pipe = ReasoningPipeline(dev, with_sat=False)
    anchors = Object_Discovery.generate_random_proposal(H, W)
    props = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(anchors, (a.images,) + anchors.shape))).to(dev)
    fields = torch.empty((a.images, 4, H, W), dtype=torch.float32, device=dev)
    ids = torch.arange(rank * a.images, (rank + 1) * a.images, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]

    def step():
        ev[0].record()
        for i0 in range(0, a.images, a.batch):
            n = min(a.batch, a.images - i0)
            prod(images_of(i0, n), out=fields[i0:i0 + n])
        ev[1].record()
        st = {}
        r = pipe.run_chunk(fields, props, stats=st)
        rows = gather_detections(pack_detections(ids, r["bbox"], r["keep_counts"], r["out"][:, :, 0].float()))
        ev[2].record()
        return rows, st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        step()
    barrier()
    l0 = ops.LAUNCHES
    tp = tr = 0.0
    t0 = time.perf_counter()
    for _ in range(a.steps):
        rows, st = step()
        torch.cuda.synchronize()
        tp += ev[0].elapsed_time(ev[1]); tr += ev[1].elapsed_time(ev[2])
    barrier()
    wall = time.perf_counter() - t0
    t = torch.tensor([tp, tr, wall * 1e3], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tp, tr, wall_ms = (float(x) / a.steps for x in t)
    if rank == 0:
        print(json.dumps({"metric": "e2e_producer_reasoning_images_per_sec", "value": world * a.images / (wall_ms / 1e3),
                          "unit": "images/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                          "config": {"workload": f"configs[4]: random-init DPT-L objectness net ({a.dtype}) -> CUDA reasoning, {H}x{W}",
                                     "images_per_rank": a.images, "proposals_per_image": int(anchors.shape[0]),
                                     "weights": "random-init, head output layers rescaled (calibrate_random_init)"},
                          "producer_ms_per_image": tp / a.images, "reasoning_ms_per_image": tr / a.images,
                          "detections": int(rows.shape[0]), "refine_in": int(st["refine_in"].sum()),
                          "proposal_rounds": int(st["refine_rounds"].sum()), "gpu_launches": ops.LAUNCHES - l0}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
