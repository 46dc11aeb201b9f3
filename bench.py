#!/usr/bin/env python
"""Benchmark of the multi-object reasoning hot path (BASELINE.json configs[1]):
a COCO-val-shaped synthetic batch — 5000 images of 480x640 fields, 4096 proposals per image —
through discovery (existence check, center reasoning, iterative boundary refinement, NMS),
scoring + mask rasterisation and the post-process predicate.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference ...                      # the CPU oracle port, bounded sample

One "step" = one pass over the whole per-rank batch.  Prints ONE JSON line (rank 0).
Scaling is weak: every rank owns a 5000-image batch (seeds rank*images + i); the only
collective is the all-gather of detections at the end of a step (inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, N_PROP = 480, 640, 4096
WORKLOAD = "configs[1]: 5000 synthetic 480x640 field stacks x 4096 proposals/image, discovery + scoring"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images", type=int, default=5000, help="images per rank and step")
    ap.add_argument("--proposals", type=int, default=N_PROP)
    ap.add_argument("--chunk", type=int, default=250, help="images per launch group")
    ap.add_argument("--cpu-sample-proposals", type=int, default=1536,
                    help="proposals of image 0 the CPU baseline / reference arm processes per sample (~7 s on 16 cores)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
def cpu_sample(n_sample: int, n_prop: int, repeats: int = 1):
    """The oracle port (a bit-identical, faster restatement of the reference's CPU path, see
    oracle/oracle.py) on image 0 with a strided subset of its proposals; scaled to images/s."""
    import numpy as np
    import torch
    from oracle import oracle as O
    from unmore_b200 import synth

    torch.set_num_threads(os.cpu_count() or 1)
    img = synth.make_fields(0, H, W)
    props = synth.make_proposals(0, n_prop, H, W)
    sel = np.unique(np.linspace(0, n_prop - 1, n_sample).round().astype(np.int64))
    args = O.make_args()
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        det = O.discover_image(img, props[sel], args)
        if len(det):
            O.score_image(img, det.tolist(), args)
        times.append(time.perf_counter() - t0)
    t = statistics.median(times)
    scale = n_prop / len(sel)
    return {"value": 1.0 / (t * scale), "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"image 0, {len(sel)} of {n_prop} proposals (strided), discovery+scoring in {t:.1f}s, "
                      f"scaled x{scale:.1f} to a full image; antialias=False"}, t


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    for _ in range(min(args.warmup, 1)):
        cpu_sample(args.cpu_sample_proposals, args.proposals)
    vals, ts = [], []
    for _ in range(args.steps):
        cb, t = cpu_sample(args.cpu_sample_proposals, args.proposals)
        vals.append(cb["value"]); ts.append(t)
    v = statistics.median(vals)
    cb["value"] = v
    line = {"impl": "reference", "metric": "object_reasoning_images_per_sec", "value": v, "unit": "images/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * statistics.median(ts), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "images_per_rank": args.images, "proposals_per_image": args.proposals,
                       "field_hw": [H, W]},
            "proposals_per_sec": v * args.proposals, "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from unmore_b200 import ops, synth
    from unmore_b200.pipeline import ReasoningPipeline
    from unmore_b200.sharding import gather_detections, pack_detections

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    n_img, n_prop, chunk = args.images, args.proposals, min(args.chunk, args.images)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")

    # ---- synthetic inputs, generated on the device (seed = global image index)
    t_gen = time.time()
    base = rank * n_img
    fields = torch.empty((n_img, 4, H, W), dtype=torch.float32, device=dev)
    for i in range(n_img):
        fields[i] = synth.render_fields(synth.scene_params(base + i, H, W), H, W, device=dev)
    anchors = synth.anchor_proposals(H, W)
    props_np = np.empty((n_img, n_prop, 4), dtype=np.float64)
    for i in range(n_img):
        if n_prop >= len(anchors):
            props_np[i, : len(anchors)] = anchors
            props_np[i, len(anchors):] = synth.random_proposals(base + i, n_prop - len(anchors), H, W)
        else:
            props_np[i] = synth.make_proposals(base + i, n_prop, H, W)
    proposals = torch.from_numpy(props_np).to(dev)
    image_ids = torch.arange(base, base + n_img, device=dev)
    torch.cuda.synchronize()
    t_gen = time.time() - t_gen

    pipe = ReasoningPipeline(dev)
    sat_buf = torch.empty((n_img, 2, H + 1, W + 1), dtype=torch.float64, device=dev)

    def step(collect_stats=None):
        rows = []
        # north-star op (a): summed-area tables of the existence / boundary-distance fields for the
        # whole batch in one HBM-streaming launch; the chunks below read their slices
        sat = pipe.build_sat(fields, out=sat_buf)
        for c0 in range(0, n_img, chunk):
            c1 = min(c0 + chunk, n_img)
            st = {} if collect_stats is not None else None
            r = pipe.run_chunk(fields[c0:c1], proposals[c0:c1], stats=st, sat=sat[c0:c1])
            rows.append(pack_detections(image_ids[c0:c1], r["bbox"], r["keep_counts"], r["out"][:, :, 0].float()))
            if collect_stats is not None:
                collect_stats["proposal_rounds"] = collect_stats.get("proposal_rounds", 0) + int(st["refine_rounds"].sum())
                collect_stats["refine_in"] = collect_stats.get("refine_in", 0) + int(st["refine_in"].sum())
                collect_stats["center_in"] = collect_stats.get("center_in", 0) + int(st["pass1"].sum()) + int(st["split_kept"].sum())
                collect_stats["exist_in"] = collect_stats.get("exist_in", 0) + (c1 - c0) * n_prop + int(st["split"].sum())
                collect_stats["detections"] = collect_stats.get("detections", 0) + int(r["keep_counts"].sum())
        local = torch.cat(rows, dim=0)
        return gather_detections(local)  # the single collective of the path (no-op at world 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 0)):
        step()
    work = {}
    step(work)  # untimed: work counters for the roofline arithmetic (device syncs inside)

    # ---- timed region: K steps, device time, max over ranks
    sampler = ClockSampler(local_rank)
    timer = ops.StageTimer()
    barrier()
    sampler.start()
    launches0 = ops.LAUNCHES
    ops.set_timer(timer)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof_range = os.environ.get("UNMORE_PROFILE_RANGE") == "1"   # ncu --profile-from-start off: timed region only
    if prof_range:
        torch.cuda.profiler.start()
    ev0.record()
    for _ in range(args.steps):
        out = step()
    ev1.record()
    barrier()
    if prof_range:
        torch.cuda.profiler.stop()
    ops.set_timer(None)
    clocks = sampler.stop()
    launches = ops.LAUNCHES - launches0
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    images_per_s = world * n_img / (ms_per_step / 1e3)
    stages = timer.summary()

    # ---- roofline of every kernel from the live CUDA-event times (per launch averages)
    n_chunks = (n_img + chunk - 1) // chunk
    per_step = {k: v["ms"] / args.steps for k, v in stages.items()}
    px = H * W
    alg = {  # algorithmic DRAM bytes per STEP: fields read once per image + boxes in/out
        "unmore_boundary_refine": n_img * px * 4 + work["refine_in"] * (32 + 16 + 4 + 4),
        "unmore_center_reasoning": n_img * px * 12 + work["center_in"] * (32 + 8 + 4 + 128),
        "unmore_existence_scores": n_img * px * 4 + work["exist_in"] * (32 + 4),
        "unmore_sat_build_fields": n_img * 2 * (px * 4 + (H + 1) * (W + 1) * 8),
        "unmore_score_and_rasterise": n_img * px * 16 + work["detections"] * (H * ((W + 31) // 32) * 4),
    }
    kernels = {}
    for name, b in alg.items():
        if name in per_step and per_step[name] > 0:
            ach = b / (per_step[name] / 1e3) / 1e9
            kernels[name] = {"ms_per_step": per_step[name], "share": per_step[name] / ms_per_step,
                             "ms_per_launch": stages[name]["ms"] / stages[name]["calls"], "achieved_gbs": ach,
                             "frac": ach / hbm_peak}
    # ncu DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum of one --set full / --metrics
    # capture of a launch of exactly this shape), committed under profiles/; None when the shape differs
    traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("images_per_launch") == chunk and tj.get("proposals_per_image") == n_prop and tj.get("field_hw") == [H, W]:
            traffic = tj.get("dram_bytes_per_launch", {})
    except Exception:
        pass
    for name, k in kernels.items():
        k["traffic"] = traffic.get(name)
    dominant = max(per_step, key=per_step.get)
    dk = kernels.get(dominant, {"achieved_gbs": 0.0, "frac": 0.0})
    roofline = {"kernel": dominant, "bound": "hbm", "achieved": dk["achieved_gbs"], "peak": hbm_peak, "unit": "GB/s",
                "frac": dk["frac"], "traffic": traffic.get(dominant), "peak_source": peak_src,
                "note": "the per-proposal kernels re-read L2-resident fields and are bound by fp32/MUFU issue, not DRAM "
                        "(SURVEY.md §8d); only unmore_sat_build streams from HBM — see `kernels`",
                "proposal_rounds_per_step": work["proposal_rounds"],
                "proposal_rounds_per_sec": work["proposal_rounds"] / (per_step.get("unmore_boundary_refine", 1e9) / 1e3)}

    # ---- end to end through the public API with HOST buffers (pinned), copies inside the timed region
    e2e = None
    if not args.no_e2e:
        import psutil
        need = n_img * 4 * px * 4
        pool = n_img
        avail = psutil.virtual_memory().available
        while pool > chunk and pool * 4 * px * 4 * world > avail // 4:   # every rank pins its own pool
            pool //= 2
        pool = max(chunk, (pool // chunk) * chunk)
        h_fields = torch.empty((pool, 4, H, W), dtype=torch.float32).pin_memory()
        for c0 in range(0, pool, chunk):
            h_fields[c0:c0 + chunk].copy_(fields[c0:c0 + chunk])
        h_props = torch.from_numpy(props_np).pin_memory()
        torch.cuda.synchronize()
        copy_stream = torch.cuda.Stream()
        bufs = [(torch.empty((chunk, 4, H, W), dtype=torch.float32, device=dev),
                 torch.empty((chunk, n_prop, 4), dtype=torch.float64, device=dev)) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]
        h2d = d2h = 0

        def e2e_step():
            nonlocal h2d, d2h
            h2d = d2h = 0
            host_rows = []
            chunks = list(range(0, n_img, chunk))

            def issue(j):
                c0 = chunks[j]
                c1 = min(c0 + chunk, n_img)
                fb, pb = bufs[j % 2]
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(freed[j % 2])
                    p0 = c0 % pool
                    fb[: c1 - c0].copy_(h_fields[p0:p0 + (c1 - c0)], non_blocking=True)
                    pb[: c1 - c0].copy_(h_props[c0:c1], non_blocking=True)
                    ready[j % 2].record(copy_stream)
                return (c1 - c0) * (4 * px * 4 + n_prop * 32)

            h2d += issue(0)
            for j, c0 in enumerate(chunks):
                c1 = min(c0 + chunk, n_img)
                if j + 1 < len(chunks):
                    h2d += issue(j + 1)
                torch.cuda.current_stream().wait_event(ready[j % 2])
                fb, pb = bufs[j % 2]
                r = pipe.run_chunk(fb[: c1 - c0], pb[: c1 - c0])
                rows = pack_detections(image_ids[c0:c1], r["bbox"], r["keep_counts"], r["out"][:, :, 0].float())
                freed[j % 2].record(torch.cuda.current_stream())
                hr = rows.cpu()  # device -> host read of the chunk's result
                d2h += hr.numel() * 4
                host_rows.append(hr)
            return torch.cat(host_rows)

        freed[0].record(); freed[1].record()
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(max(1, min(args.steps, 2))):
            e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / max(1, min(args.steps, 2))
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * n_img / dt, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "host_pool_images": pool, "ms_per_step": dt * 1e3}

    # ---- configs[2] side measurement (outside the step): mask bit-pack (HBM streaming) and the
    # mask-IoU / box NMS sweep at 1k - 16k masks per image, 480x640
    extras = {}
    if rank == 0 and not args.no_extras:
        del sat_buf
        torch.cuda.empty_cache()

        def timed(fn, reps=5):
            fn(); torch.cuda.synchronize()
            best = 1e30
            for _ in range(reps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record(); torch.cuda.synchronize()
                best = min(best, a.elapsed_time(b))
            return best

        g = torch.Generator(device=dev).manual_seed(0)
        ys = torch.arange(H, device=dev).view(1, H, 1); xs = torch.arange(W, device=dev).view(1, 1, W)
        sweep = []
        for K in (1024, 4096, 16384):
            dense = torch.empty((K, H, W), dtype=torch.uint8, device=dev)
            cx = torch.rand(K, generator=g, device=dev) * W; cy = torch.rand(K, generator=g, device=dev) * H
            rx = 20 + torch.rand(K, generator=g, device=dev) * 120; ry = 20 + torch.rand(K, generator=g, device=dev) * 120
            for k0 in range(0, K, 256):
                sl = slice(k0, k0 + 256)
                dense[sl] = ((((ys - cy[sl].view(-1, 1, 1)) / ry[sl].view(-1, 1, 1)) ** 2 +
                              ((xs - cx[sl].view(-1, 1, 1)) / rx[sl].view(-1, 1, 1)) ** 2) < 1).to(torch.uint8)
            msc = torch.rand(K, generator=g, device=dev)
            t_pack = timed(lambda: ops.mask_pack(dense))
            packed = ops.mask_pack(dense)
            del dense
            pack_bytes = K * H * W + K * H * ((W + 31) // 32) * 4
            stats_ = ops.mask_stats(packed, W)
            t_nms = timed(lambda: ops.mask_nms(packed, W, msc, 0.5, stats=stats_), reps=3)
            kept = int(ops.mask_nms(packed, W, msc, 0.5, stats=stats_).numel())
            boxes = stats_[1].to(torch.float32)
            t_box = timed(lambda: ops.box_nms_matrix(boxes, msc, 0.5), reps=3)
            sweep.append({"masks": K, "pack_ms": t_pack, "pack_gbs": pack_bytes / t_pack / 1e6,
                          "pack_frac_of_hbm_peak": pack_bytes / t_pack / 1e6 / hbm_peak,
                          "mask_nms_ms": t_nms, "mask_nms_kept": kept, "box_nms_matrix_ms": t_box})
            del packed
        # CPU side of the sweep: the dense-mask greedy restatement (oracle, north-star op C) on 256 masks
        cpu_nms = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle as O
            Kc = 256
            dense = torch.empty((Kc, H, W), dtype=torch.uint8, device=dev)
            cx = torch.rand(Kc, generator=g, device=dev) * W; cy = torch.rand(Kc, generator=g, device=dev) * H
            rx = 20 + torch.rand(Kc, generator=g, device=dev) * 120; ry = 20 + torch.rand(Kc, generator=g, device=dev) * 120
            dense[:] = ((((ys - cy.view(-1, 1, 1)) / ry.view(-1, 1, 1)) ** 2 + ((xs - cx.view(-1, 1, 1)) / rx.view(-1, 1, 1)) ** 2) < 1).to(torch.uint8)
            msc = torch.rand(Kc, generator=g, device=dev)
            packed = ops.mask_pack(dense)
            t_gpu = timed(lambda: ops.mask_nms(packed, W, msc, 0.5), reps=3)
            keep_gpu = ops.mask_nms(packed, W, msc, 0.5).cpu().numpy()
            dn, sn = dense.cpu().numpy(), msc.cpu().numpy()
            t0 = time.perf_counter()
            keep_cpu = O.mask_nms_dense(dn, sn, 0.5)
            t_cpu = (time.perf_counter() - t0) * 1e3
            cpu_nms = {"masks": Kc, "cpu_dense_ms": t_cpu, "gpu_ms_incl_stats": t_gpu, "keep_sets_equal": bool(np.array_equal(keep_gpu, keep_cpu)),
                       "kind": "port (numpy, 1 core)"}
            del dense, packed
        extras = {"cpu_mask_nms": cpu_nms, "nms_sweep_480x640": sweep,
                  "note": "configs[2]; mask_nms = rank sort + 64-wide IoU bit-matrix (popc on packed masks) + greedy scan"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline, _ = cpu_sample(args.cpu_sample_proposals, n_prop)

    if rank == 0:
        line = {"metric": "object_reasoning_images_per_sec", "value": images_per_s, "unit": "images/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "images_per_rank": n_img, "proposals_per_image": n_prop,
                           "field_hw": [H, W], "chunk_images": chunk, "n_round": 50, "resize": "bilinear, antialias=False",
                           "l2": f"inputs ({n_img * 4 * px * 4 / 1e9:.1f} GB of fields per rank) exceed the 126 MB L2; no flush needed",
                           "input_generation_s": round(t_gen, 1)},
                "proposals_per_sec": images_per_s * n_prop, "detections": int(out.shape[0]),
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "kernels": kernels,
                "stage_ms_per_step": per_step, "e2e": e2e, "cpu_baseline": cpu_baseline, "extras": extras}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
