#!/usr/bin/env python
"""BASELINE.json configs[4]: VCM (light paths + vertex connection + hash-grid merging) on the caustic scene, 1920x1080, iterations
sharded over N GPUs (scene replicated, one NCCL reduce of the film).  Same launch contract as bench.py:

  python scripts/bench_vcm.py --gpus 1 --steps K --warmup W
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/bench_vcm.py --gpus N ...

A step = --iterations-per-step VCM iterations per GPU (weak scaling).  One JSON line on stdout (rank 0): value = sub-paths/s
(vmCount light + vmCount camera paths per iteration) over all ranks, CUDA events on the render stream, max over ranks.
"""
import argparse
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from bench import ClockSampler, log  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--iterations-per-step", type=int, default=2)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    from selas_b200 import capi, distributed, scenes

    rank, world, local_rank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        real_stdout = os.dup(1)  # NCCL prints its banner to fd 1: keep stdout to the one JSON line
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    w, h, K = 1920, 1080, args.iterations_per_step
    ctx = capi.Context(local_rank)
    data = scenes.caustic_scene()
    scene = capi.Scene(ctx, data)
    cam = scenes.caustic_camera(w, h)
    sphere = scenes.bounding_sphere(data)
    film = torch.zeros((h, w, 3), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream(dev)

    def settings(step):
        first, count = distributed.shard_iterations(K * world, rank, world)
        return capi.vcm_settings(w, h, count, sphere, first_iteration=step * K * world + first, seed=1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    reducer = distributed.FilmReducer(ctx, rank, world) if world > 1 else None
    for i in range(args.warmup):
        scene.render_vcm_device(cam, settings(i), film.data_ptr(), stream.cuda_stream)
    if reducer is not None:
        reducer.reduce(film, dst=0, stream=stream.cuda_stream)
    film.zero_()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    agg = {k: 0.0 for k in ("lightPassMs", "gridBuildMs", "cameraPassMs", "mergeMs", "connectMs", "closestRays", "occlusionRays",
                            "lightVertices", "mergeCandidates", "mergesAccepted", "connections", "kernelLaunches")}
    e0.record(stream)
    for i in range(args.steps):
        st = scene.render_vcm_device(cam, settings(args.warmup + i), film.data_ptr(), stream.cuda_stream)
        for k in agg:
            agg[k] += getattr(st, k)
    if world > 1:
        reducer.reduce(film, dst=0, stream=stream.cuda_stream)
    e1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    iterations = K * args.steps * world
    paths = 2 * w * h * iterations
    if rank == 0:
        total_iters_rank = K * args.steps
        finite = torch.isfinite(film)
        doc = {"metric": "vcm_sub_paths_per_s_1080p", "value": paths / (total_ms * 1e-3), "unit": "paths/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "ms_per_iteration_per_gpu": total_ms / total_iters_rank,
               "iterations_per_s": iterations / (total_ms * 1e-3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic",
               "config": {"workload": "BASELINE configs[4]: VCM, caustic scene (floor + glass sphere + box, sun-and-sky IBL), 1920x1080, "
                                      "vmCount = 2073600 light paths per iteration, path length <= 10",
                          "triangles": data.triangle_count(), "iterations_per_step_per_gpu": K,
                          "sharding": f"iterations over {world} GPU(s), scene replicated, one NCCL reduce of the film"},
               "clocks": clocks, "gpu_launches": int(agg["kernelLaunches"]),
               "stage_ms_per_iteration_rank0": {k: agg[k] / total_iters_rank for k in ("lightPassMs", "gridBuildMs", "cameraPassMs", "mergeMs", "connectMs")},
               "per_iteration_rank0": {k: agg[k] / total_iters_rank for k in ("lightVertices", "closestRays", "occlusionRays", "mergeCandidates",
                                                                            "mergesAccepted", "connections")},
               "mean_radiance": float(film[finite].mean().item() / iterations), "nonfinite_film_values": int((~finite).sum().item())}
        line = json.dumps(doc)
        if world > 1:
            os.write(real_stdout, (line + "\n").encode())
        else:
            print(line, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    scene.close()
    ctx.close()


if __name__ == "__main__":
    main()
