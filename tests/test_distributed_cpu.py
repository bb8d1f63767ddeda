"""N > 1 host logic on CPU: two gloo ranks each render their shard of the sample passes (through the oracle, since there is
no GPU here), the films are merged with the path's single reduce, and rank 0 must hold the same image as one process
rendering all samples.  On the GPU box the same sharding + reduce runs over NCCL in bench.py."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total_spp, out_path):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist

    from selas_b200 import capi, distributed
    from tests import inputs, oracle_binding

    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = oracle_binding.load_oracle()
    data = inputs.small_scene()
    cam = orc.camera_init((0.0, 1.0, -5.0), (0.3, 0.0, 0.0), (0.0, 1.0, 0.0), 45.0, 32, 24)
    hs = orc.scene_create(data)
    first, count = distributed.shard_samples(total_spp, rank, world)
    rs = capi.render_settings(32, 24, count, max_path_length=5, rr_start=8, layers=1, seed=3, first_sample=first)
    film = orc.render(hs, cam, rs, nthreads=1)[0][0]
    t = torch.from_numpy(film.copy())
    distributed.reduce_film(t, dst=0)
    if rank == 0:
        np.save(out_path, t.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sample_sharding_and_reduce(tmp_path, oracle):
    import torch.multiprocessing as mp

    from selas_b200 import capi
    from tests import inputs

    total_spp = 5  # odd: the shards are uneven (3 + 2)
    out = tmp_path / "film.npy"
    mp.spawn(_worker, args=(2, _free_port(), total_spp, str(out)), nprocs=2, join=True)
    merged = np.load(out)
    data = inputs.small_scene()
    cam = oracle.camera_init((0.0, 1.0, -5.0), (0.3, 0.0, 0.0), (0.0, 1.0, 0.0), 45.0, 32, 24)
    hs = oracle.scene_create(data)
    full = oracle.render(hs, cam, capi.render_settings(32, 24, total_spp, 5, 8, 1, seed=3), nthreads=1)[0][0]
    oracle.scene_destroy(hs)
    assert full.sum() > 0
    np.testing.assert_allclose(merged, full, rtol=1e-5, atol=1e-6)  # fp32 sum order differs across shards


def _vcm_worker(rank, world, port, total_iterations, out_path):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist

    from selas_b200 import capi, distributed, scenes
    from tests import oracle_binding

    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = oracle_binding.load_oracle()
    data = scenes.caustic_scene(sphere_res=(12, 24), floor_cells=4, ibl_size=(64, 32))
    cam = orc.camera_init((0.0, 4.5, -9.0), (0.4, 0.6, 0.0), (0.0, 1.0, 0.0), 45.0, 48, 27)
    hs = orc.scene_create(data)
    first, count = distributed.shard_iterations(total_iterations, rank, world)
    vs = capi.vcm_settings(48, 27, count, scenes.bounding_sphere(data), first_iteration=first, seed=6)
    film = orc.render_vcm(hs, cam, vs, nthreads=1)[0]
    t = torch.from_numpy(film.copy())
    distributed.reduce_film(t, dst=0)
    if rank == 0:
        np.save(out_path, t.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_vcm_iteration_sharding(tmp_path, oracle):
    """VCM shards whole iterations: ranks render passes [0,2) and [2,3); the reduced film equals passes [0,3) in one process."""
    import torch.multiprocessing as mp

    from selas_b200 import capi, scenes

    total = 3
    out = tmp_path / "vcm.npy"
    mp.spawn(_vcm_worker, args=(2, _free_port(), total, str(out)), nprocs=2, join=True)
    merged = np.load(out)
    data = scenes.caustic_scene(sphere_res=(12, 24), floor_cells=4, ibl_size=(64, 32))
    cam = oracle.camera_init((0.0, 4.5, -9.0), (0.4, 0.6, 0.0), (0.0, 1.0, 0.0), 45.0, 48, 27)
    hs = oracle.scene_create(data)
    full = oracle.render_vcm(hs, cam, capi.vcm_settings(48, 27, total, scenes.bounding_sphere(data), seed=6), nthreads=1)[0]
    oracle.scene_destroy(hs)
    assert full.sum() > 0 and np.isfinite(full).all()
    np.testing.assert_allclose(merged, full, rtol=1e-4, atol=1e-5)
