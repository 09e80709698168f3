"""The N>1 path on CPU: world_size-2 gloo processes shard a render by sample range and reduce the
films (yet-another-raytracer_b200/sharding.py).  The CPU oracle stands in for the GPU renderer --
what is under test is the partition, the collective and the exactness of the decomposition."""
import importlib
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_shard_range_partitions_exactly(yart):
    sh = importlib.import_module("yet-another-raytracer_b200.sharding")
    for begin, end, world in ((0, 1024, 8), (3, 10, 4), (0, 3, 8), (5, 5, 2), (0, 1000, 7)):
        pieces = [sh.shard_range(begin, end, r, world) for r in range(world)]
        assert pieces[0][0] == begin and pieces[-1][1] == end
        assert all(pieces[i][1] == pieces[i + 1][0] for i in range(world - 1))
        sizes = [b - a for a, b in pieces]
        assert max(sizes) - min(sizes) <= 1 and sum(sizes) == end - begin
    assert [sh.step_sample_range(2, r, 4, 8) for r in range(4)] == [(64, 72), (72, 80), (80, 88), (88, 96)]


def _worker(rank, world, port, out_path):
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    y = importlib.import_module("yet-another-raytracer_b200")
    sh = importlib.import_module("yet-another-raytracer_b200.sharding")
    from oracle import orc
    preset = y.ScenePreset("cornell-box")
    scene = orc.Scene(preset)
    w = h = 32
    cam = preset.camera(w, h)
    film = torch.zeros((h, w, 3), dtype=torch.float64)

    def render_fn(lo, hi, t):
        scene.render(cam, w, h, lo, hi, max_depth=50, seed=7, n_threads=1, film=t.numpy())

    lo, hi = sh.render_distributed(render_fn, film, 0, 7, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi))
    if rank == 0:
        np.save(out_path, film.numpy())
        assert gathered == [(0, 4), (4, 7)]
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_render_equals_single_process(yart, orc, tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "film.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    preset = yart.ScenePreset("cornell-box")
    want, _ = orc.Scene(preset).render(preset.camera(32, 32), 32, 32, 0, 7, max_depth=50, seed=7, n_threads=2)
    assert np.allclose(got, want, rtol=1e-13, atol=1e-13)  # equal up to the order of f64 additions
    assert got.sum() > 0
