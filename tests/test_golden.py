"""Golden vectors (tests/golden/, made by tools/gen_golden.py from the oracle): the CPU suite pins
the oracle to them, the GPU suite checks the CUDA path against the committed files."""
from pathlib import Path

import numpy as np
import pytest

GOLDEN = Path(__file__).resolve().parent / "golden"
INF = float("inf")
FIELDS = ("t", "u", "v", "prim_id", "obj_id", "front_face")


@pytest.mark.parametrize("name", ["cube", "sycee", "david"])
def test_oracle_reproduces_closest_hit_golden(orc, mesh_scene, name):
    g = np.load(GOLDEN / ("closest_hit_%s.npz" % name))
    _, _, s = mesh_scene(name)
    for order in (0, 1):
        hits, _ = s.closest_hit(g["rays"], 0, 0.001, INF, order)
        for f in FIELDS:
            assert np.array_equal(hits[f], g["hits"][f]), f
    bf, ties = s.brute_force_hit(g["rays"], 0, 0.001)
    assert np.array_equal(bf["t"], g["brute_t"]) and np.array_equal(ties, g["ties"])


@pytest.mark.parametrize("scene", ["david", "cornell-box", "next-week-final"])
def test_oracle_reproduces_scene_golden(yart, orc, scene):
    g = np.load(GOLDEN / ("scene_%s.npz" % scene))
    w, h, spp, depth, seed, scene_seed = [int(x) for x in g["size"]]
    p = yart.ScenePreset(scene, seed=scene_seed)
    s = orc.Scene(p)
    cam = p.camera(w, h)
    rays, wl, tm = orc.camera_rays(cam, w, h, 0, 1, seed=seed)
    assert np.array_equal(rays["origin"], g["rays"]["origin"]) and np.array_equal(rays["direction"], g["rays"]["direction"])
    assert np.array_equal(wl, g["wavelength"]) and np.array_equal(tm, g["time"])
    hits, _ = s.closest_hit(rays, orc.abi.TARGET_WORLD, 0.001, INF, 0)
    for f in FIELDS:
        assert np.array_equal(hits[f], g["hits"][f]), f
    film, st = s.render(cam, w, h, 0, spp, max_depth=depth, seed=seed, n_threads=3)
    assert np.array_equal(film, g["film"]) and st.rays == int(g["rays_traced"][0])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cube", "sycee", "david"])
def test_gpu_matches_closest_hit_golden(yart, ctx, mesh_scene, name):
    g = np.load(GOLDEN / ("closest_hit_%s.npz" % name))
    _, ms, _ = mesh_scene(name)
    ctx.set_scene(ms.desc)
    for order in (yart.ORDER_REFERENCE, yart.ORDER_NEAR):
        hits, _ = ctx.closest_hit(g["rays"], 0, 0.001, INF, order)
        for f in FIELDS:
            assert np.array_equal(hits[f], g["hits"][f]), f


@pytest.mark.gpu
@pytest.mark.parametrize("scene", ["david", "cornell-box", "next-week-final"])
def test_gpu_matches_scene_golden(yart, ctx, scene):
    g = np.load(GOLDEN / ("scene_%s.npz" % scene))
    w, h, spp, depth, seed, scene_seed = [int(x) for x in g["size"]]
    p = yart.ScenePreset(scene, seed=scene_seed)
    ctx.set_scene(p)
    cam = p.camera(w, h)
    rays, wl, tm = ctx.camera_rays(cam, w, h, 0, 1, seed=seed)
    assert np.array_equal(rays["origin"], g["rays"]["origin"]) and np.array_equal(rays["direction"], g["rays"]["direction"])
    assert np.array_equal(wl, g["wavelength"]) and np.array_equal(tm, g["time"])
    hits, _ = ctx.closest_hit(rays, yart.TARGET_WORLD, 0.001, INF, yart.ORDER_NEAR)
    med = np.array([bool(p.desc.contents.objects[int(i)].wrap & 8) if i != yart.MISS else False for i in g["hits"]["obj_id"]])
    for f in FIELDS:
        assert np.array_equal(hits[f][~med], g["hits"][f][~med]), f
    assert np.allclose(hits["t"][med], g["hits"]["t"][med], rtol=1e-12, atol=0)
    film, st = ctx.render(cam, w, h, 0, spp, max_depth=depth, seed=seed)
    rel = np.sqrt(((film - g["film"]) ** 2).sum() / (g["film"] ** 2).sum())
    assert rel <= 0.01  # the stated image tolerance; in practice ~1e-15 when no path flips a branch
    assert abs(int(st.rays) - int(g["rays_traced"][0])) <= 8
