#!/usr/bin/env python3
"""The other BASELINE.json configs once each (GPU numbers, optional bounded CPU-oracle numbers):
C1 cornell 400x400x32, C2 bunny preset (sycee.obj stand-in) 1280x720x256, C4 next-week-final 1920x1080
(reduced spp), C5 sweeps on david and sycee (uniform / axis / path rays)."""
import importlib
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import raysets  # noqa: E402
from oracle import orc  # noqa: E402

y = importlib.import_module("yet-another-raytracer_b200")
ctx = y.Context(0)
cpu = "--cpu" in sys.argv
cores = os.cpu_count()


def render_cfg(tag, scene, w, h, spp, cpu_spp):
    p = y.ScenePreset(scene, seed=1)
    ctx.set_scene(p)
    cam = p.camera(w, h)
    ctx.render(cam, w, h, 0, min(spp, 4), 50, 1)  # warm
    film, st = ctx.render(cam, w, h, 0, spp, 50, 1)
    line = "%s %s %dx%d %d spp: %.1f Mrays/s, %.2f Msamples/s, %.2f rays/sample, %.1f ms" % (
        tag, scene, w, h, spp, st.rays / st.gpu_ms / 1e3, st.paths / st.gpu_ms / 1e3, st.rays / st.paths, st.gpu_ms)
    if cpu:
        s = orc.Scene(p)
        t0 = time.perf_counter()
        ofilm, ost = s.render(cam, w, h, 0, cpu_spp, 50, 1, 0, cores)
        dt = time.perf_counter() - t0
        gfilm, _ = ctx.render(cam, w, h, 0, cpu_spp, 50, 1)
        rel = float(np.sqrt(((gfilm - ofilm) ** 2).sum() / (ofilm ** 2).sum()))
        line += " | CPU oracle %d cores, %d spp: %.2f Mrays/s (%.1f s); film relRMSE GPU vs oracle %.2e" % (
            cores, cpu_spp, ost.rays / dt / 1e6, dt, rel)
    print(line, flush=True)


def sweep_cfg(mesh, n=1 << 24):
    m = y.TriangleMesh.from_obj(y.assets_dir() + "/%s.obj" % mesh)
    ms = orc.MeshScene(m.positions(), m.normals(), m.uvs())
    q = y.L4QBVH.from_mesh(m)
    ctx.set_scene(ms.desc)
    for rs, gen in (("uniform", raysets.uniform), ("axis", raysets.axis)):
        o, d = gen(n, q.info.bbox_min, q.info.bbox_max)
        rays = orc.abi.make_rays(o, d)
        for order, oname in ((y.ORDER_REFERENCE, "reference"), (y.ORDER_NEAR, "near")):
            best = min(ctx.closest_hit(rays, 0, 0.0, float("inf"), order)[1].gpu_ms for _ in range(3))
            _, st = ctx.closest_hit(rays, 0, 0.0, float("inf"), order, count_visits=True)
            nb = (128 * st.node_visits + 48 * st.tri_tests) / n + 88
            print("C5 %s %s %s: %.1f Mrays/s (%.2f ms), %.2f nodes %.2f tris per ray, %.0f B/ray -> %.0f GB/s" % (
                mesh, rs, oname, n / best / 1e3, best, st.node_visits / n, st.tri_tests / n, nb, nb * n / best / 1e6), flush=True)
        if cpu and rs == "uniform":
            s = orc.Scene(ms)
            sub = rays[:1 << 21]
            t0 = time.perf_counter()
            s.closest_hit(sub, 0, 0.0, float("inf"), 0, n_threads=cores)
            dt = time.perf_counter() - t0
            print("C5 %s uniform CPU oracle (%d cores, reference order, %d rays): %.2f Mrays/s" % (mesh, cores, len(sub), len(sub) / dt / 1e6), flush=True)


def path_ray_sweep(n=1 << 22):
    """the ray distribution the renderer really sees: world rays of every bounce of the david camera"""
    p = y.ScenePreset("david")
    s = orc.Scene(p)
    cam = p.camera(480, 270)
    rays = s.dump_path_rays(cam, 480, 270, 0, 16, n)
    ctx.set_scene(p)
    for order, oname in ((y.ORDER_REFERENCE, "reference"), (y.ORDER_NEAR, "near")):
        best = min(ctx.closest_hit(rays, y.TARGET_WORLD, 0.001, float("inf"), order)[1].gpu_ms for _ in range(3))
        print("C5 david path rays (world.hit, %d rays) %s: %.1f Mrays/s" % (len(rays), oname, len(rays) / best / 1e3), flush=True)


render_cfg("C1", "cornell-box", 400, 400, 32, 32)
render_cfg("C2", "bunny", 1280, 720, 256, 4)
render_cfg("C3", "david", 1920, 1080, 32, 1)
render_cfg("C4", "next-week-final", 1920, 1080, 16, 1)
sweep_cfg("david")
sweep_cfg("sycee")
path_ray_sweep()
