#!/usr/bin/env python3
"""A small, fast exercise of every kernel and every scene feature.

Run against the bounds-checked library (build.py --bounds-check) it plays the part of a memcheck run:
    YART_LIB_PATH=yet-another-raytracer_b200/libyart_b200_checked.so python tools/exercise_kernels.py
Any index outside its array traps and the next call raises.  tests/test_gpu_bounds.py runs exactly this.
"""
import importlib, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import raysets
y = importlib.import_module("yet-another-raytracer_b200")
ctx = y.Context(0)
for scene, w, h, spp in (("david", 48, 32, 2), ("cornell-box-smoke", 32, 32, 2), ("next-week-final", 32, 24, 2), ("earth", 24, 16, 2)):
    p = y.ScenePreset(scene, seed=2)
    ctx.set_scene(p)
    cam = p.camera(w, h)
    film, st = ctx.render(cam, w, h, 0, spp, 50, 1)
    rgba = ctx.film_finalize(film, spp)
    rays, _, _ = ctx.camera_rays(cam, w, h, 0, 1)
    for order in (0, 1):
        ctx.closest_hit(rays, y.TARGET_WORLD, 0.001, float("inf"), order, count_visits=True)
    print(scene, st.rays, int(rgba.sum()))
p = y.ScenePreset("david"); ctx.set_scene(p)
o, d = raysets.uniform(20000, [-60, 0, -90], [70, 200, 60])
rays = y.make_rays(o, d)
weird = y.make_rays([(0, 0, 0), (0, 0, 5), (np.nan, 0, 0), (0, 0, 5)], [(0, 0, 0), (0, 0, -np.inf), (0, 0, 1), (0, 1e-320, -1)])
for r in (rays, weird, rays[:1], rays[:33]):
    for order in (0, 1):
        ctx.closest_hit(r, 0, 0.0, float("inf"), order)
print("exercise done")
