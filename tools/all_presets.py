#!/usr/bin/env python3
"""Every scene preset once at a medium size (default 640 wide, 16 spp): rates, NaN check, image statistics."""
import importlib, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
y = importlib.import_module("yet-another-raytracer_b200")
ctx = y.Context(0)
width = int(sys.argv[1]) if len(sys.argv) > 1 else 640
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 16
for name in y.SCENE_NAMES:
    p = y.ScenePreset(name, seed=1)
    o = p.resolve_render_options(None, width, None, spp, None, None, None, None)
    ctx.set_scene(p)
    cam = p.camera(o["width"], o["height"], o["vfov"], o["aperture"])
    film, st = ctx.render(cam, o["width"], o["height"], 0, spp, o["max_depth"], 1)
    rgba = ctx.film_finalize(film, spp)
    print("%-20s %4dx%-4d %3d spp: %7.1f Mrays/s  %5.2f rays/sample  %7.1f ms  launches %4d  deepest bounce %2d  finite %s  mean rgb %s" % (
        name, o["width"], o["height"], spp, st.rays / st.gpu_ms / 1e3, st.rays / max(st.paths, 1), st.gpu_ms, st.kernel_launches,
        st.max_bounce, bool(np.isfinite(film).all()), np.round(rgba[..., :3].reshape(-1, 3).mean(0), 1)))
