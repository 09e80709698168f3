#!/usr/bin/env python3
"""One wavefront step of the bench workload (david 1920x1080, 8 spp) -- a short target for ncu."""
import importlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
scene = sys.argv[1] if len(sys.argv) > 1 else "david"
w, h, spp = (int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (1920, 1080, 8)
y = importlib.import_module("yet-another-raytracer_b200")
p = y.ScenePreset(scene)
ctx = y.Context(0)
ctx.set_scene(p)
cam = p.camera(w, h)
for rep in range(2):
    film, st = ctx.render(cam, w, h, 0, spp, max_depth=50, seed=1, order=y.ORDER_NEAR)
    print("rays %d paths %d gpu_ms %.2f trace_ms %.2f Mrays/s %.1f launches %d max_bounce %d" % (
        st.rays, st.paths, st.gpu_ms, st.trace_ms, st.rays / st.gpu_ms / 1e3, st.kernel_launches, st.max_bounce))
