#!/usr/bin/env python3
"""Per-bounce closest-hit / shade times of one batch of any preset (YART_DEBUG_BOUNCES=1 makes the library print them):
    YART_DEBUG_BOUNCES=1 python tools/scene_probe.py next-week-final 1920 1080 16"""
import importlib, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
y = importlib.import_module("yet-another-raytracer_b200")
scene = sys.argv[1] if len(sys.argv) > 1 else "david"
w, h, spp = (int(a) for a in sys.argv[2:5]) if len(sys.argv) > 4 else (1920, 1080, 16)
p = y.ScenePreset(scene); ctx = y.Context(0); ctx.set_scene(p)
cam = p.camera(w, h)
film = torch.zeros((h, w, 3), dtype=torch.float64, device="cuda")
for rep in range(2):
    st = ctx.render_device(cam, w, h, 0, spp, film.data_ptr(), 50, 1, y.ORDER_NEAR, spp)
print("%s %dx%d %d spp: rays %d gpu_ms %.2f trace_ms %.2f Mrays/s %.1f launches %d (%d closest-hit) max_bounce %d" % (
    scene, w, h, spp, st.rays, st.gpu_ms, st.trace_ms, st.rays / st.gpu_ms / 1e3, st.kernel_launches, st.trace_launches, st.max_bounce))
