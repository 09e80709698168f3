#!/usr/bin/env python3
"""Throughput of one render call of the bench workload at several batch sizes (device film)."""
import importlib, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
y = importlib.import_module("yet-another-raytracer_b200")
p = y.ScenePreset("david"); ctx = y.Context(0); ctx.set_scene(p)
w, h = 1920, 1080
cam = p.camera(w, h)
film = torch.zeros((h, w, 3), dtype=torch.float64, device="cuda")
for spp in [int(a) for a in sys.argv[1:]] or [4, 8, 16, 32]:
    for rep in range(2):
        st = ctx.render_device(cam, w, h, 0, spp, film.data_ptr(), 50, 1, y.ORDER_NEAR, spp)
    print("spp/batch %2d: rays %d gpu_ms %.2f trace_ms %.2f Mrays/s %.1f launches %d" % (spp, st.rays, st.gpu_ms, st.trace_ms, st.rays / st.gpu_ms / 1e3, st.kernel_launches))
