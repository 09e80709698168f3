#!/usr/bin/env python3
"""e2e rate of yart_closest_hit / yart_closest_hit_f32 with pinned host arrays for several YART_TUNE_HOST_CHUNK values
(16 Mi uniform rays vs david.obj).  tools/host_chunk_probe.py [chunk ...]"""
import importlib, os, sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import raysets
y = importlib.import_module("yet-another-raytracer_b200")
from bench import MeshOnlyScene, sweep_rays
ctx = y.Context(0)
sc = MeshOnlyScene(y, "david"); ctx.set_scene(sc.desc)
n = 1 << 24
src = sweep_rays(y, "david", "uniform", n)
o, d = src["origin"], src["direction"]
h_rays = torch.empty(n * 48, dtype=torch.uint8).pin_memory(); rays = h_rays.numpy().view(y.RAY_DTYPE)
rays["origin"], rays["direction"] = o, d
h_hits = torch.empty(n * 40, dtype=torch.uint8).pin_memory(); hits = h_hits.numpy().view(y.HIT_DTYPE)
h_r32 = torch.empty(n * 24, dtype=torch.uint8).pin_memory(); r32 = h_r32.numpy().view(y.abi.RAY_F32_DTYPE)
r32["origin"], r32["direction"] = o, d
h_h32 = torch.empty(n * 16, dtype=torch.uint8).pin_memory(); h32 = h_h32.numpy().view(y.abi.HIT_F32_DTYPE)
pageable = rays.copy()
pageable_hits = np.empty(n, dtype=y.HIT_DTYPE); pageable_hits[:] = 0  # (touched once: no page faults inside the timed calls)
registered = rays.copy(); registered_hits = np.zeros(n, dtype=y.HIT_DTYPE)
import time as _t
t0 = _t.perf_counter(); ctx.host_register(registered); ctx.host_register(registered_hits)
print("yart_host_register of %.0f + %.0f MB: %.1f ms" % (registered.nbytes / 1e6, registered_hits.nbytes / 1e6, (_t.perf_counter() - t0) * 1e3), flush=True)
for chunk in [int(a) for a in sys.argv[1:]] or [1 << 30, 1 << 22, 1 << 21, 1 << 20, 1 << 19]:
    os.environ["YART_TUNE_HOST_CHUNK"] = str(chunk)
    out = []
    for fn, args in ((ctx.closest_hit, (rays, hits)), (ctx.closest_hit_f32, (r32, h32)), (ctx.closest_hit, (pageable, pageable_hits)),
                     (ctx.closest_hit, (registered, registered_hits))):
        fn(args[0], 0, 0.0, float("inf"), y.ORDER_NEAR, hits=args[1])
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3): fn(args[0], 0, 0.0, float("inf"), y.ORDER_NEAR, hits=args[1])
        torch.cuda.synchronize(); out.append(3 * n / (time.perf_counter() - t0) / 1e6)
    print("chunk %10d: f64 pinned %7.1f  f32 pinned %7.1f  f64 pageable %7.1f  f64 yart_host_register'ed %7.1f Mrays/s" % (chunk, *out), flush=True)
