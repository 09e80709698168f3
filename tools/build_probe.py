#!/usr/bin/env python3
"""Host vs GPU L4QBVH build time (csrc/host_qbvh.cpp vs csrc/device_build.cu) on david.obj and on synthetic soups."""
import importlib, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
y = importlib.import_module("yet-another-raytracer_b200")
ctx = y.Context(0)


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        t = time.perf_counter(); r = fn(); best = min(best, time.perf_counter() - t)
    return best, r


mesh = y.TriangleMesh.from_obj(y.assets_dir() + "/david.obj")
cases = [("david.obj", mesh.trimesh, mesh)]
rng = np.random.default_rng(1)
for n in [int(a) for a in sys.argv[1:]] or [1_000_000, 8_000_000]:
    c = rng.random((n, 1, 3), dtype=np.float32) * 100.0
    pos = c + rng.random((n, 3, 3), dtype=np.float32) * 0.2
    t, keep = y.trimesh_from_arrays(pos)
    cases.append(("%d random triangles" % n, t, keep))
for name, t, keep in cases:
    y.L4QBVH(t, keepalive=keep, ctx=ctx)  # warm-up (CUB temp sizing, first launches)
    td, qd = timed(lambda: y.L4QBVH(t, keepalive=keep, ctx=ctx))
    th, qh = timed(lambda: y.L4QBVH(t, keepalive=keep), reps=1)
    same = qd.nodes().tobytes() == qh.nodes().tobytes() and qd.tris().tobytes() == qh.tris().tobytes()
    print("%-26s nodes %8d  host %9.1f ms   device %8.1f ms (incl. upload + copy back)   identical %s" %
          (name, qh.info.n_nodes, th * 1e3, td * 1e3, same))
preset = y.ScenePreset("david")
for name, b in (("host", y.BUILDER_HOST), ("device", y.BUILDER_DEVICE)):
    ctx.set_builder(b)
    ctx.set_scene(preset)
    t, _ = timed(lambda: (ctx.set_scene(preset), ctx.synchronize()))
    print("yart_ctx_set_scene(david), %-6s builder: %.1f ms" % (name, t * 1e3))
