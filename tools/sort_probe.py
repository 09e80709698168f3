#!/usr/bin/env python3
"""How much would ray sorting buy?  Same 8 Mi random rays: unsorted vs sorted by direction octant, vs sorted by
(octant, Morton code of the origin)."""
import importlib, os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import raysets
os.environ.setdefault("YART_TUNE_HOST_CHUNK", str(1 << 30))  # one launch per query (no host-array chunk pipeline)
y = importlib.import_module("yet-another-raytracer_b200")
from bench import MeshOnlyScene
ms = MeshOnlyScene(y, "david"); q = y.L4QBVH.from_mesh(ms.mesh)
ctx = y.Context(0); ctx.set_scene(ms.desc)
n = 1 << 23
o, d = raysets.uniform(n, q.info.bbox_min, q.info.bbox_max)
lo, hi = np.asarray(q.info.bbox_min), np.asarray(q.info.bbox_max)
octant = (d[:, 0] >= 0).astype(np.uint64) | ((d[:, 1] >= 0).astype(np.uint64) << 1) | ((d[:, 2] >= 0).astype(np.uint64) << 2)
cell = np.clip(((o - lo) / (hi - lo) * 1024).astype(np.uint64), 0, 1023)
def spread(v):
    v = (v | (v << 16)) & np.uint64(0x030000FF); v = (v | (v << 8)) & np.uint64(0x0300F00F)
    v = (v | (v << 4)) & np.uint64(0x030C30C3); v = (v | (v << 2)) & np.uint64(0x09249249); return v
morton = spread(cell[:, 0]) | (spread(cell[:, 1]) << np.uint64(1)) | (spread(cell[:, 2]) << np.uint64(2))
def run(tag, idx):
    rays = y.make_rays(o[idx], d[idx])
    best = min(ctx.closest_hit(rays, 0, 0.0, float("inf"), y.ORDER_NEAR)[1].gpu_ms for _ in range(3))
    print("%-34s %.3f ms  %.0f Mrays/s" % (tag, best, n / best / 1e3), flush=True)
run("unsorted", np.arange(n))
run("sorted by octant", np.argsort(octant, kind="stable"))
run("sorted by octant, morton(origin)", np.argsort((octant << np.uint64(30)) | morton, kind="stable"))
run("sorted by morton(origin) only", np.argsort(morton, kind="stable"))
coarse = spread(cell[:, 0] >> np.uint64(6)) | (spread(cell[:, 1] >> np.uint64(6)) << np.uint64(1)) | (spread(cell[:, 2] >> np.uint64(6)) << np.uint64(2))
run("sorted by 16^3 cell, octant", np.argsort((coarse << np.uint64(3)) | octant, kind="stable"))
