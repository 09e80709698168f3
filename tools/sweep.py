#!/usr/bin/env python3
"""Closest-hit sweep (BASELINE config 5): N random rays vs a mesh QBVH, kernel-only time from the
library's own CUDA events.  Usage: python tools/sweep.py [--n 16777216] [--mesh david] [--reps 5]"""
import argparse
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1 << 24)
    ap.add_argument("--mesh", default="david")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--sets", default="uniform,axis")
    ap.add_argument("--count", action="store_true")
    a = ap.parse_args()
    # one launch over the whole ray set (the library would otherwise pipeline host arrays in 2^19-ray chunks)
    os.environ.setdefault("YART_TUNE_HOST_CHUNK", str(1 << 30))
    y = importlib.import_module("yet-another-raytracer_b200")
    from bench import MeshOnlyScene
    ms = MeshOnlyScene(y, a.mesh)
    q = y.L4QBVH.from_mesh(ms.mesh)
    ctx = y.Context(0)
    ctx.set_scene(ms.desc)
    for rs in a.sets.split(","):
        gen = raysets.uniform if rs == "uniform" else raysets.axis
        o, d = gen(a.n, q.info.bbox_min, q.info.bbox_max)
        rays = y.make_rays(o, d)
        for order, oname in ((y.ORDER_REFERENCE, "reference"), (y.ORDER_NEAR, "near")):
            ms_list = []
            for r in range(a.reps):
                hits, st = ctx.closest_hit(rays, 0, 0.0, float("inf"), order)
                ms_list.append(st.gpu_ms)
            best, med = min(ms_list), float(np.median(ms_list))
            line = "%s %s %s n=%d: best %.3f ms (%.1f Mrays/s) median %.3f ms (%.1f Mrays/s) hit-rate %.3f" % (
                a.mesh, rs, oname, a.n, best, a.n / best / 1e3, med, a.n / med / 1e3, (hits["prim_id"] != y.MISS).mean())
            if a.count:
                _, st = ctx.closest_hit(rays, 0, 0.0, float("inf"), order, count_visits=True)
                nb = (128 * st.node_visits + 48 * st.tri_tests) / a.n + 88
                line += " nodes/ray %.2f tris/ray %.2f alg bytes/ray %.0f -> %.1f GB/s" % (
                    st.node_visits / a.n, st.tri_tests / a.n, nb, nb * a.n / med / 1e6)
            print(line, flush=True)


if __name__ == "__main__":
    main()
