#!/usr/bin/env python3
"""Regenerate tests/golden/*.npz from the CPU oracle (the Rust reference cannot run here, so the
golden vectors are the oracle's own outputs, cross-checked against brute force when made).

    python tools/gen_golden.py

Committed with the fixtures so they can be reproduced; the CPU suite checks the oracle still
produces them, the GPU suite checks the CUDA path against them."""
import importlib
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import raysets  # noqa: E402
from oracle import orc  # noqa: E402

OUT = ROOT / "tests" / "golden"
INF = float("inf")


def main():
    y = importlib.import_module("yet-another-raytracer_b200")
    OUT.mkdir(exist_ok=True)
    A = y.assets_dir()
    for name in ("cube", "sycee", "david"):
        pos, nrm, uv = orc.load_obj_numpy(os.path.join(A, name + ".obj"))
        s = orc.Scene(orc.MeshScene(pos, nrm, uv))
        info = s.qbvh_info(0)
        o1, d1 = raysets.uniform(384, info.bbox_min, info.bbox_max, 1001)
        o2, d2 = raysets.axis(128, info.bbox_min, info.bbox_max, 1002)
        rays = orc.abi.make_rays(np.concatenate([o1, o2]), np.concatenate([d1, d2]))
        hits, _ = s.closest_hit(rays, 0, 0.001, INF, 0)
        bf, ties = s.brute_force_hit(rays, 0, 0.001)
        same = hits["t"] == bf["t"]
        assert same.mean() > 0.99, name  # (holes of flat leaves aside, SURVEY/DESIGN quirk list)
        np.savez_compressed(OUT / ("closest_hit_%s.npz" % name), rays=rays, hits=hits, brute_t=bf["t"], ties=ties)
    for scene in ("david", "cornell-box", "next-week-final"):
        p = y.ScenePreset(scene, seed=2)
        s = orc.Scene(p)
        w, h = 24, 16
        cam = p.camera(w, h)
        rays, wl, tm = orc.camera_rays(cam, w, h, 0, 1, seed=3)
        hits, _ = s.closest_hit(rays, orc.abi.TARGET_WORLD, 0.001, INF, 0)
        film, st = s.render(cam, w, h, 0, 4, max_depth=50, seed=3, n_threads=4)
        np.savez_compressed(OUT / ("scene_%s.npz" % scene), rays=rays, wavelength=wl, time=tm, hits=hits, film=film,
                            rays_traced=np.array([st.rays]), size=np.array([w, h, 4, 50, 3, 2]))
    print("wrote", sorted(f.name for f in OUT.glob("*.npz")))


if __name__ == "__main__":
    main()
