#!/usr/bin/env python3
"""Reference-held parity pins: low-frequency digests of the reference's own shipped renders.

    python tools/gen_reference_pins.py            (needs /root/reference; run in the dev container only)

Writes tests/golden/ref_{david,cornell_box,sycee,earth}_png_lowfreq.npz.

The reference's tests hold no golden vector for the render path and the Rust
binary cannot be built here, so the only artefacts *produced by the reference itself* are the PNGs
under /root/reference/output/.  Two of them were written by scene code that is still what
`scenes.rs` / `main.rs` hold today:
  * output/david.png  (600x600): today's pipeline incl. the sRGB OETF (color.rs:92-107);
  * output/cornell_box.png (600x600): today's scene, but the legacy display transform -- a plain
    gamma 2.0 (`sqrt`, still visible in the comments at main.rs:684-688) instead of the sRGB OETF.
  * output/sycee.png (1000x1000): today's pipeline (sRGB OETF); the glass sycee.obj mesh on two ground triangles;
  * output/earth.png (1200x800): the image-textured sphere; legacy gamma 2.0.  Used for structure (block
    correlation: the texture lookup and get_sphere_uv orientation) more than for absolute level.
(three_spheres.png, the_next_week_final_scene.png, two_perlin_spheres.png and random_scene.png do not
correspond to today's scene code or depend on the unseeded RNG; bunny/teapot need the missing OBJ files.)
Their sample counts are unknown (far above anything a test renders), so what is kept is what a
converged image determines and noise does not: the linearised image box-averaged to 40x40 blocks
(15x15 pixels each for the 600x600 frames).  tests/refpins.py turns a render of ours into the same digest; the tests
compare region means and block correlation (tests/test_reference_pins.py, tests/test_gpu_reference_pins.py).

No reference SOURCE is copied: each output is a 40x40x3 (earth: 40x60x3) float array derived from image pixels.
"""
import sys
from pathlib import Path

import numpy as np
from PIL import Image

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))
import refpins  # noqa: E402

REF_OUT = Path("/root/reference/output")


def main():
    out = ROOT / "tests" / "golden"
    for name, linearise, dst, size, nb in (("david.png", "srgb", "ref_david_png_lowfreq.npz", (600, 600), (40, 40)),
                                           ("cornell_box.png", "gamma2", "ref_cornell_box_png_lowfreq.npz", (600, 600), (40, 40)),
                                           ("sycee.png", "srgb", "ref_sycee_png_lowfreq.npz", (1000, 1000), (40, 40)),
                                           ("earth.png", "gamma2", "ref_earth_png_lowfreq.npz", (800, 1200), (40, 60))):
        im = np.asarray(Image.open(REF_OUT / name).convert("RGBA"))
        assert im.shape == size + (4,) and (im[..., 3] == 255).all()
        lin = refpins.linearise(im[..., :3], linearise)
        blocks = refpins.block_mean(lin, nb)
        np.savez_compressed(out / dst, blocks=blocks.astype(np.float32), source=np.array("output/" + name),
                            linearise=np.array(linearise), size=np.array(im.shape[:2]))
        print(dst, blocks.shape, "mean %.4f" % blocks.mean())


if __name__ == "__main__":
    main()
