"""The oracle against artefacts the REFERENCE ITSELF produced: its shipped renders output/david.png and
output/cornell_box.png, reduced to 40x40 block digests by tools/gen_reference_pins.py
(tests/golden/ref_*_png_lowfreq.npz).  The reference's tests hold no golden vector for the render path, and
its RNG is OS-seeded, so a converged low-frequency comparison is the strongest pin the reference offers:
it exercises the presets, camera, QBVH, materials, the biased light pick (hittable.rs:113-122), the spectral
pipeline and the film normalisation end to end.  CPU only; the GPU path has the same tests at 600x600 in
tests/test_gpu_reference_pins.py."""
import os

import numpy as np

import refpins


def test_oracle_david_render_matches_the_references_david_png(yart, orc):
    preset = yart.ScenePreset("david", seed=1)
    s = orc.Scene(preset)
    w = h = 80
    spp = 128
    cam = preset.camera(w, h)
    film, st = s.render(cam, w, h, 0, spp, max_depth=50, seed=1, n_threads=os.cpu_count())
    d = refpins.film_digest(film, spp, 40, "srgb")
    ref = refpins.load_pin("ref_david_png_lowfreq.npz")
    left, right, corr = refpins.halves(d, ref)
    print("david oracle 80x80x128 vs output/david.png: left %.4f right %.4f corr %.4f rays/sample %.3f"
          % (left, right, corr, st.rays / st.paths))
    # measured 0.993 / 1.045 / 0.915 (and 1.010 / 1.055 / 0.960 at 120x120x300): the matte half agrees to ~1 %,
    # the glass half reads ~5 % high because the shipped PNG clips its speculars per 600x600 pixel before
    # averaging (SURVEY.md Appendix B saw the same 0.4525 vs 0.4305)
    assert abs(left - 1.0) <= 0.03
    assert 0.97 <= right <= 1.12
    assert corr >= 0.85
    assert 4.2 < st.rays / st.paths < 4.45  # 4.29-4.31 world rays per sample


def test_oracle_cornell_render_matches_the_references_cornell_box_png(yart, orc):
    preset = yart.ScenePreset("cornell-box", seed=1)
    s = orc.Scene(preset)
    w = h = 120
    spp = 400
    cam = preset.camera(w, h)
    film, st = s.render(cam, w, h, 0, spp, max_depth=50, seed=1, n_threads=os.cpu_count())
    d = refpins.film_digest(film, spp, 40, "gamma2")
    ref = refpins.load_pin("ref_cornell_box_png_lowfreq.npz")
    ratios = refpins.region_ratios(d, ref)
    print("cornell oracle 120x120x400 vs output/cornell_box.png:",
          {k: round(v[0], 3) for k, v in ratios.items()}, "overall %.4f" % (d.mean() / ref.mean()))
    # measured: back wall 1.001 (per channel 1.000 / 1.000 / 1.002), green wall 0.999, ceiling 0.989, floor 0.989,
    # box front 0.998, red wall 0.963.  An unbiased light pick would read 0.5-0.6.
    for k in ("back wall", "green wall", "ceiling", "floor", "box front"):
        assert abs(ratios[k][0] - 1.0) <= 0.04, (k, ratios[k])
    assert np.abs(ratios["back wall"][1] - 1.0).max() <= 0.03  # the colour pipeline, channel by channel
    assert 0.88 <= ratios["red wall"][0] <= 1.03
    assert abs(d.mean() / ref.mean() - 1.0) <= 0.03
    assert refpins.correlation(d.mean(axis=-1), ref.mean(axis=-1)) >= 0.93
    assert 2.85 < st.rays / st.paths < 3.05  # 2.93-2.98


def test_oracle_sycee_render_matches_the_references_sycee_png(yart, orc):
    """output/sycee.png (1000x1000): the SF66 glass sycee.obj mesh on two loose ground triangles, one sphere light
    (scenes.rs:433-480, main.rs:351-366)."""
    preset = yart.ScenePreset("sycee", seed=1)
    s = orc.Scene(preset)
    w = h = 80
    spp = 100
    cam = preset.camera(w, h)
    film, st = s.render(cam, w, h, 0, spp, max_depth=50, seed=1, n_threads=os.cpu_count())
    d = refpins.film_digest(film, spp, 40, "srgb")
    ref = refpins.load_pin("ref_sycee_png_lowfreq.npz")
    overall = d.mean() / ref.mean()
    corr = refpins.correlation(d.mean(axis=-1), ref.mean(axis=-1))
    print("sycee oracle 80x80x100 vs output/sycee.png: overall %.4f rgb %s corr %.4f" % (
        overall, np.round(d.mean(axis=(0, 1)) / ref.mean(axis=(0, 1)), 3), corr))
    assert abs(overall - 1.0) <= 0.04       # measured 0.996 (rgb 1.005 / 0.990 / 0.994)
    assert corr >= 0.90                     # measured 0.951


def test_oracle_earth_render_matches_the_references_earth_png(yart, orc):
    """output/earth.png (1200x800): ImageTexture + get_sphere_uv (texture.rs:313-345, sphere.rs:213-220).  The file
    is of the legacy gamma-2.0 vintage and its JPEG decoder is not ours, so the level is held to 5 % and the
    structure (block correlation: the map's orientation and the sky gradient) tightly."""
    preset = yart.ScenePreset("earth", seed=1)
    s = orc.Scene(preset)
    w, h, spp = 240, 160, 64
    cam = preset.camera(w, h)
    film, st = s.render(cam, w, h, 0, spp, max_depth=50, seed=1, n_threads=os.cpu_count())
    d = refpins.film_digest(film, spp, (40, 60), "gamma2")
    ref = refpins.load_pin("ref_earth_png_lowfreq.npz")
    overall = d.mean() / ref.mean()
    corr = refpins.correlation(d.mean(axis=-1), ref.mean(axis=-1))
    print("earth oracle 240x160x64 vs output/earth.png: overall %.4f rgb %s corr %.4f" % (
        overall, np.round(d.mean(axis=(0, 1)) / ref.mean(axis=(0, 1)), 3), corr))
    assert abs(overall - 1.0) <= 0.05       # measured 1.0015
    assert corr >= 0.99                     # measured 0.997
