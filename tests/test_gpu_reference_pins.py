"""GPU renders (yart_render + yart_film_finalize through the C ABI) against artefacts the REFERENCE ITSELF produced:
output/david.png, output/cornell_box.png, output/sycee.png, output/earth.png as 40x40 block digests
(tools/gen_reference_pins.py -> tests/golden/ref_*_png_lowfreq.npz).  Rendered at the PNGs' own resolution with
enough samples that noise is below the asserted tolerances, displayed with the product's own finalisation, and
digested exactly like the PNGs (tests/refpins.py).  See tests/test_reference_pins.py for the same pins on the oracle.
"""
import numpy as np
import pytest

import refpins

pytestmark = pytest.mark.gpu


def _render_u8(yart, ctx, scene, w, h, spp):
    preset = yart.ScenePreset(scene, seed=1)
    ctx.set_scene(preset)
    cam = preset.camera(w, h)
    film, st = ctx.render(cam, w, h, 0, spp, max_depth=50, seed=1)
    assert st.paths == w * h * spp
    return film, ctx.film_finalize(film, spp), st


def test_gpu_david_600x600_matches_the_references_david_png(yart, ctx):
    spp = 2048
    film, rgba, st = _render_u8(yart, ctx, "david", 600, 600, spp)
    ref = refpins.load_pin("ref_david_png_lowfreq.npz")
    left, right, corr = refpins.halves(refpins.digest(rgba, 40, "srgb"), ref)
    fl, fr, fc = refpins.halves(refpins.film_digest(film, spp, 40, "srgb"), ref)
    print("david GPU 600x600x%d vs output/david.png: displayed left %.4f right %.4f corr %.4f | film-space left %.4f "
          "right %.4f corr %.4f | %.3f rays/sample" % (spp, left, right, corr, fl, fr, fc, st.rays / st.paths))
    assert abs(left - 1.0) <= 0.01     # matte David + background (0.2662 vs 0.2660)
    assert abs(right - 1.0) <= 0.06    # the glass instance: per-pixel clipping of speculars depends on the sample count
    assert corr >= 0.97
    assert abs(fl - 1.0) <= 0.01
    assert 4.2 < st.rays / st.paths < 4.45


def test_gpu_cornell_600x600_matches_the_references_cornell_box_png(yart, ctx):
    spp = 1024
    film, rgba, st = _render_u8(yart, ctx, "cornell-box", 600, 600, spp)
    ref = refpins.load_pin("ref_cornell_box_png_lowfreq.npz")  # linearised with the legacy gamma 2.0
    d = refpins.digest(rgba, 40, "srgb")                        # ours is displayed with the sRGB OETF
    ratios = refpins.region_ratios(d, ref)
    corr = refpins.correlation(d.mean(axis=-1), ref.mean(axis=-1))
    print("cornell GPU 600x600x%d vs output/cornell_box.png:" % spp, {k: round(v[0], 4) for k, v in ratios.items()},
          "overall %.4f corr %.4f" % (d.mean() / ref.mean(), corr))
    for k in ("back wall", "green wall", "ceiling", "floor", "box front"):
        assert abs(ratios[k][0] - 1.0) <= 0.035, (k, ratios[k])   # +-3.5 %
    assert np.abs(ratios["back wall"][1] - 1.0).max() <= 0.02
    assert 0.88 <= ratios["red wall"][0] <= 1.02
    assert abs(d.mean() / ref.mean() - 1.0) <= 0.02
    assert corr >= 0.99
    assert 2.85 < st.rays / st.paths < 3.05


def test_gpu_sycee_1000x1000_matches_the_references_sycee_png(yart, ctx):
    spp = 512
    film, rgba, st = _render_u8(yart, ctx, "sycee", 1000, 1000, spp)
    ref = refpins.load_pin("ref_sycee_png_lowfreq.npz")
    d = refpins.digest(rgba, 40, "srgb")
    overall, corr = d.mean() / ref.mean(), refpins.correlation(d.mean(axis=-1), ref.mean(axis=-1))
    fd = refpins.film_digest(film, spp, 40, "srgb")
    print("sycee GPU 1000x1000x%d vs output/sycee.png: displayed overall %.4f corr %.4f | film-space overall %.4f corr %.4f"
          % (spp, overall, corr, fd.mean() / ref.mean(), refpins.correlation(fd.mean(axis=-1), ref.mean(axis=-1))))
    assert abs(overall - 1.0) <= 0.04
    assert corr >= 0.97


def test_gpu_earth_1200x800_matches_the_references_earth_png(yart, ctx):
    spp = 256
    film, rgba, st = _render_u8(yart, ctx, "earth", 1200, 800, spp)
    ref = refpins.load_pin("ref_earth_png_lowfreq.npz")
    d = refpins.digest(rgba, (40, 60), "srgb")
    overall, corr = d.mean() / ref.mean(), refpins.correlation(d.mean(axis=-1), ref.mean(axis=-1))
    print("earth GPU 1200x800x%d vs output/earth.png: overall %.4f rgb %s corr %.4f" % (
        spp, overall, np.round(d.mean(axis=(0, 1)) / ref.mean(axis=(0, 1)), 3), corr))
    assert abs(overall - 1.0) <= 0.05
    assert corr >= 0.995
