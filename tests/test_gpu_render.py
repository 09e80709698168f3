"""GPU parity, wavefront renderer: yart_render / yart_film_finalize / yart_generate_camera_rays
(CUDA through the C ABI) against the CPU oracle with the SAME Philox streams.

Tolerances (stated, as north_star asks): the image gate is relRMSE(film) <= 1 % and total
luminance within 0.5 %.  Because both sides consume identical random numbers and the kernels do
f64 arithmetic in the reference's order, the films actually agree to ~1e-12 except for the rare
sample whose path flips a branch on a last-bit difference of a transcendental (CUDA's sin/cos/
log vs glibc's); the tests therefore also assert a much tighter bound on almost all pixels.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rel_rmse(a, b):
    return float(np.sqrt(((a - b) ** 2).sum()) / max(np.sqrt((b ** 2).sum()), 1e-300))


def compare_films(got, want, what, tight_fraction=0.97):
    assert np.isfinite(got).all()
    r = rel_rmse(got, want)
    lum_g, lum_w = got[..., 1].sum(), want[..., 1].sum()
    assert r <= 0.01, "%s: relRMSE %.3g > 1%%" % (what, r)
    assert abs(lum_g - lum_w) <= 0.005 * abs(lum_w), "%s: luminance %.6g vs %.6g" % (what, lum_g, lum_w)
    scale = np.abs(want).max() + 1e-300
    close = (np.abs(got - want) <= 1e-9 * scale).all(axis=-1)
    assert close.mean() >= tight_fraction, "%s: only %.4f of the pixels agree to 1e-9" % (what, close.mean())
    return r, close.mean()


def test_camera_rays_are_bit_exact(yart, orc, ctx):
    preset = yart.ScenePreset("david")
    ctx.set_scene(preset)
    for (w, h, ap) in ((64, 48, None), (40, 40, 0.5)):
        cam = preset.camera(w, h, aperture=ap)
        got = ctx.camera_rays(cam, w, h, 3, 6, seed=11)
        want = orc.camera_rays(cam, w, h, 3, 6, seed=11)
        assert np.array_equal(got[0]["origin"], want[0]["origin"])
        assert np.array_equal(got[0]["direction"], want[0]["direction"])
        assert np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
        assert (got[1] >= 360).all() and (got[1] < 720).all() and (got[2] >= 0).all() and (got[2] < 1).all()


@pytest.mark.parametrize("scene,w,h,spp,depth", [
    ("cornell-box", 80, 80, 8, 50),          # BASELINE config 1, reduced
    ("david", 96, 96, 6, 50),                # BASELINE config 3, reduced
    ("bunny", 64, 64, 6, 50),                # BASELINE config 2's preset (sycee.obj stands in)
    ("next-week-final", 48, 48, 6, 50),      # BASELINE config 4, reduced
    ("cornell-box-smoke", 48, 48, 6, 50),
    ("random-scene", 48, 32, 4, 50),
    ("two-spheres", 48, 32, 4, 50),
    ("two-perlin-spheres", 48, 32, 4, 50),
    ("earth", 48, 32, 4, 50),
    ("simple-light", 48, 32, 8, 50),
    ("three-spheres", 48, 48, 6, 50),
    ("sycee", 48, 48, 4, 50),
    ("teapot", 48, 48, 4, 50),
    ("david", 40, 40, 4, 3),                 # depth exhaustion returns 1.0 (main.rs:544-546)
])
def test_render_matches_oracle(yart, orc, ctx, scene, w, h, spp, depth):
    preset = yart.ScenePreset(scene, seed=2)
    s = orc.Scene(preset)
    ctx.set_scene(preset)
    cam = preset.camera(w, h)
    want, st_w = s.render(cam, w, h, 0, spp, max_depth=depth, seed=9, n_threads=os.cpu_count())
    for order in (yart.ORDER_NEAR, yart.ORDER_REFERENCE):
        got, st = ctx.render(cam, w, h, 0, spp, max_depth=depth, seed=9, order=order)
        # libm-sensitive scenes (media: log; textures: sin/atan2/acos) get a looser "tight" share
        loose = scene in ("cornell-box-smoke", "next-week-final", "two-perlin-spheres", "simple-light")
        compare_films(got, want, "%s order %d" % (scene, order), 0.90 if loose else 0.97)
        assert st.paths == st_w.paths == w * h * spp
        assert abs(int(st.rays) - int(st_w.rays)) <= max(4, st_w.rays // 500)
        assert st.kernel_launches >= 4 and st.gpu_ms > 0
    rgba_g = ctx.film_finalize(got, spp)
    rgba_w = orc.film_finalize(want, spp)
    assert rgba_g.shape == (h, w, 4) and (rgba_g[..., 3] == 255).all()
    assert (np.abs(rgba_g.astype(int) - rgba_w.astype(int)) <= 1).mean() > 0.97


def test_film_finalize_is_exact_on_identical_input(yart, orc, ctx):
    g = np.random.Generator(np.random.Philox(1))
    film = g.random((40, 56, 3)) * 3.0
    film[0, 0] = [np.nan, 1, 1]
    film[0, 1] = [-1, -1, -1]
    film[0, 2] = [1e9, 1e9, 1e9]
    a = ctx.film_finalize(film, 7)
    b = orc.film_finalize(film, 7)
    # pow() may differ in the last bit; a u8 can flip only when 256*c sits within 1e-13 of an integer
    assert (a != b).sum() <= 1
    # widths that are not multiples of 8: the reference leaves the remainder pixels (0,0,0,0)
    film = np.ones((20, 21, 3))
    a = ctx.film_finalize(film, 1)
    b = orc.film_finalize(film, 1)
    assert np.array_equal(a, b) and (a[..., 3] == 0).any()


def test_render_is_deterministic_and_sample_ranges_add_up(yart, orc, ctx):
    """Multi-GPU sharding is by sample range: film[0,8) must equal
    film[0,3) + film[3,8) up to f64 summation order, and reruns must be bit-identical."""
    preset = yart.ScenePreset("cornell-box")
    ctx.set_scene(preset)
    w = h = 64
    cam = preset.camera(w, h)
    full, _ = ctx.render(cam, w, h, 0, 8, seed=4)
    again, _ = ctx.render(cam, w, h, 0, 8, seed=4)
    assert np.array_equal(full, again)
    small_batches, _ = ctx.render(cam, w, h, 0, 8, seed=4, batch_spp=3)
    assert np.array_equal(full, small_batches)  # per-pixel sums run in sample order either way
    a, _ = ctx.render(cam, w, h, 0, 3, seed=4)
    b, _ = ctx.render(cam, w, h, 3, 8, seed=4)
    assert np.allclose(a + b, full, rtol=1e-12, atol=1e-12)
    chained, _ = ctx.render(cam, w, h, 3, 8, seed=4, film=a.copy())
    assert np.array_equal(chained, full)
    other, _ = ctx.render(cam, w, h, 0, 8, seed=5)
    assert not np.array_equal(other, full)


def test_full_size_frame_matches_oracle_and_batches_add_up(yart, orc, ctx):
    """BASELINE.json config 3 at its real frame (david, 1920x1080, depth 50), two samples per pixel against the
    oracle on every pixel; then the size-independent property at a sample count that spans two device batches
    (batch_spp=32: a 32-spp and a 2-spp batch): film[0,34) == film[0,32) chained with [32,34), bit for bit."""
    preset = yart.ScenePreset("david")
    s = orc.Scene(preset)
    ctx.set_scene(preset)
    w, h = 1920, 1080
    cam = preset.camera(w, h)
    want, st_w = s.render(cam, w, h, 5, 7, max_depth=50, seed=3, n_threads=os.cpu_count())
    got, st = ctx.render(cam, w, h, 5, 7, max_depth=50, seed=3)
    compare_films(got, want, "david 1920x1080", 0.9995)
    assert st.paths == st_w.paths == w * h * 2
    assert abs(int(st.rays) - int(st_w.rays)) <= 64
    whole, st34 = ctx.render(cam, w, h, 0, 34, max_depth=50, seed=3, batch_spp=32)
    part, _ = ctx.render(cam, w, h, 0, 32, max_depth=50, seed=3)
    part, _ = ctx.render(cam, w, h, 32, 34, max_depth=50, seed=3, film=part)
    assert np.array_equal(whole, part)
    # SURVEY Appendix B measured ~4.3 rays / sample on square frames; 16:9 sees more background (~3.6)
    assert st34.paths == w * h * 34 and 3.2 < st34.rays / st34.paths < 4.6


def test_non_multiple_of_8_frames_follow_the_reference_tiling(yart, orc, ctx):
    preset = yart.ScenePreset("cornell-box")
    s = orc.Scene(preset)
    ctx.set_scene(preset)
    w, h = 43, 29
    cam = preset.camera(w, h)
    want, st_w = s.render(cam, w, h, 0, 4, seed=1, n_threads=4)
    got, st = ctx.render(cam, w, h, 0, 4, seed=1)
    compare_films(got, want, "43x29")
    assert st.paths == st_w.paths == (w // 8 * 8) * (h // 8 * 8) * 4


    def covered(n):  # tiles of n/8 starting at n*col/8 (main.rs:643-646)
        m = np.zeros(n, bool)
        for col in range(8):
            m[n * col // 8: n * col // 8 + n // 8] = True
        return m

    mask = covered(h)[:, None] & covered(w)[None, :]
    assert mask.sum() == st.paths // 4 and (got[~mask] == 0).all() and (got[mask].sum(axis=-1) != 0).mean() > 0.5


def test_cli_writes_a_png(yart, tmp_path):
    import importlib
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    cli = importlib.import_module("yart_cli")
    out = tmp_path / "sub" / "cornell.png"
    assert cli.main(["--scene", "cornell-box", "--width", "64", "--samples", "4", "--output", str(out)]) == 0
    from PIL import Image
    im = Image.open(out)
    assert im.size == (64, 64) and im.mode == "RGBA"
    px = np.asarray(im)
    assert (px[..., 3] == 255).all() and px[..., :3].mean() > 5


def test_cli_preview_checkpoint_and_resume_reproduce_the_one_shot_image(yart, tmp_path):
    """progressive preview and checkpoint / resume.  A render cut at 3 of 6 samples and
    resumed, and one that wrote previews every 2 samples, give the same bytes as the uninterrupted render."""
    import importlib
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    cli = importlib.import_module("yart_cli")
    from PIL import Image
    base = ["--scene", "cornell-box", "--width", "48", "--seed", "5"]
    one, prog, res = (str(tmp_path / n) for n in ("one.png", "prog.png", "res.png"))
    ck = str(tmp_path / "ck.npz")
    assert cli.main(base + ["--samples", "6", "--output", one]) == 0
    assert cli.main(base + ["--samples", "6", "--output", prog, "--preview-every", "2"]) == 0
    assert cli.main(base + ["--samples", "3", "--output", res, "--checkpoint", ck]) == 0
    half = np.asarray(Image.open(res)).copy()
    with np.load(ck) as z:
        assert int(z["samples_done"]) == 3 and z["film"].shape == (48, 48, 3)
    assert cli.main(base + ["--samples", "6", "--output", res, "--checkpoint", ck, "--resume"]) == 0
    a, b, c = (np.asarray(Image.open(p)) for p in (one, prog, res))
    assert np.array_equal(a, b) and np.array_equal(a, c) and not np.array_equal(a, half)
    with pytest.raises(SystemExit):  # a checkpoint of another image is refused
        cli.main(["--scene", "cornell-box", "--width", "48", "--seed", "6", "--samples", "6", "--output", res,
                  "--checkpoint", ck, "--resume"])
