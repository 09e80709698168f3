"""BASELINE.json's configs at the sizes they are written at, CUDA path (through the C ABI) against the oracle with
shared Philox streams.  Gates as SURVEY.md 8(d) states them: films relRMSE <= 1 %, total luminance within 0.5 %;
closest hits: primitive id equal and |dt| <= 1e-5 t -- and the MISMATCH COUNT is printed and asserted to be 0
(the kernels are in fact bit-exact, which is asserted too).

  C1  cornell-box, --width 400 (=> 400x400), 32 spp, depth 50: the whole film against the oracle.
  C2  bunny preset (sycee.obj stands in for the unshipped bunny.obj), 1280x720: closest-hit parity on (i) one primary ray
      per pixel centre, (ii) 1 Mi uniform rays, (iii) 1 Mi bounce rays dumped by the oracle from the preset's own paths.
  C3  david: the full 1920x1080 frame is in test_gpu_render.py; here the converged-image gate -- 480x272 at 64 spp.
  C4  next-week-final (volumes, Perlin / image textures, motion), 240x136 at 16 spp.
  C5  the 16 Mi-ray sweep is in test_gpu_closest_hit.py::test_full_size_sweep_properties.
"""
import os

import numpy as np
import pytest

import raysets
from test_gpu_render import compare_films

pytestmark = pytest.mark.gpu
INF = float("inf")


def hit_mismatches(got, want):
    """north_star's closest-hit criterion: prim id (and object) equal, |dt| <= 1e-5 t.  Returns (count, bit-exact?)."""
    miss_w = want["obj_id"] == 0xFFFFFFFF
    bad = (got["prim_id"] != want["prim_id"]) | (got["obj_id"] != want["obj_id"])
    both = ~miss_w & ~bad
    with np.errstate(invalid="ignore"):
        bad |= both & ~(np.abs(got["t"] - want["t"]) <= 1e-5 * np.abs(want["t"]))
    exact = all(np.array_equal(got[f], want[f]) for f in ("t", "u", "v", "prim_id", "obj_id", "front_face"))
    return int(bad.sum()), exact


def test_c1_cornell_box_400x400x32(yart, orc, ctx):
    preset = yart.ScenePreset("cornell-box", seed=1)
    s = orc.Scene(preset)
    ctx.set_scene(preset)
    w, h = yart.resolve_dimensions(preset.info.width, preset.info.height, 400, None)
    assert (w, h) == (400, 400)
    cam = preset.camera(w, h)
    want, st_w = s.render(cam, w, h, 0, 32, max_depth=50, seed=1, n_threads=os.cpu_count())
    for order in (yart.ORDER_NEAR, yart.ORDER_REFERENCE):
        got, st = ctx.render(cam, w, h, 0, 32, max_depth=50, seed=1, order=order)
        r, tight = compare_films(got, want, "C1 cornell 400x400x32 order %d" % order, 0.995)
        assert st.paths == st_w.paths == 400 * 400 * 32 and abs(int(st.rays) - int(st_w.rays)) <= 64
    print("C1 cornell-box 400x400x32: relRMSE %.3g, %.5f of pixels agree to 1e-9, %d rays (oracle %d)" % (
        r, tight, st.rays, st_w.rays))
    a, b = ctx.film_finalize(got, 32), orc.film_finalize(want, 32)
    assert (a != b).sum() <= 4  # displayed image: identical up to a last-bit pow() flip


def test_c2_bunny_preset_closest_hit_parity_1280x720(yart, orc, ctx):
    preset = yart.ScenePreset("bunny", seed=1)
    s = orc.Scene(preset)
    ctx.set_scene(preset)
    w, h = 1280, 720
    cam0 = preset.camera(w, h, aperture=0.0)
    sets = {}
    o, d = raysets.pixel_centre_rays(cam0, w, h)
    sets["(i) 1280x720 pixel-centre primaries"] = orc.abi.make_rays(o, d)
    c = np.asarray(list(preset.info.lookat))
    r = np.linalg.norm(np.asarray(list(preset.info.lookfrom)) - c)
    o, d = raysets.uniform(1 << 20, c - 0.5 * r, c + 0.5 * r, 4242)
    sets["(ii) 1 Mi uniform rays"] = orc.abi.make_rays(o, d)
    cam = preset.camera(w, h)
    bounce = s.dump_path_rays(cam, w, h, 0, 2, 1 << 22, max_depth=50, seed=1)
    # drop the camera rays: keep the secondary (bounce) rays the preset's own paths produce
    prim = orc.camera_rays(cam, w, h, 0, 2, seed=1)[0]
    is_primary = np.isin(bounce["origin"].view([("", "<f8")] * 3).reshape(-1), prim["origin"].view([("", "<f8")] * 3).reshape(-1)) & \
        np.isin(bounce["direction"].view([("", "<f8")] * 3).reshape(-1), prim["direction"].view([("", "<f8")] * 3).reshape(-1))
    bounce = bounce[~is_primary][:1 << 20]
    assert len(bounce) == 1 << 20
    sets["(iii) %d dumped bounce rays" % len(bounce)] = bounce
    total_bad = 0
    for what, rays in sets.items():
        want, _ = s.closest_hit(rays, yart.TARGET_WORLD, 0.001, INF, yart.ORDER_REFERENCE, n_threads=os.cpu_count())
        for order in (yart.ORDER_REFERENCE, yart.ORDER_NEAR):
            got, st = ctx.closest_hit(rays, yart.TARGET_WORLD, 0.001, INF, order)
            bad, exact = hit_mismatches(got, want)
            print("C2 bunny preset %s, order %d: %d rays, hit rate %.3f, MISMATCHES %d, bit-exact %s" % (
                what, order, len(rays), (want["obj_id"] != yart.MISS).mean(), bad, exact))
            total_bad += bad
            assert exact
    assert total_bad == 0


def test_c3_david_converged_image_gate_480x272x64(yart, orc, ctx):
    preset = yart.ScenePreset("david", seed=1)
    s = orc.Scene(preset)
    ctx.set_scene(preset)
    w, h, spp = 480, 272, 64
    cam = preset.camera(w, h)
    want, st_w = s.render(cam, w, h, 0, spp, max_depth=50, seed=1, n_threads=os.cpu_count())
    got, st = ctx.render(cam, w, h, 0, spp, max_depth=50, seed=1)
    r, tight = compare_films(got, want, "C3 david 480x272x64", 0.999)
    a, b = ctx.film_finalize(got, spp).astype(int), orc.film_finalize(want, spp).astype(int)
    u8_rmse = float(np.sqrt(((a - b) ** 2).mean()))
    print("C3 david 480x272x%d: relRMSE %.3g, luminance ratio %.9f, u8 sRGB RMSE %.4f, %.5f of pixels to 1e-9" % (
        spp, r, got[..., 1].sum() / want[..., 1].sum(), u8_rmse, tight))
    assert u8_rmse < 0.5 and st.paths == st_w.paths == w * h * spp


def test_c4_next_week_final_240x136x16(yart, orc, ctx):
    preset = yart.ScenePreset("next-week-final", seed=1)
    s = orc.Scene(preset)
    ctx.set_scene(preset)
    w, h, spp = 240, 136, 16
    cam = preset.camera(w, h)
    want, st_w = s.render(cam, w, h, 0, spp, max_depth=50, seed=1, n_threads=os.cpu_count())
    got, st = ctx.render(cam, w, h, 0, spp, max_depth=50, seed=1)
    # media take log() of a uniform draw and textures sin/atan2/acos: a last-bit difference can flip a branch of a
    # rare path, so the "agrees to 1e-9" share is lower here; the stated gates (1 % / 0.5 %) hold with room
    r, tight = compare_films(got, want, "C4 next-week-final 240x136x16", 0.90)
    print("C4 next-week-final 240x136x%d: relRMSE %.3g, luminance ratio %.9f, %.5f of pixels to 1e-9, %.2f rays/sample" % (
        spp, r, got[..., 1].sum() / want[..., 1].sum(), tight, st.rays / st.paths))
    assert st.paths == st_w.paths == w * h * spp
