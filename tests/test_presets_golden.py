"""Every field of the 13 `yart_preset_build` scene descriptions (csrc/host_presets.cpp) against an INDEPENDENT
description: tests/golden/presets.json, parsed out of the reference's own scenes.rs / main.rs / material.rs text by
tools/gen_preset_golden.py.  The GPU path and the oracle both consume yart_preset_build's output, so without this a
wrong constant there would change both identically; with it, it fails here.  CPU only.

Deterministic builders are compared value for value (exact doubles: both sides parse the same decimal literals).  The
two builders that draw from thread_rng (random_scene, the_next_week_final_scene) are compared on their deterministic
statements, and their loops on the parameters the reference's text states (grid, radii, ranges, thresholds).
Documented deviations asserted as such: next-week-final's light carries FlipFace and its sphere group is non-empty
(SURVEY.md A-12); teapot / bunny load sycee.obj because their OBJ files are not shipped (A-13)."""
import copy
import ctypes as C
import hashlib
import json
import math
import os
from pathlib import Path

import numpy as np
import pytest

GOLD = json.loads((Path(__file__).resolve().parent / "golden" / "presets.json").read_text())
NAMES = ["random-scene", "two-spheres", "two-perlin-spheres", "earth", "simple-light", "cornell-box",
         "cornell-box-smoke", "next-week-final", "teapot", "bunny", "three-spheres", "sycee", "david"]
NOISE = {"square": 0, "trilinear": 1, "smooth": 2, "marble": 3, "net": 4}


class Mismatch(AssertionError):
    pass


def eq(what, got, want):
    g, w = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    if g.shape != w.shape or not np.array_equal(g, w):
        raise Mismatch("%s: got %s, reference says %s" % (what, got, want))


def texture_matches(abi, desc, ti, want, what):
    t = desc.textures[ti]
    if want["kind"] == "solid":
        if t.kind != abi.TEX_SOLID:
            raise Mismatch(what + ": texture kind")
        eq(what + " rgb", list(t.rgb_a), want["rgb"])
    elif want["kind"] == "checker":
        if t.kind != abi.TEX_CHECKER:
            raise Mismatch(what + ": texture kind")
        eq(what + " odd", list(t.rgb_a), want["odd"])   # odd is taken where the sine product is negative (texture.rs:57-68)
        eq(what + " even", list(t.rgb_b), want["even"])
    elif want["kind"] == "noise":
        if t.kind != abi.TEX_NOISE or t.noise_type != NOISE[want["noise_type"]]:
            raise Mismatch(what + ": noise kind")
        eq(what + " scale", t.scale, want["scale"])
        assert t.perlin < desc.n_perlins
    elif want["kind"] == "image":
        if t.kind != abi.TEX_IMAGE:
            raise Mismatch(what + ": texture kind")
        im = desc.images[t.image]
        a = GOLD["assets"][want["path"]]
        eq(what + " image size", [im.width, im.height], [a["width"], a["height"]])
    else:
        raise Mismatch(what + ": unknown golden texture " + want["kind"])


def material_matches(abi, desc, mi, want, what):
    assert mi < desc.n_materials
    m = desc.materials[mi]
    kinds = {"none": abi.MAT_NONE, "lambertian": abi.MAT_LAMBERTIAN, "metal": abi.MAT_METAL,
             "dielectric": abi.MAT_DIELECTRIC, "diffuse_light": abi.MAT_DIFFUSE_LIGHT, "isotropic": abi.MAT_ISOTROPIC}
    if m.kind != kinds[want["kind"]]:
        raise Mismatch("%s: material kind %d, reference says %s" % (what, m.kind, want["kind"]))
    if want["kind"] == "dielectric":
        eq(what + " sellmeier b", list(m.sellmeier_b), want["b"])
        eq(what + " sellmeier c", list(m.sellmeier_c), want["c"])
    if want["kind"] == "metal":
        eq(what + " fuzz", m.fuzz, want["fuzz"])
    if "texture" in want:
        texture_matches(abi, desc, m.texture, want["texture"], what + " texture")


def object_matches(abi, desc, o, want, what, lights=False, allow_extra_flip=False):
    """`want` is the golden nesting ConstantMedium(Translate(RotateY(FlipFace(primitive)))) -- the only order the
    reference uses and the ABI encodes."""
    wrap = 0
    medium = None
    if want["kind"] == "constant_medium":
        medium, want = want, want["boundary"]
        wrap |= abi.WRAP_MEDIUM
    offset = [0.0, 0.0, 0.0]
    if want["kind"] == "translate":
        offset, want = want["offset"], want["inner"]
        wrap |= abi.WRAP_TRANSLATE
    angle = None
    if want["kind"] == "rotate_y":
        angle, want = want["angle"], want["inner"]
        wrap |= abi.WRAP_ROTATE_Y
    if want["kind"] == "flip_face":
        want = want["inner"]
        wrap |= abi.WRAP_FLIP_FACE
    if allow_extra_flip:
        wrap |= abi.WRAP_FLIP_FACE
    if o.wrap != wrap:
        raise Mismatch("%s: wrap bits %d, reference nesting gives %d" % (what, o.wrap, wrap))
    if wrap & abi.WRAP_TRANSLATE:
        eq(what + " offset", list(o.offset), offset)
    if wrap & abi.WRAP_ROTATE_Y:  # RotateY::new: radians = angle.to_radians(); sin/cos (hittable.rs:166-169)
        r = math.radians(angle)
        if abs(o.sin_theta - math.sin(r)) > 1e-15 or abs(o.cos_theta - math.cos(r)) > 1e-15:
            raise Mismatch("%s: sin/cos of %g degrees" % (what, angle))
    if medium is not None:
        eq(what + " -1/density", o.neg_inv_density, -1.0 / medium["density"])
        material_matches(abi, desc, o.material, {"kind": "isotropic", "texture": medium["texture"]}, what + " phase function")
    k = want["kind"]
    p = list(o.p)
    if k == "sphere":
        if o.kind != abi.OBJ_SPHERE:
            raise Mismatch(what + ": kind")
        eq(what + " centre+radius", p[:4], want["center"] + [want["radius"]])
    elif k == "moving_sphere":
        if o.kind != abi.OBJ_MOVING_SPHERE:
            raise Mismatch(what + ": kind")
        eq(what, p[:9], want["center0"] + want["center1"] + [want["time0"], want["time1"], want["radius"]])
    elif k in ("xy_rect", "xz_rect", "yz_rect"):
        if o.kind != {"xy_rect": abi.OBJ_XY_RECT, "xz_rect": abi.OBJ_XZ_RECT, "yz_rect": abi.OBJ_YZ_RECT}[k]:
            raise Mismatch(what + ": kind")
        eq(what, p[:5], [want["a0"], want["a1"], want["b0"], want["b1"], want["k"]])
    elif k == "box":
        if o.kind != abi.OBJ_BOX:
            raise Mismatch(what + ": kind")
        eq(what, p[:6], want["p0"] + want["p1"])
    elif k == "triangle":
        if o.kind != abi.OBJ_TRIANGLE:
            raise Mismatch(what + ": kind")
        eq(what + " vertices", p[:9], sum(want["vertices"], []))
        eq(what + " normals", p[9:18], sum(want["normals"], []))
        eq(what + " uv", p[18:24], sum(want["uv"], []))
    elif k == "mesh":
        if o.kind != abi.OBJ_MESH:
            raise Mismatch(what + ": kind")
        assert o.index < desc.n_meshes
        path = want["path"]
        if path in GOLD["assets"]["missing"]:
            path = "input/sycee.obj"  # documented substitution
        if desc.meshes[o.index].n_tris != GOLD["assets"][path]["n_tris"]:
            raise Mismatch("%s: %d triangles, %s has %d" % (what, desc.meshes[o.index].n_tris, path, GOLD["assets"][path]["n_tris"]))
    elif k == "bvh":
        if o.kind != abi.OBJ_GROUP:
            raise Mismatch(what + ": kind")
        assert o.index < desc.n_groups
    else:
        raise Mismatch(what + ": unknown golden kind " + k)
    if medium is None and k != "bvh" and not lights:
        material_matches(abi, desc, o.material, want["material"], what + " material")
    if lights and want["material"]["kind"] != "none":
        raise Mismatch(what + ": sampling lights carry NoMaterial")


def compare_preset(yart, name, gold):
    """Raises Mismatch at the first field of yart_preset_build(name) that differs from the golden description."""
    abi = yart.abi
    preset = yart.ScenePreset(name, seed=7)
    desc = preset.desc.contents
    g = gold["presets"][name]
    info = preset.info
    for k in ("width", "height", "samples_per_pixel", "max_depth", "workers", "vfov", "aperture"):
        eq("%s defaults.%s" % (name, k), getattr(info, k), g["defaults"][k])
    eq(name + " lookfrom", list(info.lookfrom), g["lookfrom"])
    eq(name + " lookat", list(info.lookat), g["lookat"])
    if info.output_filename.decode() != g["output_filename"]:
        raise Mismatch(name + " output filename")
    eq(name + " background", list(desc.background_rgb), g["background"])
    cam = preset.camera(info.width, info.height)
    eq(name + " vup", list(cam.vup), gold["camera"]["vup"])
    eq(name + " focus/time", [cam.focus_dist, cam.time0, cam.time1],
       [gold["camera"]["focus_dist"], gold["camera"]["time0"], gold["camera"]["time1"]])
    eq(name + " camera", list(cam.lookfrom) + list(cam.lookat) + [cam.vfov_degrees, cam.aperture, cam.aspect_ratio],
       g["lookfrom"] + g["lookat"] + [g["defaults"]["vfov"], g["defaults"]["aperture"],
                                      g["defaults"]["width"] / g["defaults"]["height"]])
    if desc.n_lights != len(g["lights"]):
        raise Mismatch("%s: %d sampling lights, reference has %d" % (name, desc.n_lights, len(g["lights"])))
    for i, want in enumerate(g["lights"]):
        object_matches(abi, desc, desc.lights[i], want, "%s light %d" % (name, i), lights=True)
    # the world list, with random loops expanded on our side
    objs = [desc.objects[i] for i in range(desc.n_objects)]
    fixed_after = 0
    loop_at = [i for i, w in enumerate(g["world"]) if w["kind"] == "random_loop"]
    if loop_at:
        fixed_after = len(g["world"]) - loop_at[0] - 1
        head = objs[:loop_at[0]]
        tail = objs[len(objs) - fixed_after:]
        loop_objs = objs[loop_at[0]:len(objs) - fixed_after]
        wants = g["world"][:loop_at[0]] + g["world"][loop_at[0] + 1:]
        objs = head + tail
    else:
        loop_objs, wants = [], g["world"]
    if len(objs) != len(wants):
        raise Mismatch("%s: %d world objects, reference has %d" % (name, len(objs), len(wants)))
    for i, (o, want) in enumerate(zip(objs, wants)):
        flip = name == "next-week-final" and want["kind"] == "xz_rect"  # documented deviation: the light is flipped
        object_matches(abi, desc, o, want, "%s object %d" % (name, i), allow_extra_flip=flip)
    return preset, desc, loop_objs


@pytest.mark.parametrize("name", NAMES)
def test_preset_equals_the_reference_source(yart, name):
    compare_preset(yart, name, GOLD)


def test_a_wrong_constant_is_caught(yart):
    """The comparison has teeth: perturb single values of the golden description (as a wrong constant in
    host_presets.cpp would look from here) and expect a Mismatch each time."""
    def mutate(path, new):
        g = copy.deepcopy(GOLD)
        node = g["presets"]
        for k in path[:-1]:
            node = node[k]
        node[path[-1]] = new
        return g
    cases = [
        ("david", ["david", "world", 1, "inner", "angle"], 301.0),
        ("david", ["david", "world", 1, "offset"], [50.0, 0.0, 51.0]),
        ("david", ["david", "lights", 3, "center"], [1200.0, 1300.0, -800.0]),
        ("david", ["david", "world", 0, "material", "texture", "rgb"], [1.0, 1.0, 0.99]),
        ("david", ["david", "defaults", "vfov"], 21.0),
        ("cornell-box", ["cornell-box", "world", 2, "inner", "a0"], 214.0),
        ("cornell-box", ["cornell-box", "world", 7, "material", "b"], [2.0245976, 0.470187196, 2.6]),
        ("cornell-box", ["cornell-box", "world", 0, "material", "texture", "rgb"], [0.12, 0.45, 0.16]),
        ("cornell-box-smoke", ["cornell-box-smoke", "world", 6, "density"], 0.02),
        ("next-week-final", ["next-week-final", "world", 2, "center1"], [431.0, 400.0, 200.0]),
        ("two-spheres", ["two-spheres", "world", 0, "material", "texture", "odd"], [0.9, 0.9, 0.9]),
        ("bunny", ["bunny", "lights", 0, "center"], [0.0, 6.0, -2.0]),   # A-11: the sampling light is NOT where the emitter is
        ("three-spheres", ["three-spheres", "world", 3, "radius"], 0.7),
        ("simple-light", ["simple-light", "lookat"], [0.0, 2.5, 0.0]),
    ]
    for name, path, new in cases:
        with pytest.raises(Mismatch):
            compare_preset(yart, name, mutate(path, new))


def test_random_scene_follows_the_reference_loop_parameters(yart):
    abi = yart.abi
    R = GOLD["presets"]["random-scene"]["random"]
    for seed in (1, 2):
        preset = yart.ScenePreset("random-scene", seed=seed)
        desc = preset.desc.contents
        small = [desc.objects[i] for i in range(1, desc.n_objects - 3)]
        n_cells = (R["grid"][1] - R["grid"][0]) * (R["grid_b"][1] - R["grid_b"][0])
        assert n_cells - 8 <= len(small) <= n_cells
        cells, kinds = set(), []
        for o in small:
            assert o.kind == abi.OBJ_SPHERE and o.wrap == 0 and o.p[3] == R["radius"] and o.p[1] == R["y"]
            a, b = math.floor(o.p[0]), math.floor(o.p[2])
            assert R["grid"][0] <= a < R["grid"][1] and R["grid_b"][0] <= b < R["grid_b"][1]
            assert 0.0 <= o.p[0] - a < R["jitter"] + 1e-12 and 0.0 <= o.p[2] - b < R["jitter"] + 1e-12
            assert (a, b) not in cells
            cells.add((a, b))
            k = R["keep_out_center"]
            assert math.dist((o.p[0], o.p[1], o.p[2]), k) > R["keep_out_dist"]
            m = desc.materials[o.material]
            kinds.append(m.kind)
            if m.kind == abi.MAT_LAMBERTIAN:
                rgb = list(desc.textures[m.texture].rgb_a)
                assert all(R["lambertian_albedo_range"][0] <= c < R["lambertian_albedo_range"][1] for c in rgb)
            elif m.kind == abi.MAT_METAL:
                rgb = list(desc.textures[m.texture].rgb_a)
                assert all(R["metal_albedo_range"][0] <= c < R["metal_albedo_range"][1] for c in rgb)
                assert R["metal_fuzz_range"][0] <= m.fuzz < R["metal_fuzz_range"][1]
            else:
                assert m.kind == abi.MAT_DIELECTRIC
        kinds = np.array(kinds)
        assert abs((kinds == abi.MAT_LAMBERTIAN).mean() - R["p_lambertian"]) < 0.07
        assert abs((kinds == abi.MAT_METAL).mean() - (R["p_metal"] - R["p_lambertian"])) < 0.06
        assert abs((kinds == abi.MAT_DIELECTRIC).mean() - (1.0 - R["p_metal"])) < 0.05
    # a different seed is a different scene
    a = yart.ScenePreset("random-scene", seed=1).desc.contents.objects[5].p[0]
    b = yart.ScenePreset("random-scene", seed=2).desc.contents.objects[5].p[0]
    assert a != b


def test_next_week_final_groups_follow_the_reference_loop_parameters(yart):
    abi = yart.abi
    R = GOLD["presets"]["next-week-final"]["random"]
    preset = yart.ScenePreset("next-week-final", seed=3)
    desc = preset.desc.contents
    assert desc.n_groups == 2
    boxes, spheres = desc.groups[desc.objects[0].index], desc.groups[desc.objects[desc.n_objects - 1].index]
    n = R["boxes_per_side"]
    assert boxes.n_members == n * n
    for i in range(n):
        for j in range(n):
            o = boxes.members[i * n + j]
            x0, z0 = R["box_origin"] + i * R["box_width"], R["box_origin"] + j * R["box_width"]
            assert o.kind == abi.OBJ_BOX and o.wrap == 0
            assert [o.p[0], o.p[1], o.p[2], o.p[3], o.p[5]] == [x0, R["box_y0"], z0, x0 + R["box_width"], z0 + R["box_width"]]
            assert R["box_y1_range"][0] <= o.p[4] < R["box_y1_range"][1]
            material_matches(abi, desc, o.material, R["box_material"], "ground box")
    assert spheres.n_members == R["n_spheres"]   # deviation A-12(a): the reference reads the size before filling (0)
    cs = np.array([[spheres.members[i].p[k] for k in range(4)] for i in range(spheres.n_members)])
    assert (cs[:, 3] == R["sphere_radius"]).all()
    lo, hi = R["sphere_center_range"]
    assert (cs[:, :3] >= lo).all() and (cs[:, :3] < hi).all()
    assert abs(cs[:, :3].mean() - 0.5 * (lo + hi)) < 4.0 and cs[:, :3].std() > 40.0  # uniform over the cube
    for i in (0, 500, 999):
        material_matches(abi, desc, spheres.members[i].material,
                         {"kind": "lambertian", "texture": {"kind": "solid", "rgb": [0.73, 0.73, 0.73]}}, "white sphere")


def test_perlin_tables_follow_perlin_new(yart):
    """Perlin::new (texture.rs:96-111): ranfloat in [0,1), ranvec in [-1,1)^3 and -- because `permute` draws its
    target from 0..i, never i itself (texture.rs:185-192, SURVEY A-16) -- each permutation is ONE 256-cycle
    (Sattolo's algorithm), a property an ordinary Fisher-Yates shuffle has with probability 1/256."""
    for name, n_perlins in (("two-perlin-spheres", 2), ("simple-light", 2), ("next-week-final", 1)):
        desc = yart.ScenePreset(name, seed=11).desc.contents
        assert desc.n_perlins == n_perlins
        for k in range(n_perlins):
            p = desc.perlins[k]
            rf = np.array(list(p.ranfloat))
            rv = np.array([list(v) for v in p.ranvec])
            assert (rf >= 0).all() and (rf < 1).all() and (rv >= -1).all() and (rv < 1).all() and rf.std() > 0.2
            for perm in (p.perm_x, p.perm_y, p.perm_z):
                perm = list(perm)
                assert sorted(perm) == list(range(256))
                i, steps = perm[0], 1
                while i != 0:
                    i, steps = perm[i], steps + 1
                assert steps == 256


def test_shipped_assets_are_the_references_input_files(yart, assets):
    """assets/*.gz are byte-identical copies of the reference's input/*.obj; the earth map is earthmap.jpg decoded
    to RGB8 (PIL) -- hashes recorded from /root/reference by tools/gen_preset_golden.py."""
    for rel in ("input/cube.obj", "input/david.obj", "input/sycee.obj"):
        data = open(os.path.join(assets, os.path.basename(rel)), "rb").read()
        assert hashlib.sha256(data).hexdigest() == GOLD["assets"][rel]["sha256"]
    raw = open(os.path.join(assets, "earthmap_1024x512.rgb8"), "rb").read()
    e = GOLD["assets"]["input/earthmap.jpg"]
    assert len(raw) == e["width"] * e["height"] * 3 and hashlib.sha256(raw).hexdigest() == e["sha256_rgb8_pil"]
    assert sorted(GOLD["assets"]["missing"]) == ["input/bunny.obj", "input/teapot.obj"]
