"""CPU-only tests of the product's host side: the C-ABI library loads and exports every symbol
include/yart.h declares, the OBJ front end and the QBVH flattening agree with independent
restatements, presets carry the reference's values, and compute calls fail loudly without a GPU.
No compute calls are made here."""
import ctypes as C
import importlib
import os
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol(yart):
    header = (ROOT / "include" / "yart.h").read_text()
    declared = set(re.findall(r"\b(yart_[a-z_0-9]+)\s*\(", header))
    declared -= {"yart_status"}
    lib = C.CDLL(str(yart.LIB_PATH))
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, "declared in yart.h but not exported: %s" % missing
    assert declared == set(yart.EXPORTED_SYMBOLS)
    assert lib.yart_version  # and the binding lists nothing the header lacks
    # no torch types at the boundary: the library must not even link against torch / python
    needed = os.popen("ldd %s" % yart.LIB_PATH).read()
    assert "libtorch" not in needed and "libpython" not in needed and "libc10" not in needed


def test_compute_fails_loudly_without_a_gpu(yart):
    """There is no CPU fallback: without a device, context creation is a hard YART_ERR_CUDA."""
    if yart.device_count() > 0:
        pytest.skip("a GPU is visible here")
    with pytest.raises(yart.YartError) as e:
        yart.Context(0)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


@pytest.mark.parametrize("name,n_tris", [("cube", 12), ("sycee", 31642), ("david", 46664)])
def test_obj_loader_matches_independent_reader(yart, orc, assets, name, n_tris):
    """tobj semantics: fan triangulation, per-corner v/vt/vn, f32 attributes."""
    m = yart.TriangleMesh.from_obj(os.path.join(assets, name + ".obj"))
    pos, nrm, uv = orc.load_obj_numpy(os.path.join(assets, name + ".obj"))
    assert m.n_tris == n_tris == pos.shape[0]
    assert np.array_equal(m.positions(), pos)
    assert np.array_equal(m.normals(), nrm)
    assert np.array_equal(m.uvs(), uv)


def test_obj_loader_edge_cases(yart, tmp_path):
    p = tmp_path / "poly.obj"
    p.write_text("# pentagon, negative indices, no normals / uvs\n"
                 "v 0 0 0\nv 1 0 0\nv 1.5 1 0\nv 0.5 2 0\nv -0.5 1 0\n"
                 "f -5 -4 -3 -2 -1\n"
                 "l 1 2\np 1\n"
                 "v 0 0 1\nvt 0.25 0.75\nvn 0 0 1\nf 1/1/1 2//1 6/1\n")
    m = yart.TriangleMesh.from_obj(p)
    assert m.n_tris == 4  # fan (0,1,2)(0,2,3)(0,3,4) + one triangle; lines and points ignored
    pos = m.positions()
    assert np.array_equal(pos[1], [[0, 0, 0], [1.5, 1, 0], [0.5, 2, 0]])
    assert np.allclose(m.normals()[0], [[0, 0, 1]] * 3)  # face-normal fallback, f64
    assert np.array_equal(m.uvs()[0], np.zeros((3, 2)))
    assert np.array_equal(m.uvs()[3], [[0.25, 0.75], [0, 0], [0.25, 0.75]])
    # the third corner of the last face has no vn -> face normal of that triangle
    n = m.normals()[3]
    assert np.array_equal(n[0], [0, 0, 1]) and abs(np.linalg.norm(n[2]) - 1) < 1e-15
    with pytest.raises(yart.YartError) as e:
        yart.TriangleMesh.from_obj(tmp_path / "missing.obj")
    assert e.value.code == -4
    (tmp_path / "bad.obj").write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 9\n")
    with pytest.raises(yart.YartError):
        yart.TriangleMesh.from_obj(tmp_path / "bad.obj")


@pytest.mark.parametrize("name", ["cube", "sycee", "david"])
def test_flat_qbvh_equals_oracle_tree(yart, orc, mesh_scene, name):
    """The product's host build (flat f32 layout) and the oracle's (reference-style f64 nodes)
    must be the same tree: same node order, child ids, axes and -- losslessly -- the same boxes."""
    m, _, s = mesh_scene(name)
    q = yart.L4QBVH.from_mesh(m)
    info = s.qbvh_info(0)
    assert (q.info.n_nodes, q.info.n_leaves, q.info.n_tris) == (info.n_nodes, info.n_leaves, info.n_tris)
    assert q.info.root == info.n_nodes - 1  # post-order: the root is the last node (qbvh.rs:383-384)
    assert np.array_equal(np.asarray(q.info.bbox_min), np.asarray(info.bbox_min))
    assert np.array_equal(np.asarray(q.info.bbox_max), np.asarray(info.bbox_max))
    nodes, tris = q.nodes(), q.tris()
    assert np.array_equal(tris["orig"], s.tri_order(0))
    assert sorted(tris["orig"].tolist()) == list(range(info.n_tris))
    step = max(1, info.n_nodes // 400)
    for i in list(range(0, info.n_nodes, step)) + [info.n_nodes - 1]:
        boxes, children, axes = s.qbvh_node(0, i)
        nd = nodes[i]
        flat = np.concatenate([nd["min_x"], nd["min_y"], nd["min_z"], nd["max_x"], nd["max_y"], nd["max_z"]])
        assert np.array_equal(flat.astype(np.float64), boxes), "node %d boxes" % i
        assert np.array_equal(nd["child"], children)
        assert nd["axes"] == axes[0] | (axes[1] << 2) | (axes[2] << 4)
    # the traversal stack bound covers what the oracle actually needs
    o, d = [], []
    rng = np.random.Generator(np.random.Philox(5))
    lo, hi = np.asarray(info.bbox_min), np.asarray(info.bbox_max)
    rays = orc.abi.make_rays(lo + (hi - lo) * rng.random((2000, 3)), rng.standard_normal((2000, 3)))
    _, cnt = s.closest_hit(rays, 0, 0.0, float("inf"), 0)
    assert cnt.max_stack <= q.info.max_stack <= 64
    # vertices are stored verbatim
    P = m.positions()
    k = np.arange(0, info.n_tris, max(1, info.n_tris // 500))
    assert np.array_equal(tris["v0"][k], P[tris["orig"][k], 0])
    assert np.array_equal(tris["v2"][k], P[tris["orig"][k], 2])


def test_qbvh_rejects_what_the_reference_panics_on(yart, orc):
    pos = np.zeros((4, 3, 3), np.float32)
    ms = orc.MeshScene(pos, np.zeros((4, 3, 3)), np.zeros((4, 3, 2), np.float32))
    with pytest.raises(yart.YartError) as e:  # <= 4 triangles: zero nodes, hit() underflows (SURVEY A-17)
        yart.L4QBVH(ms.trimesh)
    assert e.value.code == -1


# ---- CLI / option resolution, mirroring main.rs:831-915 ---------------------------------------
def test_scene_names_match_the_reference_cli(yart):
    lib = yart.load_library()
    names = [lib.yart_preset_name(i).decode() for i in range(lib.yart_preset_count())]
    assert names == yart.SCENE_NAMES and "david" in names  # cli_accepts_david_scene
    with pytest.raises(yart.YartError) as e:  # cli_rejects_unknown_scene
        yart.ScenePreset("not-a-scene")
    assert e.value.code == -1


def test_resolve_dimensions(yart):  # main.rs:845-865
    assert yart.resolve_dimensions(1200, 800) == (1200, 800)
    assert yart.resolve_dimensions(1200, 800, 600, None) == (600, 400)
    assert yart.resolve_dimensions(1200, 800, None, 200) == (300, 200)
    assert yart.resolve_dimensions(1200, 800, 320, 240) == (320, 240)
    assert yart.resolve_dimensions(600, 600, 400, None) == (400, 400)  # BASELINE config 1
    assert yart.resolve_dimensions(1200, 800, 1, None) == (1, 1)


def test_resolve_render_options(yart):  # main.rs:868-915
    p = yart.ScenePreset("two-spheres")
    o = p.resolve_render_options()
    assert o["output_path"] == os.path.join("output", "two_spheres.png")
    assert (o["width"], o["height"], o["samples_per_pixel"], o["max_depth"], o["workers"]) == (1200, 800, 100, 50, 30)
    assert o["vfov"] == 20.0 and o["aperture"] == 0.0
    o = p.resolve_render_options(output="x.png", width=300, samples=7, max_depth=3, workers=2, vfov=35.0, aperture=0.25)
    assert o == {"output_path": "x.png", "width": 300, "height": 200, "samples_per_pixel": 7, "max_depth": 3,
                 "workers": 2, "vfov": 35.0, "aperture": 0.25}


def test_preset_values_follow_the_reference(yart):
    """Spot checks of build_scene_preset (main.rs:211-432) and scenes.rs."""
    d = yart.ScenePreset("david")
    sc = d.desc.contents
    assert (d.info.width, d.info.height, d.info.samples_per_pixel, d.info.vfov, d.info.aperture) == (600, 600, 10000, 20.0, 0.001)
    assert list(d.info.lookfrom) == [50.0, 120.0, 300.0] and list(d.info.lookat) == [0.0, 120.0, 0.0]
    assert sc.n_objects == 7 and sc.n_lights == 5 and sc.n_meshes == 1 and sc.meshes[0].n_tris == 46664
    inst = sc.objects[1]  # Translate(RotateY(mesh, 300), (50,0,50)) scenes.rs:590-596
    assert inst.kind == 7 and inst.wrap == 3 and list(inst.offset) == [50.0, 0.0, 50.0]
    assert abs(inst.sin_theta - np.sin(np.deg2rad(300.0))) < 1e-15 and abs(inst.cos_theta - 0.5) < 1e-15
    assert sc.materials[sc.objects[0].material].kind == 1 and sc.materials[inst.material].kind == 3
    lights = [tuple(sc.lights[i].p[:4]) for i in range(5)]
    assert lights[2] == lights[4] == (1200.0, 1300.0, -800.0, 700.0)  # the duplicated light (main.rs:396-410)
    glass = sc.materials[inst.material]
    assert glass.sellmeier_c[0] == 0.0147053225 * 1e6 and glass.sellmeier_b[2] == 2.59970433

    c = yart.ScenePreset("cornell-box")
    sc = c.desc.contents
    assert (c.info.width, c.info.height, c.info.vfov) == (600, 600, 40.0) and sc.n_objects == 8 and sc.n_lights == 2
    assert sc.objects[2].kind == 3 and sc.objects[2].wrap == 4  # FlipFace(XZRect light) scenes.rs:181-183
    assert list(sc.objects[2].p[:5]) == [213.0, 343.0, 227.0, 332.0, 554.0]
    assert sc.lights[0].kind == 3 and sc.lights[1].kind == 0  # rect first, glass sphere second
    box = sc.objects[6]
    assert box.kind == 5 and box.wrap == 3 and list(box.p[:6]) == [0, 0, 0, 165.0, 330.0, 165.0]
    cam = c.camera(400, 400)
    assert cam.focus_dist == 10.0 and cam.aspect_ratio == 1.0 and list(cam.vup) == [0, 1, 0] and (cam.time0, cam.time1) == (0.0, 1.0)

    b = yart.ScenePreset("bunny")
    sc = b.desc.contents
    assert sc.objects[3].p[2] == -2.0 and sc.lights[0].p[2] == 2.0  # emitter z=-2, sampling light z=+2 (A-11)

    n = yart.ScenePreset("next-week-final", seed=4)
    sc = n.desc.contents
    assert sc.n_groups == 2 and sc.groups[0].n_members == 400 and sc.groups[1].n_members == 1000
    assert sc.objects[1].wrap & 4  # documented deviation: the light is flipped (SURVEY A-12)
    assert sum(1 for i in range(sc.n_objects) if sc.objects[i].wrap & 8) == 2  # two constant media
    p2, p3 = yart.ScenePreset("next-week-final", seed=4), yart.ScenePreset("next-week-final", seed=5)
    n2, n3 = p2.desc.contents, p3.desc.contents  # (borrowed views: the presets must stay alive)
    assert [n2.groups[1].members[i].p[0] for i in range(5)] == [sc.groups[1].members[i].p[0] for i in range(5)]
    assert n3.groups[1].members[0].p[0] != sc.groups[1].members[0].p[0]  # seeded, reproducible


def test_perlin_tables_are_permutations(yart):
    p = yart.ScenePreset("two-perlin-spheres", seed=9).desc.contents
    assert p.n_perlins == 2
    for k in range(2):
        for perm in (p.perlins[k].perm_x, p.perlins[k].perm_y, p.perlins[k].perm_z):
            assert sorted(perm) == list(range(256))
        v = np.array([list(p.perlins[k].ranvec[i]) for i in range(256)])
        assert (np.abs(v) < 1).all() and (np.array(list(p.perlins[k].ranfloat)) < 1).all()


def test_cli_parses_like_the_reference():  # main.rs:831-842 and parse_positive_usize :148-158
    import importlib
    import sys
    sys.path.insert(0, str(ROOT))
    cli = importlib.import_module("yart_cli")
    y = importlib.import_module("yet-another-raytracer_b200")
    p = cli.build_parser(y.SCENE_NAMES)
    a = p.parse_args(["--scene", "david"])
    assert a.scene == "david" and a.width is None and a.samples is None
    a = p.parse_args(["--scene", "cornell-box", "--width", "400", "--samples", "32", "--max-depth", "50"])
    assert (a.width, a.samples, a.max_depth) == (400, 32, 50)
    for bad in (["--scene", "not-a-scene"], ["--scene", "david", "--samples", "0"], ["--scene", "david", "--width", "-3"], []):
        with pytest.raises(SystemExit):
            p.parse_args(bad)


def test_rust_sys_crate_source_matches_the_header(yart):
    """bindings/rust/yart-sys cannot be compiled here (no Rust toolchain), so keep it honest textually: it declares
    exactly the functions of yart.h, and its #[repr(C)] structs list the same fields in the same order as the
    ctypes structs the tests run with (sizes of the Rust scalar types give the same struct sizes)."""
    src = (ROOT / "bindings" / "rust" / "yart-sys" / "src" / "lib.rs").read_text()
    header = (ROOT / "include" / "yart.h").read_text()
    declared = set(re.findall(r"\b(yart_[a-z_0-9]+)\s*\(", header)) - {"yart_status"}
    rust_fns = set(re.findall(r"pub fn (yart_[a-z_0-9]+)\s*\(", src))
    assert rust_fns == declared
    abi = importlib.import_module("yet-another-raytracer_b200._abi")
    pairs = {"yart_trimesh": abi.Trimesh, "yart_object": abi.Object, "yart_group": abi.Group, "yart_material": abi.Material,
             "yart_texture": abi.Texture, "yart_perlin": abi.Perlin, "yart_image": abi.Image, "yart_scene_desc": abi.SceneDesc,
             "yart_camera": abi.Camera, "yart_stats": abi.Stats, "yart_render_opts": abi.RenderOpts,
             "yart_qbvh_info": abi.QbvhInfo, "yart_preset_info": abi.PresetInfo}
    scalar = {"u8": 1, "u32": 4, "i32": 4, "u64": 8, "f32": 4, "f64": 8, "c_char": 1}

    def size_of(ty):
        ty = ty.strip()
        if ty.startswith("*"):
            return 8, 8
        m = re.fullmatch(r"\[(.+);\s*(\d+)\]", ty)
        if m:
            s, a = size_of(m.group(1))
            return s * int(m.group(2)), a
        return scalar[ty], scalar[ty]

    for name, cls in pairs.items():
        m = re.search(r"pub struct %s \{(.*?)\n?\}" % name, src, re.S)
        assert m, name
        fields = re.findall(r"pub (\w+):\s*([^,}]+?)\s*(?:,|$)", m.group(1).replace("\n", " "))
        assert [f for f, _ in fields] == [f for f, _ in cls._fields_], name
        off = 0
        align = 1
        for _, ty in fields:  # C layout rules
            s, a = size_of(ty)
            off = (off + a - 1) // a * a + s
            align = max(align, a)
        assert (off + align - 1) // align * align == C.sizeof(cls), name
    for name, size in (("yart_ray", 48), ("yart_hit", 40)):
        assert re.search(r"pub struct %s \{" % name, src)
