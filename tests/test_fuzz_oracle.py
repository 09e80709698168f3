"""CPU side of the fuzz parity tests (tests/scene_fuzz.py): the oracle's near-first traversal must return what its
verbatim reference-order traversal returns on random scenes -- including coincident meshes, where a hit AT the t_max
left by an earlier object must stay out (the reference's `t_max > t`, qbvh.rs:478, is strict)."""
import ctypes as C
import os

import numpy as np
import pytest

from scene_fuzz import FuzzScene

INF = float("inf")


@pytest.mark.parametrize("seed", range(1000, 1024))
def test_oracle_near_order_equals_reference_order_on_random_scenes(yart, orc, seed):
    sc = FuzzScene(yart, seed)
    s = orc.Scene(sc)
    o, d = sc.rays(12000)
    rays = yart.make_rays(o, d)
    for t_min, t_max in ((0.001, INF), (0.0, 3.5)):
        a, _ = s.closest_hit(rays, yart.TARGET_WORLD, t_min, t_max, yart.ORDER_REFERENCE, n_threads=os.cpu_count())
        b, _ = s.closest_hit(rays, yart.TARGET_WORLD, t_min, t_max, yart.ORDER_NEAR, n_threads=os.cpu_count())
        assert a.tobytes() == b.tobytes()


def test_coincident_meshes_first_in_list_order_wins(yart, orc):
    """Three copies of the same cube at the same place: every hit is an exact tie between the objects, and
    HittableList::hit (hittable.rs:66-79) keeps the first -- in either traversal order."""
    abi = yart.abi
    mesh = yart.TriangleMesh.from_obj(os.path.join(yart.assets_dir(), "cube.obj"))
    tex, mat = abi.Texture(), abi.Material()
    tex.kind, mat.kind = abi.TEX_SOLID, abi.MAT_LAMBERTIAN
    objs = (abi.Object * 3)()
    for o in objs:
        o.kind, o.cos_theta = abi.OBJ_MESH, 1.0
    sd = abi.SceneDesc()
    sd.objects, sd.n_objects = C.cast(objs, C.POINTER(abi.Object)), 3
    sd.meshes, sd.n_meshes = C.pointer(mesh.trimesh), 1
    sd.materials, sd.n_materials = C.pointer(mat), 1
    sd.textures, sd.n_textures = C.pointer(tex), 1
    s = orc.Scene(C.pointer(sd))
    rng = np.random.default_rng(5)
    d = rng.normal(size=(5000, 3))
    o = -4.0 * d / np.linalg.norm(d, axis=1, keepdims=True) + rng.uniform(-0.3, 0.3, size=(5000, 3))
    rays = yart.make_rays(o, d)
    for order in (yart.ORDER_REFERENCE, yart.ORDER_NEAR):
        hits, _ = s.closest_hit(rays, yart.TARGET_WORLD, 0.001, INF, order)
        hit = hits["prim_id"] != yart.MISS
        assert hit.mean() > 0.5 and (hits["obj_id"][hit] == 0).all()


def _random_obj_text(seed):
    """A random OBJ file in the dialects tobj accepts: interleaved v / vt / vn and f lines (negative indices count back
    from what has been read so far), polygons, all four corner forms, ignored statements, odd number spellings, tabs,
    trailing blanks, CRLF line ends."""
    r = np.random.default_rng(seed)
    eol = "\r\n" if seed % 3 == 0 else "\n"
    lines, nv, nt, nn = ["# fuzz %d" % seed, "mtllib nothing.mtl"], 0, 0, 0

    def num():
        x = float(r.choice([r.uniform(-100, 100), r.integers(-5, 6), r.uniform(-1e-3, 1e-3), r.uniform(-1e6, 1e6)]))
        return str(r.choice(["%.9g" % x, "%.3f" % x, "%e" % x, "%.17g" % x, "%+.4E" % x, repr(np.float32(x).item())]))

    for block in range(int(r.integers(1, 6))):
        if r.random() < 0.5:
            lines.append(str(r.choice(["o part%d" % block, "g grp%d" % block, "s off", "usemtl m%d" % block])))
        for _ in range(int(r.integers(3, 12))):
            lines.append("v" + str(r.choice([" ", "\t", "  "])) + " ".join(num() for _ in range(3)) + str(r.choice(["", " ", "\t"])))
            nv += 1
        for _ in range(int(r.integers(0, 6))):
            lines.append("vt " + " ".join(num() for _ in range(int(r.choice([2, 2, 3])))))
            nt += 1
        for _ in range(int(r.integers(0, 6))):
            lines.append("vn " + " ".join(num() for _ in range(3)))
            nn += 1
        if r.random() < 0.3:
            lines += ["", "   ", "# a comment between blocks", "l 1 2", "p 1"]
        for _ in range(int(r.integers(1, 8))):
            k = int(r.choice([3, 3, 3, 4, 4, 5, 8]))
            form = int(r.integers(0, 4))
            if form in (1, 3) and nt == 0:
                form = 0
            if form in (2, 3) and nn == 0:
                form = 0 if form == 2 else 1 if nt else 0
            neg = r.random() < 0.4
            corners = []
            for _ in range(k):
                def idx(n):
                    i = int(r.integers(1, n + 1))
                    return str(i - n - 1) if neg else str(i)
                c = idx(nv)
                if form == 1:
                    c += "/" + idx(nt)
                elif form == 2:
                    c += "//" + idx(nn)
                elif form == 3:
                    c += "/" + idx(nt) + "/" + idx(nn)
                corners.append(c)
            lines.append("f " + str(r.choice([" ", "  ", "\t"])).join(corners))
    return eol.join(lines) + (eol if seed % 2 else "")


@pytest.mark.parametrize("seed", range(40))
def test_obj_loader_on_random_files_matches_the_independent_reader(yart, orc, tmp_path, seed):
    """TriangleMesh::from_obj (triangle.rs:110-175, tobj 4.0.2 GPU_LOAD_OPTIONS): csrc/host_obj.cpp against
    oracle/orc.py's independent numpy reader on random files."""
    p = tmp_path / ("fuzz%d.obj" % seed)
    p.write_bytes(_random_obj_text(seed).encode())
    m = yart.TriangleMesh.from_obj(p)
    pos, nrm, uv = orc.load_obj_numpy(p)
    assert m.n_tris == pos.shape[0] > 0
    assert np.array_equal(m.positions(), pos)
    assert np.array_equal(m.uvs(), uv)
    assert np.array_equal(m.normals(), nrm, equal_nan=True)


@pytest.mark.parametrize("seed,n_tris", [(s, n) for s, n in enumerate([5, 6, 7, 8, 9, 12, 13, 16, 17, 31, 33, 64, 100, 257, 1000, 3001])])
def test_host_tree_equals_oracle_tree_on_random_soups(yart, orc, seed, n_tris):
    """L4QBVH::new (qbvh.rs:251-361, 636-693): csrc/host_qbvh.cpp and the oracle build the same tree -- node for node,
    with every box, child id and axis -- from soups full of equal centroids, degenerate and duplicated triangles."""
    from scene_fuzz import fuzz_soup
    pos, nrm, uv = fuzz_soup(seed, n_tris)
    ms = orc.MeshScene(pos, nrm, uv)
    s = orc.Scene(ms)
    q = yart.L4QBVH(ms.trimesh)
    info = s.qbvh_info(0)
    assert (q.info.n_nodes, q.info.n_leaves, q.info.n_tris, q.info.root) == (info.n_nodes, info.n_leaves, info.n_tris, info.n_nodes - 1)
    assert list(q.info.bbox_min) == list(info.bbox_min) and list(q.info.bbox_max) == list(info.bbox_max)
    nodes, tris = q.nodes(), q.tris()
    assert np.array_equal(tris["orig"], s.tri_order(0))
    for i in range(info.n_nodes):
        boxes, children, axes = s.qbvh_node(0, i)
        nd = nodes[i]
        flat = np.concatenate([nd["min_x"], nd["min_y"], nd["min_z"], nd["max_x"], nd["max_y"], nd["max_z"]])
        assert np.array_equal(flat.astype(np.float64), boxes), "node %d boxes" % i
        assert np.array_equal(nd["child"], children), "node %d children" % i
        assert nd["axes"] == axes[0] | (axes[1] << 2) | (axes[2] << 4), "node %d axes" % i
    assert np.array_equal(tris["v0"], pos[tris["orig"], 0]) and np.array_equal(tris["v1"], pos[tris["orig"], 1])
