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
