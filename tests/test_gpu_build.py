"""the GPU L4QBVH builder (csrc/device_build.cu) against the host builder
(csrc/host_qbvh.cpp, itself checked against the oracle's tree in test_host.py): the same nodes, triangle
order and shading records, byte for byte -- including meshes full of equal centroids, where only the
tie rule (original index) decides the permutation."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def assert_same_tree(yart, ctx, trimesh, keep):
    host = yart.L4QBVH(trimesh, keepalive=keep)
    dev = yart.L4QBVH(trimesh, keepalive=keep, ctx=ctx)
    for f in ("n_nodes", "n_leaves", "n_tris", "height", "root", "max_stack"):
        assert getattr(host.info, f) == getattr(dev.info, f), f
    assert list(host.info.bbox_min) == list(dev.info.bbox_min) and list(host.info.bbox_max) == list(dev.info.bbox_max)
    ht, dt = host.tris(), dev.tris()
    assert np.array_equal(ht["orig"], dt["orig"]), "permutation differs at %s" % np.flatnonzero(ht["orig"] != dt["orig"])[:8]
    assert ht.tobytes() == dt.tobytes()
    assert host.nodes().tobytes() == dev.nodes().tobytes()
    assert np.array_equal(host.shade(), dev.shade())
    return host


@pytest.mark.parametrize("name", ["cube", "sycee", "david"])
def test_device_build_equals_host_build(yart, ctx, assets, name):
    mesh = yart.TriangleMesh.from_obj("%s/%s.obj" % (assets, name))
    host = assert_same_tree(yart, ctx, mesh.trimesh, mesh)
    if name == "david":
        assert host.info.n_tris == 46664


@pytest.mark.parametrize("n,seed", [(5, 1), (6, 2), (17, 3), (64, 4), (1000, 5), (4097, 6), (30011, 7)])
def test_device_build_on_meshes_with_many_equal_centroids(yart, ctx, n, seed):
    rng = np.random.default_rng(seed)
    # vertices on a coarse integer lattice (and signed zeros): lots of identical centroids along every axis
    pos = rng.integers(-3, 4, size=(n, 3, 3)).astype(np.float32)
    pos[rng.random((n, 3, 3)) < 0.1] = -0.0
    nrm = rng.standard_normal((n, 3, 3))
    uv = rng.random((n, 3, 2)).astype(np.float32)
    t, keep = yart.trimesh_from_arrays(pos, nrm, uv)
    assert_same_tree(yart, ctx, t, keep)


def test_scene_built_on_the_device_traces_identically(yart, orc, ctx, assets):
    """yart_ctx_set_builder(DEVICE): set_scene builds on the GPU; closest hits and a render are unchanged."""
    import raysets
    preset = yart.ScenePreset("david")
    cam = preset.camera(64, 48)
    o, d = raysets.uniform(50000, [-60, 0, -90], [70, 200, 60], seed=3)
    rays = yart.make_rays(o, d)
    ctx.set_builder(yart.BUILDER_HOST)
    ctx.set_scene(preset)
    hits_h, _ = ctx.closest_hit(rays, yart.TARGET_WORLD, 0.001, float("inf"), yart.ORDER_NEAR)
    film_h, _ = ctx.render(cam, 64, 48, 0, 4, seed=2)
    try:
        ctx.set_builder(yart.BUILDER_DEVICE)
        ctx.set_scene(preset)
        hits_d, _ = ctx.closest_hit(rays, yart.TARGET_WORLD, 0.001, float("inf"), yart.ORDER_NEAR)
        film_d, _ = ctx.render(cam, 64, 48, 0, 4, seed=2)
    finally:
        ctx.set_builder(yart.BUILDER_DEVICE)  # the default
    assert hits_h.tobytes() == hits_d.tobytes() and np.array_equal(film_h, film_d)


def test_device_build_refuses_what_the_host_build_refuses(yart, ctx):
    t, keep = yart.trimesh_from_arrays(np.zeros((4, 3, 3), np.float32))
    with pytest.raises(yart.YartError):
        yart.L4QBVH(t, keepalive=keep, ctx=ctx)
