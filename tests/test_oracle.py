"""The oracle pinned against everything the reference offers for this path:
known answers (Appendix D, derived independently by brute force), the 4-triangle fixture of
qbvh.rs:1168-1246, brute force vs tree, Philox known-answer vectors, and the reference's own
unit tests restated (main.rs:783-916, color.rs:1984-2007, qbvh.rs:801-817).  CPU only."""
import ctypes as C

import numpy as np
import pytest

import raysets

INF = float("inf")

# (origin, direction, triangle, t, u, v), t_min = 0.001
APPENDIX_D = {
    "david": [
        ((50, 120, 300), (-50, 0, -300), 4617, 0.898004531629719, 0.115053499024, 0.125230910863),
        ((50, 120, 300), (-45, 10, -300), 4881, 0.913677506906396, 0.066186596336, 0.565977447081),
        ((50, 120, 300), (-60, -30, -300), 500, 0.913745944578526, 0.201501851285, 0.0408707379663),
        ((0, 100, 200), (0, 0, -1), 3494, 171.531272979231, 0.222295218534, 0.488353778865),
        ((-100, 150, 0), (1, 0, 0), 15160, 47.9782348839455, 0.363261807749, 0.416020533512),
        ((6.5, 250, -16), (0, -1, 0), 24194, 57.9046893628717, 0.305742983857, 0.539779423117),
        ((10, 60, -200), (0.1, 0.3, 1), 42848, 134.439396385678, 0.209821232005, 0.44725619781),
        ((300, 300, 300), (-1, -1, -1), 46593, 273.047073066032, 0.00887197304, 0.577471250458),
    ],
    "sycee": [
        ((1, 5, -8), (-1, -4.5, 8), 6044, 0.947010155709422, 0.234617837313, 0.0832507557203),
        ((-3, 0.4, 0.1), (1, 0, 0), 23189, 2.13676520102015, 0.142561628768, 0.62982679537),
        ((0.3, 0.2, -3), (0, 0.05, 1), 9999, 2.35496145451006, 0.117118560528, 0.650650770834),
        ((2, 2, 2), (-1, -0.9, -1), 3272, 1.4640082014909, 0.0220449167767, 0.304854359185),
    ],
    "cube": [
        ((0.25, 0.5, 5), (0, 0, -1), 3, 3.99999940395325, 0.12499968335, 0.25),
        ((3, 2, 1), (-3, -2, -1), 2, 0.66666683554627, 0.499999339382, 0.166666835546),
        ((0, 0, 0), (1, 0.5, 0.25), 2, 0.999999620020908, 0.374999488704, 0.250000094995),
    ],
}


def test_philox_known_answers(orc):
    """Random123 kat_vectors for philox4x32-10."""
    cases = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in cases:
        c, k, o = np.array(ctr, np.uint32), np.array(key, np.uint32), np.zeros(4, np.uint32)
        orc.lib().orc_philox4x32_10(c.ctypes.data, k.ctypes.data, o.ctypes.data)
        assert o.tolist() == want


def test_uniform_draws_are_in_unit_interval(orc):
    u = np.zeros(2)
    seen = []
    for slot in range(64):
        orc.lib().orc_uniform2(1, 5, 7, 2, slot, u.ctypes.data)
        assert 0.0 <= u[0] < 1.0 and 0.0 <= u[1] < 1.0
        seen.extend(u.tolist())
    assert len(set(seen)) == 128 and 0.3 < np.mean(seen) < 0.7


@pytest.mark.parametrize("name", ["david", "sycee", "cube"])
def test_appendix_d_known_answers(orc, mesh_scene, name):
    _, _, s = mesh_scene(name)
    rows = APPENDIX_D[name]
    rays = orc.abi.make_rays([r[0] for r in rows], [r[1] for r in rows])
    bf, ties = s.brute_force_hit(rays, 0, 0.001)
    for order in (0, 1):
        hits, _ = s.closest_hit(rays, 0, 0.001, INF, order)
        for i, (_, _, tri, t, u, v) in enumerate(rows):
            assert ties[i] == 1
            assert bf[i]["prim_id"] == tri and hits[i]["prim_id"] == tri
            assert hits[i]["t"] == bf[i]["t"] and hits[i]["u"] == bf[i]["u"] and hits[i]["v"] == bf[i]["v"]
            assert abs(hits[i]["t"] - t) <= 1e-12 * t
            assert abs(hits[i]["u"] - u) < 1e-11 and abs(hits[i]["v"] - v) < 1e-11


def test_sycee_axis_ray_through_shared_vertex(orc, mesh_scene):
    """SURVEY.md Appendix D's tie row.  Brute force sees several triangles at the same t; the
    reference's QBVH itself returns NO hit for this ray: the ray runs exactly along x = 0 = the
    split plane, so (bound - origin) * (1/0) is 0 * inf = NaN for one bound and -inf/+inf for the
    other, and the minNum/maxNum folds of qbvh.rs:506-519 close the interval.  The oracle keeps
    that behaviour (the GPU must too)."""
    _, _, s = mesh_scene("sycee")
    ray = orc.abi.make_rays([(0, 3, 0)], [(0, -1, 0)])
    bf, ties = s.brute_force_hit(ray, 0, 0.001)
    assert abs(bf[0]["t"] - 2.08414000272751) < 1e-12 and ties[0] >= 5
    assert 4553 in s.tie_set(ray[0], 0).tolist()
    for order in (0, 1):
        hits, _ = s.closest_hit(ray, 0, 0.001, INF, order)
        assert hits[0]["prim_id"] == orc.abi.MISS and hits[0]["t"] == INF


def test_four_triangle_fixture(orc):
    """get_triangles_and_ray_for_hit_test (qbvh.rs:1168-1246): all four triangles are hit; the
    closest is #1 at t = 0.954084586."""
    v = [[(-1.076726, -0.017016, 0.613202), (-1.117708, -0.041064, 0.593336), (-1.124824, -0.040218, 0.613324)],
         [(-1.076726, -0.017016, 0.613202), (-1.124824, -0.040218, 0.613324), (-1.07417, -0.017264, 0.633146)],
         [(-1.074288, -0.017314, 0.593228), (-1.117708, -0.041064, 0.593336), (-1.076726, -0.017016, 0.613202)],
         [(-1.105662, -0.042332, 0.573312), (-1.117708, -0.041064, 0.593336), (-1.074288, -0.017314, 0.593228)]]
    # the fixture's f64 literals are not f32-exact; the mesh path stores f32, so compare against
    # brute force on the same (rounded) vertices and only loosely against the f64 expectation
    pos = np.array(v + v, dtype=np.float32)  # 8 triangles: L4QBVH needs more than 4
    nrm = np.zeros_like(pos, dtype=np.float64)
    nrm[..., 1] = 1.0
    ms = orc.MeshScene(pos, nrm, np.zeros((8, 3, 2), np.float32))
    s = orc.Scene(ms)
    ray = orc.abi.make_rays([(-0.003898251, 2.0127985, 9.99872)], [(-1.1280149, -2.129233, -9.836952)])
    bf, ties = s.brute_force_hit(ray, 0, 0.0)
    assert bf[0]["prim_id"] == 1 and ties[0] == 2  # the duplicate (#5) ties exactly
    assert abs(bf[0]["t"] - 0.954084586) < 2e-6
    assert abs(bf[0]["u"] - 0.071161399) < 2e-4 and abs(bf[0]["v"] - 0.011283724) < 2e-4
    for order in (0, 1):
        hits, _ = s.closest_hit(ray, 0, 0.0, INF, order)
        assert hits[0]["t"] == bf[0]["t"] and hits[0]["prim_id"] in (1, 5)


@pytest.mark.parametrize("name,n", [("cube", 4000), ("sycee", 1500), ("david", 1500)])
def test_tree_equals_brute_force(orc, mesh_scene, name, n):
    """L4QBVH::hit (reference order and mirrored order) against testing every triangle."""
    m, _, s = mesh_scene(name)
    info = s.qbvh_info(0)
    o1, d1 = raysets.uniform(n, info.bbox_min, info.bbox_max)
    o2, d2 = raysets.axis(n // 2, info.bbox_min, info.bbox_max)
    rays = orc.abi.make_rays(np.concatenate([o1, o2]), np.concatenate([d1, d2]))
    for t_min in (0.0, 0.001):
        bf, ties = s.brute_force_hit(rays, 0, t_min)
        ref, c_ref = s.closest_hit(rays, 0, t_min, INF, 0)
        near, c_near = s.closest_hit(rays, 0, t_min, INF, 1)
        single = ties <= 1
        # The reference's box test is strict (`tfar > tnear`, qbvh.rs:532), so a leaf whose box has
        # zero thickness -- only axis-aligned flat triangles -- is never entered: the tree has holes
        # there (sycee has 4 such triangles).  Everywhere else tree == brute force, bit for bit.
        P = m.positions()
        flat = set(np.nonzero((np.ptp(P, axis=1) == 0).any(axis=1))[0].tolist())
        differs = np.nonzero(ref["t"] != bf["t"])[0]
        assert all(int(bf["prim_id"][i]) in flat for i in differs) and len(differs) <= 2
        same = ref["t"] == bf["t"]
        assert np.array_equal(ref["prim_id"][single & same], bf["prim_id"][single & same])
        # the mirrored order returns the very same record, ties included
        for f in ("t", "u", "v", "prim_id", "front_face"):
            assert np.array_equal(ref[f], near[f]), f
        assert c_near.node_visits <= c_ref.node_visits  # near-first prunes more
        if name != "cube":  # (the cube has a single node)
            assert c_near.node_visits < 0.85 * c_ref.node_visits
        assert c_ref.max_stack <= 64


def test_equal_t_ties_resolve_identically_in_both_orders(orc):
    pos, nrm, uv, h = raysets.grid_mesh()
    s = orc.Scene(orc.MeshScene(pos, nrm, uv))
    o, d = raysets.grid_tie_rays(h, 3000)
    rays = orc.abi.make_rays(o, d)
    bf, ties = s.brute_force_hit(rays, 0, 0.001)
    assert (ties > 1).sum() > 500  # the set really exercises ties
    ref, _ = s.closest_hit(rays, 0, 0.001, INF, 0)
    near, _ = s.closest_hit(rays, 0, 0.001, INF, 1)
    for f in ("t", "u", "v", "prim_id", "front_face"):
        assert np.array_equal(ref[f], near[f]), f
    # (the integer height field has flat leaves and axis-parallel rays, i.e. the strict-box holes
    # and 0*inf NaNs of the reference, so tree != brute force for some rays; what matters here is
    # that many genuine ties are resolved, and resolved identically)
    tied = (ties > 1) & (ref["t"] == bf["t"])
    assert tied.sum() > 300
    for i in np.nonzero(tied)[0][:200]:
        assert ref[i]["prim_id"] in s.tie_set(rays[i], 0).tolist()


def test_tree_shape_matches_survey(orc, mesh_scene):
    """5,461 nodes = (4^7-1)/3 and 16,384 leaves for both meshes."""
    for name, by_count in (("david", [0, 0, 2488, 13896, 0]), ("sycee", [0, 1126, 15258, 0, 0])):
        info = mesh_scene(name)[2].qbvh_info(0)
        assert (info.n_nodes, info.n_leaves, info.empty_children) == (5461, 16384, 0)
        assert list(info.leaves_by_count) == by_count
    info = mesh_scene("cube")[2].qbvh_info(0)
    assert (info.n_nodes, info.n_leaves, info.n_tris) == (1, 4, 12)
    assert list(info.leaves_by_count) == [0, 0, 0, 4, 0]


# ---- the reference's unit tests, restated ------------------------------------------------------
def test_push_hit_children_pushes_only_hit_lanes_in_order(orc):  # qbvh.rs:801-817
    stack = np.zeros(8, np.uint32)
    cursor = C.c_uint32(0)
    children = np.array([10, 20, 30, 40], np.uint32)
    order = np.array([2, 0, 3, 1], np.uint32)
    hits = np.array([1, 0, 1, 0], np.uint8)
    orc.lib().orc_push_hit_children(stack.ctypes.data, C.byref(cursor), children.ctypes.data, order.ctypes.data,
                                    hits.ctypes.data)
    assert cursor.value == 2 and stack[0] == 30 and stack[1] == 10


def _sanitize(orc, xyz):
    a, o = np.array(xyz, np.float64), np.zeros(3)
    orc.lib().orc_sanitize_sample_xyz(a.ctypes.data, o.ctypes.data)
    return o


def test_sanitize_sample_xyz_discards_non_finite_samples(orc):  # main.rs:809-818
    assert _sanitize(orc, [1.0, float("nan"), 0.5]).tolist() == [0, 0, 0]
    assert _sanitize(orc, [float("inf"), 1.0, 0.5]).tolist() == [0, 0, 0]


def test_sanitize_sample_xyz_clamps_fireflies_without_changing_chromaticity(orc):  # main.rs:821-828
    o = _sanitize(orc, [10.0, 40.0, 20.0])
    assert abs(o[1] - 20.0) < 1e-12 and abs(o[0] - 5.0) < 1e-12 and abs(o[2] - 10.0) < 1e-12
    assert _sanitize(orc, [1.0, 2.0, 3.0]).tolist() == [1.0, 2.0, 3.0]
    assert _sanitize(orc, [1.0, -2.0, 3.0]).tolist() == [1.0, -2.0, 3.0]


def test_gamma_correction(orc):  # color.rs:1988-2006
    a, o = np.array([-0.25, 0.18, -1.0]), np.zeros(3)
    orc.lib().orc_gamma_corrected(a.ctypes.data, o.ctypes.data)
    assert o[0] == 0.0 and o[2] == 0.0 and o[1] > 0.0 and np.isfinite(o).all()
    a = np.array([0.001, 0.002, 0.003])
    orc.lib().orc_gamma_corrected(a.ctypes.data, o.ctypes.data)
    assert np.allclose(o, [0.01292, 0.02584, 0.03876], atol=1e-6)


def test_clamp_display_channel(orc):  # main.rs:461-463
    f = orc.lib().orc_clamp_display_channel
    assert f(-1.0) == 0 and f(0.0) == 0 and f(0.5) == 128 and f(1.0) == 255 and f(7.0) == 255
    assert f(float("nan")) == 0


def test_spectral_pipeline_facts(orc):
    """white reflects ~0.9995; SF66 index 2.072 / 1.932 / 1.902;
    CIE lookups truncate to 1 nm and vanish outside [360, 831)."""
    white = np.array([1.0, 1.0, 1.0])
    r = orc.lib().orc_rgb_reflect(white.ctypes.data, 550.0)
    assert 0.999 < r < 1.0
    red = np.array([1.0, 0.0, 0.0])
    assert orc.lib().orc_rgb_reflect(red.ctypes.data, 650.0) > 0.9 > 0.1 > orc.lib().orc_rgb_reflect(red.ctypes.data, 450.0)
    # the bin index clamps for out-of-range wavelengths (color.rs:279-283)
    assert orc.lib().orc_rgb_reflect(red.ctypes.data, 100.0) == orc.lib().orc_rgb_reflect(red.ctypes.data, 360.0)
    assert orc.lib().orc_rgb_reflect(red.ctypes.data, 9999.0) == orc.lib().orc_rgb_reflect(red.ctypes.data, 719.0)
    m = orc.abi.Material()
    m.sellmeier_b[:] = [2.0245976, 0.470187196, 2.59970433]
    m.sellmeier_c[:] = [0.0147053225 * 1e6, 0.0692998276 * 1e6, 161.817601 * 1e6]
    n = [orc.lib().orc_sellmeier_index(C.byref(m), wl) for wl in (360.0, 550.0, 720.0)]
    assert np.allclose(n, [2.072, 1.932, 1.902], atol=1e-3)
    xyz = np.zeros(3)
    orc.lib().orc_xyz_from_wavelength(555.9, xyz.ctypes.data)
    a = xyz.copy()
    orc.lib().orc_xyz_from_wavelength(555.0, xyz.ctypes.data)
    assert np.array_equal(a, xyz) and xyz[1] > 0.99
    orc.lib().orc_xyz_from_wavelength(359.0, xyz.ctypes.data)
    assert xyz.tolist() == [0, 0, 0]
    # 1 nm Riemann sum of the Y table over the 360 nm sampling window ~ CIE_Y_INTEGRAL (color.rs:12)
    tot = 0.0
    for wl in range(360, 831):
        orc.lib().orc_xyz_from_wavelength(float(wl), xyz.ctypes.data)
        tot += xyz[1]
    assert abs(tot - 106.856895) < 1e-3


def test_closed_form_push_slots_equal_order_table():
    """k_traverse does not walk ORDER_TABLE (qbvh.rs:14-31); it computes every hit child's stack slot in closed
    form from the three sign bits (device_trace.cuh).  Check the formula against the table for all 8 x 16 cases."""
    ORDER_TABLE = [0x0123, 0x0132, 0x1023, 0x1032, 0x2301, 0x3201, 0x2310, 0x3210]
    for idx in range(8):
        T, L, R = (idx >> 2) & 1, (idx >> 1) & 1, idx & 1
        order = [(ORDER_TABLE[idx] >> (4 * j)) & 0xF for j in range(4)]
        for hitmask in range(16):
            h = [(hitmask >> k) & 1 for k in range(4)]
            want = {}  # child -> slot, by push_hit_children
            cursor = 0
            for i in order:
                if h[i]:
                    want[i] = cursor
                    cursor += 1
            nl, nr = h[0] + h[1], h[2] + h[3]
            bl, br = (0 if T else nr), (nl if T else 0)
            got = {}
            if h[0]: got[0] = bl + (0 if L else h[1])
            if h[1]: got[1] = bl + (h[0] if L else 0)
            if h[2]: got[2] = br + (0 if R else h[3])
            if h[3]: got[3] = br + (h[2] if R else 0)
            assert got == want, (idx, hitmask)
