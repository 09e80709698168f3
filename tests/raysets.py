"""Deterministic ray sets shared by the CPU and GPU tests and by bench.py.

`uniform` and `axis` follow the reference's bench generators (qbvh.rs:973-986 + vec3.rs:214-220,
and qbvh.rs:949-971): origins uniform in the mesh AABB; directions uniform on the sphere, or one
of +x/+y/+z.  numpy's Philox replaces thread_rng.
"""
import numpy as np

SEED_UNIFORM = 0x5EED0001
SEED_AXIS = 0x5EED0002


def _rng(seed):
    return np.random.Generator(np.random.Philox(seed))


def uniform(n, bbox_min, bbox_max, seed=SEED_UNIFORM):
    g = _rng(seed)
    lo, hi = np.asarray(bbox_min, dtype=np.float64), np.asarray(bbox_max, dtype=np.float64)
    o = lo + (hi - lo) * g.random((n, 3))
    a = g.random(n) * (2.0 * np.pi)
    z = -1.0 + 2.0 * g.random(n)
    r = np.sqrt(1.0 - z * z)
    d = np.stack([r * np.cos(a), r * np.sin(a), z], axis=1)
    return o, d


def axis(n, bbox_min, bbox_max, seed=SEED_AXIS):
    g = _rng(seed)
    lo, hi = np.asarray(bbox_min, dtype=np.float64), np.asarray(bbox_max, dtype=np.float64)
    o = lo + (hi - lo) * g.random((n, 3))
    d = np.eye(3)[g.integers(0, 3, n)]
    return o, d


def grid_mesh(nx=24, ny=24, seed=7):
    """An integer height field: every coordinate is a small integer, so Moller-Trumbore is exact
    and rays aimed at grid vertices produce genuine equal-t ties between up to six triangles."""
    g = _rng(seed)
    h = g.integers(0, 4, (nx + 1, ny + 1)).astype(np.float64)
    tris = []
    for i in range(nx):
        for j in range(ny):
            p00, p10 = (i, j, h[i, j]), (i + 1, j, h[i + 1, j])
            p01, p11 = (i, j + 1, h[i, j + 1]), (i + 1, j + 1, h[i + 1, j + 1])
            tris.append((p00, p10, p11))
            tris.append((p00, p11, p01))
    pos = np.array(tris, dtype=np.float32)
    nrm = np.zeros_like(pos, dtype=np.float64)
    nrm[..., 2] = 1.0
    uv = np.zeros((pos.shape[0], 3, 2), dtype=np.float32)
    return pos, nrm, uv, h


def grid_tie_rays(h, n, seed=11):
    """Rays that pass exactly through grid vertices: origin = vertex + 8*k, direction = -k with
    small-integer k, so t = 8 exactly for every triangle sharing the vertex."""
    g = _rng(seed)
    nx, ny = h.shape[0] - 1, h.shape[1] - 1
    i = g.integers(1, nx, n)
    j = g.integers(1, ny, n)
    k = np.stack([g.integers(-2, 3, n), g.integers(-2, 3, n), g.integers(1, 4, n)], axis=1).astype(np.float64)
    v = np.stack([i, j, h[i, j]], axis=1).astype(np.float64)
    return v + 8.0 * k, -k


def pixel_centre_rays(cam, width, height):
    """One primary ray through the centre of every pixel, no lens, no jitter -- Camera::new / get_ray
    (camera.rs:41-94) restated in numpy with s = (x + 0.5) / (W - 1), t = 1 - (y + 0.5) / (H - 1)
    (the render loop's own mapping with the jitter at 0.5, main.rs:693-696).  `cam` is a yart_camera."""
    lookfrom, lookat, vup = (np.array(list(v), dtype=np.float64) for v in (cam.lookfrom, cam.lookat, cam.vup))
    theta = cam.vfov_degrees * np.pi / 180.0
    vh = 2.0 * np.tan(theta / 2.0)
    vw = cam.aspect_ratio * vh
    w = lookfrom - lookat
    w /= np.linalg.norm(w)
    u = np.cross(vup, w)
    u /= np.linalg.norm(u)
    v = np.cross(w, u)
    hor, ver = cam.focus_dist * vw * u, cam.focus_dist * vh * v
    llc = lookfrom - hor / 2.0 - ver / 2.0 - cam.focus_dist * w
    ys, xs = np.mgrid[0:height, 0:width]
    s = ((xs + 0.5) / (width - 1)).reshape(-1, 1)
    t = (1.0 - (ys + 0.5) / (height - 1)).reshape(-1, 1)
    d = llc + s * hor + t * ver - lookfrom
    o = np.broadcast_to(lookfrom, d.shape).copy()
    return o, d
