"""Random scene descriptions for the fuzz parity tests: every object kind, wrapper nesting, material, texture and light
list the ABI (include/yart.h) can express, in combinations the 13 presets never use.  The GPU path and the oracle read
the SAME description through the same structs, so whatever differs between them is a bug in one of the two."""
import ctypes as C
import math
import os

import numpy as np


class FuzzScene:
    """A random scene behind `.desc` (what `Context.set_scene` and `orc.Scene` take).  Keeps every array alive."""

    def __init__(self, pkg, seed, n_objects=None, allow_mesh=True, allow_media=True, allow_groups=True):
        abi = self.abi = pkg.abi
        r = self.rng = np.random.default_rng(seed)
        self.keep = []
        self.seed = seed

        # ---- textures ----
        self.perlins = [self._perlin() for _ in range(int(r.integers(1, 3)))]
        self.images = [self._image() for _ in range(int(r.integers(1, 3)))]
        textures = []
        for _ in range(int(r.integers(3, 9))):
            t = abi.Texture()
            t.kind = int(r.choice([abi.TEX_SOLID, abi.TEX_SOLID, abi.TEX_CHECKER, abi.TEX_NOISE, abi.TEX_IMAGE]))
            t.noise_type = int(r.integers(0, 5))
            t.perlin = int(r.integers(0, len(self.perlins)))
            t.image = int(r.integers(0, len(self.images)))
            for k in range(3):
                t.rgb_a[k] = float(r.uniform(0.05, 0.95))
                t.rgb_b[k] = float(r.uniform(0.05, 0.95))
            t.scale = float(r.uniform(0.2, 4.0))
            textures.append(t)
        bright = abi.Texture()
        bright.kind = abi.TEX_SOLID
        for k in range(3):
            bright.rgb_a[k] = float(r.uniform(4.0, 15.0))
        textures.append(bright)
        self.bright_tex = len(textures) - 1

        # ---- materials ----
        glass = self._preset_glass(pkg)
        materials = []
        for _ in range(int(r.integers(4, 10))):
            m = abi.Material()
            m.kind = int(r.choice([abi.MAT_LAMBERTIAN, abi.MAT_LAMBERTIAN, abi.MAT_METAL, abi.MAT_DIELECTRIC, abi.MAT_NONE]))
            m.texture = int(r.integers(0, len(textures) - 1))
            m.fuzz = float(r.choice([0.0, r.uniform(0.0, 1.0)]))
            if m.kind == abi.MAT_DIELECTRIC:
                for k in range(3):
                    m.sellmeier_b[k], m.sellmeier_c[k] = glass.sellmeier_b[k], glass.sellmeier_c[k]
            materials.append(m)
        light = abi.Material()
        light.kind, light.texture = abi.MAT_DIFFUSE_LIGHT, self.bright_tex
        materials.append(light)
        self.light_mat = len(materials) - 1
        iso = abi.Material()
        iso.kind, iso.texture = abi.MAT_ISOTROPIC, int(r.integers(0, len(textures) - 1))
        materials.append(iso)
        self.iso_mat = len(materials) - 1
        self.n_plain_materials = len(materials) - 2

        # ---- meshes (the cube asset: 12 triangles, the smallest mesh L4QBVH::new accepts) ----
        self.meshes = []
        if allow_mesh and r.random() < 0.6:
            self.meshes.append(pkg.TriangleMesh.from_obj(os.path.join(pkg.assets_dir(), "cube.obj")))

        # ---- groups ----
        groups = []
        if allow_groups and r.random() < 0.5:
            for _ in range(int(r.integers(1, 3))):
                members = []
                for _ in range(int(r.integers(1, 40))):
                    o = self._sphere() if r.random() < 0.5 else self._box()
                    o.material = int(r.integers(0, self.n_plain_materials))
                    members.append(o)
                arr = (abi.Object * len(members))(*members)
                self.keep.append(arr)
                g = abi.Group()
                g.members, g.n_members = C.cast(arr, C.POINTER(abi.Object)), len(members)
                groups.append(g)

        # ---- the world list ----
        n = int(n_objects if n_objects is not None else r.integers(3, 14))
        objects, lights = [], []
        for i in range(n):
            kinds = ["sphere", "sphere", "moving", "xy", "xz", "yz", "box", "tri"]
            if self.meshes:
                kinds += ["mesh", "mesh"]
            if groups:
                kinds.append("group")
            kind = str(r.choice(kinds))
            o = {"sphere": self._sphere, "moving": self._moving_sphere, "xy": lambda: self._rect(abi.OBJ_XY_RECT),
                 "xz": lambda: self._rect(abi.OBJ_XZ_RECT), "yz": lambda: self._rect(abi.OBJ_YZ_RECT), "box": self._box,
                 "tri": self._triangle, "mesh": self._mesh, "group": lambda: self._group(len(groups))}[kind]()
            o.material = int(r.integers(0, self.n_plain_materials))
            emissive = kind in ("sphere", "xz") and r.random() < 0.35
            if emissive:
                o.material = self.light_mat
            plain = abi.Object.from_buffer_copy(o)  # (the sampling-lights list holds the un-wrapped primitive)
            # wrappers, outermost first: ConstantMedium(Translate(RotateY(FlipFace(primitive))))
            if r.random() < 0.25:
                o.wrap |= abi.WRAP_FLIP_FACE
            if r.random() < 0.3:
                a = math.radians(float(r.uniform(-180, 180)))
                o.wrap |= abi.WRAP_ROTATE_Y
                o.sin_theta, o.cos_theta = math.sin(a), math.cos(a)
            if r.random() < 0.3:
                o.wrap |= abi.WRAP_TRANSLATE
                for k in range(3):
                    o.offset[k] = float(r.uniform(-2, 2))
            if allow_media and kind in ("sphere", "box") and not emissive and r.random() < 0.3:
                o.wrap |= abi.WRAP_MEDIUM
                o.neg_inv_density = -1.0 / float(r.uniform(0.05, 2.0))
                o.material = self.iso_mat
            objects.append(o)
            if emissive and not (o.wrap & (abi.WRAP_ROTATE_Y | abi.WRAP_TRANSLATE)) and r.random() < 0.8:
                lights.append(plain)
        if r.random() < 0.2:  # a light-list entry of a kind that cannot be sampled: the trait defaults (pdf 0, direction (1,0,0))
            lights.append(self._box())
        if r.random() < 0.3:
            lights.append(self._sphere())  # aiming at something that does not emit is legal too (cornell's glass sphere)
        r.shuffle(lights)

        sd = self.sd = abi.SceneDesc()
        for field, cls, items in (("objects", abi.Object, objects), ("lights", abi.Object, lights), ("groups", abi.Group, groups),
                                  ("materials", abi.Material, materials), ("textures", abi.Texture, textures),
                                  ("perlins", abi.Perlin, self.perlins), ("images", abi.Image, self.images)):
            arr = (cls * max(len(items), 1))(*items)
            self.keep.append(arr)
            setattr(sd, field, C.cast(arr, C.POINTER(cls)))
            setattr(sd, "n_" + field, len(items))
        if self.meshes:
            arr = (abi.Trimesh * len(self.meshes))(*[m.trimesh for m in self.meshes])
            self.keep.append(arr)
            sd.meshes, sd.n_meshes = C.cast(arr, C.POINTER(abi.Trimesh)), len(self.meshes)
        sky = r.random() < 0.6 or not any(o.material == self.light_mat for o in objects)  # (never a scene without any light)
        for k in range(3):
            sd.background_rgb[k] = float(r.uniform(0.3, 1.0)) if sky else 0.0
        self.desc = C.pointer(sd)
        self.n_objects, self.n_lights = len(objects), len(lights)

    # ---- pieces ----
    def _u(self, lo, hi):
        return float(self.rng.uniform(lo, hi))

    def _obj(self, kind):
        o = self.abi.Object()
        o.kind, o.cos_theta = kind, 1.0
        return o

    def _sphere(self):
        o = self._obj(self.abi.OBJ_SPHERE)
        o.p[0], o.p[1], o.p[2] = self._u(-5, 5), self._u(-5, 5), self._u(-5, 5)
        o.p[3] = self._u(0.3, 2.0) * (-1.0 if self.rng.random() < 0.1 else 1.0)  # a negative radius = inward normals (sphere.rs)
        return o

    def _moving_sphere(self):
        o = self._obj(self.abi.OBJ_MOVING_SPHERE)
        for k in range(3):
            o.p[k] = self._u(-5, 5)
            o.p[3 + k] = o.p[k] + self._u(-1, 1)
        o.p[6], o.p[7], o.p[8] = 0.0, 1.0, self._u(0.3, 1.5)
        return o

    def _rect(self, kind):
        o = self._obj(kind)
        a0, b0 = self._u(-5, 3), self._u(-5, 3)
        o.p[0], o.p[1], o.p[2], o.p[3], o.p[4] = a0, a0 + self._u(0.5, 4), b0, b0 + self._u(0.5, 4), self._u(-5, 5)
        return o

    def _box(self):
        o = self._obj(self.abi.OBJ_BOX)
        for k in range(3):
            o.p[k] = self._u(-5, 3)
            o.p[3 + k] = o.p[k] + self._u(0.3, 3)
        return o

    def _triangle(self):
        o = self._obj(self.abi.OBJ_TRIANGLE)
        v = self.rng.uniform(-5, 5, size=(3, 3))
        v[1:] = v[0] + self.rng.uniform(-3, 3, size=(2, 3))
        nrm = np.cross(v[1] - v[0], v[2] - v[0])
        nrm = nrm / (np.linalg.norm(nrm) + 1e-300)
        vals = list(v.reshape(-1)) + list(np.tile(nrm, 3) + self.rng.uniform(-0.1, 0.1, size=9)) + list(self.rng.uniform(0, 1, size=6))
        for k, x in enumerate(vals):
            o.p[k] = float(x)
        return o

    def _mesh(self):
        o = self._obj(self.abi.OBJ_MESH)
        o.index = 0
        return o

    def _group(self, n_groups):
        o = self._obj(self.abi.OBJ_GROUP)
        o.index = int(self.rng.integers(0, n_groups))
        return o

    def _perlin(self):
        p = self.abi.Perlin()
        r = self.rng
        for i in range(256):
            p.ranfloat[i] = float(r.random())
            v = r.uniform(-1, 1, size=3)
            v = v / (np.linalg.norm(v) + 1e-300)
            for k in range(3):
                p.ranvec[i][k] = float(v[k])
        for name in ("perm_x", "perm_y", "perm_z"):
            perm = r.permutation(256)
            arr = getattr(p, name)
            for i in range(256):
                arr[i] = int(perm[i])
        return p

    def _image(self):
        w, h = int(self.rng.integers(1, 9)), int(self.rng.integers(1, 9))
        px = (C.c_uint8 * (w * h * 3))(*[int(x) for x in self.rng.integers(0, 256, size=w * h * 3)])
        self.keep.append(px)
        im = self.abi.Image()
        im.rgb8, im.width, im.height = C.cast(px, C.POINTER(C.c_uint8)), w, h
        return im

    _glass = None

    @classmethod
    def _preset_glass(cls, pkg):
        if cls._glass is None:
            p = pkg.ScenePreset("cornell-box")
            sd = p.desc.contents
            for i in range(sd.n_materials):
                if sd.materials[i].kind == pkg.abi.MAT_DIELECTRIC:
                    cls._glass = pkg.abi.Material.from_buffer_copy(sd.materials[i])
            assert cls._glass is not None
        return cls._glass

    def camera(self, width, height):
        cam = self.abi.Camera()
        r = self.rng
        src = [self._u(-14, 14), self._u(-6, 10), -18.0 + self._u(-3, 3)]
        for k in range(3):
            cam.lookfrom[k], cam.lookat[k], cam.vup[k] = src[k], self._u(-1, 1), (0.0, 1.0, 0.0)[k]
        cam.vfov_degrees, cam.aspect_ratio = self._u(25, 60), width / height
        cam.aperture, cam.focus_dist, cam.time0, cam.time1 = float(r.choice([0.0, 0.2])), 10.0, 0.0, 1.0
        return cam

    def rays(self, n):
        """Rays from a shell around the scene towards its inside, plus axis-parallel ones (rect / box edge cases)."""
        r = self.rng
        d = r.normal(size=(n, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        o = -d * r.uniform(0, 14, size=(n, 1)) + r.uniform(-4, 4, size=(n, 3))
        d *= r.uniform(0.1, 12, size=(n, 1))  # the reference never normalises directions
        k = n // 8
        axis = r.integers(0, 3, size=k)
        d[:k] = 0.0
        d[np.arange(k), axis] = r.choice([-1.0, 1.0], size=k) * r.uniform(0.5, 3, size=k)
        return o, d


def fuzz_soup(seed, n_tris):
    """A random triangle soup (positions f32, normals f64, uvs f32) with the awkward cases mixed in: lattice vertices
    (equal centroids, exact ties, flat boxes), degenerate and sliver triangles, axis-aligned triangles, duplicated
    triangles, far-away clusters and signed zeros."""
    r = np.random.default_rng(seed)
    style = seed % 4
    if style == 0:
        pos = r.uniform(-10, 10, size=(n_tris, 1, 3)) + r.uniform(-1.5, 1.5, size=(n_tris, 3, 3))
    elif style == 1:
        pos = r.integers(-4, 5, size=(n_tris, 3, 3)).astype(np.float64)
    elif style == 2:
        pos = r.uniform(-3, 3, size=(n_tris, 1, 3)) + r.uniform(-0.4, 0.4, size=(n_tris, 3, 3))
        far = r.random(n_tris) < 0.2
        pos[far] += np.array([4096.0, -2048.0, 1e4])
    else:
        pos = r.uniform(-6, 6, size=(n_tris, 1, 3)) + r.uniform(-2, 2, size=(n_tris, 3, 3))
        flat = r.random(n_tris) < 0.3                       # axis-aligned: one coordinate shared by the three vertices
        ax = r.integers(0, 3, size=n_tris)
        idx = np.flatnonzero(flat)
        pos[idx, :, ax[idx]] = np.round(pos[idx, 0, ax[idx]])[:, None]
    k = max(1, n_tris // 12)
    deg = r.choice(n_tris, size=k, replace=False)
    pos[deg, 2] = pos[deg, 1]                                # degenerate: two equal vertices
    dup = r.choice(n_tris, size=k, replace=False)
    pos[dup] = pos[r.choice(n_tris, size=k)]                 # exact duplicates (ties between different triangles)
    pos[r.random(pos.shape) < 0.02] = -0.0
    pos = pos.astype(np.float32)
    nrm = r.standard_normal((n_tris, 3, 3))
    uv = r.random((n_tris, 3, 2)).astype(np.float32)
    return pos, nrm, uv


def fuzz_mesh_rays(seed, pos, n):
    """Rays for a soup: random ones, rays through vertices and edge midpoints (ties), axis-parallel rays (zero
    components: NaN slabs), rays starting ON a vertex, denormal and huge direction components."""
    r = np.random.default_rng(seed + 77)
    p = pos.astype(np.float64)
    lo, hi = p.reshape(-1, 3).min(0), p.reshape(-1, 3).max(0)
    span = np.maximum(hi - lo, 1.0)
    o = r.uniform(lo - 0.5 * span, hi + 0.5 * span, size=(n, 3))
    d = r.standard_normal((n, 3)) * r.uniform(0.01, 20, size=(n, 1))
    q = n // 6
    tri = r.integers(0, len(p), size=3 * q)
    vert = p[tri, r.integers(0, 3, size=3 * q)]
    d[:q] = (vert[:q] - o[:q]) * r.choice([0.5, 1.0, 2.0], size=(q, 1))                    # through a vertex
    mid = 0.5 * (p[tri[q:2 * q], 0] + p[tri[q:2 * q], 1])
    d[q:2 * q] = mid - o[q:2 * q]                                                          # through an edge midpoint
    o[2 * q:3 * q] = vert[2 * q:3 * q]                                                     # starting on a vertex
    a = r.integers(0, 3, size=q)
    d[3 * q:4 * q] = 0.0
    d[np.arange(3 * q, 4 * q), a] = r.choice([-1.0, 1.0, 3.5], size=q)                     # axis-parallel
    o[3 * q:4 * q] = vert[:q] + 0.0
    o[np.arange(3 * q, 4 * q), a] -= 5.0 * np.sign(d[np.arange(3 * q, 4 * q), a])
    w = slice(4 * q, 4 * q + q // 2)
    d[w, r.integers(0, 3)] = r.choice([5e-324, -1e-310, 1e-200, 1e200], size=q // 2)       # denormal / huge components
    return o, d
