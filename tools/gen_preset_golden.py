#!/usr/bin/env python3
"""An INDEPENDENT description of the reference's 13 scene presets, parsed out of the reference's own source text.

    python tools/gen_preset_golden.py        (needs /root/reference; run in the dev container only)

Reads /root/reference/raytracer/src/{scenes.rs, main.rs, material.rs}, interprets the small subset of Rust the
scene builders are written in (let bindings, `objects.add_object(...)`, constructor calls, struct literals,
`.clone()`, `Arc::new`, Vec3 addition) and writes tests/golden/presets.json.  tests/test_presets_golden.py
then checks that every object / material / texture / light / camera / default field that
`yart_preset_build` (csrc/host_presets.cpp) produces equals what the reference's text says -- so a wrong
constant in host_presets.cpp fails a test instead of changing the GPU path and the oracle identically
(both consume yart_preset_build's output).

Nothing of host_presets.cpp is consulted here, and no reference source is copied into the repository: the
output is data (numbers and kind names).  The two builders that draw from `thread_rng` (random_scene,
the_next_week_final_scene) are described by their deterministic statements plus the PARAMETERS of their random
loops (grid ranges, radii, thresholds, value ranges), which the test checks as properties of the seeded scenes.
"""
import json
import math
import re
import sys
from pathlib import Path

REF = Path("/root/reference/raytracer/src")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "presets.json"


# ---------------------------------------------------------------------------------------------------------
# tokenizer / expression parser for the Rust subset
# ---------------------------------------------------------------------------------------------------------
def strip_comments(src):
    return re.sub(r"//[^\n]*", "", src)


TOKEN = re.compile(r"""\s*(?:
    (?P<num>\d+\.\d+(?:e-?\d+)?|\d+\.(?!\.)|\d+(?:e-?\d+)?)
  | (?P<str>"[^"]*")
  | (?P<id>[A-Za-z_][A-Za-z0-9_]*(?:::[A-Za-z_][A-Za-z0-9_]*)*)
  | (?P<op>\.\.|[()\[\]{},;:.+\-*/=&<>!])
)""", re.X)


def tokenize(text):
    pos, out = 0, []
    text = text.rstrip()
    while pos < len(text):
        m = TOKEN.match(text, pos)
        if not m:
            raise SyntaxError("cannot tokenize at %r" % text[pos:pos + 40])
        pos = m.end()
        for kind in ("num", "str", "id", "op"):
            if m.group(kind) is not None:
                out.append((kind, m.group(kind)))
                break
    return out


class Parser:
    """Evaluates expressions straight to Python values (lists for Vec3/RGB, dicts for scene items)."""

    def __init__(self, tokens, env):
        self.t, self.i, self.env = tokens, 0, env

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else ("eof", "")

    def take(self, val=None):
        tok = self.peek()
        if val is not None and tok[1] != val:
            raise SyntaxError("expected %r, got %r (token %d)" % (val, tok, self.i))
        self.i += 1
        return tok

    def args(self, close=")"):
        out = []
        while self.peek()[1] != close:
            out.append(self.expr())
            if self.peek()[1] == ",":
                self.take()
        self.take(close)
        return out

    def expr(self):
        v = self.term()
        while self.peek()[1] in ("+", "-", "*"):
            op = self.take()[1]
            r = self.term()
            v = binop(op, v, r)
        return v

    def term(self):
        kind, val = self.peek()
        if val == "-":
            self.take()
            return binop("neg", self.term(), None)
        if val == "&":
            self.take()
            if self.peek()[1] == "mut":
                self.take()
            return self.term()
        if val == "(":
            self.take()
            items = self.args(")")
            v = items[0] if len(items) == 1 else items
            return self.postfix(v)
        if val == "[":
            self.take()
            return self.postfix(self.args("]"))
        if kind == "num":
            self.take()
            return self.postfix(float(val) if ("." in val or "e" in val) else int(val))
        if kind == "str":
            self.take()
            return self.postfix(val[1:-1])
        if kind == "id":
            self.take()
            if self.peek()[1] == "(":
                self.take()
                return self.postfix(construct(val, self.args(")")))
            if self.peek()[1] == "{" and val[0].isupper():  # struct literal
                self.take()
                fields = {}
                while self.peek()[1] != "}":
                    name = self.take()[1]
                    self.take(":")
                    fields[name] = self.expr()
                    if self.peek()[1] == ",":
                        self.take()
                self.take("}")
                return self.postfix(construct_struct(val, fields))
            if val in self.env:
                return self.postfix(self.env[val])
            return self.postfix(constant(val))
        raise SyntaxError("unexpected token %r" % (self.peek(),))

    def postfix(self, v):
        while self.peek()[1] == ".":
            name = self.peek(1)[1]
            if name in ("clone", "unwrap", "size"):
                self.take(); self.take(); self.take("("); self.take(")")
                if name == "size":
                    v = {"size_of": v}
            elif name == "objects":
                self.take(); self.take()
            else:
                break
        if self.peek()[1] == "as":  # `x as f64`
            self.take(); self.take()
        return v


def binop(op, a, b):
    if op == "neg":
        return [-x for x in a] if isinstance(a, list) else -a
    if isinstance(a, list) and isinstance(b, list):
        return [binop(op, x, y) for x, y in zip(a, b)]
    if isinstance(a, list):
        return [binop(op, x, b) for x in a]
    return {"+": a + b, "-": a - b, "*": a * b}[op]


GLASS = {}


def constant(name):
    if name in GLASS:
        return {"kind": "dielectric", "glass": name, "b": GLASS[name]["b"], "c": GLASS[name]["c"]}
    if name == "NoMaterial":
        return {"kind": "none"}
    if name.startswith("NoiseType::"):
        return name.split("::")[1].lower()
    raise NameError("unknown identifier %s" % name)


def f(x):
    return float(x)


def construct(name, a):
    if name in ("Arc::new", "Path::new"):
        return a[0]
    if name in ("Vec3::new", "RGB::new"):
        return [f(x) for x in a]
    if name == "RGB::default":
        return [0.0, 0.0, 0.0]
    if name == "Vec3::random":
        return {"random_vec3": [f(a[0]), f(a[1])]}
    if name == "HittableList::new":
        return []
    if name == "SolidColor::new":
        return {"kind": "solid", "rgb": a[0]}
    if name == "CheckerTexture::new":
        return {"kind": "checker", "odd": a[0], "even": a[1]}
    if name == "NoiseTexture::new":
        return {"kind": "noise", "noise_type": a[0], "scale": f(a[1])}
    if name == "ImageTexture::new":
        return {"kind": "image", "path": a[0]}
    if name == "Lambertian::new":
        return {"kind": "lambertian", "texture": a[0]}
    if name == "Metal::new":
        return {"kind": "metal", "texture": a[0], "fuzz": f(a[1])}
    if name == "DiffuseLight::new":
        return {"kind": "diffuse_light", "texture": a[0]}
    if name == "StillSphere::new":
        return {"kind": "sphere", "center": a[0], "radius": f(a[1]), "material": a[2]}
    if name == "MovingSphere::new":
        return {"kind": "moving_sphere", "center0": a[0], "center1": a[1], "time0": f(a[2]), "time1": f(a[3]),
                "radius": f(a[4]), "material": a[5]}
    if name in ("XYRect::new", "XZRect::new", "YZRect::new"):
        return {"kind": name[:2].lower() + "_rect", "a0": f(a[0]), "a1": f(a[1]), "b0": f(a[2]), "b1": f(a[3]),
                "k": f(a[4]), "material": a[5]}
    if name == "BoxEntity::new":
        return {"kind": "box", "p0": a[0], "p1": a[1], "material": a[2]}
    if name == "Translate::new":
        return {"kind": "translate", "inner": a[0], "offset": a[1]}
    if name == "RotateY::new":
        return {"kind": "rotate_y", "inner": a[0], "angle": f(a[1])}
    if name == "FlipFace::new":
        return {"kind": "flip_face", "inner": a[0]}
    if name == "ConstantMedium::new":
        return {"kind": "constant_medium", "boundary": a[0], "density": f(a[1]), "texture": a[2]}
    if name == "TriangleMesh::from_obj":
        return {"kind": "mesh", "path": a[0], "material": a[1]}
    if name == "BVHNode::new":
        return {"kind": "bvh", "members": a[0], "time0": f(a[3]), "time1": f(a[4])}
    raise NameError("unknown constructor %s" % name)


def construct_struct(name, fields):
    if name == "Triangle":
        return {"kind": "triangle", "vertices": fields["vertices"], "normals": fields["normals"],
                "uv": [[f(x) for x in p] for p in fields["uv"]], "material": fields["material"]}
    raise NameError("unknown struct %s" % name)


# ---------------------------------------------------------------------------------------------------------
# statements
# ---------------------------------------------------------------------------------------------------------
def split_statements(body):
    """Top-level statements of a block: `...;` runs and `for ... { ... }` blocks."""
    out, depth, cur, i = [], 0, "", 0
    while i < len(body):
        ch = body[i]
        if depth == 0 and re.match(r"for\b", body[i:]) and not cur.strip():
            j = body.index("{", i)
            d, k = 1, j + 1
            while d:
                d += {"{": 1, "}": -1}.get(body[k], 0)
                k += 1
            out.append(("for", body[i:k]))
            i = k
            continue
        cur += ch
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        elif ch == ";" and depth == 0:
            out.append(("stmt", cur.strip()[:-1].strip()))
            cur = ""
        i += 1
    if cur.strip():
        out.append(("stmt", cur.strip()))
    return out


def run_builder(body, result_names=("objects", "world")):
    env, loops = {}, []
    for kind, text in split_statements(body):
        if kind == "for":
            # a loop over thread_rng draws: leave a placeholder where its objects go
            target = re.search(r"(\w+)\.add_object", text).group(1)
            env[target].append({"kind": "random_loop", "loop": len(loops)})
            loops.append(text)
            continue
        m = re.match(r"let\s+(?:mut\s+)?(\w+)\s*(?::\s*[\w<>]+\s*)?=\s*(.*)$", text, re.S)
        if m:
            if "thread_rng" in m.group(2):
                continue
            env[m.group(1)] = Parser(tokenize(m.group(2)), env).expr()
            continue
        m = re.match(r"(\w+)\.add_object\((.*)\)$", text, re.S)
        if m:
            env[m.group(1)].append(Parser(tokenize(m.group(2)), env).expr())
            continue
        m = re.match(r"(\w+)\s*=\s*(.*)$", text, re.S)
        if m:
            env[m.group(1)] = Parser(tokenize(m.group(2)), env).expr()
            continue
        m = re.match(r"(?:return\s+)?(\w+)$", text)
        if m and m.group(1) in result_names:
            return env[m.group(1)], loops, env
        raise SyntaxError("statement not understood: %r" % text[:80])
    raise SyntaxError("builder returned nothing")


def function_bodies(src):
    out = {}
    for m in re.finditer(r"pub fn (\w+)\(\)\s*->\s*HittableList\s*\{", src):
        d, k = 1, m.end()
        while d:
            d += {"{": 1, "}": -1}.get(src[k], 0)
            k += 1
        out[m.group(1)] = src[m.end():k - 1]
    return out


def num(pattern, text, cast=float):
    m = re.search(pattern, text, re.S)
    if not m:
        raise SyntaxError("pattern %r not found" % pattern)
    return [cast(g) for g in m.groups()] if m.lastindex and m.lastindex > 1 else cast(m.group(1))


def random_scene_params(loops):
    t = loops[0]
    return {
        "grid": num(r"for a in (-?\d+)\.\.(-?\d+)", t, int), "grid_b": num(r"for b in (-?\d+)\.\.(-?\d+)", t, int),
        "jitter": num(r"a as f64 \+ ([\d.]+) \* rng", t), "y": num(r"a as f64 \+ [\d.]+ \* rng\.gen::<f64>\(\),\s*([\d.]+),", t),
        "keep_out_center": num(r"center - Vec3::new\(([\d.]+), ([\d.]+), ([\d.]+)\)", t), "keep_out_dist": num(r"\.length\(\) > ([\d.]+)", t),
        "p_lambertian": num(r"if choose_mat < ([\d.]+)", t), "p_metal": num(r"else if choose_mat < ([\d.]+)", t),
        "lambertian_albedo_range": num(r"RGB::new\(\s*rng\.gen_range\((-?[\d.]+)\.\.([\d.]+)\)", t),
        "metal_albedo_range": num(r"choose_mat < 0\.95.*?gen_range\(([\d.]+)\.\.([\d.]+)\)", t),
        "metal_fuzz_range": num(r"fuzz: f64 = rng\.gen_range\(([\d.]+)\.\.([\d.]+)\)", t),
        "radius": num(r"StillSphere::new\(center, ([\d.]+), sphere_material", t),
    }


def next_week_params(loops):
    b1, b2 = loops[0], loops[1]
    return {
        "boxes_per_side": num(r"for i in 0\.\.(\d+)", b1, int), "box_width": num(r"let w = ([\d.]+)", b1),
        "box_origin": num(r"let x0: f64 = (-?[\d.]+) \+", b1), "box_y0": num(r"let y0: f64 = ([\d.]+)", b1),
        "box_y1_range": num(r"let y1: f64 = rng\.gen_range\(([\d.]+)\.\.([\d.]+)\)", b1),
        "box_material": Parser(tokenize("Lambertian::new(SolidColor::new(RGB::new(0.48, 0.83, 0.53)))"), {}).expr(),
        "n_spheres": num(r"for _ in 0\.\.(\d+)", b2, int), "sphere_center_range": num(r"Vec3::random\(([\d.]+), ([\d.]+)\)", b2),
        "sphere_radius": num(r"Vec3::random\([\d., ]+\),\s*([\d.]+),", b2),
    }


# ---------------------------------------------------------------------------------------------------------
def parse_main(main_src, builders):
    m = re.search(r"fn build_scene_preset.*?\{(.*?)\n\}\n", main_src, re.S)
    body = m.group(1)
    d = re.search(r"let mut defaults = RenderDefaults \{(.*?)\};", body, re.S).group(1)
    base = {k: (float(v) if "." in v else int(v)) for k, v in re.findall(r"(\w+):\s*([\d.]+)", d)}
    presets = {}
    for arm in re.finditer(r"SceneName::(\w+) => \{(.*?)\n        \}", body, re.S):
        variant, text = arm.group(1), arm.group(2)
        name = re.sub(r"(?<!^)([A-Z])", r"-\1", variant).lower()  # clap ValueEnum: kebab-case
        p = {"defaults": dict(base), "lights": []}
        env = {"lights": p["lights"]}
        for kind, st in split_statements(text):
            mm = re.match(r"world = Arc::new\((\w+)\(\)\)$", st)
            if mm:
                p["world_fn"] = mm.group(1)
                continue
            mm = re.match(r"defaults\.(\w+) = ([\d.]+)$", st)
            if mm:
                p["defaults"][mm.group(1)] = float(mm.group(2)) if "." in mm.group(2) else int(mm.group(2))
                continue
            mm = re.match(r"lights\.add_object\((.*)\)$", st, re.S)
            if mm:
                p["lights"].append(Parser(tokenize(mm.group(1)), env).expr())
                continue
            mm = re.match(r"(background|lookfrom|lookat|output_filename) = (.*)$", st, re.S)
            if mm:
                p[mm.group(1)] = Parser(tokenize(mm.group(2)), env).expr()
                continue
            raise SyntaxError("preset statement not understood: %r" % st[:80])
        world, loops, _ = builders[p["world_fn"]]
        p["world"] = world
        presets[name] = p
    return presets


def main():
    mat_src = strip_comments((REF / "material.rs").read_text())
    for m in re.finditer(r"pub static (\w+): Dielectric = Dielectric \{(.*?)\};", mat_src, re.S):
        vals = {k: Parser(tokenize(v), {}).expr() for k, v in re.findall(r"(\w\d):\s*([^,]+),", m.group(2))}
        GLASS[m.group(1)] = {"b": [vals["b1"], vals["b2"], vals["b3"]], "c": [vals["c1"], vals["c2"], vals["c3"]]}
    scenes_src = strip_comments((REF / "scenes.rs").read_text())
    builders = {}
    random_params = {}
    for name, body in function_bodies(scenes_src).items():
        world, loops, env = run_builder(body)
        builders[name] = (world, loops, env)
        if name == "random_scene":
            random_params["random-scene"] = random_scene_params(loops)
        if name == "the_next_week_final_scene":
            random_params["next-week-final"] = next_week_params(loops)
    main_src = strip_comments((REF / "main.rs").read_text())
    presets = parse_main(main_src, builders)
    for k, v in random_params.items():
        presets[k]["random"] = v
    # the camera arguments render() adds (main.rs:605-625)
    cam = re.search(r"let vup: Vec3 = Vec3::new\(([\d.]+), ([\d.]+), ([\d.]+)\);\s*let dist_to_focus(?:: f64)? = ([\d.]+);", main_src)
    t = re.search(r"dist_to_focus,\s*([\d.]+),\s*([\d.]+),\s*\)", main_src)
    import hashlib
    from PIL import Image
    assets = {}
    for rel in ("input/cube.obj", "input/david.obj", "input/sycee.obj"):
        data = (REF.parent.parent / rel).read_bytes()
        n_tris = sum(len(line.split()) - 3 for line in data.decode().splitlines() if line.startswith("f "))
        assets[rel] = {"sha256": hashlib.sha256(data).hexdigest(), "n_tris": n_tris}
    im = Image.open(REF.parent.parent / "input/earthmap.jpg").convert("RGB")
    assets["input/earthmap.jpg"] = {"width": im.size[0], "height": im.size[1],
                                    "sha256_rgb8_pil": hashlib.sha256(im.tobytes()).hexdigest()}
    assets["missing"] = (REF.parent.parent / ".MISSING_LARGE_BLOBS").read_text().split() if (REF.parent.parent / ".MISSING_LARGE_BLOBS").exists() else []
    doc = {"assets": assets, "source": "parsed from the reference's raytracer/src/{scenes.rs,main.rs,material.rs} by tools/gen_preset_golden.py",
           "camera": {"vup": [float(x) for x in cam.groups()[:3]], "focus_dist": float(cam.group(4)),
                      "time0": float(t.group(1)), "time1": float(t.group(2))},
           "presets": presets}
    OUT.write_text(json.dumps(doc, indent=1, sort_keys=True) + "\n")
    print("wrote", OUT, "presets:", ", ".join("%s(%d objects, %d lights)" % (k, len(v["world"]), len(v["lights"]))
                                              for k, v in presets.items()))


if __name__ == "__main__":
    main()
