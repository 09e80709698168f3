"""The C ABI from a plain C99 host (examples/yart_main.c): no Python, no torch in the process -- the way a Rust
`-sys` crate of the reference would use libyart_b200.so."""
import importlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "yet-another-raytracer_b200"


@pytest.fixture(scope="module")
def c_host(tmp_path_factory):
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = tmp_path_factory.mktemp("c_host") / "yart_main"
    cmd = ["gcc", "-std=c99", "-O2", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", str(ROOT / "include"),
           str(ROOT / "examples" / "yart_main.c"), "-o", str(exe), "-L", str(PKG), "-lyart_b200", "-Wl,-rpath," + str(PKG)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr  # yart.h is valid, warning-free C99
    return str(exe)


def run(exe, args, assets):
    return subprocess.run([exe] + args, capture_output=True, text=True, timeout=300, env=dict(os.environ, YART_ASSETS=assets))


def test_c_host_front_end_without_a_gpu(c_host, assets, yart):
    res = run(c_host, ["--info"], assets)
    assert res.returncode == 0, res.stderr
    assert "13 presets" in res.stdout and " david" in res.stdout
    assert "= 1920x1920" in res.stdout  # resolve_dimensions keeps the aspect (main.rs:845-865)
    assert "cube.obj: 12 triangles -> 1 nodes, 4 leaves" in res.stdout
    ldd = subprocess.run(["ldd", c_host], capture_output=True, text=True).stdout
    assert "libyart_b200" in ldd and "libtorch" not in ldd and "libpython" not in ldd


def test_c_host_fails_loudly_without_a_gpu(c_host, assets, yart, tmp_path):
    if yart.load_library().yart_device_count() > 0:
        pytest.skip("a GPU is present")
    res = run(c_host, ["cornell-box", "32", "32", "1", str(tmp_path / "x.ppm")], assets)
    assert res.returncode == 1 and "no CPU fallback" in res.stderr and not (tmp_path / "x.ppm").exists()


@pytest.mark.gpu
def test_c_host_renders_the_same_image_as_the_python_binding(c_host, assets, yart, ctx, tmp_path):
    out = tmp_path / "cornell.ppm"
    res = run(c_host, ["cornell-box", "64", "48", "5", str(out), "7"], assets)
    assert res.returncode == 0, res.stderr
    assert "15360 paths" in res.stdout
    raw = out.read_bytes()
    header = b"P6\n64 48\n255\n"
    assert raw.startswith(header)
    got = np.frombuffer(raw[len(header):], dtype=np.uint8).reshape(48, 64, 3)
    preset = yart.ScenePreset("cornell-box", seed=7)
    ctx.set_scene(preset)
    film, _ = ctx.render(preset.camera(64, 48), 64, 48, 0, 5, max_depth=50, seed=7)
    want = ctx.film_finalize(film, 5)[..., :3]
    assert np.array_equal(got, want)
