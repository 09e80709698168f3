"""The native command-line host (yet-another-raytracer_b200/host/yart_main.cpp -> `yart`): the reference's compiled
binary front end (clap `Cli`, main.rs:78-107; resolve_render_options, main.rs:188-209) on the C ABI alone -- no
Python, no torch in the process.  CPU tests restate the reference's own CLI unit tests (main.rs:831-915) against
`--dry-run`; the GPU test compares its PNG with the Python front end's."""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
EXE = ROOT / "yet-another-raytracer_b200" / "yart"


def run(args, assets, **kw):
    return subprocess.run([str(EXE)] + args + ["--assets", assets], capture_output=True, text=True, timeout=600, **kw)


def test_native_host_is_built_without_python_or_torch(yart):
    assert EXE.exists(), "build.py builds it next to libyart_b200.so"
    ldd = subprocess.run(["ldd", str(EXE)], capture_output=True, text=True).stdout
    assert "libyart_b200" in ldd and "libtorch" not in ldd and "libpython" not in ldd and "libnccl" not in ldd


def test_cli_accepts_named_scene_values_and_rejects_unknown(yart, assets):  # main.rs:831-842
    names = run(["--list-scenes"], assets).stdout.split()
    assert names == yart.SCENE_NAMES
    ok = run(["--scene", "david", "--dry-run"], assets)
    assert ok.returncode == 0
    bad = run(["--scene", "unknown-scene", "--dry-run"], assets)
    assert bad.returncode == 2 and "invalid value 'unknown-scene'" in bad.stderr and "possible values" in bad.stderr
    assert run(["--dry-run"], assets).returncode == 2                      # --scene is required
    for flag in ("--samples", "--max-depth", "--workers", "--width", "--height"):   # parse_positive_usize / range(1..)
        r = run(["--scene", "david", flag, "0", "--dry-run"], assets)
        assert r.returncode == 2 and "greater than 0" in r.stderr, flag
        r = run(["--scene", "david", flag, "abc", "--dry-run"], assets)
        assert r.returncode == 2 and "invalid integer" in r.stderr, flag
    assert run(["--scene", "david", "--bogus"], assets).returncode == 2


def test_resolve_render_options_like_the_reference(yart, assets):  # main.rs:868-915
    o = json.loads(run(["--scene", "david", "--dry-run"], assets).stdout)
    assert o["output_path"] == os.path.join("output", "david.png")       # default_output_path
    assert (o["width"], o["height"], o["samples_per_pixel"], o["max_depth"], o["workers"]) == (600, 600, 10000, 50, 30)
    o = json.loads(run(["--scene", "two-spheres", "--output", "custom/output.png", "--width", "600", "--samples", "32",
                        "--max-depth", "12", "--workers", "8", "--vfov", "45.0", "--aperture=0.25", "--dry-run"], assets).stdout)
    assert o == {"output_path": "custom/output.png", "width": 600, "height": 400, "samples_per_pixel": 32, "max_depth": 12,
                 "workers": 8, "vfov": 45.0, "aperture": 0.25}
    o = json.loads(run(["--scene", "two-spheres", "--height", "400", "--dry-run"], assets).stdout)   # main.rs:855-858
    assert (o["width"], o["height"]) == (600, 400)
    o = json.loads(run(["--scene", "two-spheres", "--width", "1024", "--height", "512", "--dry-run"], assets).stdout)
    assert (o["width"], o["height"]) == (1024, 512)
    # every preset's defaults equal the independent description parsed from the reference's source
    gold = json.loads((ROOT / "tests" / "golden" / "presets.json").read_text())["presets"]
    for name, g in gold.items():
        o = json.loads(run(["--scene", name, "--dry-run"], assets).stdout)
        d = g["defaults"]
        assert (o["width"], o["height"], o["samples_per_pixel"], o["max_depth"], o["workers"], o["vfov"], o["aperture"]) == (
            d["width"], d["height"], d["samples_per_pixel"], d["max_depth"], d["workers"], d["vfov"], d["aperture"]), name
        assert o["output_path"] == os.path.join("output", g["output_filename"])


def test_native_host_fails_loudly_without_a_gpu(yart, assets, tmp_path):
    if yart.device_count() > 0:
        pytest.skip("a GPU is visible here")
    r = run(["--scene", "cornell-box", "--width", "32", "--samples", "1", "--output", str(tmp_path / "x.png")], assets)
    assert r.returncode == 1 and "no CPU mode" in r.stderr and not (tmp_path / "x.png").exists()


@pytest.mark.gpu
def test_native_host_png_equals_the_python_front_end(yart, assets, tmp_path):
    import importlib
    from PIL import Image
    sys.path.insert(0, str(ROOT))
    cli = importlib.import_module("yart_cli")
    a, b, c = (str(tmp_path / "sub" / n) for n in ("native.png", "python.png", "multi.png"))
    common = ["--scene", "cornell-box", "--width", "80", "--samples", "6", "--seed", "4"]
    r = run(common + ["--output", a], assets)
    assert r.returncode == 0, r.stderr
    assert "rendered in" in r.stdout and "38400 paths" in r.stdout
    assert cli.main(common + ["--output", b]) == 0
    ia, ib = np.asarray(Image.open(a)), np.asarray(Image.open(b))
    assert ia.shape == (80, 80, 4) and np.array_equal(ia, ib)
    n = min(yart.device_count(), 2)       # every visible GPU (1 on the default box; `gpurun --gpus 2` exercises NCCL)
    r = run(common + ["--output", c, "--gpus", str(n), "--unbiased-light-pick"], assets)
    assert r.returncode == 0, r.stderr
    ic = np.asarray(Image.open(c)).astype(float)
    assert ic[..., :3].mean() < 0.95 * ia[..., :3].astype(float).mean()   # the unbiased pick is darker (SURVEY A-2)
    r = run(common + ["--output", c, "--gpus", "64"], assets)
    assert r.returncode == 1 and "GPU(s) visible" in r.stderr
