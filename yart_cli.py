#!/usr/bin/env python3
"""Command-line front end with the reference's flags (reference raytracer/src/main.rs:78-107):

    python yart_cli.py --scene david [--output out.png] [--width W] [--height H] [--samples N]
                       [--max-depth D] [--vfov F] [--aperture A]            (reference flags)
                       [--seed S] [--gpus G] [--order near|reference]        (additions)

`--workers` is accepted for compatibility and ignored (the render runs on the GPU).  Option resolution
follows resolve_render_options / resolve_dimensions (main.rs:166-209); the image is finalised exactly like
main.rs:710-718 and written as PNG (main.rs:769-774).  There is no CPU mode.
"""
import argparse
import importlib
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def positive_int(text):  # parse_positive_usize (main.rs:148-158)
    try:
        v = int(text)
    except ValueError as e:
        raise argparse.ArgumentTypeError("invalid integer `%s`: %s" % (text, e))
    if v <= 0:
        raise argparse.ArgumentTypeError("value must be greater than 0")
    return v


def build_parser(scene_names):
    ap = argparse.ArgumentParser(description="Render predefined raytracer scenes")
    ap.add_argument("--scene", required=True, choices=scene_names)
    ap.add_argument("--output", default=None)
    ap.add_argument("--width", type=positive_int, default=None)
    ap.add_argument("--height", type=positive_int, default=None)
    ap.add_argument("--samples", type=positive_int, default=None)
    ap.add_argument("--max-depth", type=positive_int, default=None, dest="max_depth")
    ap.add_argument("--workers", type=positive_int, default=None)
    ap.add_argument("--vfov", type=float, default=None)
    ap.add_argument("--aperture", type=float, default=None)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--order", choices=["near", "reference"], default="near")
    ap.add_argument("--device", type=int, default=0)
    return ap


def main(argv=None):
    y = importlib.import_module("yet-another-raytracer_b200")
    args = build_parser(y.SCENE_NAMES).parse_args(argv)
    preset = y.ScenePreset(args.scene, seed=args.seed)
    o = preset.resolve_render_options(args.output, args.width, args.height, args.samples, args.max_depth,
                                      args.workers, args.vfov, args.aperture)
    start = time.time()  # the reference's timer starts in render(), after the scene is built (main.rs:591)
    ctx = y.Context(args.device)
    ctx.set_scene(preset)
    cam = preset.camera(o["width"], o["height"], o["vfov"], o["aperture"])
    order = y.ORDER_NEAR if args.order == "near" else y.ORDER_REFERENCE
    film, st = ctx.render(cam, o["width"], o["height"], 0, o["samples_per_pixel"], o["max_depth"], args.seed, order)
    rgba = ctx.film_finalize(film, o["samples_per_pixel"])
    print("%s rendered in %d seconds" % (o["output_path"], int(time.time() - start)))  # main.rs:763-767
    print("  %d paths, %d rays, %.1f Mrays/s on the device" % (st.paths, st.rays, st.rays / max(st.gpu_ms, 1e-9) / 1e3))
    parent = os.path.dirname(o["output_path"])
    if parent:
        os.makedirs(parent, exist_ok=True)
    from PIL import Image
    Image.fromarray(rgba, "RGBA").save(o["output_path"])
    return 0


if __name__ == "__main__":
    sys.exit(main())
