"""Builds libyart_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python yet-another-raytracer_b200/build.py [--force] [--verbose]
    python yet-another-raytracer_b200/build.py --bounds-check    (-> libyart_b200_checked.so, see below)

Flags that matter:
  -gencode arch=compute_100a,code=sm_100a   B200 only, no PTX for other archs
  -fmad=false                               the reference is f64 Rust without FMA contraction;
                                            bit-exact closest-hit parity needs the same roundings
  -Xcompiler -ffp-contract=off              same for the host-side QBVH build / camera maths
  -lineinfo                                 so ncu's source page maps SASS back to these files

--bounds-check builds a second library with -DYART_BOUNDS_CHECK: every index the kernels form (tree nodes,
triangle records, traversal stack, ray / hit / queue slots) is checked and a violation traps.  It stands in
for compute-sanitizer, which the GPU pool does not offer; run it with
YART_LIB_PATH=yet-another-raytracer_b200/libyart_b200_checked.so (tests/test_gpu_bounds.py does).
"""
import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
ROOT = PKG.parent
LIB = PKG / "libyart_b200.so"
LIB_CHECKED = PKG / "libyart_b200_checked.so"
STAMP = PKG / ".build_stamp"

SOURCES = ["host_obj.cpp", "host_qbvh.cpp", "host_presets.cpp", "host_api.cpp", "yart_device.cu", "device_build.cu"]
HEADERS = ["host_common.h", "device_common.cuh", "device_trace.cuh", "device_shade.cuh", "device_build.h"]
INCLUDES = ["yart.h", "yart_rng.h", "yart_spectral_tables.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall",
    "-Xptxas", "-v",
    "--shared", "-cudart", "shared",
]
for knob in ("YART_TRAVERSE_MIN_BLOCKS", "YART_SHADE_MIN_BLOCKS", "YART_TRACE_THREADS", "YART_SHADE_THREADS"):  # tuning: resident blocks per SM compiled for
    if os.environ.get(knob):
        NVCC_FLAGS += ["-D%s=%s" % (knob, os.environ[knob])]


def _digest():
    h = hashlib.sha256()
    for f in [CSRC / s for s in SOURCES + HEADERS] + [ROOT / "include" / i for i in INCLUDES] + [Path(__file__)]:
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def build(force=False, verbose=False, bounds_check=False):
    digest = _digest()
    lib = LIB_CHECKED if bounds_check else LIB
    stamp = PKG / (".build_stamp_checked" if bounds_check else ".build_stamp")
    if not force and lib.exists() and stamp.exists() and stamp.read_text().strip() == digest:
        return str(lib)
    extra = ["-DYART_BOUNDS_CHECK"] if bounds_check else []
    cmd = [nvcc_path()] + NVCC_FLAGS + extra + ["-o", str(lib)] + [str(CSRC / s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    (PKG / ("build_checked.log" if bounds_check else "build.log")).write_text(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed (see %s)" % (PKG / "build.log"))
    if verbose:
        print(log)
    stamp.write_text(digest)
    return str(lib)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, bounds_check="--bounds-check" in sys.argv))
