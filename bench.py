#!/usr/bin/env python3
"""bench.py -- the headline measurement of the path-tracing hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], the one its `metric` is quoted on): the `david` preset
(david.obj, 46,664 triangles, white Lambertian mesh + a rotated glass instance + 5 sphere lights)
at 1920x1080, max-depth 50, 1024 spp.  A STEP is one wavefront pass of the hot path over one batch:
`spp_per_step` (32) samples of every pixel of the frame = 66 M camera paths, ~237 M world rays,
with the scene resident in HBM.  The default K = 32 steps is the whole 1024-spp job.
With N GPUs every rank renders its own sample range each step (scene replicated, weak scaling:
per-GPU work is fixed) and the f64 XYZ films are combined with one NCCL reduce inside the timed
region.  All timing is on the device (CUDA events on the launching stream), max over ranks.

Rank 0 prints ONE JSON line; see the keys in main().  `--impl reference` times the CPU oracle --
our C++ restatement of the reference's renderer (the Rust original cannot be built here) -- on all
host cores on bounded samples of the same workload.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

SCENE = "david"
WIDTH, HEIGHT, MAX_DEPTH, TOTAL_SPP = 1920, 1080, 50, 1024
SPP_PER_STEP = 128  # one wavefront batch: 265 M paths in flight, 36 GB of path state (of 180 GB)
SEED = 1
METRIC = "Mrays/s (primary+secondary) on david.obj 1920x1080 max-depth 50"
# bytes one ray moves besides node/triangle fetches: 48 B ray + 8 B time + 4 B queue entry read,
# 32 B hit record written (DESIGN.md "Algorithmic bytes")
STREAM_BYTES_PER_RAY = 92
NODE_BYTES, TRI_BYTES = 128, 48


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.idx = device_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [x for x in sm if x > 0.5 * max(mx + [1.0])] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def profiled_traffic_per_ray():
    """DRAM bytes per ray per k_traverse launch from the committed ncu capture (profiles/r1_traffic.json)."""
    try:
        return float(json.loads((ROOT / "profiles" / "r1_traffic.json").read_text())["dram_bytes_per_ray_per_launch"])
    except Exception:
        return None


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------------------
# the reference arm: the CPU oracle (port of the reference's renderer) on all host cores
# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    pkg = importlib.import_module("yet-another-raytracer_b200")
    from oracle import orc
    orc.build()
    cores = os.cpu_count() or 1
    preset = pkg.ScenePreset(SCENE, seed=SEED)  # host-side scene description only; no GPU involved
    scene = orc.Scene(preset)
    cam = preset.camera(WIDTH, HEIGHT)
    film = np.zeros((HEIGHT, WIDTH, 3))
    # size a step so that (K + W) steps take about two minutes: probe the rate on two tiles
    t0 = time.perf_counter()
    _, st = scene.render(cam, WIDTH, HEIGHT, 0, 1, MAX_DEPTH, SEED, 0, cores, film, tiles=(27, 29))
    rate_paths = st.paths / (time.perf_counter() - t0)
    paths_per_tile = (WIDTH // 8) * (HEIGHT // 8)
    budget_s = 120.0
    tiles_per_step = int(max(1, min(64, round(budget_s * rate_paths / (args.steps + args.warmup) / paths_per_tile))))

    def step(i):
        t = (i * tiles_per_step) % 64
        spp0 = (i * tiles_per_step) // 64
        lo, hi = t, min(64, t + tiles_per_step)
        _, s = scene.render(cam, WIDTH, HEIGHT, spp0, spp0 + 1, MAX_DEPTH, SEED, 0, cores, film, tiles=(lo, hi))
        return s.rays, s.paths

    for i in range(args.warmup):
        step(i)
    rays = paths = 0
    t0 = time.perf_counter()
    for i in range(args.steps):
        r, p = step(args.warmup + i)
        rays += r
        paths += p
    dt = time.perf_counter() - t0
    val = rays / dt / 1e6
    sample = "%d of the 64 tiles of one 1920x1080 sample pass per step (%d paths/step), reference traversal order" % (
        tiles_per_step, tiles_per_step * paths_per_tile)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "david 1920x1080 max-depth 50 (BASELINE configs[2]); CPU arm: " + sample,
                   "scene": SCENE, "width": WIDTH, "height": HEIGHT, "max_depth": MAX_DEPTH, "seed": SEED},
        "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "spp_per_s": paths / dt / (WIDTH * HEIGHT),
        "note": "C++ restatement of the reference (oracle/), the Rust original cannot be built here",
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    rank, world, local = dist_env()
    if args.gpus != world and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if world == 1 and args.gpus > 1:
        raise SystemExit("launch N>1 with torch.distributed.run (one process per GPU)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module("yet-another-raytracer_b200")
    pkg.load_library()
    preset = pkg.ScenePreset(SCENE, seed=SEED)
    ctx = pkg.Context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)  # torch events then time exactly the stream the kernels run on
    ctx.set_scene(preset)
    cam = preset.camera(WIDTH, HEIGHT)
    order = pkg.ORDER_NEAR
    K, W, N = args.steps, args.warmup, world
    film = torch.zeros((HEIGHT, WIDTH, 3), dtype=torch.float64, device="cuda")

    sharding = importlib.import_module("yet-another-raytracer_b200.sharding")

    def sample_range(step):  # weak scaling: every rank renders its own SPP_PER_STEP samples per step
        return sharding.step_sample_range(step % (1 << 20), rank, N, SPP_PER_STEP)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- algorithmic bytes per ray of this workload: visit counts from one untimed counted step ----
    s0, s1 = sample_range(0)
    stc = ctx.render_device(cam, WIDTH, HEIGHT, s0, s1, film.data_ptr(), MAX_DEPTH, SEED, order, SPP_PER_STEP,
                            count_visits=True)
    nodes_per_ray = stc.node_visits / stc.rays
    tris_per_ray = stc.tri_tests / stc.rays
    bytes_per_ray = NODE_BYTES * nodes_per_ray + TRI_BYTES * tris_per_ray + STREAM_BYTES_PER_RAY
    film.zero_()
    for i in range(W):
        s0, s1 = sample_range(i)
        ctx.render_device(cam, WIDTH, HEIGHT, s0, s1, film.data_ptr(), MAX_DEPTH, SEED, order, SPP_PER_STEP)
    film.zero_()

    # ---- timed region 1: device-resident (`value`) ----
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    rays = paths = launches = trace_launches = 0
    trace_ms = 0.0
    for i in range(K):
        s0, s1 = sample_range(W + i)
        st = ctx.render_device(cam, WIDTH, HEIGHT, s0, s1, film.data_ptr(), MAX_DEPTH, SEED, order, SPP_PER_STEP)
        rays += st.rays
        paths += st.paths
        launches += st.kernel_launches
        trace_launches += st.trace_launches
        trace_ms += st.trace_ms
    sharding.reduce_film(film, dst=0)  # the one real exchange step: NCCL sum of the per-rank f64 XYZ films
    ev1.record(stream)
    barrier()
    clocks = sampler.stop() if sampler else None
    ms = ev0.elapsed_time(ev1)

    # ---- timed region 2: end to end through the C ABI with HOST buffers (`e2e`) ----
    K2 = min(K, 8)
    host_film = torch.zeros((HEIGHT, WIDTH, 3), dtype=torch.float64).pin_memory()
    hf = host_film.numpy()
    ctx.render(cam, WIDTH, HEIGHT, 0, SPP_PER_STEP, MAX_DEPTH, SEED, order, SPP_PER_STEP, film=hf)  # warm the path
    hf[...] = 0.0
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    rays2 = 0
    for i in range(K2):
        s0, s1 = sample_range(W + K + i)
        _, st = ctx.render(cam, WIDTH, HEIGHT, s0, s1, MAX_DEPTH, SEED, order, SPP_PER_STEP, film=hf)
        rays2 += st.rays
    e1.record(stream)
    barrier()
    ms2 = e0.elapsed_time(e1)
    luminance = float(hf[..., 1].sum())  # the host-side result the caller reads

    # ---- combine over ranks: totals summed, time = max ----
    tot = torch.tensor([rays, paths, rays2, launches], dtype=torch.float64, device="cuda")
    tmax = torch.tensor([ms, ms2, trace_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tot)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    all_rays, all_paths, all_rays2, all_launches = tot.tolist()
    ms, ms2, trace_ms_max = tmax.tolist()
    value = all_rays / ms / 1e3
    e2e = all_rays2 / ms2 / 1e3
    film_bytes = HEIGHT * WIDTH * 3 * 8

    cpu_baseline = None
    if rank == 0 and N == 1 and not args.no_cpu_baseline:
        from oracle import orc
        orc.build()
        cores = os.cpu_count() or 1
        oscene = orc.Scene(preset)
        ofilm = np.zeros((HEIGHT, WIDTH, 3))
        t0 = time.perf_counter()
        _, ost = oscene.render(cam, WIDTH, HEIGHT, 0, 1, MAX_DEPTH, SEED, 0, cores, ofilm)
        probe = time.perf_counter() - t0  # one full 1920x1080 sample pass
        n_spp = int(max(1, min(32, round(15.0 / max(probe, 1e-3)))))  # ~15 s of CPU work
        t0 = time.perf_counter()
        _, ost = oscene.render(cam, WIDTH, HEIGHT, 1, 1 + n_spp, MAX_DEPTH, SEED, 0, cores, ofilm)
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": ost.rays / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
                        "sample": "%d spp of the 1920x1080 frame (%d paths, %d rays, %.1f s of the 1024-spp job), "
                                  "C++ restatement of the reference on %d threads, reference traversal order" % (
                                      n_spp, ost.paths, ost.rays, dt, cores)}

    if rank == 0:
        peak, peak_src = measured_peaks()
        # the fetch rooflines of SURVEY.md 8(d), measured live (untimed, ~50 ms): random whole-line (128 B = one
        # node visit) fetches from an L1-resident table and from a table the size of the scene (L2-resident)
        l1_fetch = max(ctx.measure_fetch_peak(128 * 1024, 4096, 0) for _ in range(3))
        l2_fetch = max(ctx.measure_fetch_peak(7_400_000, 4096, 0) for _ in range(3))
        # dominant kernel: k_trace.  algorithmic bytes of all its launches / their summed CUDA-event time
        trace_launches = max(1, trace_launches)
        # closest-hit launches per bounce: one k_traverse per mesh instance (+ one k_analytic per run of analytic
        # objects that k_shade does not take over); david: 2
        passes = max(1, round(trace_launches / float(K * MAX_DEPTH)))
        achieved = bytes_per_ray * rays / (trace_ms * 1e-3) / 1e9 if trace_ms > 0 else None
        line = {
            "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": N, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": "david 1920x1080 max-depth 50, %d spp per step per GPU (BASELINE configs[2]; K=%d "
                            "steps = the 1024-spp job)" % (SPP_PER_STEP, TOTAL_SPP // SPP_PER_STEP),
                "scene": SCENE, "width": WIDTH, "height": HEIGHT, "max_depth": MAX_DEPTH, "spp_per_step": SPP_PER_STEP,
                "spp_total": SPP_PER_STEP * K * N, "seed": SEED, "traversal_order": "near (bit-identical hits)",
                "l2": "inputs larger than L2: %.0f GB of path state streams per step; the 7 MB scene is meant to "
                      "stay L2-resident" % (WIDTH * HEIGHT * SPP_PER_STEP * 136 / 1e9),
                "parallelism": "sample-range sharding x%d, scene replicated, one NCCL reduce of the f64 film" % N,
            },
            "spp_per_s": all_paths / (ms * 1e-3) / (WIDTH * HEIGHT),
            "rays_per_sample": all_rays / max(all_paths, 1),
            "e2e": {"value": e2e, "unit": "Mrays/s", "h2d_bytes_per_step": film_bytes + 224,
                    "d2h_bytes_per_step": film_bytes, "steps": K2, "ms_per_step": ms2 / K2,
                    "host_result_luminance_sum": luminance},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {
                "bound": "hbm", "kernel": "closest-hit stage: k_traverse, %d launches per bounce (one per mesh instance)" % passes, "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                "traffic": (profiled_traffic_per_ray() * rays / (trace_launches / float(passes))) if profiled_traffic_per_ray() else None,
                "traffic_note": "bytes per k_traverse launch = 91.9 B/ray (dram read+write from the ncu --set full capture in "
                                "profiles/r1_traffic.json) x this run's average rays per k_traverse launch; ~20x below the algorithmic bytes because "
                                "node/triangle fetches hit L1/L2",
                "fetch_peaks": {"l1_resident_gbs": l1_fetch, "scene_sized_l2_resident_gbs": l2_fetch,
                                "frac_of_l1_resident": (achieved / l1_fetch) if achieved else None,
                                "frac_of_l2_resident": (achieved / l2_fetch) if achieved else None,
                                "note": "yart_measure_fetch_peak: every lane fetches whole 128-B lines (four "
                                        "LDG.E.256) at independent random positions, 16 warps per SM like k_traverse; "
                                        "the closest-hit stage's fetches are ~70 % L1 hits, the rest L2 hits"},
                "peak_source": peak_src, "bytes_per_ray": bytes_per_ray, "nodes_per_ray": nodes_per_ray,
                "tris_per_ray": tris_per_ray, "trace_launches": int(trace_launches),
                "avg_launch_ms": trace_ms / trace_launches, "trace_share_of_step": trace_ms / ms if ms else None,
                "note": "algorithmic bytes = 128 B x nodes + 48 B x triangles visited (counted on this workload, "
                        "equal to the oracle's counts) + 92 B of ray/hit stream; node and triangle fetches are "
                        "served by L1/L2 (the tree is 5.9 MB), so the HBM fraction is a conservative denominator",
            },
            "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=TOTAL_SPP // SPP_PER_STEP)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.steps < 1:
        raise SystemExit("--steps must be >= 1")
    args.warmup = max(args.warmup, 0)
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
