#!/usr/bin/env python3
"""Random 128-byte line fetch rate vs table size (yart_measure_fetch_peak): the L1/L2-resident fetch roofline
of the closest-hit stage."""
import importlib, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
y = importlib.import_module("yet-another-raytracer_b200")
ctx = y.Context(0)
for mb in (0.125, 0.7, 2, 7.4, 32, 96, 512, 4096):
    a = max(ctx.measure_fetch_peak(int(mb * 1e6), 4096, 0) for _ in range(3))
    b = max(ctx.measure_fetch_peak(int(mb * 1e6), 4096, 1) for _ in range(3))
    c = max(ctx.measure_fetch_peak(int(mb * 1e6), 4096, 2) for _ in range(3))
    d = max(ctx.measure_fetch_peak(int(mb * 1e6), 4096, 3) for _ in range(3))
    print("table %8.3f MB: one line per LANE (four loads) %6.0f GB/s at the lean kernel's 20 warps/SM, %6.0f at full occupancy | "
          "one line per QUAD (one load per lane) %6.0f / %6.0f GB/s" % (mb, a, b, c, d))
