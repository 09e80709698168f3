#!/bin/bash
# sweep the traversal kernel's scheduling thresholds (env-tunable) on the sweep and the render probe
for rt in 4 8 16; do for nt in 8 12 16 20; do
  a=$(YART_TUNE_RT=$rt YART_TUNE_NT=$nt python tools/sweep.py --n 8388608 --reps 3 --sets uniform 2>&1 | grep near | sed 's/.*best \([0-9.]*\) ms.*/\1/')
  b=$(YART_TUNE_RT=$rt YART_TUNE_NT=$nt python tools/bounce_probe.py 16 2>&1 | sed 's/.*trace_ms \([0-9.]*\).*/\1/')
  echo "RT=$rt NT=$nt: sweep ${a} ms, render trace ${b} ms"
done; done
