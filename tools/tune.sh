#!/bin/bash
# sweep the traversal kernel's scheduling thresholds (env-tunable) on the closest-hit sweep
for rt in 1 4 8 16; do for nt in 1 12 20 28; do
  echo "RT=$rt NT=$nt: $(YART_TUNE_RT=$rt YART_TUNE_NT=$nt python tools/sweep.py --n 8388608 --reps 3 --sets uniform 2>&1 | grep near | sed 's/.*best/best/')"
done; done
