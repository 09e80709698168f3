#!/bin/bash
# A/B runner: tools/ab_run.sh "<label>|<env assignments>" ...   -> one line per config: render Mrays/s, k_traverse share, sweep Mrays/s
for spec in "$@"; do
  label="${spec%%|*}"; envs="${spec#*|}"
  out=$(env $envs python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-time-to-image --no-other-configs 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('render %.1f Mrays/s  step %.1f ms  trace share %.3f' % (d['value'], d['ms_per_step'], d['roofline']['trace_share_of_step']))")
  sw=$(env $envs python tools/sweep.py --sets uniform --reps 4 2>/dev/null | grep near | sed -e 's/.*best [0-9.]* ms (\([0-9.]*\) Mrays.*/\1/')
  echo "$label: $out  | sweep near $sw Mrays/s"
done
