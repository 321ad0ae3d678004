#!/bin/bash
# Scratch tuning sweep (kept for the record): occupancy variants x fuse factors of the headline kernel.
for minb in 4 5 6 8; do
  for fuse in 64; do
    RBS_MINB=$minb python bench.py --no-cpu-baseline --steps 3 --warmup 2 --fuse $fuse 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('minb=$minb fuse=$fuse value=%.3e ms=%.2f k1_GBps=%.0f k1_us=%.1f e2e=%.3e' % (d['value'], d['ms_per_step'], d['roofline_k1']['achieved'], d['roofline_k1']['launch_ms']*1e3, d['e2e']['value']))"
  done
done
