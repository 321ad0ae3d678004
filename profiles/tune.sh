#!/bin/bash
# Scratch tuning sweep (kept for the record): arithmetic policy x occupancy x fuse factor of the headline kernel.
for arith in fast strict; do
for minb in 4 6 8; do
  for fuse in 64 256; do
    RBS_MINB=$minb python bench.py --no-cpu-baseline --steps 3 --warmup 2 --fuse $fuse --arith $arith 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('arith=$arith minb=$minb fuse=$fuse value=%.3e ms=%.2f frac=%.3f k1_GBps=%.0f k1_us=%.1f e2e=%.3e' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline_k1']['achieved'], d['roofline_k1']['launch_ms']*1e3, d['e2e']['value']))"
  done
done
done
