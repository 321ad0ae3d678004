#!/bin/bash
# Scratch tuning sweep (kept for the record): occupancy variants of the fast headline kernel.
for minb in 4 5 6 8; do
    RBS_MINB=$minb python bench.py --no-cpu-baseline --steps 4 --warmup 3 --arith fast 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('fast minb=$minb value=%.4e ms=%.3f frac=%.3f k1_GBps=%.0f e2e=%.3e' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline_k1']['achieved'], d['e2e']['value']))"
done
