#!/bin/bash
# usage: scratch/sweep_variants.sh lib1.so lib2.so ...   (files under pyclaw_b200/csrc)
# Euler 4096^2, developed field (value) and the application's quiescent field, per-kernel times.
for lib in "$@"; do
  CLAWB200_LIB=$PWD/pyclaw_b200/csrc/$lib python bench.py --n 4096 --steps 6 --warmup 3 --no-cpu --no-e2e --no-other-build --no-other-workloads --arithmetic strict ${VARIANT_ARGS} 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; q=d['config'].get('quiescent') or {}
        print('$lib'.ljust(28), 'developed %.3e (%.2f ms)'%(d['value'], d['ms_per_step']), 'quiescent %.3e'%(q.get('value') or 0), {k[:6]:round(v,3) for k,v in r['all_kernels_ms'].items()})
"
done
