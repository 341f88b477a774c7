#!/bin/bash
# usage: scratch/sweep_variants.sh lib1.so lib2.so ...
for lib in "$@"; do
  for args in "--n 4096 --steps 5 --warmup 3" "--n 4096 --steps 5 --warmup 3 --perturb" "--workload acoustics --steps 10" "--workload shallow --n 4096 --steps 3"; do
    CLAWB200_LIB=$PWD/pyclaw_b200/csrc/$lib python bench.py $args --no-cpu --no-e2e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('$lib', '$args'[:34].ljust(34), 'value %.3e'%d['value'], 'ms/step %.2f'%d['ms_per_step'], {k[:6]:round(v,3) for k,v in r['all_kernels_ms'].items()})
"
  done
done
