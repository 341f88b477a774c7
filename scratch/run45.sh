cd $GRAFT_REPO_ROOT
for lib in libclawb200_fma.so libclawb200_fma_xng.so; do
  bash scratch/sweep_variants.sh $lib
  for w in acoustics sphere; do
  CLAWB200_LIB=$PWD/pyclaw_b200/csrc/$lib python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-e2e --no-other-build --arithmetic strict 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('   $w %.3e (%.3f ms)'%(d['value'], d['ms_per_step']), {k[:6]:round(v,3) for k,v in r['all_kernels_ms'].items()})
"
  done
done
