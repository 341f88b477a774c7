cd $GRAFT_REPO_ROOT
for lib in libclawb200.so libclawb200_sc3.so libclawb200_fma_sc3.so; do
CLAWB200_LIB=$PWD/pyclaw_b200/csrc/$lib python bench.py --workload shallow --steps 5 --warmup 3 --no-cpu --no-e2e --no-other-build --no-quiescent-leg --arithmetic strict 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('$lib shallow %.3e (%.3f ms)'%(d['value'], d['ms_per_step']), r['all_kernels_ms'])
"
done
