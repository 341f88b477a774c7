cd $GRAFT_REPO_ROOT
O="--steps 10 --warmup 3 --no-other-workloads --no-other-build --no-quiescent-leg --no-e2e --no-cpu"
for h in 32 64 96 128 256; do
CLAWB200_ROWS_PER_CTA=$h python bench.py $O 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('rows $h', '%.4e'%d['value'], round(d['ms_per_step'],3), d['roofline']['all_kernels_ms'])
"
done
