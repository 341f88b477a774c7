cd $GRAFT_REPO_ROOT
O="--steps 10 --warmup 3 --no-other-workloads --no-other-build --no-quiescent-leg --no-e2e --no-cpu"
for h in 24 32 48 64 96 128; do
CLAWB200_ROWS_PER_CTA=$h python bench.py --n 4096 $O 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('euler4096 rows $h', '%.4e'%d['value'], round(d['ms_per_step'],3), d['roofline']['all_kernels_ms'])
"
CLAWB200_ROWS_PER_CTA=$h python bench.py --workload acoustics $O 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('acoustics4096 rows $h', '%.4e'%d['value'], round(d['ms_per_step'],3), d['roofline']['all_kernels_ms'])
"
done
for h in 64 96 128; do
CLAWB200_ROWS_PER_CTA=$h python bench.py --workload sphere $O 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('sphere rows $h', '%.4e'%d['value'], round(d['ms_per_step'],3), d['roofline']['all_kernels_ms'])
"
done
