cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py tests/test_gpu_baseline_sizes.py tests/test_gpu_rp_properties.py -q -x 2>&1 | tail -3
for ar in fma strict; do
python bench.py --workload sphere --steps 10 --warmup 3 --no-cpu --no-e2e --no-other-build --arithmetic $ar 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('sphere $ar %.3e (%.3f ms)'%(d['value'], d['ms_per_step']), {k[:6]:round(v,3) for k,v in r['all_kernels_ms'].items()})
"
done
