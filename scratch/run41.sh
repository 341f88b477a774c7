cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py tests/test_gpu_baseline_sizes.py tests/test_gpu_regressions.py tests/test_gpu_examples.py -q -x 2>&1 | tail -3
for ar in strict fma; do
python bench.py --workload shallow --steps 5 --warmup 3 --no-cpu --no-e2e --no-other-build --no-quiescent-leg --arithmetic $ar 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('shallow $ar %.3e (%.3f ms)'%(d['value'], d['ms_per_step']), r['all_kernels_ms'])
"
done
