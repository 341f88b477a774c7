cd $GRAFT_REPO_ROOT
python -m pytest tests -q -m gpu -x 2>&1 | tail -5 > gpurun_out/r2_pytest9.log; cat gpurun_out/r2_pytest9.log
python profiles/fma_study.py > gpurun_out/fma_study_r2f.json 2> gpurun_out/fma_study_r2f.err; grep -A3 "sharpclaw\|shallow" gpurun_out/fma_study_r2f.json | grep "rel_linf\|\": {"
for ar in strict fma; do
python bench.py --workload shallow --steps 5 --warmup 3 --no-cpu --no-e2e --no-other-build --no-quiescent-leg --arithmetic $ar 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('shallow $ar %.3e (%.3f ms)'%(d['value'], d['ms_per_step']), r['all_kernels_ms'])
"
done
