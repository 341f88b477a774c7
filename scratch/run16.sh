cd $GRAFT_REPO_ROOT
bash scratch/sweep_variants.sh libclawb200_fma.so
for w in acoustics sphere; do
python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-e2e --no-other-build 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('   $w %.3e (%.3f ms)'%(d['value'], d['ms_per_step']), {k[:6]:round(v,3) for k,v in r['all_kernels_ms'].items()})
"
done
python profiles/fma_study.py > gpurun_out/fma_study_r2g.json 2> gpurun_out/fma_study_r2g.err; grep "rel_linf\|\": {$\|fma\":" gpurun_out/fma_study_r2g.json
