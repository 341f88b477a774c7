cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py -q -x 2>&1 | tail -3
for lib in libclawb200_fma.so libclawb200_ac_noregs.so libclawb200_ac_regs_mb3.so libclawb200_ac_regs_x4.so; do
CLAWB200_LIB=$PWD/pyclaw_b200/csrc/$lib python bench.py --workload acoustics --arithmetic strict --steps 20 --no-cpu --no-e2e --no-other-build 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('$lib'.ljust(30), 'value %.3e (%.3f ms)'%(d['value'], d['ms_per_step']), {k[:6]:round(v,4) for k,v in r['all_kernels_ms'].items()})
"
done
