cd $GRAFT_REPO_ROOT
CLAWB200_LIB=$PWD/pyclaw_b200/csrc/libclawb200_noguard.so python -m pytest tests/test_gpu_kernels.py tests/test_gpu_baseline_sizes.py tests/test_gpu_golden.py -q -x 2>&1 | tail -4
for lib in libclawb200_fma.so libclawb200_fma_noguard.so; do
  bash scratch/sweep_variants.sh $lib
  CLAWB200_LIB=$PWD/pyclaw_b200/csrc/$lib python bench.py --workload acoustics --steps 10 --warmup 3 --no-cpu --no-e2e --no-other-build --arithmetic strict 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('   acoustics %.3e (%.3f ms)'%(d['value'], d['ms_per_step']), {k[:6]:round(v,3) for k,v in r['all_kernels_ms'].items()})
"
  CLAWB200_LIB=$PWD/pyclaw_b200/csrc/$lib python bench.py --workload sphere --steps 5 --warmup 3 --no-cpu --no-e2e --no-other-build --arithmetic strict 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('   sphere %.3e (%.3f ms)'%(d['value'], d['ms_per_step']), {k[:6]:round(v,3) for k,v in r['all_kernels_ms'].items()})
"
done
python scratch/time3d.py 256
python scratch/time3d.py 64
