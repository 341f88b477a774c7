cd $GRAFT_REPO_ROOT
CLAWB200_LIB=$PWD/pyclaw_b200/csrc/libclawb200_rot.so python -m pytest tests/test_gpu_kernels.py -q -x -k "step2_unsplit" 2>&1 | tail -3
for lib in libclawb200_fma.so libclawb200_fma_rot.so; do
  CLAWB200_LIB=$PWD/pyclaw_b200/csrc/$lib python bench.py --workload acoustics --steps 10 --warmup 3 --no-cpu --no-e2e --no-other-build --arithmetic strict 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('$lib acoustics %.3e (%.3f ms)'%(d['value'], d['ms_per_step']), {k[:6]:round(v,3) for k,v in r['all_kernels_ms'].items()})
"
done
python scratch/time3d.py 128 > /dev/null && ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r02_step3_128.csv python scratch/time3d.py 128 > gpurun_out/ncu_3d.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/launches_r02_step3_128.csv')) if len(r) > 5]
h = rows[0]; ik = h.index('Kernel Name'); iv = h.index('Metric Value')
agg = collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ik][:60]].append(float(r[iv].replace(',', '')))
    except ValueError: pass
for k, v in agg.items(): print('%-62s n=%3d mean %.1f us' % (k, len(v), sum(v) / len(v) / 1e3))
PY
