set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -q -m gpu 2>&1 | tail -12 > gpurun_out/r2_pytest5.log; cat gpurun_out/r2_pytest5.log
for wl in euler acoustics shallow sphere; do
  python bench.py --workload $wl --no-cpu > gpurun_out/bench_${wl}_r2d.json 2> gpurun_out/bench_${wl}_r2d.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_${wl}_r2d.json').read().strip().split('\n')[-1]); r=d['roofline']; c=d['config']
print('$wl', c['arithmetic'], 'value %.4e ms %.3f'%(d['value'], d['ms_per_step']), 'other', c['other_build'] and '%.4e'%c['other_build']['value'], 'quiescent', c['quiescent_value'], r['all_kernels_ms'], 'e2e %.3e'%d['e2e']['value'])
"
done
