cd $GRAFT_REPO_ROOT
python bench.py --impl reference > gpurun_out/bench_reference_r2j.json 2> gpurun_out/bench_reference_r2j.err
python bench.py > gpurun_out/bench_euler_r2j.json 2> gpurun_out/bench_euler_r2j.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_euler_r2j.json').read().strip().split('\n')[-1]); r=d['roofline']; c=d['config']
print('euler', c['arithmetic'], 'value %.4e ms %.3f'%(d['value'], d['ms_per_step']), 'other', c['other_build'] and '%.4e'%c['other_build']['value'], 'quiescent %.4e'%c['quiescent_value'], r['all_kernels_ms'], 'frac %.3f step_frac %.3f'%(r['frac'], r['step_frac']), r['fp64_pipe'], 'e2e %.3e'%d['e2e']['value'], 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'], d['clocks'])
print({k:(v['value'], v['ms_per_step']) for k,v in d['other_workloads'].items()})
"
for wl in acoustics shallow sphere; do
  python bench.py --workload $wl --no-cpu > gpurun_out/bench_${wl}_r2j.json 2> gpurun_out/bench_${wl}_r2j.err
  python -c "
import json
d=json.loads(open('gpurun_out/bench_${wl}_r2j.json').read().strip().split('\n')[-1]); r=d['roofline']; c=d['config']
print('$wl', c['arithmetic'], 'value %.4e ms %.3f'%(d['value'], d['ms_per_step']), 'other', c['other_build'] and '%.4e'%c['other_build']['value'], r['all_kernels_ms'], 'frac %.3f step_frac %.3f'%(r['frac'], r['step_frac']), r['fp64_pipe'] and round(r['fp64_pipe']['frac'],3), 'e2e %.3e'%d['e2e']['value'])
"
done
