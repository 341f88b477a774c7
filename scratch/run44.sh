cd $GRAFT_REPO_ROOT
python bench.py > gpurun_out/bench_euler_r2l.json 2> gpurun_out/bench_euler_r2l.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_euler_r2l.json').read().strip().split('\n')[-1]); r=d['roofline']; c=d['config']
print('euler', c['arithmetic'], 'value %.4e ms %.3f'%(d['value'], d['ms_per_step']), 'other', c['other_build'] and '%.4e'%c['other_build']['value'], 'quiescent %.4e'%c['quiescent_value'], r['all_kernels_ms'], 'frac %.3f step_frac %.3f'%(r['frac'], r['step_frac']), r['fp64_pipe'], 'e2e %.3e'%d['e2e']['value'], 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'], d['gpu_launches'])
print({k:(v['value'], v['ms_per_step']) for k,v in d['other_workloads'].items()})
"
