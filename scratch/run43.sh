cd $GRAFT_REPO_ROOT
python -m pytest tests -q -m gpu 2>&1 | tail -3
python bench.py --workload shallow --no-cpu > gpurun_out/bench_shallow_r2l.json 2> gpurun_out/bench_shallow_r2l.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_shallow_r2l.json').read().strip().split('\n')[-1]); r=d['roofline']; c=d['config']
print('shallow', c['arithmetic'], 'value %.4e ms %.3f'%(d['value'], d['ms_per_step']), 'other', c['other_build'] and '%.4e'%c['other_build']['value'], r['all_kernels_ms'], 'frac %.3f step_frac %.3f'%(r['frac'], r['step_frac']), 'e2e %.3e'%d['e2e']['value'], 'quiescent', c['quiescent_value'])
"
B="python bench.py --workload shallow --n 2048 --steps 2 --warmup 3 --no-cpu --no-e2e --no-other-build --no-quiescent-leg"
$B > gpurun_out/plain_s43.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sc2d_kernel -s 6 -c 1 -o gpurun_out/prof_r02_shallow2048_final3 $B > gpurun_out/ncu_s43.log 2>&1
