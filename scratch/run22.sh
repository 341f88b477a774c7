cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_euler_n8_r2h.json 2> gpurun_out/bench_euler_n8_r2h.err; tail -3 gpurun_out/bench_euler_n8_r2h.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_euler_n8_r2h.json').read().strip().split('\n')[-1]); print(d['n_gpus'], '%.4e'%d['value'], d['ms_per_step'], d.get('partition_parity',{}).get('ok'), '%.3e'%d['e2e']['value'], {k:(round(v['ms_per_step'],3), '%.3e'%v['value']) for k,v in (d.get('other_workloads') or {}).items()}, d['config']['other_build'])
"
