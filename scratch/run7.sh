cd $GRAFT_REPO_ROOT
python -m pytest tests -q -m gpu 2>&1 | tail -6
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-other-workloads > gpurun_out/bench_euler_n2_r2b.json 2> gpurun_out/bench_euler_n2_r2b.err
python bench.py --steps 20 --warmup 5 --no-cpu --no-other-workloads > gpurun_out/bench_euler_n1_r2b.json 2>/dev/null; python -c "
import json
for f in ('gpurun_out/bench_euler_n1_r2b.json','gpurun_out/bench_euler_n2_r2b.json'):
    d=json.loads(open(f).read().strip().split('\n')[-1]); print(f, d['n_gpus'], '%.4e'%d['value'], d['ms_per_step'], (d.get('partition_parity') or {}).get('ok'), '%.3e'%d['e2e']['value'], d['config']['other_build'])
"
