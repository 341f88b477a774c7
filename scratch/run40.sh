cd $GRAFT_REPO_ROOT
python -m pytest tests -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_euler_n4_r2k.json 2> gpurun_out/bench_euler_n4_r2k.err; tail -2 gpurun_out/bench_euler_n4_r2k.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_euler_n4_r2k.json').read().strip().split('\n')[-1]); print(d['n_gpus'], '%.4e'%d['value'], round(d['ms_per_step'],3), (d.get('partition_parity') or {}).get('ok'), '%.3e'%d['e2e']['value'], {k:(round(v['ms_per_step'],3), '%.3e'%v['value']) for k,v in (d.get('other_workloads') or {}).items()})
"
