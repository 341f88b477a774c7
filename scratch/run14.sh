cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mp_partition_check.py 2>&1 | grep -E "case|app|Error|error|Traceback" | tail -24 > gpurun_out/mp2_r2b.log; cat gpurun_out/mp2_r2b.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_euler_n2_r2c.json 2> gpurun_out/bench_euler_n2_r2c.err; tail -5 gpurun_out/bench_euler_n2_r2c.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_euler_n2_r2c.json').read().strip().split('\n')[-1]); print(d['n_gpus'], '%.4e'%d['value'], d['ms_per_step'], d.get('partition_parity',{}).get('ok'), '%.3e'%d['e2e']['value'], {k:(round(v['ms_per_step'],3)) for k,v in (d.get('other_workloads') or {}).items()})
"
