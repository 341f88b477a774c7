cd $GRAFT_REPO_ROOT
python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/r2_pytest35.log; cat gpurun_out/r2_pytest35.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mp_partition_check.py 2>&1 | grep -E "case|app|Error|error|Traceback" | tail -24 > gpurun_out/mp2_r2c.log; cat gpurun_out/mp2_r2c.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_euler_n2_r2k.json 2> gpurun_out/bench_euler_n2_r2k.err
python bench.py > gpurun_out/bench_euler_r2k.json 2> gpurun_out/bench_euler_r2k.err
python -c "
import json
for f in ('gpurun_out/bench_euler_r2k.json','gpurun_out/bench_euler_n2_r2k.json'):
    d=json.loads(open(f).read().strip().split('\n')[-1]); print(d['n_gpus'], '%.4e'%d['value'], round(d['ms_per_step'],3), (d.get('partition_parity') or {}).get('ok'), '%.3e'%d['e2e']['value'], {k:(round(v['ms_per_step'],3), '%.3e'%v['value']) for k,v in (d.get('other_workloads') or {}).items()})
"
