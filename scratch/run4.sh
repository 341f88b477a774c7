set -x
cd $GRAFT_REPO_ROOT
python scratch/debug_rpt.py 2>&1 | tail -60 > gpurun_out/debug_rpt2.log; cat gpurun_out/debug_rpt2.log
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py tests/test_gpu_regressions.py -q -x 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mp_partition_check.py 2>&1 | grep -E "case|app|Error|error" | tail -20 > gpurun_out/mp2_r2.log; cat gpurun_out/mp2_r2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_euler_n2_r2.json 2> gpurun_out/bench_euler_n2_r2.err; tail -c 1800 gpurun_out/bench_euler_n2_r2.json; tail -5 gpurun_out/bench_euler_n2_r2.err
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_euler_n1_r2.json 2>/dev/null; python -c "
import json
for f in ('gpurun_out/bench_euler_n1_r2.json','gpurun_out/bench_euler_n2_r2.json'):
    d=json.loads(open(f).read().strip().split('\n')[-1]); print(f, d['n_gpus'], '%.4e'%d['value'], d['ms_per_step'], d.get('partition_parity'), d['e2e']['value'])
"
