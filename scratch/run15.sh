cd $GRAFT_REPO_ROOT
O="--steps 20 --warmup 5 --no-other-workloads --no-other-build --no-quiescent-leg --no-e2e --no-cpu --no-parity"
p() { python -c "
import json,sys
for l in open('$1'):
    if l.startswith('{'):
        d=json.loads(l); print('$1', d['n_gpus'], '%.4e'%d['value'], round(d['ms_per_step'],3), d['roofline']['all_kernels_ms'])
"; }
python bench.py $O > gpurun_out/t_n1a.json 2>/dev/null; p gpurun_out/t_n1a.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 $O > gpurun_out/t_n2a.json 2>/dev/null; p gpurun_out/t_n2a.json
python bench.py $O > gpurun_out/t_n1b.json 2>/dev/null; p gpurun_out/t_n1b.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 $O > gpurun_out/t_n2b.json 2>/dev/null; p gpurun_out/t_n2b.json
CUDA_VISIBLE_DEVICES=1 python bench.py $O > gpurun_out/t_n1c.json 2>/dev/null; p gpurun_out/t_n1c.json
