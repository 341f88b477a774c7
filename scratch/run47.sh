cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tests/mp_partition_check.py 2>&1 | grep -E "case|app|Error|error|Traceback|thinner" | tail -24 > gpurun_out/mp8_r2.log; cat gpurun_out/mp8_r2.log
