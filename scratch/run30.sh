cd $GRAFT_REPO_ROOT
B="python bench.py --n 2048 --steps 2 --warmup 3 --no-cpu --no-e2e --no-other-build --no-quiescent-leg --no-other-workloads"
$B > gpurun_out/plain_e30.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 6 -c 2 -o gpurun_out/prof_r02_euler2048_fma_final2 $B > gpurun_out/ncu_e30.log 2>&1
B="python bench.py --workload sphere --n 1024 --steps 2 --warmup 3 --no-cpu --no-e2e --no-other-build"
$B > gpurun_out/plain_p30.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 6 -c 2 -o gpurun_out/prof_r02_sphere1024_fma_final2 $B > gpurun_out/ncu_p30.log 2>&1
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-other-build --no-quiescent-leg --no-other-workloads"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_euler8192_final.csv $B > gpurun_out/ncu_l30.log 2>&1
ls -la gpurun_out/*final2.ncu-rep gpurun_out/launches_r02_euler8192_final.csv
