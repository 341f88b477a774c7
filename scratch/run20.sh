cd $GRAFT_REPO_ROOT
O="--steps 5 --warmup 3 --no-other-workloads --no-other-build --no-quiescent-leg --no-e2e --no-cpu"
for h in 64 128; do
CLAWB200_ROWS_PER_CTA=$h python bench.py --workload shallow $O 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('shallow rows $h', '%.4e'%d['value'], round(d['ms_per_step'],3), d['roofline']['all_kernels_ms'])
"
done
B="python bench.py --n 2048 --steps 2 --warmup 3 --no-cpu --no-e2e --no-other-build --no-quiescent-leg --no-other-workloads"
$B > gpurun_out/plain_e20.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 6 -c 2 -o gpurun_out/prof_r02_euler2048_fma_final $B > gpurun_out/ncu_e20.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_euler2048_fma_final.csv $B > gpurun_out/ncu_l20.log 2>&1
B="python bench.py --workload acoustics --steps 2 --warmup 3 --no-cpu --no-e2e --no-other-build"
$B > gpurun_out/plain_a20.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fused_step2 -s 6 -c 1 -o gpurun_out/prof_r02_acoustics4096_fused_final $B > gpurun_out/ncu_a20.log 2>&1
B="python bench.py --workload shallow --n 2048 --steps 2 --warmup 3 --no-cpu --no-e2e --no-other-build --no-quiescent-leg"
$B > gpurun_out/plain_s20.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sc2d_kernel -s 6 -c 1 -o gpurun_out/prof_r02_shallow2048_final $B > gpurun_out/ncu_s20.log 2>&1
B="python bench.py --workload sphere --n 1024 --steps 2 --warmup 3 --no-cpu --no-e2e --no-other-build"
$B > gpurun_out/plain_p20.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 6 -c 2 -o gpurun_out/prof_r02_sphere1024_fma_final $B > gpurun_out/ncu_p20.log 2>&1
python scratch/time3d.py 128 > gpurun_out/plain_3d20.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"flux3_kernel|apply3_kernel" -s 6 -c 6 -o gpurun_out/prof_r02_step3_128 python scratch/time3d.py 128 > gpurun_out/ncu_3d20.log 2>&1
ls -la gpurun_out/*final*.ncu-rep gpurun_out/prof_r02_step3_128.ncu-rep
