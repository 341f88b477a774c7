set -x
cd $GRAFT_REPO_ROOT
B="python bench.py --workload acoustics --steps 2 --warmup 3 --no-cpu --no-e2e --no-other-build"
$B > gpurun_out/plain_ac7.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:fused_step2 -s 6 -c 1 -o gpurun_out/prof_r02_acoustics4096_fused $B > gpurun_out/ncu_ac7.log 2>&1
CLAWB200_TWO_PASS=1 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 6 -c 2 -o gpurun_out/prof_r02_acoustics4096_twopass $B > gpurun_out/ncu_ac7b.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
