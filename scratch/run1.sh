set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
nproc; free -g | head -2
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_pytest1.log
cat gpurun_out/r2_pytest1.log
python profiles/fma_study.py > gpurun_out/fma_study.json 2> gpurun_out/fma_study.err; tail -3 gpurun_out/fma_study.err
for wl in euler acoustics shallow sphere; do
  for ar in strict fma; do
    python bench.py --workload $wl --arithmetic $ar > gpurun_out/bench_${wl}_${ar}_r2a.json 2> gpurun_out/bench_${wl}_${ar}_r2a.err
    tail -c 600 gpurun_out/bench_${wl}_${ar}_r2a.json
  done
done
bash scratch/sweep_variants.sh libclawb200.so libclawb200_mb3.so libclawb200_mb4.so libclawb200_fmaonly.so libclawb200_fma.so libclawb200_fma_mb3.so libclawb200_fma_mb4.so 2>&1 | tee gpurun_out/variants_r2a.log
