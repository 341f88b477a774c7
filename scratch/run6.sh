set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -q -m gpu -x 2>&1 | tail -12 > gpurun_out/r2_pytest6.log; cat gpurun_out/r2_pytest6.log
for tp in 0 1; do
  CLAWB200_TWO_PASS=$tp python bench.py --workload acoustics --no-cpu --no-other-build > gpurun_out/bench_acoustics_r2e_tp$tp.json 2> gpurun_out/bench_acoustics_r2e_tp$tp.err
  tail -c 3000 gpurun_out/bench_acoustics_r2e_tp$tp.json
done
