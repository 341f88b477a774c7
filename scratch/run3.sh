set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -q -m gpu 2>&1 | tail -30 > gpurun_out/r2_pytest3.log; cat gpurun_out/r2_pytest3.log
for i in 1 2; do python -m pytest tests/test_gpu_rp_properties.py -q 2>&1 | tail -3; done
python bench.py > gpurun_out/bench_euler_r2c.json 2> gpurun_out/bench_euler_r2c.err; tail -c 1500 gpurun_out/bench_euler_r2c.json
B="python bench.py --n 2048 --steps 2 --warmup 3 --no-cpu --no-e2e --no-other-build --no-quiescent-leg"
for ar in fma strict; do
  $B --arithmetic $ar > gpurun_out/plain_$ar.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_euler2048_$ar.csv $B --arithmetic $ar > gpurun_out/ncu_l_$ar.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 6 -c 2 -o gpurun_out/prof_r02_euler2048_$ar $B --arithmetic $ar > gpurun_out/ncu_f_$ar.log 2>&1
  tail -2 gpurun_out/ncu_f_$ar.log
done
ls -la gpurun_out/*.ncu-rep | tail -3
