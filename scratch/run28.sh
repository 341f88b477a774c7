cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py tests/test_gpu_rp_properties.py -q -x 2>&1 | tail -3
bash scratch/sweep_variants.sh libclawb200_fma.so
python profiles/fma_study.py 2>/dev/null | grep -A3 "shockbubble" | grep "rel_linf\|fma\":"
