cd $GRAFT_REPO_ROOT
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -x -k "step2_unsplit and (37 or 130) and (2-2 or 2-0) and not euler" 2>&1 | tail -6
echo "memcheck fused rc=$?"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_3d.py -q -x -k "step3_unsplit and (9 or 20) and 22" 2>&1 | tail -6
echo "memcheck step3 rc=$?"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -x -k "step2_unsplit and 37 and 2-2 and acoustics" 2>&1 | tail -6
echo "racecheck fused rc=$?"
