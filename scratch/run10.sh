cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_3d.py -q -x 2>&1 | tail -15
