cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_3d.py tests/test_gpu_examples.py -q -x 2>&1 | tail -3
python scratch/time3d.py 256 | tail -1
CLAWB200_STEP3DS_PLANES=1 python scratch/time3d.py 256 | tail -1
python scratch/time3d.py 64 | tail -1
