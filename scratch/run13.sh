cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_3d.py -q -x 2>&1 | tail -3
python scratch/time3d.py 256
python -c "
import pyclaw_b200._lib as L; L.set_variant('fma')
import runpy, sys; sys.argv=['time3d.py','256']; runpy.run_path('scratch/time3d.py', run_name='__main__')"
