set -x
cd $GRAFT_REPO_ROOT
python scratch/debug_rpt.py 2>&1 | tail -40 > gpurun_out/debug_rpt.log; cat gpurun_out/debug_rpt.log
CLAWB200_LIB=$PWD/pyclaw_b200/csrc/libclawb200_ystrip.so python -m pytest tests -q -m gpu --deselect tests/test_gpu_rp_properties.py 2>&1 | tail -25 > gpurun_out/r2_pytest_ystrip.log; cat gpurun_out/r2_pytest_ystrip.log
bash scratch/sweep_variants.sh libclawb200.so libclawb200_ystrip.so libclawb200_ystrip_mb3.so libclawb200_ystrip_mb4.so libclawb200_fma.so libclawb200_fma_ystrip_mb3.so libclawb200_fma_ystrip_mb4.so 2>&1 | tee gpurun_out/variants_r2b.log
python - <<'PY' 2>&1 | tail -5 | tee gpurun_out/fma_sharpclaw.log
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np
from pyclaw_b200 import _lib
_lib.LIB_PATHS['fmaonly'] = os.path.join(os.getcwd(), 'pyclaw_b200/csrc/libclawb200_fmaonly.so')
import test_gpu_golden as tg
o = dict(time_integrator="SSP33", cfl_max=0.6, cfl_desired=0.5)
r = {v: tg._shallow("sharpclaw", arithmetic=v, **o) for v in ("strict", "fma", "fmaonly")}
for v in ("fma", "fmaonly"):
    print(v, max(np.abs(r[v][m] - r["strict"][m]).max() / np.abs(r["strict"][m]).max() for m in range(3)))
PY
