cd $GRAFT_REPO_ROOT
python profiles/fma_study.py > gpurun_out/fma_study_r2h.json 2> gpurun_out/fma_study_r2h.err; tail -3 gpurun_out/fma_study_r2h.err; grep -A4 "shallow_sharpclaw" gpurun_out/fma_study_r2h.json
