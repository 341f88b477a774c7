cd $GRAFT_REPO_ROOT
python scratch/pcie_bw.py
