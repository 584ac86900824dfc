cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q 2>&1 | tail -25
python tools/gpu_c2_probe.py 2>&1 | tail -2
