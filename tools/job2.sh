cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q 2>&1 | tail -15
KP_DEBUG=1 python tools/gpu_c3_probe.py 100x1000 1000x1000 2>&1 | tail -8
