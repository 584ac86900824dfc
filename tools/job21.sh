cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
timeout 300 python tools/gpu_c3_probe.py 300x1000 2>&1 | tail -1
timeout 300 python tools/gpu_c2_probe.py 2>&1 | tail -1
