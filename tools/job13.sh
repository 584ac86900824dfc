cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -12
KP_DEBUG=1 timeout 300 python tools/gpu_c2_probe.py 2>&1 | tail -4
KP_DEBUG=1 timeout 300 python tools/gpu_c3_probe.py 200x1000 2>&1 | tail -3
KP_DEBUG=1 timeout 300 python tools/gpu_c3_probe.py 1000x1000 2>&1 | tail -3
