cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
KP_DEBUG=1 timeout 300 python tools/gpu_c3_probe.py 300x1000 2>&1 | tail -2
timeout 300 python tools/gpu_c2_probe.py 2>&1 | tail -1
timeout 600 python tools/gpu_dep_probe.py 300x1000:t 100x1000:s 1000x1000:t 2>&1 | tail -8
