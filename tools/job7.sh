cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python tools/gpu_c2_probe.py 2>&1 | tail -2
KP_DEBUG=1 python tools/gpu_c3_probe.py 1000x1000 2>&1 | tail -3
python tools/gpu_c5_probe.py 10000000 batch 2>&1 | tail -3
