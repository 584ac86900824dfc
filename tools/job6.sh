cd $GRAFT_REPO_ROOT
KP_DEBUG=1 python tools/gpu_c5_probe.py 1600000 both 2>&1 | tail -12
KP_DEBUG=1 python tools/gpu_c5_probe.py 10000000 single 2>&1 | tail -6
