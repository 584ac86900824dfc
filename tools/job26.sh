cd $GRAFT_REPO_ROOT
timeout 300 python tools/gpu_c3_probe.py 200x1000 1000x1000 2>&1 | tail -2
timeout 900 python -m pytest tests/test_fuzz_parity.py tests/test_gpu_parity.py tests/test_reference_topology.py tests/test_preferences.py -m gpu -q -x 2>&1 | tail -3
