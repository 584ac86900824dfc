cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q -x 2>&1 | tail -5
echo "== default"; KP_DEBUG=1 python tools/gpu_c3_probe.py 1000x1000 2>&1 | tail -3
echo "== CR=0 CS_CAP=2048"; KP_CR=0 KP_CS_CAP=2048 KP_DEBUG=1 python tools/gpu_c3_probe.py 1000x1000 2>&1 | tail -3
echo "== CR=0 CS_CAP=1280"; KP_CR=0 KP_CS_CAP=1280 KP_DEBUG=1 python tools/gpu_c3_probe.py 1000x1000 2>&1 | tail -3
