cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q -x 2>&1 | tail -6
python tools/gpu_c2_probe.py 2>&1 | tail -2
KP_NO_LEAN=1 python tools/gpu_c2_probe.py 2>&1 | tail -1
python - <<'PY'
import sys, time
sys.path.insert(0, '.')
from karpenter_b200 import _abi, _native, workloads
enc, consol = workloads.config_c4()
h = _native.Handle()
ci = _abi.ConsolInput(**consol)
for i in range(4):
    t = time.time(); r = h.consolidate(enc.problem, ci); dt = time.time() - t
    print("C4 device ms", round(r["solve_ms"], 2), "e2e ms", round(dt * 1000, 1))
h.close()
PY
