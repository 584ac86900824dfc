import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from karpenter_b200 import _abi, _native, workloads
enc, consol = workloads.config_c4()
h = _native.Handle()
for _ in range(2):
    res = h.consolidate(enc.problem, _abi.ConsolInput(**consol))
    print(res["solve_ms"], h.stats())
h.close()
