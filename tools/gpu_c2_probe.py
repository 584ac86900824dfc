"""Probe: C2 (100k pods, selectors + tolerations, 500 types) resident solve time on the CUDA path."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from karpenter_b200 import _native, workloads

enc = workloads.config_c2()
h = _native.Handle()
h.upload(enc.problem)
ms = []
for i in range(6):
    res = h.solve_resident()
    ms.append(h.stats()["solve_ms"])
print("C2 100k resident ms:", [round(x, 2) for x in ms], "claims", res["n_claims"])
h.close()
