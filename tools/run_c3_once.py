import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from karpenter_b200 import _native, workloads
enc = workloads.config_c3(n_apps=100, replicas=1000, n_its=1000)
h = _native.Handle()
res = h.solve(enc.problem)
print(res["n_claims"], h.stats()["solve_ms"])
h.close()
