import sys
sys.path.insert(0, "/root/repo")
import numpy as np
from karpenter_b200 import _native, workloads
from tests import oracle_lib
from tests.parity import assert_same
enc = workloads.config_c3(40, 50, 300)
h = _native.Handle()
g = h.solve(enc.problem)
o = oracle_lib.solve(enc.problem)
print("claims", g["n_claims"], o["n_claims"])
bad = np.nonzero(g["pod_target"] != o["pod_target"])[0]
print("mismatch", len(bad), bad[:10], g["pod_target"][bad[:10]], o["pod_target"][bad[:10]])
print("rank", g["claim_rank"][:20], o["claim_rank"][:20])
print("npods", g["claim_npods"][:20], o["claim_npods"][:20])
