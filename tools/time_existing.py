"""Time kp_solve on a live-cluster workload: 10k existing nodes + 100k pending pods (existing-node stage + claims)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time

import numpy as np

from karpenter_b200 import _native, workloads
from tests import oracle_lib
from tests.parity import assert_same

h = _native.Handle()
for nn, npods, check in [(2000, 20000, True), (10000, 100000, False)]:
    enc = workloads.config_existing(n_nodes=nn, n_pods=npods, fill=0.7)
    t = time.time()
    res = h.solve(enc.problem)
    dt = time.time() - t
    tgt = res["pod_target"]
    print(f"existing {nn} nodes / {npods} pods: on nodes {(tgt >= 0).sum()}, claims {res['n_claims']}, unsched {(tgt == -1).sum()}, "
          f"e2e {dt*1000:.0f} ms, kernels {h.stats()['solve_ms']:.0f} ms", flush=True)
    if check:
        assert_same(res, oracle_lib.solve(enc.problem), "existing ")
        print("  parity ok", flush=True)
h.close()
