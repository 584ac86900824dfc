"""Time the CUDA path on C3-shaped workloads (apps x replicas, zone spread + hostname anti-affinity) and check parity
against the oracle where it finishes quickly."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time

from karpenter_b200 import _native, workloads
from tests import oracle_lib
from tests.parity import assert_same

h = _native.Handle()
for apps, reps, check in [(100, 1000, True), (300, 1000, False)]:
    enc = workloads.config_c3(n_apps=apps, replicas=reps, n_its=1000)
    t = time.time()
    res = h.solve(enc.problem)
    dt = time.time() - t
    st = h.stats()
    print(f"C3 {apps}x{reps}: {apps*reps} pods, {res['n_claims']} claims, e2e {dt*1000:.0f} ms, kernels {st['solve_ms']:.0f} ms, "
          f"{apps*reps/dt:.0f} pods/s", flush=True)
    if check:
        t = time.time()
        orc = oracle_lib.solve(enc.problem)
        print(f"  oracle {time.time()-t:.1f} s", flush=True)
        assert_same(res, orc, "C3 ")
        print("  parity ok", flush=True)
h.close()
