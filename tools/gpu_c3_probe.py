"""Probe: C3-shaped workloads on the CUDA path, with the solver's debug counters (KP_DEBUG=1)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time

from karpenter_b200 import _native, workloads

sizes = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]] or [(100, 1000)]
h = _native.Handle()
for apps, reps in sizes:
    t = time.time()
    enc = workloads.config_c3(n_apps=apps, replicas=reps, n_its=1000)
    te = time.time() - t
    t = time.time()
    res = h.solve(enc.problem)
    dt = time.time() - t
    st = h.stats()
    print(f"C3 {apps}x{reps}: {apps*reps} pods, {res['n_claims']} claims, encode {te:.1f}s e2e {dt*1000:.0f} ms, kernels "
          f"{st['solve_ms']:.0f} ms, {apps*reps/dt:.0f} pods/s, {st['solve_ms']*1000/(apps*reps):.2f} us/pod, "
          f"inflight_evals {res['n_inflight_evals']} max_npods {res['claim_npods'].max()}", flush=True)
h.close()
