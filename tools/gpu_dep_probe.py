"""Probe: deployment-shaped workloads (workloads.config_deployments) on the CUDA path, cohorts on and off."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time

from karpenter_b200 import _native, workloads

for arg in sys.argv[1:] or ["100x1000:t"]:
    size, kind = arg.split(":")
    deps, reps = (int(x) for x in size.split("x"))
    enc = workloads.config_deployments(deps, reps, 1000 if kind == "t" else 500, topology=kind == "t")
    ref = None
    for mode in ("cohort", "plain"):
        os.environ.pop("KP_NO_COHORT", None)
        if mode == "plain":
            os.environ["KP_NO_COHORT"] = "1"
        h = _native.Handle()
        t = time.time()
        res = h.solve(enc.problem)
        dt = time.time() - t
        st = h.stats()
        h.close()
        same = ""
        if ref is None:
            ref = res
        else:
            import numpy as np
            same = " identical=" + str(all(np.array_equal(res[k], ref[k]) for k in ("pod_target", "claim_rank", "claim_npods", "claim_its")))
        print(f"deployments {deps}x{reps} {'topology' if kind == 't' else 'selectors'} [{mode}]: {res['n_claims']} claims, kernels "
              f"{st['solve_ms']:.1f} ms, {st['solve_ms']*1000/(deps*reps):.3f} us/pod, unscheduled {(res['pod_target'] == -1).sum()}{same}",
              flush=True)
