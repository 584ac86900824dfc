"""Probe: C4 (10k nodes / 200k pods / 166 750 subsets) consolidation time on the CUDA path."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time

import numpy as np

from karpenter_b200 import _abi, _native, workloads

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
enc, consol = workloads.config_c4()
h = _native.Handle()
ci = _abi.ConsolInput(**consol)
for i in range(reps):
    t = time.time()
    r = h.consolidate(enc.problem, ci)
    dt = time.time() - t
    print("C4 device ms", round(r["solve_ms"], 2), "e2e ms", round(dt * 1000, 1), "decisions", np.bincount(r["decision"], minlength=3).tolist(), flush=True)
h.close()
