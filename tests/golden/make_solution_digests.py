"""Writes tests/golden/solution_digests.json: sha256 over the parity arrays (PARITY_KEYS) of the ORACLE's solution of
fixed, seeded problems -- the synthetic KWOK configs at sizes the oracle finishes in seconds and a band of fuzz seeds
with and without soft constraints.  The oracle itself is pinned to the reference by the known-answer tables and the
restated reference scenarios; these digests freeze its answers so that (a) any later change of the oracle shows up in
the CPU tier and (b) the GPU tier has committed vectors to compare the CUDA path with, oracle or no oracle.

    python tests/golden/make_solution_digests.py      # regenerate after an intended behaviour change
"""
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402

from karpenter_b200 import _abi, workloads  # noqa: E402


def digest(res: dict) -> str:
    h = hashlib.sha256()
    for k in _abi.PARITY_KEYS:
        v = res[k]
        if k in ("claim_reservations", "claim_dropped") and not np.any(v):
            continue  # added in ABI 3: a solution without reservations keeps its earlier digest
        if isinstance(v, np.ndarray):
            h.update(k.encode() + str(v.dtype).encode() + str(v.shape).encode() + np.ascontiguousarray(v).tobytes())
        else:
            h.update(f"{k}={int(v)}".encode())
    return h.hexdigest()


def cases():
    """name -> problem (an _abi.Problem)"""
    from tests.test_fuzz_parity import encode, encode_soft
    yield "c1_1000x50", workloads.config_c1().problem
    yield "c2_5000x500", workloads.config_c2(n_pods=5000).problem
    yield "c3_20x50x200", workloads.config_c3(n_apps=20, replicas=50, n_its=200).problem
    yield "existing_200x3000", workloads.config_existing().problem
    for seed in range(0, 400, 8):
        yield f"fuzz_{seed}", encode(seed).problem
    for seed in range(0, 300, 6):
        yield f"fuzz_soft_{seed}", encode_soft(seed).problem


def main():
    from tests import oracle_lib
    out = {}
    for name, problem in cases():
        try:
            out[name] = digest(oracle_lib.solve(problem))
        except RuntimeError as e:
            out[name] = f"refused: {e}"
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "solution_digests.json")
    json.dump(out, open(path, "w"), indent=0, sort_keys=True)
    print(len(out), "digests ->", path)


if __name__ == "__main__":
    main()
