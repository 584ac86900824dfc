"""Committed golden vectors of the solver path (tests/golden/solution_digests.json, written by
tests/golden/make_solution_digests.py): the oracle must still produce them (CPU tier) and the CUDA path must produce
them too (GPU tier, through the C ABI) -- without consulting the oracle at run time."""
import json
import os

import pytest

from tests.golden.make_solution_digests import cases, digest

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "solution_digests.json")))


def test_oracle_reproduces_the_golden_solutions():
    from tests import oracle_lib
    seen = 0
    for name, problem in cases():
        want = GOLD[name]
        try:
            got = digest(oracle_lib.solve(problem))
        except RuntimeError as e:
            got = f"refused: {e}"
        assert got == want, name
        seen += 1
    assert seen == len(GOLD)


@pytest.mark.gpu
def test_cuda_path_reproduces_the_golden_solutions():
    from karpenter_b200 import _native
    h = _native.Handle()
    bad = []
    try:
        for name, problem in cases():
            if GOLD[name].startswith("refused"):
                continue
            if digest(h.solve(problem)) != GOLD[name]:
                bad.append(name)
    finally:
        h.close()
    assert not bad, bad
