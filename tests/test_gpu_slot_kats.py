"""-m gpu: the reference's own known-answer tables for the requirement algebra (tests/golden/requirement_kats.json, from
pkg/scheduling/requirement_test.go:103-1084 and requirements_test.go:57-543) run through the DEVICE code the kernels use
(karpenter_b200/csrc/kp_slot.hpp, via kp_debug_slot_algebra) -- not only through the oracle."""
import json
import os

import numpy as np
import pytest

from karpenter_b200 import _native, encode
from karpenter_b200.model import NodeSelectorRequirement

pytestmark = pytest.mark.gpu

KATS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "requirement_kats.json")))
UNIVERSE = ["1", "2", "9", "A", "B", "a", "b", "c", "d", "e", "f", "test"]  # every value the tables mention
VI = [encode.go_atoi(v) or 0 for v in UNIVERSE]
IS = sum(1 << i for i, v in enumerate(UNIVERSE) if encode.go_atoi(v) is not None)
UNIV = (1 << len(UNIVERSE)) - 1
PRESENT = 0x10
OPS = {0: "In", 1: "NotIn", 2: "Exists", 3: "DoesNotExist"}


def slot_of(sym):
    """operator form -> (flags, mask, gte, lte) through the caller-side canonicalisation (encode.canonical_requirement ==
    NewRequirementWithFlexibility, requirement.go:48-102); None -> the key is undefined."""
    if sym is None:
        return (0, 0, 0, 0)
    _, comp, vals, gte, lte, _ = encode.canonical_requirement(NodeSelectorRequirement("key", sym["op"], tuple(sym["values"]), sym["min_values"]))
    flags = PRESENT | (1 if comp else 0) | (2 if gte is not None else 0) | (4 if lte is not None else 0)
    return (flags, sum(1 << UNIVERSE.index(v) for v in vals), gte or 0, lte or 0)


def literal(e):
    flags = PRESENT | (1 if e["complement"] else 0) | (2 if e["gte"] is not None else 0) | (4 if e["lte"] is not None else 0)
    return (flags, sum(1 << UNIVERSE.index(v) for v in e["values"]), e["gte"] or 0, e["lte"] or 0)


def case(a, b, value=-1, well_known=1, allow=1):
    c = np.zeros(1, _native.SLOT_CASE)[0]
    c["flags_a"], c["mask_a"], c["gte_a"], c["lte_a"] = a
    c["flags_b"], c["mask_b"], c["gte_b"], c["lte_b"] = b
    c["value"], c["well_known"], c["allow_undefined"] = value, well_known, allow
    return c


@pytest.fixture(scope="module")
def handle():
    h = _native.Handle()
    yield h
    h.close()


def same(o, exp):
    f, m, g, l = exp
    if int(o["flags"]) != f or int(o["mask"]) != m:
        return False
    if f & 2 and int(o["gte"]) != g:
        return False
    if f & 4 and int(o["lte"]) != l:
        return False
    return True


def test_device_intersection_table(handle):  # requirement_test.go:103-748, 953-1084: 393 entries
    sym = {k: slot_of(v) for k, v in KATS["symbols"].items()}
    rows = KATS["intersection"]
    assert len(rows) >= 392
    out = handle.slot_algebra(VI, IS, UNIV, np.array([case(sym[e["a"]], sym[e["b"]]) for e in rows], _native.SLOT_CASE))
    for e, o in zip(rows, out):
        exp = sym[e["expected_symbol"]] if "expected_symbol" in e else literal(e["expected_literal"])
        assert same(o, exp), (e, o, exp)
        nonempty = bool(int(o["flags"]) & 1) or int(o["mask"]) != 0
        assert bool(o["has_intersection"]) == nonempty  # HasIntersection == cheaper Intersection (requirement.go:208-211)


def test_device_has_and_operator_tables(handle):  # requirement_test.go:749-874
    sym = {k: slot_of(v) for k, v in KATS["symbols"].items()}
    rows = KATS["has"]
    assert len(rows) == 70
    out = handle.slot_algebra(VI, IS, UNIV, np.array([case(sym[e["r"]], sym[e["r"]], value=UNIVERSE.index(e["value"])) for e in rows],
                                                     _native.SLOT_CASE))
    for e, o in zip(rows, out):
        assert bool(o["has_value"]) == e["expected"], e
    rows = KATS["operator"]
    out = handle.slot_algebra(VI, IS, UNIV, np.array([case(sym[e["r"]], sym[e["r"]]) for e in rows], _native.SLOT_CASE))
    for e, o in zip(rows, out):  # r ∩ r == r: the operator of the intersection is the operator of r
        assert OPS[int(o["op"])] == e["expected"], e


@pytest.mark.parametrize("allow", [True, False])
def test_device_compatible_matrices(handle, allow):  # requirements_test.go:57-543: 225 pairs per mode
    sym = {k: slot_of(v) for k, v in KATS["compat_symbols"].items()}
    rows = [e for e in KATS["compatible"] if e["allow_undefined"] == allow]
    assert len(rows) == 225
    # Requirements{zone: a}.Compatible(Requirements{zone: b}): zone is a well-known label
    out = handle.slot_algebra(VI, IS, UNIV, np.array([case(sym[e["a"]], sym[e["b"]], allow=int(allow)) for e in rows], _native.SLOT_CASE))
    for e, o in zip(rows, out):
        assert bool(o["compatible"]) == e["ok"], e
