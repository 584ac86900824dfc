"""Pins the oracle's requirement algebra against the reference's own known-answer tables
(tests/golden/requirement_kats.json, extracted by tests/golden/extract_kats.py from
pkg/scheduling/requirement_test.go:103-874 and requirements_test.go:57-543)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from karpenter_b200 import _abi, encode
from karpenter_b200.model import NodeSelectorRequirement, ZONE_LABEL
from tests import oracle_lib

KATS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "requirement_kats.json")))
UNIVERSE = ["1", "2", "9", "A", "B"]
VI = np.array([encode.go_atoi(v) or 0 for v in UNIVERSE], np.int64)
IS = np.array([0 if encode.go_atoi(v) is None else 1 for v in UNIVERSE], np.uint8)
OPC = {"In": 0, "NotIn": 1, "Exists": 2, "DoesNotExist": 3, "Gt": 4, "Lt": 5, "Gte": 6, "Lte": 7}
OPN = {v: k for k, v in OPC.items()}


def canon(sym):
    """operator form -> canonical (flags, gte, lte, min, values) through the ORACLE's NewRequirementWithFlexibility."""
    L = oracle_lib.lib()
    vals = np.array([UNIVERSE.index(v) for v in sym["values"]] if sym["op"] in ("In", "NotIn") else [], np.int32)
    operand = int(sym["values"][0]) if sym["op"] in ("Gt", "Lt", "Gte", "Lte") else 0
    fl, g, l, mv, n = C.c_uint8(), C.c_int64(), C.c_int64(), C.c_int32(), C.c_int32()
    ov = np.zeros(8, np.int32)
    op = L.orc_kat_new_requirement(OPC[sym["op"]], C.c_int64(operand), 1 if sym["min_values"] is not None else 0,
                                   sym["min_values"] or 0, vals.ctypes.data_as(C.c_void_p), len(vals), C.byref(fl),
                                   C.byref(g), C.byref(l), C.byref(mv), ov.ctypes.data_as(C.c_void_p), C.byref(n))
    return dict(flags=fl.value, gte=g.value, lte=l.value, min=mv.value, values=sorted(ov[:n.value].tolist()), op=OPN[op])


def intersect(a, b):
    L = oracle_lib.lib()
    va, vb = np.array(a["values"], np.int32), np.array(b["values"], np.int32)
    fl, g, l, mv, n, hi = C.c_uint8(), C.c_int64(), C.c_int64(), C.c_int32(), C.c_int32(), C.c_int32()
    ov = np.zeros(16, np.int32)
    op = L.orc_kat_intersection(VI.ctypes.data_as(C.c_void_p), IS.ctypes.data_as(C.c_void_p), len(UNIVERSE),
                                C.c_uint8(a["flags"]), C.c_int64(a["gte"]), C.c_int64(a["lte"]), a["min"],
                                va.ctypes.data_as(C.c_void_p), len(va), C.c_uint8(b["flags"]), C.c_int64(b["gte"]),
                                C.c_int64(b["lte"]), b["min"], vb.ctypes.data_as(C.c_void_p), len(vb), C.byref(fl),
                                C.byref(g), C.byref(l), C.byref(mv), ov.ctypes.data_as(C.c_void_p), C.byref(n),
                                C.byref(hi))
    return dict(flags=fl.value, gte=g.value, lte=l.value, min=mv.value, values=sorted(ov[:n.value].tolist()),
                op=OPN[op]), bool(hi.value)


def same(r, e):
    f = r["flags"]
    if f != e["flags"] or r["values"] != e["values"]:
        return False
    if f & 2 and r["gte"] != e["gte"]:
        return False
    if f & 4 and r["lte"] != e["lte"]:
        return False
    if f & 8 and r["min"] != e["min"]:
        return False
    return True


def literal(e):
    flags = (1 if e["complement"] else 0) | (2 if e["gte"] is not None else 0) | (4 if e["lte"] is not None else 0) | \
            (8 if e["min_values"] is not None else 0)
    return dict(flags=flags, gte=e["gte"] or 0, lte=e["lte"] or 0, min=e["min_values"] or 0,
                values=sorted(UNIVERSE.index(v) for v in e["values"]))


def test_intersection_tables():
    assert len(KATS["intersection"]) >= 392
    sym = {k: canon(v) for k, v in KATS["symbols"].items()}
    for e in KATS["intersection"]:
        got, _ = intersect(sym[e["a"]], sym[e["b"]])
        exp = sym[e["expected_symbol"]] if "expected_symbol" in e else literal(e["expected_literal"])
        assert same(got, exp), (e, got, exp)


def test_has_intersection_agrees_with_intersection():
    """HasIntersection is documented as a cheaper Intersection (requirement.go:208-211): non-empty <=> true."""
    sym = {k: canon(v) for k, v in KATS["symbols"].items()}
    for a in sym.values():
        for b in sym.values():
            got, hi = intersect(a, b)
            nonempty = bool(got["flags"] & 1) or len(got["values"]) > 0
            assert hi == nonempty


def test_has_table():
    L = oracle_lib.lib()
    sym = {k: canon(v) for k, v in KATS["symbols"].items()}
    assert len(KATS["has"]) == 70
    for e in KATS["has"]:
        r = sym[e["r"]]
        va = np.array(r["values"], np.int32)
        got = L.orc_kat_has(VI.ctypes.data_as(C.c_void_p), IS.ctypes.data_as(C.c_void_p), len(UNIVERSE),
                            C.c_uint8(r["flags"]), C.c_int64(r["gte"]), C.c_int64(r["lte"]),
                            va.ctypes.data_as(C.c_void_p), len(va), UNIVERSE.index(e["value"]))
        assert bool(got) == e["expected"], e


def test_operator_and_len_tables():
    sym = {k: canon(v) for k, v in KATS["symbols"].items()}
    for e in KATS["operator"]:
        assert sym[e["r"]]["op"] == e["expected"], e
    for e in KATS["len"]:
        r = sym[e["r"]]
        n = (2**63 - 1 - len(r["values"])) if r["flags"] & 1 else len(r["values"])
        exp = e["expected"].replace("math.MaxInt64", str(2**63 - 1))
        assert n == eval(exp), e


def test_python_canonicalisation_matches_oracle_constructor():
    """encode.canonical_requirement (the caller-side NewRequirementWithFlexibility) against the oracle's."""
    for name, s in KATS["symbols"].items():
        o = canon(s)
        key, comp, vals, gte, lte, mv = encode.canonical_requirement(
            NodeSelectorRequirement("key", s["op"], tuple(s["values"]), s["min_values"]))
        flags = (1 if comp else 0) | (2 if gte is not None else 0) | (4 if lte is not None else 0) | (8 if mv is not None else 0)
        assert flags == o["flags"], name
        assert sorted(UNIVERSE.index(v) for v in vals) == o["values"], name
        if gte is not None:
            assert gte == o["gte"]
        if lte is not None:
            assert lte == o["lte"]


def _compat_problem():
    """One reqset per symbol of requirements_test.go (zone key, well known) for orc_kat_compatible."""
    b = encode.ProblemBuilder()
    ids = {}
    for name, s in KATS["compat_symbols"].items():
        reqs = [] if s is None else [encode.canonical_requirement(
            NodeSelectorRequirement(ZONE_LABEL, s["op"], tuple(s["values"]), s["min_values"]))]
        ids[name] = b.reqset(reqs)
    b.extra_keys.add(ZONE_LABEL)
    # reference every reqset from a template so none is pruned away
    enc = b.build()
    return enc, ids


@pytest.mark.parametrize("allow", [True, False])
def test_compatible_matrices(allow):
    enc, ids = _compat_problem()
    L = oracle_lib.lib()
    L.orc_kat_compatible.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    rows = [e for e in KATS["compatible"] if e["allow_undefined"] == allow]
    assert len(rows) == 225
    for e in rows:
        got = L.orc_kat_compatible(enc.problem.ref(), ids[e["a"]], ids[e["b"]], 1 if allow else 0)
        assert bool(got) == e["ok"], e


def test_gosort_is_a_sort_and_deterministic():
    L = oracle_lib.lib()
    rng = np.random.default_rng(7)
    for n in [0, 1, 5, 12, 13, 49, 50, 51, 200, 1000]:
        keys = rng.integers(0, 6, n).astype(np.int64)
        p1, p2 = np.zeros(n, np.int32), np.zeros(n, np.int32)
        L.orc_kat_gosort(keys.ctypes.data_as(C.c_void_p), n, p1.ctypes.data_as(C.c_void_p))
        L.orc_kat_gosort(keys.ctypes.data_as(C.c_void_p), n, p2.ctypes.data_as(C.c_void_p))
        assert np.array_equal(p1, p2)
        assert sorted(p1.tolist()) == list(range(n))
        assert np.all(np.diff(keys[p1]) >= 0)
    # n <= 12 is insertion sort == stable (sort.go insertionSort)
    keys = np.array([2, 1, 2, 1, 0, 2, 1], np.int64)
    p = np.zeros(7, np.int32)
    L.orc_kat_gosort(keys.ctypes.data_as(C.c_void_p), 7, p.ctypes.data_as(C.c_void_p))
    assert p.tolist() == [4, 1, 3, 6, 0, 2, 5]
