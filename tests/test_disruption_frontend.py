"""The disruption front-end (karpenter_b200/disruption.py): DisruptionCost, sortCandidates (Go sort.Slice order), budgets,
the <=100 cap, firstNConsolidationOption with all prefixes in one kp_consolidate call, the single-node pass and
validateCommand.  Restated reference cases: pkg/controllers/disruption/suite_test.go:850-900 (Pod Eviction Cost),
singlenodeconsolidation_test.go:100-190 (Candidate Shuffling), consolidation_test.go:3729-4207 (Multi-NodeClaim),
multinodeconsolidation.go:118-171.  CPU tier: the oracle behind the same encode / decode path; GPU tier: the CUDA path."""
import numpy as np
import pytest

from karpenter_b200 import _native, disruption, fake
from karpenter_b200.disruption import (Consolidation, MultiNodeConsolidation, SingleNodeConsolidation, disruption_cost,
                                       eviction_cost, interweave_by_nodepool, sort_candidates, validate_command)
from karpenter_b200.model import (ARCH_LABEL, CAPACITY_TYPE_LABEL, HOSTNAME_LABEL, INSTANCE_TYPE_LABEL, NODEPOOL_LABEL,
                                  OS_LABEL, ZONE_LABEL, NodePool, NodeSelectorRequirement, Pod, StateNode, quantity_units)
from tests import oracle_lib

BACKENDS = [pytest.param("oracle", id="oracle"), pytest.param("gpu", id="gpu", marks=pytest.mark.gpu)]


# ---- Pod Eviction Cost (suite_test.go:850-900) -------------------------------------------------------------------
def test_eviction_cost_kats():
    assert eviction_cost(Pod()) == 1.0
    assert eviction_cost(Pod(deletion_cost="100")) > 1.0
    assert eviction_cost(Pod(deletion_cost="-100")) < 1.0
    assert eviction_cost(Pod(deletion_cost="101")) > eviction_cost(Pod(deletion_cost="100")) > eviction_cost(Pod(deletion_cost="99"))
    assert eviction_cost(Pod(priority=1)) > 1.0
    assert eviction_cost(Pod(priority=-1)) < 1.0
    # utils/disruption/disruption.go:48-70: deletion cost / 2^27, priority / 2^25, clamp to [-10, 10]
    assert eviction_cost(Pod(deletion_cost="134217728")) == 2.0 and eviction_cost(Pod(priority=33554432)) == 2.0
    assert eviction_cost(Pod(deletion_cost="2147483647")) == 10.0
    assert eviction_cost(Pod(priority=1_000_000_000)) == 10.0
    assert eviction_cost(Pod(priority=-2147483648)) == -10.0
    assert eviction_cost(Pod(deletion_cost="not-a-number")) == 1.0


def _bare_node(name, pool, n_pods, **kw):
    return StateNode(name=name, nodepool=pool, pods=[Pod(name=f"{name}-{i}", uid=hash((name, i)) & 0xffff) for i in range(n_pods)], **kw)


def test_disruption_cost_scales_with_remaining_lifetime():  # types.go:131-135, disruption.go:36-46
    assert disruption_cost(_bare_node("a", "p", 3)) == 3.0
    assert disruption_cost(_bare_node("a", "p", 3, expire_after_s=100.0, age_s=25.0)) == 3.0 * 0.75
    assert disruption_cost(_bare_node("a", "p", 3, expire_after_s=100.0, age_s=500.0)) == 0.0


def test_sort_candidates_is_go_sort_slice():
    """sort.Slice is unstable: the order among equal costs is pdqsort's, not insertion order (60 candidates > 12, so the
    insertion-sort shortcut does not apply)."""
    nodes = [_bare_node(f"n{i:02d}", "p", 1 + (i * 7) % 4) for i in range(60)]
    got = sort_candidates(nodes)
    costs = [disruption_cost(n) for n in got]
    assert costs == sorted(costs) and sorted(n.name for n in got) == sorted(n.name for n in nodes)
    # the library's port and the oracle's independent port of Go's pdqsort agree on the permutation
    import ctypes as C
    keys = np.array([int(disruption_cost(n)) for n in nodes], np.int64)
    want = np.zeros(len(keys), np.int32)
    oracle_lib.lib().orc_kat_gosort(keys.ctypes.data_as(C.c_void_p), len(keys), want.ctypes.data_as(C.c_void_p))
    assert [n.name for n in got] == [nodes[i].name for i in want]
    stable = [n.name for n in sorted(nodes, key=disruption_cost)]
    assert [n.name for n in got] != stable  # ... and it is NOT the stable order


def test_go_sort_port_matches_oracle_port_on_random_keys():
    import ctypes as C
    rng = np.random.default_rng(3)
    for trial in range(300):
        n = int(rng.integers(0, 300))
        keys = rng.integers(0, max(2, n // int(rng.integers(1, 9)) + 1), n).astype(np.int64)
        if trial % 5 == 0:
            keys = np.sort(keys)
        if trial % 7 == 0:
            keys = np.sort(keys)[::-1].copy()
        want = np.zeros(n, np.int32)
        oracle_lib.lib().orc_kat_gosort(keys.ctypes.data_as(C.c_void_p), n, want.ctypes.data_as(C.c_void_p))
        assert np.array_equal(_native.go_sort_order(keys), want), trial
        assert np.array_equal(_native.go_sort_order(keys.astype(np.float64) * 0.5), want), trial


def test_single_node_candidates_interweave_by_nodepool():  # singlenodeconsolidation_test.go:100-190
    nodes = []
    for cost in (3, 2, 1):
        for pool in ("nodepool-1", "nodepool-2", "nodepool-3"):
            nodes.append(_bare_node(f"{pool}-c{cost}", pool, cost))
    got = interweave_by_nodepool(sort_candidates(nodes))
    assert len(got) == 9
    for grp, cost in ((got[0:3], 1.0), (got[3:6], 2.0), (got[6:9], 3.0)):
        assert {n.nodepool for n in grp} == {"nodepool-1", "nodepool-2", "nodepool-3"}
        assert all(disruption_cost(n) == cost for n in grp)
    first = interweave_by_nodepool(sort_candidates(nodes), previously_unseen=["nodepool-2"])
    assert first[0].nodepool == "nodepool-2"


# ---- Multi-NodeClaim consolidation (consolidation_test.go:3729-4207, multinodeconsolidation.go:52-171) ----------------
def _node(name, it, pod_list, zone="test-zone-1", ct="on-demand", pool="default", **kw):
    res = ["cpu", "memory", "pods"]
    used = {r: 0 for r in res}
    for p in pod_list:
        for r in res:
            used[r] += quantity_units(r, p.requests.get(r, 0)) if r != "pods" else 1
    avail = {}
    for r in res:
        a = quantity_units(r, it.capacity[r]) - quantity_units(r, it.overhead.get(r, 0)) - used[r]
        avail[r] = f"{a}m" if r == "cpu" else a
    arch = [x for x in it.requirements if x.key == ARCH_LABEL][0].values[0]
    labels = {HOSTNAME_LABEL: name, ZONE_LABEL: zone, CAPACITY_TYPE_LABEL: ct, OS_LABEL: "linux", ARCH_LABEL: arch,
              NODEPOOL_LABEL: pool, INSTANCE_TYPE_LABEL: it.name}
    cap = dict(it.capacity)
    cap["nodes"] = 1
    return StateNode(name=name, labels=labels, available=avail, capacity=cap, nodepool=pool, instance_type=it.name,
                     pods=list(pod_list), **kw)


def _engine(which, pools, its, nodes, **kw):
    if which == "oracle":
        return Consolidation(pools, {p.name: its for p in pools}, nodes, backend=oracle_lib.consolidate,
                             solve_backend=oracle_lib.solve, **kw)
    return Consolidation(pools, {p.name: its for p in pools}, nodes, **kw)


def _pods(n, uid0, cpu="1"):
    return [Pod(name=f"p{uid0 + i}", uid=uid0 + i, requests={"cpu": cpu}) for i in range(n)]


def _pool(name="default"):
    return NodePool(name=name, requirements=[NodeSelectorRequirement(CAPACITY_TYPE_LABEL, "In", ("on-demand",))])


@pytest.mark.parametrize("which", BACKENDS)
def test_multi_node_merges_three_nodes_into_one(which):  # consolidation_test.go:3834-3884 "can merge 3 nodes into 1"
    its = fake.default_instance_types()
    by = {it.name: it for it in its}
    big = by["arm-instance-type"]  # the most expensive default type
    nodes = [_node(f"node-{i}", big, _pods(1, 10 * i)) for i in range(3)]
    eng = _engine(which, [_pool()], its, nodes)
    try:
        cmd, names, constrained = MultiNodeConsolidation(eng).compute_command(nodes, {"default": 10})
    finally:
        eng.close()
    assert cmd is not None and cmd.decision == "replace" and sorted(names) == ["node-0", "node-1", "node-2"] and not constrained
    assert cmd.n_new_node_claims == 1 and "arm-instance-type" not in cmd.replacement_instance_types


@pytest.mark.parametrize("which", BACKENDS)
def test_multi_node_respects_budgets_and_skips_empty_nodes(which):  # multinodeconsolidation.go:66-82
    its = fake.default_instance_types()
    by = {it.name: it for it in its}
    big = by["arm-instance-type"]
    nodes = [_node(f"node-{i}", big, _pods(1, 10 * i)) for i in range(4)] + [_node("node-empty", big, [])]
    eng = _engine(which, [_pool()], its, nodes)
    try:
        m = MultiNodeConsolidation(eng)
        cmd, names, constrained = m.compute_command(nodes, {"default": 2})
        assert constrained and cmd is not None and len(names) == 2 and "node-empty" not in names
        cmd0, names0, constrained0 = m.compute_command(nodes, {"default": 0})
        assert cmd0 is None and names0 == [] and constrained0
    finally:
        eng.close()


@pytest.mark.parametrize("which", BACKENDS)
def test_multi_node_binary_search_equals_sequential_search(which):
    """firstNConsolidationOption reads the all-prefix table exactly as the reference's sequential binary search would: the
    chosen prefix is the one a search that simulates mid after mid ends on (multinodeconsolidation.go:132-169)."""
    its = fake.default_instance_types()
    by = {it.name: it for it in its}
    d, big = by["default-instance-type"], by["arm-instance-type"]
    nodes = []
    for i in range(9):  # cheap nodes with growing pod counts (cost order = index order), two big ones at the end
        nodes.append(_node(f"node-{i:02d}", d if i < 7 else big, _pods(1 + i // 3, 100 * i, cpu="1")))
    eng = _engine(which, [_pool()], its, nodes)
    try:
        m = MultiNodeConsolidation(eng)
        cmd, names, _ = m.compute_command(nodes, {"default": 100})
        table = m.last_prefix_table
        order = [n.name for n in sort_candidates([n for n in nodes if n.pods])]
        lo_, hi, last = 1, len(order) - 1, None
        while lo_ <= hi:  # the sequential form, one simulation per step
            mid = (lo_ + hi) // 2
            (c,) = eng.compute([order[:mid + 1]])
            if c.decision in ("delete", "replace"):
                last, lo_ = (c, order[:mid + 1]), mid + 1
            else:
                hi = mid - 1
        assert (cmd, names) == ((last[0], last[1]) if last else (None, []))
        assert len(table) == len(order) - 1
    finally:
        eng.close()


@pytest.mark.parametrize("which", BACKENDS)
def test_single_node_pass_picks_first_consolidatable_candidate(which):  # singlenodeconsolidation.go:56-131
    its = fake.default_instance_types()
    by = {it.name: it for it in its}
    small, big = by["small-instance-type"], by["arm-instance-type"]
    pa = _pods(1, 1, cpu="500m")
    pa[0].node_selector = {ARCH_LABEL: "amd64"}                 # cannot move to the arm node
    nodes = [_node("node-a", small, pa),                        # already on the cheapest type, nowhere to go: no-op
             _node("node-b", big, _pods(2, 10, cpu="500m"))]   # replaceable by something cheaper
    eng = _engine(which, [_pool()], its, nodes)
    try:
        s = SingleNodeConsolidation(eng)
        cmd, names, constrained = s.compute_command(nodes, {"default": 5})
        assert cmd is not None and names == ["node-b"] and cmd.decision in ("delete", "replace") and not constrained
        cmd, names, constrained = s.compute_command(nodes, {"default": 0})
        assert cmd is None and constrained
    finally:
        eng.close()


@pytest.mark.parametrize("which", BACKENDS)
def test_validate_command_resimulates_on_current_state(which):  # validation.go:296-356
    its = fake.default_instance_types()
    by = {it.name: it for it in its}
    d, big = by["default-instance-type"], by["arm-instance-type"]
    p = _pods(3, 1)
    nodes = [_node("node-1", d, p[:2]), _node("node-2", d, p[2:])]
    eng = _engine(which, [_pool()], its, nodes)
    try:
        (cmd,) = eng.compute([["node-2"]])
        assert cmd.decision == "delete" and validate_command(eng, cmd, ["node-2"])
    finally:
        eng.close()
    # the cluster changed while the command waited: node-1 filled up, the pod of node-2 now needs a NodeClaim
    nodes2 = [_node("node-1", d, _pods(3, 50)), _node("node-2", d, p[2:])]
    eng2 = _engine(which, [_pool()], its, nodes2)
    try:
        assert not validate_command(eng2, cmd, ["node-2"])
        # a replace command stays valid while the simulation's single NodeClaim still offers every type it would launch
        nodes3 = [_node("node-9", big, _pods(1, 70))]
        eng3 = _engine(which, [_pool()], its, nodes3)
        try:
            (rep,) = eng3.compute([["node-9"]])
            assert rep.decision == "replace" and validate_command(eng3, rep, ["node-9"])
            rep.replacement_instance_types = rep.replacement_instance_types + ["no-such-type"]
            assert not validate_command(eng3, rep, ["node-9"])
        finally:
            eng3.close()
    finally:
        eng2.close()
