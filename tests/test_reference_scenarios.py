"""Scenarios of the reference's own scheduling test-suite, replayed with the reference's fake cloud-provider catalog.

Each case restates one `It(...)` of pkg/controllers/provisioning/scheduling/{suite_test.go,topology_test.go} with the
expected outcome the reference asserts (instance type picked, number of nodes, zonal skew, unschedulable pods).  They
pin the ORACLE (oracle/) to reference behaviour beyond the requirement-algebra known-answer tables: the CPU tier runs
the scenarios through the oracle, the GPU tier runs the same scenarios through the CUDA path and additionally demands
bit-identical results.  The fake provider launches the cheapest instance type of a NodeClaim's options
(fake/cloudprovider.go Create), which is what `cheapest()` evaluates.
"""
from collections import Counter

import pytest

from karpenter_b200 import fake
from karpenter_b200.model import (ARCH_LABEL, CAPACITY_TYPE_LABEL, HOSTNAME_LABEL, ZONE_LABEL, LabelSelector, NodePool,
                                  NodeSelectorRequirement, Pod, PodAffinityTerm, TopologySpreadConstraint)
from karpenter_b200.scheduler import Scheduler
from tests import oracle_lib

PRICE = {it.name: fake.price_from_resources(it.capacity) for it in fake.default_instance_types()}


def nodepool(**kw):
    """suite_test.go:135-149: capacity-type In [spot, on-demand, reserved]; test.NodePool() default limit cpu 2000."""
    reqs = [NodeSelectorRequirement(CAPACITY_TYPE_LABEL, "In", ("spot", "on-demand", "reserved"))]
    reqs += kw.pop("requirements", [])
    return NodePool(name="default", requirements=reqs, limits={"cpu": "2000"}, **kw)


def pods(n, uid0=1, **kw):
    return [Pod(name=f"p{uid0 + i}", uid=uid0 + i, **kw) for i in range(n)]


def cheapest(claim):
    return min(claim.instance_type_options, key=lambda n: (PRICE[n], n))


def zone_of(claim):
    z = claim.requirements[ZONE_LABEL]
    assert not z["complement"] and len(z["values"]) == 1, z
    return z["values"][0]


def solve(backend, pod_list, np_=None, its=None):
    np_ = np_ or nodepool()
    its = its or fake.default_instance_types()
    s = Scheduler([np_], {np_.name: its}, backend=backend)
    try:
        return s.solve(pod_list)
    finally:
        s.close()


def gpu_and_oracle(pod_list, np_=None, its=None):
    """CUDA path result, checked bit for bit against the oracle on the same encoding."""
    from tests.parity import assert_same
    gpu = solve(None, pod_list, np_, its)
    orc = solve(oracle_lib.solve, pod_list, np_, its)
    assert_same(gpu.raw, orc.raw, "scenario ")
    return gpu


BACKENDS = [pytest.param("oracle", id="oracle"), pytest.param("gpu", id="gpu", marks=pytest.mark.gpu)]


def run(which, pod_list, np_=None, its=None):
    return solve(oracle_lib.solve, pod_list, np_, its) if which == "oracle" else gpu_and_oracle(pod_list, np_, its)


# ---- Binpacking (suite_test.go:1520-1836) -------------------------------------------------------------------------
@pytest.mark.parametrize("which", BACKENDS)
def test_small_pod_on_smallest_instance(which):  # suite_test.go:1521-1533
    r = run(which, pods(1, requests={"memory": "100M"}))
    assert len(r.new_node_claims) == 1 and not r.pod_errors
    assert cheapest(r.new_node_claims[0]) == "small-instance-type"


@pytest.mark.parametrize("which", BACKENDS)
def test_small_pod_on_smallest_possible_instance(which):  # suite_test.go:1534-1546
    r = run(which, pods(1, requests={"memory": "2000M"}))
    assert cheapest(r.new_node_claims[0]) == "small-instance-type"


@pytest.mark.parametrize("which", BACKENDS)
def test_multiple_small_pods_share_one_small_node(which):  # suite_test.go:1573-1592
    r = run(which, pods(5, requests={"memory": "10M"}))
    assert len(r.new_node_claims) == 1 and len(r.new_node_claims[0].pods) == 5
    assert cheapest(r.new_node_claims[0]) == "small-instance-type"


@pytest.mark.parametrize("which", BACKENDS)
def test_new_nodes_when_at_capacity(which):  # suite_test.go:1593-1611: 40 x 1.8G on 4Gi types -> 20 nodes
    r = run(which, pods(40, requests={"memory": "1.8G"}, node_selector={ARCH_LABEL: "amd64"}))
    assert len(r.new_node_claims) == 20 and not r.pod_errors
    assert all(cheapest(c) == "default-instance-type" and len(c.pods) == 2 for c in r.new_node_claims)


@pytest.mark.parametrize("which", BACKENDS)
def test_pack_small_and_large_together(which):  # suite_test.go:1612-1643: 40 large + 20 small -> still 20 nodes
    sel = {ARCH_LABEL: "amd64"}
    pl = pods(40, requests={"memory": "1.8G"}, node_selector=sel) + pods(20, uid0=100, requests={"memory": "400M"},
                                                                          node_selector=sel)
    r = run(which, pl)
    assert len(r.new_node_claims) == 20 and not r.pod_errors
    assert all(cheapest(c) == "default-instance-type" for c in r.new_node_claims)


@pytest.mark.parametrize("which", BACKENDS)
def test_zero_quantity_unknown_resource_schedules(which):  # suite_test.go:1670-1681
    r = run(which, pods(1, requests={"foo.com/weird-resources": "0"}))
    assert len(r.new_node_claims) == 1 and not r.pod_errors


@pytest.mark.parametrize("which", BACKENDS)
def test_pod_exceeding_every_instance_type(which):  # suite_test.go:1682-1692
    p = pods(1, requests={"memory": "2Ti"})
    r = run(which, p)
    assert not r.new_node_claims and id(p[0]) in r.pod_errors


@pytest.mark.parametrize("which", BACKENDS)
def test_pod_limit_per_node(which):  # suite_test.go:1693-1714: 25 tiny pods, 5 pods per node -> 5 small nodes
    # the reference asks for memory "1m" (a milli-byte); the encoder's exact unit for memory is the byte, so "1" here
    r = run(which, pods(25, requests={"memory": "1", "cpu": "1m"}, node_selector={ARCH_LABEL: "amd64"}))
    assert len(r.new_node_claims) == 5 and all(len(c.pods) == 5 for c in r.new_node_claims)
    assert all(cheapest(c) == "small-instance-type" for c in r.new_node_claims)


# ---- Topology (topology_test.go:107-651) ---------------------------------------------------------------------------
LABELS = {"test": "test"}


def topo_nodepool(requirements=()):
    """topology_test.go:41-56: capacity-type Exists (so the domain universe is what the instance types offer)."""
    return NodePool(name="default", requirements=[NodeSelectorRequirement(CAPACITY_TYPE_LABEL, "Exists")] +
                    list(requirements), limits={"cpu": "2000"})


def spread(key=ZONE_LABEL, max_skew=1, selector=None):
    return [TopologySpreadConstraint(max_skew, key, selector or LabelSelector.of(LABELS))]


@pytest.mark.parametrize("which", BACKENDS)
def test_zonal_balance_match_labels(which):  # topology_test.go:108-121: 4 pods over 3 zones -> skew (1, 1, 2)
    r = run(which, pods(4, labels=LABELS, topology_spread_constraints=spread()), topo_nodepool())
    skew = Counter(zone_of(c) for c in r.new_node_claims for _ in c.pods)
    assert sorted(skew.values()) == [1, 1, 2] and not r.pod_errors


@pytest.mark.parametrize("which", BACKENDS)
def test_zonal_balance_match_expressions(which):  # topology_test.go:122-143
    sel = LabelSelector.of(None, [("test", "In", ("test",))])
    r = run(which, pods(4, labels=LABELS, topology_spread_constraints=spread(selector=sel)), topo_nodepool())
    skew = Counter(zone_of(c) for c in r.new_node_claims for _ in c.pods)
    assert sorted(skew.values()) == [1, 1, 2]


@pytest.mark.parametrize("which", BACKENDS)
def test_zonal_respects_nodepool_subset(which):  # topology_test.go:160-176: two zones allowed -> (2, 2)
    np_ = topo_nodepool([NodeSelectorRequirement(ZONE_LABEL, "In", ("test-zone-1", "test-zone-2"))])
    r = run(which, pods(4, labels=LABELS, topology_spread_constraints=spread()), np_)
    skew = Counter(zone_of(c) for c in r.new_node_claims for _ in c.pods)
    assert sorted(skew.values()) == [2, 2]


@pytest.mark.parametrize("which", BACKENDS)
def test_hostname_spread_one_pod_per_node(which):  # topology_test.go:544-557: hostname spread maxSkew 1 -> (1, 1, 1, 1)
    r = run(which, pods(4, labels=LABELS, topology_spread_constraints=spread(key=HOSTNAME_LABEL)), topo_nodepool())
    assert sorted(len(c.pods) for c in r.new_node_claims) == [1, 1, 1, 1]


@pytest.mark.parametrize("which", BACKENDS)
def test_hostname_anti_affinity_separates_pods(which):  # topology_test.go:2240-2260 (anti-affinity on hostname)
    anti = [PodAffinityTerm(LabelSelector.of(LABELS), HOSTNAME_LABEL)]
    r = run(which, pods(3, labels=LABELS, pod_anti_affinity=anti), topo_nodepool())
    assert sorted(len(c.pods) for c in r.new_node_claims) == [1, 1, 1] and not r.pod_errors


@pytest.mark.parametrize("which", BACKENDS)
def test_capacity_type_spread(which):  # topology_test.go:652-666: spread over capacity types -> (2, 2)
    r = run(which, pods(4, labels=LABELS, topology_spread_constraints=spread(key=CAPACITY_TYPE_LABEL)), topo_nodepool())
    skew = Counter()
    for c in r.new_node_claims:
        ct = c.requirements[CAPACITY_TYPE_LABEL]
        assert len(ct["values"]) == 1
        skew[ct["values"][0]] += len(c.pods)
    assert sorted(skew.values()) == [2, 2]


# ---- Custom constraints / well-known labels / operators (suite_test.go:152-420) -----------------------------------
def req(key, op, *values):
    return NodeSelectorRequirement(key, op, tuple(values))


def scheduled_label(r, key):
    assert len(r.new_node_claims) == 1 and not r.pod_errors
    v = r.new_node_claims[0].requirements[key]
    assert not v["complement"] and len(v["values"]) == 1
    return v["values"][0]


def int_label_of_cheapest(claim):
    """the integer label of the instance type the fake provider would launch"""
    its = {it.name: it for it in fake.default_instance_types()}
    r = [x for x in its[cheapest(claim)].requirements if x.key == fake.INTEGER_INSTANCE_LABEL][0]
    return r.values[0]


@pytest.mark.parametrize("which", BACKENDS)
def test_nodepool_labels_unconstrained_pod(which):  # suite_test.go:154-161
    r = run(which, pods(1), nodepool(labels={"test-key": "test-value"}))
    assert scheduled_label(r, "test-key") == "test-value"


@pytest.mark.parametrize("which", BACKENDS)
def test_conflicting_node_selector(which):  # suite_test.go:162-170
    p = pods(1, node_selector={"test-key": "different-value"})
    r = run(which, p, nodepool(labels={"test-key": "test-value"}))
    assert not r.new_node_claims and id(p[0]) in r.pod_errors


@pytest.mark.parametrize("which", BACKENDS)
def test_node_selector_with_undefined_key(which):  # suite_test.go:171-178
    p = pods(1, node_selector={"test-key": "test-value"})
    r = run(which, p)
    assert not r.new_node_claims and id(p[0]) in r.pod_errors


@pytest.mark.parametrize("which", BACKENDS)
def test_matching_requirements(which):  # suite_test.go:179-190
    r = run(which, pods(1, node_affinity_required=[[req("test-key", "In", "test-value", "another-value")]]),
            nodepool(labels={"test-key": "test-value"}))
    assert scheduled_label(r, "test-key") == "test-value"


@pytest.mark.parametrize("which", BACKENDS)
def test_conflicting_requirements(which):  # suite_test.go:191-201
    p = pods(1, node_affinity_required=[[req("test-key", "In", "another-value")]])
    r = run(which, p, nodepool(labels={"test-key": "test-value"}))
    assert not r.new_node_claims and id(p[0]) in r.pod_errors


@pytest.mark.parametrize("which", BACKENDS)
def test_nodepool_zone_constraint(which):  # suite_test.go:204-212
    r = run(which, pods(1), nodepool(requirements=[req(ZONE_LABEL, "In", "test-zone-2")]))
    assert scheduled_label(r, ZONE_LABEL) == "test-zone-2"


@pytest.mark.parametrize("which", BACKENDS)
def test_node_selector_zone(which):  # suite_test.go:213-223
    r = run(which, pods(1, node_selector={ZONE_LABEL: "test-zone-2"}),
            nodepool(requirements=[req(ZONE_LABEL, "In", "test-zone-1", "test-zone-2")]))
    assert scheduled_label(r, ZONE_LABEL) == "test-zone-2"


@pytest.mark.parametrize("which", BACKENDS)
@pytest.mark.parametrize("zone", ["unknown", "test-zone-2"])
def test_node_selector_outside_nodepool(which, zone):  # suite_test.go:232-251
    p = pods(1, node_selector={ZONE_LABEL: zone})
    r = run(which, p, nodepool(requirements=[req(ZONE_LABEL, "In", "test-zone-1")]))
    assert not r.new_node_claims and id(p[0]) in r.pod_errors


@pytest.mark.parametrize("which", BACKENDS)
def test_requirement_operator_in(which):  # suite_test.go:252-262
    r = run(which, pods(1, node_affinity_required=[[req(ZONE_LABEL, "In", "test-zone-3")]]))
    assert scheduled_label(r, ZONE_LABEL) == "test-zone-3"


@pytest.mark.parametrize("which", BACKENDS)
@pytest.mark.parametrize("op,operand,expect", [("Gt", "8", "16"), ("Lt", "8", "2"), ("Gte", "16", "16"), ("Lte", "2", "2")])
def test_requirement_integer_operators(which, op, operand, expect):  # suite_test.go:263-298
    r = run(which, pods(1), nodepool(requirements=[req(fake.INTEGER_INSTANCE_LABEL, op, operand)]))
    assert len(r.new_node_claims) == 1 and not r.pod_errors
    assert int_label_of_cheapest(r.new_node_claims[0]) == expect


@pytest.mark.parametrize("which", BACKENDS)
def test_requirement_operator_not_in(which):  # suite_test.go:309-319
    r = run(which, pods(1, node_affinity_required=[[req(ZONE_LABEL, "NotIn", "test-zone-1", "test-zone-2", "unknown")]]))
    assert len(r.new_node_claims) == 1 and not r.pod_errors
    z = r.new_node_claims[0].requirements[ZONE_LABEL]  # the NodeClaim keeps NotIn [...]; only test-zone-3 offerings remain
    assert z["complement"] and {"test-zone-1", "test-zone-2"} <= set(z["values"]) and "test-zone-3" not in z["values"]


@pytest.mark.parametrize("which", BACKENDS)
def test_incompatible_requirements_in(which):  # suite_test.go:299-308
    p = pods(1, node_affinity_required=[[req(ZONE_LABEL, "In", "unknown")]])
    r = run(which, p)
    assert not r.new_node_claims and id(p[0]) in r.pod_errors


@pytest.mark.parametrize("which", BACKENDS)
def test_compatible_pods_share_a_node_incompatible_do_not(which):  # suite_test.go:625-664
    a = pods(1, node_affinity_required=[[req(ZONE_LABEL, "In", "test-zone-1", "test-zone-2")]])
    b = pods(1, uid0=2, node_affinity_required=[[req(ZONE_LABEL, "In", "test-zone-2", "test-zone-3")]])
    r = run(which, a + b)
    assert len(r.new_node_claims) == 1 and scheduled_label(r, ZONE_LABEL) == "test-zone-2"
    c = pods(1, uid0=3, node_affinity_required=[[req(ZONE_LABEL, "In", "test-zone-1")]])
    d = pods(1, uid0=4, node_affinity_required=[[req(ZONE_LABEL, "In", "test-zone-3")]])
    r = run(which, c + d)
    assert len(r.new_node_claims) == 2


# ---- Consolidation (pkg/controllers/disruption/consolidation_test.go) ----------------------------------------------
def _node(name, it, zone="test-zone-1", ct="on-demand", pod_list=(), initialized=True):
    """A managed, initialized node of instance type `it` holding `pod_list` (StateNode.Available = allocatable - requests)."""
    from karpenter_b200.model import INSTANCE_TYPE_LABEL, NODEPOOL_LABEL, OS_LABEL, StateNode, quantity_units
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
              NODEPOOL_LABEL: "default", INSTANCE_TYPE_LABEL: it.name}
    cap = dict(it.capacity)
    cap["nodes"] = 1
    return StateNode(name=name, labels=labels, available=avail, capacity=cap, nodepool="default", instance_type=it.name,
                     pods=list(pod_list), initialized=initialized)


def consolidate(which, nodes, candidate_sets, its=None, np_=None, **kw):
    from karpenter_b200 import _abi
    from karpenter_b200.disruption import Consolidation
    import numpy as np
    its = its or fake.default_instance_types()
    np_ = np_ or nodepool()
    orc = Consolidation([np_], {np_.name: its}, nodes, backend=oracle_lib.consolidate, **kw)
    cmds = orc.compute(candidate_sets)
    if which == "gpu":
        gpu = Consolidation([np_], {np_.name: its}, nodes, **kw)
        try:
            got = gpu.compute(candidate_sets)
        finally:
            gpu.close()
        for k in _abi.CONSOL_PARITY_KEYS:
            assert np.array_equal(gpu.raw[k], orc.raw[k]), k
        assert got == cmds
        return got
    return cmds


# SimulateScheduling schedules pending pods and the pods of deleting nodes together with the candidates' (helpers.go:65-91)
@pytest.mark.parametrize("which", BACKENDS)
def test_simulation_pending_pod_takes_the_room(which):
    its = {it.name: it for it in fake.default_instance_types()}
    d = its["default-instance-type"]  # 4 cpu: 3.9 allocatable
    p = pods(4, requests={"cpu": "1"})
    nodes = [_node("node-1", d, pod_list=p[:2]), _node("node-2", d, pod_list=p[2:3])]
    # alone, node-2's pod moves to node-1: delete.  With a pending pod of the same size (older uid: queued first) the room
    # on node-1 is gone and the simulation has to open a NodeClaim.
    (alone,) = consolidate(which, nodes, [["node-2"]])
    assert alone.decision == "delete"
    pending = pods(1, uid0=0, requests={"cpu": "1"})
    (cmd,) = consolidate(which, nodes, [["node-2"]], pending_pods=pending)
    assert cmd.n_new_node_claims == 1 and cmd.decision != "delete"


@pytest.mark.parametrize("which", BACKENDS)
def test_simulation_ignores_errors_of_pending_pods_only(which):  # AllNonPendingPodsScheduled, scheduler.go:330-334
    its = {it.name: it for it in fake.default_instance_types()}
    d = its["default-instance-type"]
    p = pods(3, requests={"cpu": "1"})
    nodes = [_node("node-1", d, pod_list=p[:2]), _node("node-2", d, pod_list=p[2:])]
    stuck = pods(1, uid0=90, requests={"cpu": "1"}, node_selector={ZONE_LABEL: "no-such-zone"})
    (cmd,) = consolidate(which, nodes, [["node-2"]], pending_pods=stuck)
    assert cmd.decision == "delete" and cmd.n_unscheduled == 0       # a pending pod that cannot schedule blocks nothing
    (cmd,) = consolidate(which, nodes, [["node-2"]], deleting_node_pods=stuck)
    assert cmd.decision == "noop" and cmd.n_unscheduled == 1          # the same pod coming off a deleting node does


@pytest.mark.parametrize("which", BACKENDS)
def test_simulation_deleting_node_pod_on_uninitialized_node_is_fine(which):  # helpers.go:121-140
    its = {it.name: it for it in fake.default_instance_types()}
    d = its["default-instance-type"]
    p = pods(2, requests={"cpu": "3"})
    nodes = [_node("node-1", d, pod_list=[], initialized=False), _node("node-2", d, pod_list=p[:1])]
    # node-2's pod can only go to the uninitialized node-1: that is an error for a candidate's pod ...
    (cmd,) = consolidate(which, nodes, [["node-2"]])
    assert cmd.decision == "noop" and cmd.n_unscheduled == 1
    # ... but not for the pod of a node that is already being deleted (it lands there in the simulation; the candidate's
    # own pod then needs a NodeClaim)
    (cmd,) = consolidate(which, nodes, [["node-2"]], deleting_node_pods=pods(1, uid0=0, requests={"cpu": "3"}))
    assert cmd.n_unscheduled == 0 and cmd.n_new_node_claims == 1


@pytest.mark.parametrize("which", BACKENDS)
def test_replacement_is_pinned_to_spot_when_od_goes_to_od_or_spot(which):  # consolidation.go:206-214
    from karpenter_b200.model import NodePool
    its = fake.default_instance_types()
    by = {it.name: it for it in its}
    np_ = NodePool(name="default", requirements=[req(CAPACITY_TYPE_LABEL, "In", "on-demand", "spot")])
    p = pods(1, requests={"cpu": "1"})
    nodes = [_node("node-1", by["default-instance-type"], ct="on-demand", pod_list=p)]
    (cmd,) = consolidate(which, nodes, [["node-1"]], np_=np_, price_order=True)
    assert cmd.decision == "replace" and cmd.replacement_nodepool == "default"
    ct = cmd.replacement_requirements[CAPACITY_TYPE_LABEL]
    assert not ct["complement"] and ct["values"] == ["spot"]
    assert cmd.replacement_requests["cpu"] >= 1000 and cmd.replacement_requests["pods"] == 1
    prices = [min(o.price for o in by[n].offerings if o.available) for n in cmd.replacement_instance_types]
    assert prices == sorted(prices) and len(prices) >= 1               # OrderByPrice order survives the round trip


@pytest.mark.parametrize("which", BACKENDS)
def test_consolidation_can_delete_nodes(which):  # consolidation_test.go:2407-2447: the lone pod fits on the other node
    its = {it.name: it for it in fake.default_instance_types()}
    d = its["default-instance-type"]
    p = pods(3, requests={"cpu": "1"})
    nodes = [_node("node-1", d, pod_list=p[:2]), _node("node-2", d, pod_list=p[2:])]
    (cmd,) = consolidate(which, nodes, [["node-2"]])
    assert cmd.decision == "delete" and cmd.n_new_node_claims == 0 and cmd.n_unscheduled == 0


@pytest.mark.parametrize("which", BACKENDS)
def test_consolidation_replaces_with_cheaper_node(which):  # consolidation_test.go "can replace node" (977-1032 family)
    its = {it.name: it for it in fake.default_instance_types()}
    big = its["arm-instance-type"]  # 16 cpu / 128Gi: price 0.1*16 + 0.1*137.4 = 15.3
    p = pods(1, requests={"cpu": "1"}, node_selector={ARCH_LABEL: "amd64"})
    # an amd64 pod cannot really sit on the arm node; what matters is the simulation: it needs one new, cheaper amd64 node
    nodes = [_node("node-1", big, pod_list=p)]
    (cmd,) = consolidate(which, nodes, [["node-1"]])
    assert cmd.decision == "replace" and cmd.n_new_node_claims == 1
    assert "small-instance-type" in cmd.replacement_instance_types
    assert "arm-instance-type" not in cmd.replacement_instance_types


@pytest.mark.parametrize("which", BACKENDS)
def test_consolidation_wont_replace_with_more_expensive(which):  # consolidation_test.go:2283-2406
    def it(name, offerings):
        return fake.new_instance_type(name, offerings=[
            __import__("karpenter_b200.model", fromlist=["Offering"]).Offering(
                [req(CAPACITY_TYPE_LABEL, "In", ct), req(ZONE_LABEL, "In", z)], price, avail) for ct, z, price, avail in offerings])
    current = it("current-on-demand", [("on-demand", "test-zone-1a", 0.5, False)])
    replacement = it("on-demand-replacement", [("on-demand", "test-zone-1a", 0.6, True), ("on-demand", "test-zone-1b", 0.6, True),
                                               ("spot", "test-zone-1b", 0.2, True), ("spot", "test-zone-1c", 0.3, True)])
    p = pods(1, requests={"cpu": "1"})
    nodes = [_node("node-1", current, zone="test-zone-1a", pod_list=p)]
    # the reference pins the NodePool to on-demand for this test (consolidation_test.go:2335-2345)
    from karpenter_b200.disruption import Consolidation
    np_ = NodePool(name="default", requirements=[req(CAPACITY_TYPE_LABEL, "In", "on-demand")], limits={"cpu": "2000"})
    c = Consolidation([np_], {"default": [current, replacement]}, nodes, backend=oracle_lib.consolidate)
    (cmd,) = c.compute([["node-1"]])
    assert cmd.decision == "noop" and cmd.n_new_node_claims == 1  # 0.6 is not cheaper than 0.5
    if which == "gpu":
        g = Consolidation([np_], {"default": [current, replacement]}, nodes)
        try:
            (gc,) = g.compute([["node-1"]])
        finally:
            g.close()
        assert gc == cmd


@pytest.mark.parametrize("which", BACKENDS)
def test_consolidation_wont_delete_onto_uninitialized_node(which):  # consolidation_test.go:2990-3035
    its = {it.name: it for it in fake.default_instance_types()}
    d = its["default-instance-type"]
    p = pods(3, requests={"cpu": "1"})
    nodes = [_node("node-1", d, pod_list=p[:2], initialized=False), _node("node-2", d, pod_list=p[2:])]
    (cmd,) = consolidate(which, nodes, [["node-2"]])
    assert cmd.decision == "noop" and cmd.n_unscheduled == 1


@pytest.mark.parametrize("which", BACKENDS)
def test_consolidation_merges_three_nodes_into_one(which):  # consolidation_test.go:3823 family: 3 nodes -> 1 replacement
    its = {it.name: it for it in fake.default_instance_types()}
    d = its["default-instance-type"]
    p = pods(3, requests={"cpu": "1"})
    nodes = [_node(f"node-{i+1}", d, pod_list=p[i:i + 1]) for i in range(3)]
    cmds = consolidate(which, nodes, [["node-1", "node-2", "node-3"], ["node-1"], ["node-1", "node-2"]])
    assert cmds[0].decision == "replace" and cmds[0].n_new_node_claims == 1  # three 0.83-priced nodes -> one of them
    assert cmds[1].decision == "delete" and cmds[2].decision == "delete"    # the others still have room


# ---- edge cases: empty input, no usable NodePool ---------------------------------------------------------------------
@pytest.mark.parametrize("which", BACKENDS)
def test_empty_pod_list(which):
    r = run(which, [])
    assert not r.new_node_claims and not r.pod_errors and not r.existing_nodes


@pytest.mark.parametrize("which", BACKENDS)
def test_nodepool_requirements_filter_out_every_instance_type(which):  # scheduler.go:147-160, 510-512
    p = pods(3, requests={"cpu": "1"})
    r = run(which, p, nodepool(requirements=[req(ZONE_LABEL, "In", "no-such-zone")]))
    assert not r.new_node_claims and len(r.pod_errors) == 3
    assert all("filtered out all" in m for m in r.pod_errors.values())


@pytest.mark.parametrize("which", BACKENDS)
def test_mixed_schedulable_and_unschedulable(which):
    ok = pods(4, requests={"cpu": "1"})
    bad = pods(2, uid0=50, requests={"cpu": "100"})  # larger than every instance type
    r = run(which, ok + bad)
    assert {id(x) for x in bad} == set(r.pod_errors) and sum(len(c.pods) for c in r.new_node_claims) == 4


@pytest.mark.parametrize("which", BACKENDS)
def test_consolidation_replace_keeps_zonal_spread(which):  # consolidation_test.go:4333-4406
    its = {it.name: it for it in fake.default_instance_types()}
    d = its["default-instance-type"]
    tsc = spread()  # zone, maxSkew 1, selector LABELS
    p = pods(3, labels=LABELS, requests={"cpu": "1"}, topology_spread_constraints=tsc)
    nodes = [_node(f"node-{i+1}", d, zone=f"test-zone-{i+1}", pod_list=p[i:i + 1]) for i in range(3)]
    (cmd,) = consolidate(which, nodes, [["node-3"]])
    # the evicted pod may only land in test-zone-3 (the other zones would skew 2:1:0): not on node-1 / node-2, so a
    # cheaper replacement in that zone is launched
    assert cmd.decision == "replace" and cmd.n_new_node_claims == 1 and cmd.n_unscheduled == 0
    assert "small-instance-type" in cmd.replacement_instance_types


@pytest.mark.parametrize("which", BACKENDS)
def test_consolidation_wont_delete_when_anti_affinity_forbids(which):  # consolidation_test.go:4407-4467
    its = {it.name: it for it in fake.default_instance_types()}
    small = its["small-instance-type"]
    anti = [PodAffinityTerm(LabelSelector.of(LABELS), HOSTNAME_LABEL)]
    p = pods(3, labels=LABELS, requests={"cpu": "1"}, pod_anti_affinity=anti)
    nodes = [_node(f"node-{i+1}", small, pod_list=p[i:i + 1]) for i in range(3)]
    cmds = consolidate(which, nodes, [["node-1"], ["node-1", "node-2"]])
    # the pod cannot join its peers; a new node of the cheapest type is not cheaper than the one it leaves
    assert cmds[0].decision == "noop" and cmds[0].n_new_node_claims == 1
    assert cmds[1].decision == "noop" and cmds[1].n_new_node_claims == 2
