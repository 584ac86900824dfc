"""Further cases of the reference's scheduling suite (pkg/controllers/provisioning/scheduling/suite_test.go): instance
type compatibility, bin-packing, in-flight / existing nodes across provisioning passes.  Same harness as
test_reference_topology.py (`Cluster`); CPU tier through the oracle, GPU tier bit-identical to it.
"""
import random

import pytest

from karpenter_b200 import fake
from karpenter_b200.model import (ARCH_LABEL, CAPACITY_TYPE_LABEL, HOSTNAME_LABEL, INSTANCE_TYPE_LABEL, OS_LABEL, ZONE_LABEL,
                                  LabelSelector, NodePool, Offering, Taint, Toleration)
from tests.test_reference_scenarios import BACKENDS, req
from tests.test_reference_topology import Cluster, spread, _pool

GPU1, GPU2 = "karpenter.sh/super-great-gpu", "karpenter.sh/even-better-gpu"


def nodes_of(c, pods_):
    return {c.bound[id(p)] for p in pods_}


# ---- Instance Type Compatibility (suite_test.go:1246-1519) -----------------------------------------------------------
@pytest.mark.parametrize("which", BACKENDS)
def test_more_resources_than_any_instance_type(which):  # suite_test.go:1247-1257
    c = Cluster(which)
    pl = c.pods(1, requests={"cpu": "512"})
    c.provision(pl)
    assert not c.scheduled(pl[0])


@pytest.mark.parametrize("which", BACKENDS)
def test_different_archs_on_different_instances(which):  # suite_test.go:1258-1282
    c = Cluster(which, pools=[_pool(requirements=[req(ARCH_LABEL, "In", "arm64", "amd64")])])
    pl = c.pods(1, node_selector={ARCH_LABEL: "amd64"}) + c.pods(1, node_selector={ARCH_LABEL: "arm64"})
    c.provision(pl)
    assert len(nodes_of(c, pl)) == 2


@pytest.mark.parametrize("which", BACKENDS)
def test_node_affinity_on_instance_type(which):  # suite_test.go:1283-1303
    c = Cluster(which, pools=[_pool(requirements=[req(ARCH_LABEL, "In", "arm64", "amd64")])])
    pl = c.pods(1, node_affinity_required=[[req(INSTANCE_TYPE_LABEL, "In", "arm-instance-type")]])
    c.provision(pl)
    assert c.node_of(pl[0]).labels[INSTANCE_TYPE_LABEL] == "arm-instance-type"


@pytest.mark.parametrize("which", BACKENDS)
def test_node_affinity_on_operating_system(which):  # suite_test.go:1304-1325: only the arm type offers ios
    c = Cluster(which, pools=[_pool(requirements=[req(ARCH_LABEL, "In", "arm64", "amd64")])])
    pl = c.pods(1, node_affinity_required=[[req(OS_LABEL, "In", "ios")]])
    c.provision(pl)
    assert c.node_of(pl[0]).labels[INSTANCE_TYPE_LABEL] == "arm-instance-type"


@pytest.mark.parametrize("which", BACKENDS)
def test_different_instance_type_selectors(which):  # suite_test.go:1366-1390
    c = Cluster(which, pools=[_pool(requirements=[req(ARCH_LABEL, "In", "arm64", "amd64")])])
    pl = c.pods(1, node_selector={INSTANCE_TYPE_LABEL: "small-instance-type"})
    pl += c.pods(1, node_selector={INSTANCE_TYPE_LABEL: "default-instance-type"})
    c.provision(pl)
    assert {c.node_of(p).labels[INSTANCE_TYPE_LABEL] for p in pl} == {"small-instance-type", "default-instance-type"}


@pytest.mark.parametrize("which", BACKENDS)
def test_different_zone_selectors(which):  # suite_test.go:1391-1415
    c = Cluster(which)
    pl = c.pods(1, node_selector={ZONE_LABEL: "test-zone-1"}) + c.pods(1, node_selector={ZONE_LABEL: "test-zone-2"})
    c.provision(pl)
    assert [c.node_of(p).labels[ZONE_LABEL] for p in pl] == ["test-zone-1", "test-zone-2"]


def _gpu_types():
    its = fake.instance_types(5)
    its[0].capacity[GPU1] = "25"
    its[1].capacity[GPU2] = "25"
    return its


@pytest.mark.parametrize("which", BACKENDS)
def test_resources_that_no_single_type_has(which):  # suite_test.go:1416-1462
    c = Cluster(which, its=_gpu_types())
    pl = c.pods(1, requests={GPU1: "1"}) + c.pods(1, requests={GPU2: "1"})
    c.provision(pl)
    assert len(nodes_of(c, pl)) == 2
    both = c.pods(1, requests={GPU1: "1", GPU2: "1"})
    c.provision(both)
    assert not c.scheduled(both[0])


# ---- Binpacking (suite_test.go:1644-1836) ----------------------------------------------------------------------------
@pytest.mark.parametrize("which", BACKENDS)
def test_pack_nodes_tightly(which):  # suite_test.go:1644-1669
    c = Cluster(which, its=fake.instance_types(5))
    pl = c.pods(1, requests={"cpu": "4.5"}) + c.pods(1, requests={"cpu": "1"})
    c.provision(pl)
    types = [c.node_of(p).labels[INSTANCE_TYPE_LABEL] for p in pl]
    assert len(nodes_of(c, pl)) == 2 and types[0] != types[1]


@pytest.mark.parametrize("which", BACKENDS)
def test_valid_instance_types_regardless_of_price(which):  # suite_test.go:1762-1835
    def it(name, cpu, price):
        off = [Offering([req(CAPACITY_TYPE_LABEL, "In", "on-demand"), req(ZONE_LABEL, "In", "test-zone-1a")], price, True)]
        return fake.new_instance_type(name, {"cpu": str(cpu), "memory": f"{cpu}Gi"}, offerings=off)
    c = Cluster(which, its=[it("medium", 2, 3.0), it("small", 1, 2.0), it("large", 4, 1.0)])
    pl = c.pods(1, requests={"cpu": "1m", "memory": "1Mi"})
    r = c.provision(pl)
    assert sorted(r.new_node_claims[0].instance_type_options) == ["large", "medium", "small"]


# ---- In-Flight Nodes (suite_test.go:1837-2478) -----------------------------------------------------------------------
@pytest.mark.parametrize("which", BACKENDS)
def test_in_flight_node_takes_the_next_pod(which):  # suite_test.go:1838-1854
    c = Cluster(which)
    first = c.pods(1, requests={"cpu": "10m"})
    c.provision(first)
    second = c.pods(1, requests={"cpu": "10m"})
    c.provision(second)
    assert nodes_of(c, first) == nodes_of(c, second)


@pytest.mark.parametrize("which", BACKENDS)
def test_second_node_when_the_pod_does_not_fit(which):  # suite_test.go:1904-1922: the node has 2000m
    c = Cluster(which)
    first = c.pods(1, requests={"cpu": "1001m"})
    c.provision(first)
    second = c.pods(1, requests={"cpu": "1"})
    c.provision(second)
    assert nodes_of(c, first) != nodes_of(c, second)


@pytest.mark.parametrize("which", BACKENDS)
def test_second_node_when_the_selector_does_not_match(which):  # suite_test.go:1923-1939
    c = Cluster(which, pools=[_pool(requirements=[req(ARCH_LABEL, "In", "arm64", "amd64")])])
    first = c.pods(1, requests={"cpu": "10m"})
    c.provision(first)
    second = c.pods(1, node_selector={ARCH_LABEL: "arm64"})
    c.provision(second)
    assert nodes_of(c, first) != nodes_of(c, second)


@pytest.mark.parametrize("which", BACKENDS)
def test_zonal_balance_with_in_flight_nodes(which):  # suite_test.go:1967-1997
    c = Cluster(which)
    lab = {"foo": "bar"}
    tsc = spread(ZONE_LABEL, 1, LabelSelector.of(lab))
    c.provision(c.pods(4, labels=lab, topology_spread_constraints=tsc))
    assert c.skew(ZONE_LABEL, LabelSelector.of(lab)) == [1, 1, 2]
    n = len(c.nodes)
    c.provision(c.pods(5, labels=lab, topology_spread_constraints=tsc))
    assert c.skew(ZONE_LABEL, LabelSelector.of(lab)) == [3, 3, 3] and len(c.nodes) == n  # no new nodes


@pytest.mark.parametrize("which", BACKENDS)
def test_hostname_balance_with_in_flight_nodes(which):  # suite_test.go:1998-2023
    c = Cluster(which)
    lab = {"foo": "bar"}
    tsc = spread(HOSTNAME_LABEL, 1, LabelSelector.of(lab))
    c.provision(c.pods(4, labels=lab, topology_spread_constraints=tsc))
    c.provision(c.pods(5, labels=lab, topology_spread_constraints=tsc))
    assert c.skew(HOSTNAME_LABEL, LabelSelector.of(lab)) == [1] * 9  # new nodes although the old ones have room


@pytest.mark.parametrize("which", BACKENDS)
def test_pack_in_flight_nodes_before_launching_new_ones(which):  # suite_test.go:2375-2414
    medium = fake.new_instance_type("medium", {"cpu": "4.25", "pods": "4"})
    c = Cluster(which, its=[medium])
    rng = random.Random(5)
    for _ in range(10):
        batch = c.pods(rng.randrange(10), requests={"cpu": "1"})
        c.provision(batch)
        assert all(c.scheduled(p) for p in batch)
    free = sum(1 for n in c.nodes if int(str(n.available["cpu"]).rstrip("m")) >= 1000)
    assert free <= 1


@pytest.mark.parametrize("which", BACKENDS)
def test_tainted_in_flight_node_is_not_assumed(which):  # suite_test.go:2086-2117
    c = Cluster(which)
    first = c.pods(1, requests={"cpu": "10m"})
    c.provision(first)
    c.node_of(first[0]).taints = [Taint("foo.com/taint", "tainted", "NoSchedule")]
    second = c.pods(1, requests={"cpu": "10m"})
    c.provision(second)
    assert nodes_of(c, first) != nodes_of(c, second)


@pytest.mark.parametrize("which", BACKENDS)
def test_nodepool_taint_needs_a_toleration(which):  # topology_test.go:2991-3017
    c = Cluster(which, pools=[_pool(taints=[Taint("test-key", "test-value", "NoSchedule")])])
    ok = c.pods(1, tolerations=[Toleration("test-key", "Equal", "test-value", "NoSchedule")])
    ok += c.pods(1, tolerations=[Toleration("test-key", "Exists", "", "NoSchedule")])
    ok += c.pods(1, tolerations=[Toleration("test-key", "Exists", "", "")])
    ok += c.pods(1, tolerations=[Toleration("", "Exists", "", "")])
    bad = c.pods(1)
    bad += c.pods(1, tolerations=[Toleration("invalid", "Exists", "", "")])
    bad += c.pods(1, tolerations=[Toleration("test-key", "Equal", "other", "NoSchedule")])
    c.provision(ok + bad)
    assert all(c.scheduled(p) for p in ok) and not any(c.scheduled(p) for p in bad)


# ---- Existing Nodes (suite_test.go:2479-2659) ------------------------------------------------------------------------
@pytest.mark.parametrize("which", BACKENDS)
def test_existing_node_not_owned_by_karpenter(which):  # suite_test.go:2480-2534
    c = Cluster(which)
    c.add_node("unowned", {ZONE_LABEL: "test-zone-1", ARCH_LABEL: "amd64", OS_LABEL: "linux"},
               available={"cpu": "10", "memory": "10Gi", "pods": 110})
    pl = c.pods(100, requests={"cpu": "10m"})
    r = c.provision(pl)
    assert not r.new_node_claims and nodes_of(c, pl) == {"unowned"}


@pytest.mark.parametrize("which", BACKENDS)
def test_pod_incompatible_with_the_existing_node_gets_a_new_one(which):  # suite_test.go:2568-2600
    c = Cluster(which)
    c.add_node("unowned", {ZONE_LABEL: "test-zone-1", ARCH_LABEL: "amd64", OS_LABEL: "linux"},
               available={"cpu": "10", "memory": "10Gi", "pods": 110})
    pl = c.pods(1, requests={"cpu": "10m"}, node_selector={ZONE_LABEL: "test-zone-2"})
    r = c.provision(pl)
    assert len(r.new_node_claims) == 1 and c.node_of(pl[0]).labels[ZONE_LABEL] == "test-zone-2"


# ---- consolidation (pkg/controllers/disruption/consolidation_test.go) ------------------------------------------------
def _consolidate(which, nodes, sets, its=None, **kw):
    import numpy as np
    from karpenter_b200.disruption import Consolidation
    from tests import oracle_lib
    from tests.test_reference_scenarios import nodepool
    its = its or fake.default_instance_types()
    np_ = nodepool()
    orc = Consolidation([np_], {np_.name: its}, nodes, backend=oracle_lib.consolidate, **kw)
    cmds = orc.compute(sets)
    if which == "gpu":
        gpu = Consolidation([np_], {np_.name: its}, nodes, **kw)
        try:
            cmds = gpu.compute(sets)
        finally:
            gpu.close()
        for k in ("decision", "n_new_claims", "n_unscheduled", "replacement_its"):
            assert np.array_equal(gpu.raw[k], orc.raw[k]), k
    return cmds


def _its():
    return {i.name: i for i in fake.default_instance_types()}


@pytest.mark.parametrize("which", BACKENDS)
@pytest.mark.parametrize("policy,decision", [("Ignore", "delete"), ("Respect", "replace")])
def test_consolidation_delete_with_preferred_anti_affinity(which, policy, decision):  # consolidation_test.go:4760-4824
    from karpenter_b200.model import Pod, PodAffinityTerm, WeightedPodAffinityTerm
    from tests.test_reference_scenarios import _node
    foo = {"app": "foo"}
    anti = [WeightedPodAffinityTerm(1, PodAffinityTerm(LabelSelector.of(foo), HOSTNAME_LABEL))]
    p = [Pod(name=f"p{i}", uid=i + 1, labels=foo, requests={"cpu": "1"}, pod_anti_affinity_preferred=anti) for i in range(3)]
    d = _its()["default-instance-type"]
    nodes = [_node("node-1", d, pod_list=p[:2]), _node("node-2", d, pod_list=p[2:])]
    (cmd,) = _consolidate(which, nodes, [["node-2"]], preference_policy=policy)
    # ignoring the preference the pod simply moves next to the others; respecting it, a fresh (cheaper) node satisfies the
    # preference before any relaxation is tried, so the command becomes a replacement
    assert cmd.decision == decision and cmd.n_new_node_claims == (0 if decision == "delete" else 1)


@pytest.mark.parametrize("which", BACKENDS)
@pytest.mark.parametrize("policy,decision", [("Ignore", "replace"), ("Respect", "noop")])
def test_consolidation_replace_with_preferred_instance_type(which, policy, decision):  # consolidation_test.go:4825-4870
    from karpenter_b200.model import Pod, PreferredSchedulingTerm
    from tests.test_reference_scenarios import _node
    its = _its()
    exp = its["arm-instance-type"]  # the most expensive type of the fake catalog
    p = [Pod(name="p", uid=1, labels={"app": "foo"}, requests={"cpu": "1"},
             node_affinity_preferred=[PreferredSchedulingTerm(1, (req(INSTANCE_TYPE_LABEL, "In", exp.name),))])]
    nodes = [_node("node-1", exp, pod_list=p)]
    (cmd,) = _consolidate(which, nodes, [["node-1"]], preference_policy=policy)
    assert cmd.decision == decision
    if decision == "replace":
        assert exp.name not in cmd.replacement_instance_types and cmd.replacement_instance_types


@pytest.mark.parametrize("which", BACKENDS)
def test_consolidation_delete_onto_unmanaged_capacity(which):  # consolidation_test.go:2525-2572
    from karpenter_b200.model import StateNode
    from tests.test_reference_scenarios import _node, pods
    d = _its()["default-instance-type"]
    managed = _node("node-1", d, pod_list=pods(3, requests={"cpu": "1"}))
    unmanaged = StateNode(name="unmanaged", labels={HOSTNAME_LABEL: "unmanaged"}, available={"cpu": "32", "pods": 100},
                          capacity={"cpu": "32", "pods": 100}, managed=False)
    (cmd,) = _consolidate(which, [managed, unmanaged], [["node-1"]])
    assert cmd.decision == "delete" and cmd.n_new_node_claims == 0


@pytest.mark.parametrize("which", BACKENDS)
def test_consolidation_would_make_a_pod_pending(which):  # consolidation_test.go:3235-3274
    from tests.test_reference_scenarios import _node, pods
    small = fake.new_instance_type("only", {"cpu": "32", "pods": "100"})
    n1 = _node("node-1", small, pod_list=pods(2, requests={"cpu": "1"}, node_selector={"foo": "1"}))
    n2 = _node("node-2", small, pod_list=pods(1, uid0=10, requests={"cpu": "1"}, node_selector={"foo": "2"}))
    n1.labels["foo"], n2.labels["foo"] = "1", "2"
    cmds = _consolidate(which, [n1, n2], [["node-1"], ["node-2"], ["node-1", "node-2"]], its=[small])
    # the NodePool knows no label foo: a pod evicted from either node has nowhere to go
    assert [c.decision for c in cmds] == ["noop", "noop", "noop"]
    assert all(c.n_unscheduled > 0 for c in cmds)


@pytest.mark.parametrize("which", BACKENDS)
def test_multi_node_replacement_by_one_of_the_removed_types(which):  # multinodeconsolidation.go:165-188 (the two examples)
    from tests.test_reference_scenarios import _node, pods
    its = _its()
    d, small = its["default-instance-type"], its["small-instance-type"]
    p = pods(3, requests={"cpu": "500m"})
    # [default, default, small] -> the pods fit one small node, but a small node is being removed: that is a deletion of
    # the two default nodes in disguise, not a replacement
    nodes = [_node("n1", d, pod_list=p[:1]), _node("n2", d, pod_list=p[1:2]), _node("n3", small, pod_list=p[2:])]
    sets = [["n1", "n2", "n3"]]
    (plain,) = _consolidate(which, nodes, sets)
    assert plain.decision == "replace" and "small-instance-type" in plain.replacement_instance_types
    (multi,) = _consolidate(which, nodes, sets, filter_same_instance_type=True)
    assert multi.decision == "noop"
    # [default, default, default] -> options cheaper than a default node stay
    nodes = [_node(f"n{i + 1}", d, pod_list=p[i:i + 1]) for i in range(3)]
    (multi,) = _consolidate(which, nodes, sets, filter_same_instance_type=True)
    assert multi.decision == "replace" and multi.replacement_instance_types == ["small-instance-type"]
    # a single node is never filtered (firstNConsolidationOption only sees >= 2 candidates)
    (single,) = _consolidate(which, nodes, [["n1"]], filter_same_instance_type=True)
    assert single.decision == "delete"


# ---- NodePool limits (provisioning/suite_test.go:742-935) -------------------------------------------------------------
def _limited(which, **limits):
    return Cluster(which, pools=[NodePool(name="default", requirements=[req(CAPACITY_TYPE_LABEL, "In", "spot", "on-demand", "reserved")],
                                          limits=limits)])


FOO = {"app": "foo"}


def _anti_foo(c, n=1):
    from karpenter_b200.model import PodAffinityTerm
    return c.pods(n, labels=FOO, requests={"cpu": "1.5"},
                  pod_anti_affinity=[PodAffinityTerm(LabelSelector.of(FOO), HOSTNAME_LABEL)])


@pytest.mark.parametrize("which", BACKENDS)
def test_limits_already_exceeded_by_existing_capacity(which):  # provisioning/suite_test.go:743-765
    from karpenter_b200.model import StateNode
    c = _limited(which, cpu="20")
    c.nodes.append(StateNode(name="big", labels={HOSTNAME_LABEL: "big"}, available={"cpu": "0", "pods": 0},
                             capacity={"cpu": "100"}, nodepool="default", initialized=True))
    pl = c.pods(1)
    c.provision(pl)
    assert not c.scheduled(pl[0])


@pytest.mark.parametrize("which", BACKENDS)
@pytest.mark.parametrize("cpu,ok", [("1.75", True), ("2.1", False)])
def test_limits_would_be_met_or_exceeded(which, cpu, ok):  # provisioning/suite_test.go:766-782, 833-847
    c = _limited(which, cpu="2")
    pl = c.pods(1, requests={"cpu": cpu})
    c.provision(pl)
    assert c.scheduled(pl[0]) == ok


@pytest.mark.parametrize("which", BACKENDS)
def test_limits_partially_schedule(which):  # provisioning/suite_test.go:783-832: the first node eats the limit
    c = _limited(which, cpu="3")
    pl = _anti_foo(c, 2)
    c.provision(pl)
    assert sum(c.scheduled(p) for p in pl) == 1


@pytest.mark.parametrize("which", BACKENDS)
def test_limit_on_pods_blocks_every_instance_type(which):  # provisioning/suite_test.go:848-863
    c = _limited(which, pods="1")
    pl = c.pods(1, requests={fake.GPU_VENDOR_A: "1"})
    c.provision(pl)
    assert not c.scheduled(pl[0])


@pytest.mark.parametrize("which", BACKENDS)
def test_limits_across_scheduling_rounds(which):  # provisioning/suite_test.go:864-891
    c = _limited(which, cpu="2")
    first = c.pods(1, requests={"cpu": "1.75"})
    c.provision(first)
    second = c.pods(1, requests={"cpu": "1.75"})
    c.provision(second)
    assert c.scheduled(first[0]) and not c.scheduled(second[0])


@pytest.mark.parametrize("which", BACKENDS)
def test_limit_on_node_count(which):  # provisioning/suite_test.go:892-934
    c = _limited(which, nodes="2")
    placed = []
    for _ in range(3):
        p = _anti_foo(c)
        c.provision(p)
        placed.append(c.scheduled(p[0]))
    assert placed == [True, True, False]


# ---- Daemonsets (provisioning/suite_test.go:936-1057) ----------------------------------------------------------------
@pytest.mark.parametrize("which", BACKENDS)
def test_daemonset_overhead_moves_the_pod_to_a_bigger_node(which):  # provisioning/suite_test.go:936-956
    c = Cluster(which, daemon_overhead={"default": {"cpu": "2", "memory": "1Gi"}})
    pl = c.pods(1, requests={"cpu": "1", "memory": "1Gi"})
    c.provision(pl)
    assert c.node_of(pl[0]).labels[INSTANCE_TYPE_LABEL] == "default-instance-type"  # 4 cpu / 4Gi; the 2-cpu type is out
    c2 = Cluster(which)
    pl = c2.pods(1, requests={"cpu": "1", "memory": "1Gi"})
    c2.provision(pl)
    assert c2.node_of(pl[0]).labels[INSTANCE_TYPE_LABEL] == "small-instance-type"


@pytest.mark.parametrize("which", BACKENDS)
def test_daemonset_overhead_too_large(which):  # provisioning/suite_test.go:1005-1014
    c = Cluster(which, daemon_overhead={"default": {"cpu": "10000", "memory": "10000Gi"}})
    pl = c.pods(1)
    c.provision(pl)
    assert not c.scheduled(pl[0])


# ---- NodePool selection (provisioning/suite_test.go:2637-2711) -------------------------------------------------------
def _pools(*specs):
    return [NodePool(name=n, weight=w, requirements=[req(CAPACITY_TYPE_LABEL, "In", "spot", "on-demand", "reserved")],
                     labels=dict(lab), taints=list(t), limits={"cpu": "2000"}) for n, w, lab, t in specs]


@pytest.mark.parametrize("which", BACKENDS)
def test_explicitly_selected_nodepool(which):  # provisioning/suite_test.go:2638-2645, 2696-2710
    from karpenter_b200.model import NODEPOOL_LABEL
    c = Cluster(which, pools=_pools(("target", 0, {}, ()), ("w20", 20, {}, ()), ("w100", 100, {}, ())))
    pl = c.pods(1, node_selector={NODEPOOL_LABEL: "target"})
    r = c.provision(pl)
    assert r.new_node_claims[0].nodepool == "target"


@pytest.mark.parametrize("which", BACKENDS)
def test_nodepool_by_labels(which):  # provisioning/suite_test.go:2646-2661
    c = Cluster(which, pools=_pools(("plain", 0, {}, ()), ("labelled", 0, {"foo": "bar"}, ())))
    r = c.provision(c.pods(1, node_selector={"foo": "bar"}))
    assert r.new_node_claims[0].nodepool == "labelled"


@pytest.mark.parametrize("which", BACKENDS)
def test_prefer_no_schedule_pool_is_avoided_when_another_matches(which):  # provisioning/suite_test.go:2662-2677
    soft = Taint("foo", "bar", "PreferNoSchedule")
    for order in (("soft", "plain"), ("plain", "soft")):  # whatever the name order (OrderByWeight ties break on names)
        specs = [(n, 0, {}, (soft,) if n == "soft" else ()) for n in order]
        c = Cluster(which, pools=_pools(*specs))
        r = c.provision(c.pods(1))
        assert r.new_node_claims[0].nodepool == "plain"


@pytest.mark.parametrize("which", BACKENDS)
def test_highest_weight_nodepool_always(which):  # provisioning/suite_test.go:2680-2695
    c = Cluster(which, pools=_pools(("w0", 0, {}, ()), ("w20", 20, {}, ()), ("w100", 100, {}, ())))
    r = c.provision(c.pods(3))
    assert {cl.nodepool for cl in r.new_node_claims} == {"w100"}


@pytest.mark.parametrize("which", BACKENDS)
def test_prefer_no_schedule_tolerated_after_affinity_terms_relaxed(which):  # provisioning/suite_test.go:2352-2378
    from karpenter_b200.model import PreferredSchedulingTerm
    c = Cluster(which, pools=_pools(("default", 0, {}, (Taint("foo", "bar", "PreferNoSchedule"),))))
    pl = c.pods(1, node_affinity_preferred=[PreferredSchedulingTerm(1, (req(ZONE_LABEL, "In", "invalid"),)),
                                            PreferredSchedulingTerm(1, (req(INSTANCE_TYPE_LABEL, "In", "invalid"),))])
    c.provision(pl)
    assert c.scheduled(pl[0]) and c.node_of(pl[0]).taints == [Taint("foo", "bar", "PreferNoSchedule")]


@pytest.mark.parametrize("which", BACKENDS)
def test_unsatisfiable_node_preference_is_dropped(which):  # provisioning/suite_test.go:2379-2424 (zone entry)
    from karpenter_b200.model import PreferredSchedulingTerm
    c = Cluster(which)
    big = c.pods(1, labels={"app": "foo"}, requests={"cpu": "2"})
    pref = c.pods(1, labels={"app": "baz"}, requests={"cpu": "1"},
                  node_affinity_preferred=[PreferredSchedulingTerm(1, (req(ZONE_LABEL, "In", "value-1"),))])
    c.provision(big + pref)
    assert nodes_of(c, big) == nodes_of(c, pref)
