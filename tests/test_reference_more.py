"""Further reference cases on both tiers (the file started as oracle-only pins at the end of round 1; every case now has its
CUDA-path twin): capacity-type / architecture spreads seen through node affinity, in-flight nodes with taints
(topology_test.go:815-940, suite_test.go:2025-2209), the spot-to-spot consolidation rules (consolidation_test.go:1033-1218)."""
import pytest

from karpenter_b200.model import ARCH_LABEL, CAPACITY_TYPE_LABEL, INSTANCE_TYPE_LABEL, ZONE_LABEL, Taint, Toleration
from tests.test_reference_scenarios import req
from tests.test_reference_topology import LABELS, Cluster, _pool, spread

BACKENDS = [pytest.param("oracle", id="oracle"), pytest.param("gpu", id="gpu", marks=pytest.mark.gpu)]


@pytest.mark.parametrize("W", BACKENDS)
def test_capacity_type_spread_excludes_pods_outside_the_node_affinity(W):  # topology_test.go:815-850
    c = Cluster(W)
    c.provision(c.pods(1, labels=LABELS, node_affinity_required=[[req(ZONE_LABEL, "In", "test-zone-1"),
                                                                  req(CAPACITY_TYPE_LABEL, "In", "on-demand")]]))
    c.provision(c.pods(5, labels=LABELS, topology_spread_constraints=spread(CAPACITY_TYPE_LABEL),
                       node_affinity_required=[[req(ZONE_LABEL, "In", "test-zone-2"), req(CAPACITY_TYPE_LABEL, "In", "spot")]]))
    # the on-demand pod in zone 1 is outside the new pods' node affinity, so it does not count: all five go to spot
    assert c.skew(CAPACITY_TYPE_LABEL) == [1, 5]


@pytest.mark.parametrize("W", BACKENDS)
def test_capacity_type_spread_sees_the_existing_on_demand_node(W):  # topology_test.go:852-894
    c = Cluster(W, pools=[_pool(requirements=[req(ARCH_LABEL, "In", "amd64", "arm64")])])
    c.provision(c.pods(1, labels=LABELS, node_selector={INSTANCE_TYPE_LABEL: "single-pod-instance-type"},
                       node_affinity_required=[[req(CAPACITY_TYPE_LABEL, "In", "on-demand")]]))
    c.pools[0].requirements = [req(CAPACITY_TYPE_LABEL, "In", "spot")]
    c.provision(c.pods(5, labels=LABELS, requests={"cpu": "2"}, topology_spread_constraints=spread(CAPACITY_TYPE_LABEL)))
    assert c.skew(CAPACITY_TYPE_LABEL) == [1, 2]


@pytest.mark.parametrize("W", BACKENDS)
def test_arch_spread_sees_the_existing_amd64_node(W):  # topology_test.go:895-938
    c = Cluster(W, pools=[_pool(requirements=[req(ARCH_LABEL, "In", "amd64", "arm64")])])
    c.provision(c.pods(1, labels=LABELS, node_selector={INSTANCE_TYPE_LABEL: "single-pod-instance-type"},
                       node_affinity_required=[[req(ARCH_LABEL, "In", "amd64")]]))
    c.pools[0].requirements = [req(CAPACITY_TYPE_LABEL, "In", "spot", "on-demand", "reserved"), req(ARCH_LABEL, "In", "arm64")]
    c.provision(c.pods(5, labels=LABELS, requests={"cpu": "2"}, topology_spread_constraints=spread(ARCH_LABEL)))
    assert c.skew(ARCH_LABEL) == [1, 2]


@pytest.mark.parametrize("W", BACKENDS)
def test_untainted_in_flight_node_is_assumed(W):  # suite_test.go:2026-2047
    c = Cluster(W)
    first = c.pods(1, requests={"cpu": "10m"})
    c.provision(first)
    second = c.pods(1, requests={"cpu": "10m"})
    c.provision(second)
    assert c.bound[id(first[0])] == c.bound[id(second[0])]


@pytest.mark.parametrize("W", BACKENDS)
def test_tainted_in_flight_node_with_a_toleration(W):  # suite_test.go:2086-2117, the tolerating twin
    c = Cluster(W)
    first = c.pods(1, requests={"cpu": "10m"})
    c.provision(first)
    c.node_of(first[0]).taints = [Taint("foo.com/taint", "tainted", "NoSchedule")]
    tol = c.pods(1, requests={"cpu": "10m"}, tolerations=[Toleration("foo.com/taint", "Exists", "", "")])
    c.provision(tol)
    assert c.bound[id(first[0])] == c.bound[id(tol[0])]


# ---- spot-to-spot consolidation rules (consolidation_test.go:1033-1218, consolidation.go:236-316) ---------------------
def _assorted(n):
    """the first n of fake.InstanceTypesAssorted() (fake/instancetype.go:156-192): 1 cpu / 1 Gi in every combination of
    zone x capacity type x os x arch, one offering each, all at the same price"""
    from karpenter_b200 import fake
    from karpenter_b200.model import Offering
    out = []
    for zone in ("test-zone-1", "test-zone-2", "test-zone-3"):
        for ct in ("spot", "on-demand"):
            for os_ in ("linux", "windows"):
                for arch in ("amd64", "arm64"):
                    res = {"cpu": "1", "memory": "1Gi"}
                    off = [Offering([req(CAPACITY_TYPE_LABEL, "In", ct), req(ZONE_LABEL, "In", zone)],
                                    fake.price_from_resources(res), True)]
                    out.append(fake.new_instance_type(f"1-cpu-1-mem-{arch}-{os_}-{zone}-{ct}", res, architecture=arch,
                                                      operating_systems=(os_,), offerings=off))
    return out[:n]


def _spot_case(W, n_types, spot_to_spot=True):
    from karpenter_b200.disruption import Consolidation
    from karpenter_b200.model import NodePool, Offering
    from tests import oracle_lib
    from tests.test_reference_scenarios import _node, pods
    its = _assorted(n_types)
    # the first type becomes a dirt-cheap spot offering: one option that is cheaper than the node
    its[0].offerings = [Offering([req(CAPACITY_TYPE_LABEL, "In", "spot"), req(ZONE_LABEL, "In", "test-zone-1")], 0.001, True)]
    spot = [i for i in its[1:] if i.offerings[0].requirements[0].values == ("spot",)]
    node_it = spot[-1]
    zone = node_it.offerings[0].requirements[1].values[0]
    np_ = NodePool(name="default", requirements=[req(CAPACITY_TYPE_LABEL, "In", "spot", "on-demand"),
                                                 req(ARCH_LABEL, "In", "amd64", "arm64")], limits={"cpu": "2000"})
    n = _node("spot-node", node_it, zone=zone, ct="spot", pod_list=pods(1, requests={"cpu": "100m"}))
    c = Consolidation([np_], {np_.name: its}, [n], spot_to_spot=spot_to_spot,
                      backend=oracle_lib.consolidate if W == "oracle" else None)
    try:
        return c.compute([["spot-node"]])[0]
    finally:
        c.close()


@pytest.mark.parametrize("W", BACKENDS)
@pytest.mark.parametrize("n_types", [5, 20])
def test_spot_to_spot_needs_fifteen_cheaper_types(W, n_types):  # consolidation_test.go:1033-1107, 1149-1218
    cmd = _spot_case(W, n_types)
    assert cmd.decision == "noop" and cmd.n_new_node_claims == 1  # a cheaper spot type exists, but only one of them


@pytest.mark.parametrize("W", BACKENDS)
def test_spot_to_spot_needs_the_feature_gate(W):  # consolidation_test.go:1108-1148
    assert _spot_case(W, 20, spot_to_spot=False).decision == "noop"
