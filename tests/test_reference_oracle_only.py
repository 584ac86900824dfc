"""Further reference cases restated against the ORACLE only (CPU tier).  They were written after this round's GPU budget
was spent, so their CUDA-path twins (the `gpu` parametrisation the other scenario files carry) wait for the next round;
what they pin is the oracle itself: capacity-type / architecture spreads seen through node affinity, in-flight nodes
with taints (topology_test.go:815-940, suite_test.go:2025-2209)."""
import pytest

from karpenter_b200.model import ARCH_LABEL, CAPACITY_TYPE_LABEL, INSTANCE_TYPE_LABEL, ZONE_LABEL, Taint, Toleration
from tests.test_reference_scenarios import req
from tests.test_reference_topology import LABELS, Cluster, _pool, spread

W = "oracle"


def test_capacity_type_spread_excludes_pods_outside_the_node_affinity():  # topology_test.go:815-850
    c = Cluster(W)
    c.provision(c.pods(1, labels=LABELS, node_affinity_required=[[req(ZONE_LABEL, "In", "test-zone-1"),
                                                                  req(CAPACITY_TYPE_LABEL, "In", "on-demand")]]))
    c.provision(c.pods(5, labels=LABELS, topology_spread_constraints=spread(CAPACITY_TYPE_LABEL),
                       node_affinity_required=[[req(ZONE_LABEL, "In", "test-zone-2"), req(CAPACITY_TYPE_LABEL, "In", "spot")]]))
    # the on-demand pod in zone 1 is outside the new pods' node affinity, so it does not count: all five go to spot
    assert c.skew(CAPACITY_TYPE_LABEL) == [1, 5]


def test_capacity_type_spread_sees_the_existing_on_demand_node():  # topology_test.go:852-894
    c = Cluster(W, pools=[_pool(requirements=[req(ARCH_LABEL, "In", "amd64", "arm64")])])
    c.provision(c.pods(1, labels=LABELS, node_selector={INSTANCE_TYPE_LABEL: "single-pod-instance-type"},
                       node_affinity_required=[[req(CAPACITY_TYPE_LABEL, "In", "on-demand")]]))
    c.pools[0].requirements = [req(CAPACITY_TYPE_LABEL, "In", "spot")]
    c.provision(c.pods(5, labels=LABELS, requests={"cpu": "2"}, topology_spread_constraints=spread(CAPACITY_TYPE_LABEL)))
    assert c.skew(CAPACITY_TYPE_LABEL) == [1, 2]


def test_arch_spread_sees_the_existing_amd64_node():  # topology_test.go:895-938
    c = Cluster(W, pools=[_pool(requirements=[req(ARCH_LABEL, "In", "amd64", "arm64")])])
    c.provision(c.pods(1, labels=LABELS, node_selector={INSTANCE_TYPE_LABEL: "single-pod-instance-type"},
                       node_affinity_required=[[req(ARCH_LABEL, "In", "amd64")]]))
    c.pools[0].requirements = [req(CAPACITY_TYPE_LABEL, "In", "spot", "on-demand", "reserved"), req(ARCH_LABEL, "In", "arm64")]
    c.provision(c.pods(5, labels=LABELS, requests={"cpu": "2"}, topology_spread_constraints=spread(ARCH_LABEL)))
    assert c.skew(ARCH_LABEL) == [1, 2]


def test_untainted_in_flight_node_is_assumed():  # suite_test.go:2026-2047
    c = Cluster(W)
    first = c.pods(1, requests={"cpu": "10m"})
    c.provision(first)
    second = c.pods(1, requests={"cpu": "10m"})
    c.provision(second)
    assert c.bound[id(first[0])] == c.bound[id(second[0])]


def test_tainted_in_flight_node_with_a_toleration():  # suite_test.go:2086-2117, the tolerating twin
    c = Cluster(W)
    first = c.pods(1, requests={"cpu": "10m"})
    c.provision(first)
    c.node_of(first[0]).taints = [Taint("foo.com/taint", "tainted", "NoSchedule")]
    tol = c.pods(1, requests={"cpu": "10m"}, tolerations=[Toleration("foo.com/taint", "Exists", "", "")])
    c.provision(tol)
    assert c.bound[id(first[0])] == c.bound[id(tol[0])]
