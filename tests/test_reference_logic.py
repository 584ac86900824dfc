"""Requirement "Scheduling Logic" cases (pkg/controllers/provisioning/scheduling/suite_test.go:951-1124) and the rest of
the spread cases of topology_test.go that the other scenario files do not restate.  Same harness (`Cluster`); CPU tier
through the oracle, GPU tier bit-identical to it."""
import pytest

from karpenter_b200.model import HOSTNAME_LABEL, ZONE_LABEL, LabelSelector, TopologySpreadConstraint
from tests.test_reference_scenarios import BACKENDS, req
from tests.test_reference_topology import LABELS, SEL, Cluster, _pool, spread

KEY = "test-key"


def one(which, pod_reqs, pool_reqs=()):
    c = Cluster(which, pools=[_pool(requirements=list(pool_reqs))])
    pl = c.pods(1, node_affinity_required=[list(pod_reqs)] if pod_reqs else [])
    c.provision(pl)
    return c, pl[0]


DEFINED = [req(KEY, "In", "test-value")]


# ---- undefined key (suite_test.go:952-989) ---------------------------------------------------------------------------
@pytest.mark.parametrize("which", BACKENDS)
@pytest.mark.parametrize("op,values,ok", [("In", ("test-value",), False), ("NotIn", ("test-value",), True),
                                          ("Exists", (), False), ("DoesNotExist", (), True)])
def test_operator_on_an_undefined_key(which, op, values, ok):
    c, p = one(which, [req(KEY, op, *values)])
    assert c.scheduled(p) == ok
    if ok:
        assert KEY not in c.node_of(p).labels


# ---- key defined by the NodePool (suite_test.go:990-1068) ------------------------------------------------------------
@pytest.mark.parametrize("which", BACKENDS)
@pytest.mark.parametrize("pod_req,ok", [(None, True), (("In", "test-value"), True), (("NotIn", "test-value"), False),
                                        (("Exists",), True), (("DoesNotExist",), False), (("In", "another-value"), False),
                                        (("NotIn", "another-value"), True)])
def test_operator_on_a_key_the_nodepool_defines(which, pod_req, ok):
    reqs = [req(KEY, pod_req[0], *pod_req[1:])] if pod_req else []
    c, p = one(which, reqs, DEFINED)
    assert c.scheduled(p) == ok
    if ok:
        assert c.node_of(p).labels[KEY] == "test-value"


@pytest.mark.parametrize("which", BACKENDS)
def test_compatible_pods_share_a_node(which):  # suite_test.go:1069-1088
    c = Cluster(which, pools=[_pool(requirements=[req(KEY, "In", "test-value", "another-value")])])
    pl = c.pods(1, node_affinity_required=[[req(KEY, "In", "test-value")]])
    pl += c.pods(1, node_affinity_required=[[req(KEY, "NotIn", "another-value")]])
    c.provision(pl)
    assert c.bound[id(pl[0])] == c.bound[id(pl[1])] and c.node_of(pl[0]).labels[KEY] == "test-value"


@pytest.mark.parametrize("which", BACKENDS)
def test_incompatible_pods_get_different_nodes(which):  # suite_test.go:1089-1109
    c = Cluster(which, pools=[_pool(requirements=[req(KEY, "In", "test-value", "another-value")])])
    pl = c.pods(1, node_affinity_required=[[req(KEY, "In", "test-value")]])
    pl += c.pods(1, node_affinity_required=[[req(KEY, "In", "another-value")]])
    c.provision(pl)
    assert [c.node_of(p).labels[KEY] for p in pl] == ["test-value", "another-value"]


@pytest.mark.parametrize("which", BACKENDS)
def test_exists_does_not_overwrite_a_value(which):  # suite_test.go:1110-1122
    c, p = one(which, [req(ZONE_LABEL, "In", "non-existent-zone"), req(ZONE_LABEL, "Exists")])
    assert not c.scheduled(p)


# ---- spread odds and ends (topology_test.go:445-481, 1139-1194, 1740-1830) --------------------------------------------
@pytest.mark.parametrize("which", BACKENDS)
def test_nil_label_selector_spread_still_schedules(which):  # topology_test.go:92-105, 445-456
    c = Cluster(which)
    pl = c.pods(1, topology_spread_constraints=[TopologySpreadConstraint(1, ZONE_LABEL, None)])
    c.provision(pl)
    assert c.scheduled(pl[0])


@pytest.mark.parametrize("which", BACKENDS)
def test_interdependent_selectors(which):  # topology_test.go:457-480: the pods do not match their own selector
    c = Cluster(which)
    pl = c.pods(5, topology_spread_constraints=spread(HOSTNAME_LABEL))  # no labels: they never count toward the skew
    c.provision(pl)
    assert len({c.bound[id(p)] for p in pl}) == 1


@pytest.mark.parametrize("which", BACKENDS)
def test_match_label_keys(which):  # topology_test.go:1139-1167
    c = Cluster(which)
    tsc = [TopologySpreadConstraint(1, HOSTNAME_LABEL, SEL, match_label_keys=("test-label",))]
    pl = c.pods(2, labels={**LABELS, "test-label": "value-a"}, topology_spread_constraints=tsc)
    pl += c.pods(2, labels={**LABELS, "test-label": "value-b"}, topology_spread_constraints=tsc)
    c.provision(pl)
    assert c.skew(HOSTNAME_LABEL) == [2, 2]  # one pod of each "deployment" per node; 4 nodes without matchLabelKeys


@pytest.mark.parametrize("which", BACKENDS)
def test_unknown_match_label_keys_are_ignored(which):  # topology_test.go:1168-1193
    c = Cluster(which)
    tsc = [TopologySpreadConstraint(1, HOSTNAME_LABEL, SEL, match_label_keys=("test-label",))]
    c.provision(c.pods(4, labels=LABELS, topology_spread_constraints=tsc))
    assert c.skew(HOSTNAME_LABEL) == [1, 1, 1, 1]


@pytest.mark.parametrize("which", BACKENDS)
def test_spread_limited_by_node_selector(which):  # topology_test.go:1740-1765
    c = Cluster(which)
    pl = c.pods(5, labels=LABELS, topology_spread_constraints=spread(ZONE_LABEL), node_selector={ZONE_LABEL: "test-zone-1"})
    pl += c.pods(10, labels=LABELS, topology_spread_constraints=spread(ZONE_LABEL), node_selector={ZONE_LABEL: "test-zone-2"})
    c.provision(pl)
    assert c.skew(ZONE_LABEL) == [5, 10]


@pytest.mark.parametrize("which", BACKENDS)
def test_spread_limited_by_node_requirements(which):  # topology_test.go:1766-1787
    c = Cluster(which)
    c.provision(c.pods(10, labels=LABELS, topology_spread_constraints=spread(ZONE_LABEL),
                       node_affinity_required=[[req(ZONE_LABEL, "In", "test-zone-1", "test-zone-2")]]))
    assert c.skew(ZONE_LABEL) == [5, 5]


@pytest.mark.parametrize("which", BACKENDS)
def test_spread_limited_by_required_node_affinity_then_opened(which):  # topology_test.go:1788-1830
    c = Cluster(which)
    two = [[req(ZONE_LABEL, "In", "test-zone-1", "test-zone-2")]]
    c.provision(c.pods(6, labels=LABELS, topology_spread_constraints=spread(ZONE_LABEL), node_affinity_required=two))
    assert c.skew(ZONE_LABEL) == [3, 3]
    c.provision(c.pods(1, labels=LABELS, topology_spread_constraints=spread(ZONE_LABEL),
                       node_affinity_required=[[req(ZONE_LABEL, "In", "test-zone-2", "test-zone-3")]]))
    assert c.skew(ZONE_LABEL) == [1, 3, 3]  # the empty zone 3, although the skew is violated: it improves it
    c.provision(c.pods(5, labels=LABELS, topology_spread_constraints=spread(ZONE_LABEL)))
    assert c.skew(ZONE_LABEL) == [4, 4, 4]
