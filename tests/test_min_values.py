"""minValues on NodePool requirements (InstanceTypes.SatisfiesMinValues, cloudprovider/types.go:301-337;
filterInstanceTypesByRequirements, nodeclaim.go:412-480).  Restates the minValues cases of
pkg/controllers/provisioning/scheduling/instance_selection_test.go:624-1540 with the outcomes the reference asserts.
CPU tier: the oracle.  GPU tier: the CUDA path, bit-identical to the oracle.
"""
import pytest

from karpenter_b200 import fake
from karpenter_b200.model import (ARCH_LABEL, CAPACITY_TYPE_LABEL, INSTANCE_TYPE_LABEL, ZONE_LABEL, NodePool,
                                  NodeSelectorRequirement, Offering, Pod)
from karpenter_b200.scheduler import Scheduler
from tests import oracle_lib
from tests.test_reference_scenarios import BACKENDS, pods, req

GEN = "karpenter/numerical-value"


def mreq(key, op, *values, min_values=None):
    return NodeSelectorRequirement(key, op, tuple(values), min_values)


def it(name, cpu, arch="arm64", price=1.0, extra=()):
    off = [Offering([req(CAPACITY_TYPE_LABEL, "In", "spot"), req(ZONE_LABEL, "In", "test-zone-1-spot")], price, True)]
    t = fake.new_instance_type(name, {"cpu": str(cpu), "memory": f"{cpu}Gi"}, architecture=arch,
                               operating_systems=("linux",), offerings=off)
    t.requirements = list(t.requirements) + list(extra)
    return t


def two_types(extra1=(), extra2=(), arch2="arm64", cpu2=4):
    return [it("instance-type-1", 1, price=0.52, extra=extra1), it("instance-type-2", cpu2, arch=arch2, price=1.0, extra=extra2)]


def solve(which, pod_list, reqs, its, policy="Strict"):
    np_ = NodePool(name="default", requirements=[req(CAPACITY_TYPE_LABEL, "In", "spot", "on-demand", "reserved")] + reqs,
                   limits={"cpu": "2000"})

    def go(backend):
        s = Scheduler([np_], {np_.name: its}, backend=backend, min_values_policy=policy)
        try:
            return s.solve(pod_list)
        finally:
            s.close()
    r = go(oracle_lib.solve)
    if which == "gpu":
        from tests.parity import assert_same
        g = go(None)
        assert_same(g.raw, r.raw, "minValues ")
        return g
    return r


SMALL = dict(requests={"cpu": "0.9", "memory": "900Mi"})


@pytest.mark.parametrize("which", BACKENDS)
def test_min_values_on_instance_type(which):  # instance_selection_test.go:624-699
    r = solve(which, pods(2, **SMALL), [mreq(INSTANCE_TYPE_LABEL, "In", "instance-type-1", "instance-type-2", min_values=2)],
              two_types())
    # without minValues both pods share an instance-type-2 node; with it the second pod would leave one type only
    assert not r.pod_errors and len(r.new_node_claims) == 2
    assert all(len(c.instance_type_options) >= 2 for c in r.new_node_claims)
    assert all(c.requirements[INSTANCE_TYPE_LABEL]["min_values"] == 2 for c in r.new_node_claims)


@pytest.mark.parametrize("which", BACKENDS)
def test_without_min_values_pods_share_a_node(which):
    r = solve(which, pods(2, **SMALL), [mreq(INSTANCE_TYPE_LABEL, "In", "instance-type-1", "instance-type-2")], two_types())
    assert not r.pod_errors and len(r.new_node_claims) == 1 and r.new_node_claims[0].instance_type_options == ["instance-type-2"]


@pytest.mark.parametrize("which", BACKENDS)
def test_min_values_with_gt(which):  # instance_selection_test.go:700-794: generation Gt 1 keeps both types, minValues 2 holds
    its = two_types(extra1=[req(GEN, "In", "2")], extra2=[req(GEN, "In", "3")])
    pl = pods(2, node_affinity_required=[[req(GEN, "Gt", "1")]], **SMALL)
    r = solve(which, pl, [mreq(GEN, "Exists", min_values=2)], its)
    assert not r.pod_errors and len(r.new_node_claims) == 2


@pytest.mark.parametrize("which", BACKENDS)
def test_min_values_with_gt_not_satisfied(which):  # instance_selection_test.go:795-882: Gt 2 leaves one generation
    its = two_types(extra1=[req(GEN, "In", "2")], extra2=[req(GEN, "In", "3")])
    pl = pods(2, node_affinity_required=[[req(GEN, "Gt", "2")]], **SMALL)
    r = solve(which, pl, [mreq(GEN, "Exists", min_values=2)], its)
    assert len(r.pod_errors) == 2 and not r.new_node_claims


@pytest.mark.parametrize("which", BACKENDS)
def test_min_values_with_lt_not_satisfied(which):  # instance_selection_test.go:977-1046
    its = two_types(extra1=[req(GEN, "In", "2")], extra2=[req(GEN, "In", "3")])
    pl = pods(2, node_affinity_required=[[req(GEN, "Lt", "3")]], **SMALL)
    r = solve(which, pl, [mreq(GEN, "Exists", min_values=2)], its)
    assert len(r.pod_errors) == 2


@pytest.mark.parametrize("which", BACKENDS)
def test_max_of_min_values_of_in_and_not_in(which):  # instance_selection_test.go:1047-1144
    its = two_types(cpu2=2) + [it("instance-type-3", 4, price=2.0)]
    reqs = [mreq(INSTANCE_TYPE_LABEL, "In", "instance-type-1", "instance-type-2", "instance-type-3", min_values=1),
            mreq(INSTANCE_TYPE_LABEL, "NotIn", "instance-type-3", min_values=2)]
    r = solve(which, pods(2, **SMALL), reqs, its)
    assert not r.pod_errors and len(r.new_node_claims) == 2
    assert all(len(c.instance_type_options) >= 2 for c in r.new_node_claims)
    assert all(c.requirements[INSTANCE_TYPE_LABEL]["min_values"] == 2 for c in r.new_node_claims)  # the larger one


@pytest.mark.parametrize("which", BACKENDS)
def test_more_min_values_than_instance_types(which):  # instance_selection_test.go:1262-1288: 10 types, minValues 11
    r = solve(which, pods(1), [mreq(INSTANCE_TYPE_LABEL, "Exists", min_values=11)], fake.instance_types(10))
    assert len(r.pod_errors) == 1 and not r.new_node_claims
    r = solve(which, pods(1), [mreq(INSTANCE_TYPE_LABEL, "Exists", min_values=10)], fake.instance_types(10))
    assert not r.pod_errors and len(r.new_node_claims[0].instance_type_options) == 10


@pytest.mark.parametrize("which", BACKENDS)
def test_several_keys_with_min_values(which):  # instance_selection_test.go:1446-1540: arch Exists minValues 2
    reqs = [mreq(ARCH_LABEL, "Exists", min_values=2),
            mreq(INSTANCE_TYPE_LABEL, "In", "instance-type-1", "instance-type-2", min_values=1)]
    r = solve(which, pods(2, **SMALL), reqs, two_types(arch2="amd64"))
    assert not r.pod_errors and len(r.new_node_claims) == 2
    assert all(len(c.instance_type_options) >= 2 for c in r.new_node_claims)


@pytest.mark.parametrize("which", BACKENDS)
def test_best_effort_relaxes_min_values(which):  # provisioning/suite_test.go:2834-2902 family; nodeclaim.go:186-191
    reqs = [mreq(INSTANCE_TYPE_LABEL, "In", "instance-type-1", "instance-type-2", min_values=2)]
    r = solve(which, pods(2, **SMALL), reqs, two_types(), policy="BestEffort")
    # BestEffort never refuses: both pods share the instance-type-2 node and the claim reports the relaxed value
    assert not r.pod_errors and len(r.new_node_claims) == 1
    mv = r.new_node_claims[0].requirements[INSTANCE_TYPE_LABEL]
    assert mv["min_values"] == 1 and mv["relaxed"]
    r = solve(which, pods(1, **SMALL), [mreq(INSTANCE_TYPE_LABEL, "Exists", min_values=11)], fake.instance_types(10),
              policy="BestEffort")
    assert not r.pod_errors and r.new_node_claims[0].requirements[INSTANCE_TYPE_LABEL]["min_values"] == 10


@pytest.mark.parametrize("which", BACKENDS)
def test_min_values_many_pods_wide_catalog(which):
    """400 fake types, family-like key with 8 values, minValues 3 on it and 20 on the instance type: every claim keeps
    at least 20 types of at least 3 families while 300 pods bin-pack."""
    its = fake.instance_types(120)
    for i, t in enumerate(its):
        t.requirements = list(t.requirements) + [req("example.com/family", "In", f"f{i % 8}")]
    reqs = [mreq("example.com/family", "Exists", min_values=3), mreq(INSTANCE_TYPE_LABEL, "Exists", min_values=20)]
    pl = pods(150, requests={"cpu": "3", "memory": "2Gi"}) + pods(150, uid0=1000, requests={"cpu": "7", "memory": "1Gi"})
    r = solve(which, pl, reqs, its)
    assert not r.pod_errors
    fam = {t.name: i % 8 for i, t in enumerate(its)}
    for c in r.new_node_claims:
        assert len(c.instance_type_options) >= 20 and len({fam[n] for n in c.instance_type_options}) >= 3
    loose = solve(which, pl, [], its)
    assert len(loose.new_node_claims) < len(r.new_node_claims)  # minValues costs nodes: claims stop growing earlier


def test_consolidation_with_best_effort_min_values_is_refused_by_the_oracle_too():
    """Strict minValues are served by kp_consolidate (tests/test_consolidation_min_values.py); BestEffort is not."""
    from karpenter_b200.disruption import Consolidation
    from tests.test_reference_scenarios import _node
    its = two_types()
    np_ = NodePool(name="default", requirements=[mreq(INSTANCE_TYPE_LABEL, "Exists", min_values=2)])
    n = _node("n1", its[1], zone="test-zone-1-spot", ct="spot", pod_list=pods(1, **SMALL))
    c = Consolidation([np_], {np_.name: its}, [n], backend=oracle_lib.consolidate, min_values_policy="BestEffort")
    with pytest.raises(RuntimeError):
        c.compute([["n1"]])
    assert Consolidation([np_], {np_.name: its}, [n], backend=oracle_lib.consolidate).compute([["n1"]])[0].decision == "noop"
