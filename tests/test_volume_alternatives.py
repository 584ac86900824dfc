"""Several volume-topology alternatives for one pod (PodData.VolumeRequirements; NodeClaim.CanAdd / ExistingNode.CanAdd try
them in turn on every candidate, nodeclaim.go:136-153, existingnode.go:98-113).  The alternatives are the encoder's input
(VolumeTopology.GetRequirements derives them, volumetopology.go:44-125); the reference's StorageClass case
(suite_test.go:2994-3038) and the order / first-success rules on both tiers."""
import pytest

from karpenter_b200 import fake
from karpenter_b200.model import (CAPACITY_TYPE_LABEL, HOSTNAME_LABEL, ZONE_LABEL, LabelSelector, NodePool,
                                  NodeSelectorRequirement, Pod, PodAffinityTerm, StateNode)
from karpenter_b200.scheduler import Scheduler
from tests import oracle_lib

BACKENDS = [pytest.param("oracle", id="oracle"), pytest.param("gpu", id="gpu", marks=pytest.mark.gpu)]


def req(key, op, *values):
    return NodeSelectorRequirement(key, op, tuple(values))


def zone(*z):
    return [req(ZONE_LABEL, "In", *z)]


def _solve(which, pods, state_nodes=(), pool=None):
    pool = pool or NodePool(name="default", requirements=[req(CAPACITY_TYPE_LABEL, "In", "on-demand")])
    def run(backend):
        s = Scheduler([pool], {pool.name: fake.default_instance_types()}, state_nodes=state_nodes, backend=backend)
        try:
            return s.solve(pods)
        finally:
            s.close()
    r = run(oracle_lib.solve)
    if which == "gpu":
        from tests.parity import assert_same
        g = run(None)
        assert_same(g.raw, r.raw, "volume alternatives ")
        r = g
    return r


def _zone_of(claim):
    v = claim.requirements[ZONE_LABEL]
    assert not v["complement"] and len(v["values"]) == 1
    return v["values"][0]


@pytest.mark.parametrize("which", BACKENDS)
def test_storage_class_with_two_zones_and_zonal_anti_affinity(which):  # suite_test.go:2994-3038
    labels = {"app": "multi-zone-sc-app"}
    anti = [PodAffinityTerm(LabelSelector.of(labels), ZONE_LABEL)]
    pods = [Pod(name=f"sc-pod-{i}", uid=i + 1, labels=labels, requests={"cpu": "100m"}, pod_anti_affinity=anti,
                volume_requirements=[zone("test-zone-1"), zone("test-zone-2")]) for i in range(2)]
    r = _solve(which, pods)
    assert not r.pod_errors and len(r.new_node_claims) == 2
    assert sorted(_zone_of(c) for c in r.new_node_claims) == ["test-zone-1", "test-zone-2"]


@pytest.mark.parametrize("which", BACKENDS)
def test_first_alternative_that_works_wins_per_candidate(which):
    # zone-3 first, zone-1 second: a fresh NodeClaim takes the first alternative ...
    p = Pod(name="p", uid=1, requests={"cpu": "100m"}, volume_requirements=[zone("test-zone-3"), zone("test-zone-1")])
    r = _solve(which, [p])
    assert _zone_of(r.new_node_claims[0]) == "test-zone-3"
    # ... but an in-flight NodeClaim pinned to zone-1 by an earlier pod is tried first, and its second alternative fits
    first = Pod(name="a", uid=1, requests={"cpu": "200m"}, node_selector={ZONE_LABEL: "test-zone-1"})
    p = Pod(name="p", uid=2, requests={"cpu": "100m"}, volume_requirements=[zone("test-zone-3"), zone("test-zone-1")])
    r = _solve(which, [first, p])
    assert len(r.new_node_claims) == 1 and _zone_of(r.new_node_claims[0]) == "test-zone-1"
    # no alternative fits the claim: a second NodeClaim, first alternative again
    p = Pod(name="p", uid=2, requests={"cpu": "100m"}, volume_requirements=[zone("test-zone-3"), zone("test-zone-2")])
    r = _solve(which, [first, p])
    assert sorted(_zone_of(c) for c in r.new_node_claims) == ["test-zone-1", "test-zone-3"]
    # every alternative impossible: the pod fails
    p = Pod(name="p", uid=2, requests={"cpu": "100m"}, volume_requirements=[zone("no-such-zone"), zone("nor-this-one")])
    r = _solve(which, [p])
    assert len(r.pod_errors) == 1


@pytest.mark.parametrize("which", BACKENDS)
def test_existing_node_accepts_through_a_later_alternative(which):  # existingnode.go:98-113
    it = fake.default_instance_types()[0]
    n = StateNode(name="n1", labels={HOSTNAME_LABEL: "n1", ZONE_LABEL: "test-zone-2", CAPACITY_TYPE_LABEL: "on-demand"},
                  available={"cpu": "4", "memory": "4Gi", "pods": 10}, capacity=dict(it.capacity), managed=False)
    p = Pod(name="p", uid=1, requests={"cpu": "100m"}, volume_requirements=[zone("test-zone-1"), zone("test-zone-2")])
    r = _solve(which, [p], state_nodes=[n])
    assert list(r.existing_nodes) == ["n1"] and not r.new_node_claims
    p = Pod(name="p", uid=1, requests={"cpu": "100m"}, volume_requirements=[zone("test-zone-1"), zone("test-zone-3")])
    r = _solve(which, [p], state_nodes=[n])
    assert not r.existing_nodes and _zone_of(r.new_node_claims[0]) == "test-zone-1"
