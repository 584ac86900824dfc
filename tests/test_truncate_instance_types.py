"""Results.TruncateInstanceTypes inside the solve (scheduler.go:361-379; provisioner.go:380 calls it with MaxInstanceTypes = 600
right after Solve): kp_problem.max_instance_types.  The reference's case (instance_selection_test.go:1289-1345, MaxInstanceTypes
lowered to 1) and the price order on both tiers; the CUDA path bit-identical to the oracle on a 1 000-type catalog."""
import numpy as np
import pytest

from karpenter_b200 import fake
from karpenter_b200.model import (ARCH_LABEL, CAPACITY_TYPE_LABEL, INSTANCE_TYPE_LABEL, ZONE_LABEL, NodePool,
                                  NodeSelectorRequirement, Offering, Pod)
from karpenter_b200.scheduler import Scheduler
from tests import oracle_lib

BACKENDS = [pytest.param("oracle", id="oracle"), pytest.param("gpu", id="gpu", marks=pytest.mark.gpu)]


def req(key, op, *values, min_values=None):
    return NodeSelectorRequirement(key, op, tuple(values), min_values=min_values)


def _two_types():
    mk = lambda name, cpu, price: fake.new_instance_type(
        name, {"cpu": cpu, "memory": f"{cpu}Gi"}, architecture="arm64", operating_systems=("linux",),
        offerings=[Offering([req(CAPACITY_TYPE_LABEL, "In", "spot"), req(ZONE_LABEL, "In", "test-zone-1-spot")], price, True)])
    return [mk("instance-type-1", "1", 0.52), mk("instance-type-2", "4", 1.0)]


def _solve(which, pool, its, pods, **kw):
    def run(backend):
        s = Scheduler([pool], {pool.name: its}, backend=backend, **kw)
        try:
            return s.solve(pods)
        finally:
            s.close()
    r = run(oracle_lib.solve)
    if which == "gpu":
        from tests.parity import assert_same
        g = run(None)
        assert_same(g.raw, r.raw, "truncate ")
        r = g
    return r


@pytest.mark.parametrize("which", BACKENDS)
def test_truncation_that_breaks_min_values_fails_the_pod(which):  # instance_selection_test.go:1289-1345
    pool = NodePool(name="default", requirements=[req(INSTANCE_TYPE_LABEL, "In", "instance-type-1", "instance-type-2", min_values=2),
                                                  req(ARCH_LABEL, "In", "arm64")])
    pod = [Pod(name="p", uid=1, requests={"cpu": "100m"})]
    r = _solve(which, pool, _two_types(), pod, max_instance_types=1)
    assert not r.new_node_claims and list(r.pod_errors.values()) == ["pod didn't schedule because NodePool couldn't meet minValues requirements"]
    assert int(r.raw["claim_dropped"][0]) == 1 and int(r.raw["pod_error"][0]) == 4
    # without the cap, or under BestEffort, the pod schedules
    assert len(_solve(which, pool, _two_types(), pod).new_node_claims) == 1
    r = _solve(which, pool, _two_types(), pod, max_instance_types=1, min_values_policy="BestEffort")
    assert len(r.new_node_claims) == 1 and r.new_node_claims[0].instance_type_options == ["instance-type-1"]


@pytest.mark.parametrize("which", BACKENDS)
def test_the_cheapest_types_survive(which):  # types.go:238-257,339-351
    its = []
    for i in range(40):  # price falls with the index: the order is not the provider's
        its.append(fake.new_instance_type(f"t-{i:02d}", {"cpu": "4", "memory": "8Gi"}, offerings=[
            Offering([req(CAPACITY_TYPE_LABEL, "In", "on-demand"), req(ZONE_LABEL, "In", "test-zone-1")], 2.0 - 0.01 * i, True),
            Offering([req(CAPACITY_TYPE_LABEL, "In", "spot"), req(ZONE_LABEL, "In", "test-zone-2")], 0.5 + 0.01 * i, True)]))
    pool = NodePool(name="default", requirements=[req(ARCH_LABEL, "In", "amd64", "arm64")])
    pods = [Pod(name="p", uid=1, requests={"cpu": "1"}, node_selector={CAPACITY_TYPE_LABEL: "on-demand"})]
    r = _solve(which, pool, its, pods, max_instance_types=10)
    # on-demand only: the cheapest on-demand offerings belong to the LAST types
    assert sorted(r.new_node_claims[0].instance_type_options) == [f"t-{i:02d}" for i in range(30, 40)]
    pods = [Pod(name="p", uid=1, requests={"cpu": "1"})]
    r = _solve(which, pool, its, pods, max_instance_types=10)
    assert sorted(r.new_node_claims[0].instance_type_options) == [f"t-{i:02d}" for i in range(10)]   # spot is cheaper: the first ones


@pytest.mark.gpu
def test_truncate_on_the_1000_type_catalog_parity():
    """C3's shape on the 1 000-type AWS-KWOK catalog: the NodeClaims end with 240 - 430 types; the 600 (nothing to cut), 300 and 60
    cheapest kept."""
    from karpenter_b200 import _native, workloads
    from tests.parity import assert_same
    h = _native.Handle()
    try:
        for cap in (600, 300, 60):
            enc = workloads.config_c3(n_apps=20, replicas=200, n_its=1000)
            enc.problem.set("max_instance_types", cap)
            gpu, orc = h.solve(enc.problem), oracle_lib.solve(enc.problem, threads=8)
            assert_same(gpu, orc, f"truncate {cap} ")
            n = np.array([sum(bin(int(w)).count("1") for w in row) for row in gpu["claim_its"]])
            assert n.max() <= cap and (cap == 600 or (n == cap).any())
    finally:
        h.close()
