"""resources.Ceiling / RequestsForPods (pkg/utils/resources/resources.go:30-39,113-118) on the host side of the boundary
(karpenter_b200/model.py ceiling / effective_requests): the reference's own 15 known-answer cases
(pkg/utils/resources/suite_test.go:40-651 -> tests/golden/ceiling_kats.json via tests/golden/extract_ceiling_kats.py), and
that the ceiled requests are what reaches the solver."""
import json
import os

import pytest

from karpenter_b200 import fake
from karpenter_b200.model import Container, NodePool, Pod, ceiling, effective_requests, quantity_units
from karpenter_b200.scheduler import Scheduler
from tests import oracle_lib

KATS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ceiling_kats.json")))["cases"]


def _pod(spec):
    return Pod(requests=spec["requests"], limits=spec["limits"], overhead=spec["overhead"],
               pod_level_requests=spec["pod_level_requests"], pod_level_limits=spec["pod_level_limits"],
               init_containers=[Container(c["requests"], c["limits"], c["restart_always"]) for c in spec["init_containers"]])


def test_there_are_fifteen_reference_cases():
    assert len(KATS) == 15 and all(k["expected"]["requests"] for k in KATS)


@pytest.mark.parametrize("kat", KATS, ids=[f"suite_test.go:{k['line']}" for k in KATS])
def test_ceiling_kat(kat):
    req, lim = ceiling(_pod(kat["pod"]))
    assert req == {k: quantity_units(k, v) for k, v in kat["expected"]["requests"].items()}, kat["name"]
    assert lim == {k: quantity_units(k, v) for k, v in kat["expected"]["limits"].items()}, kat["name"]


def test_ceiled_requests_reach_the_solver():
    """updateCachedPodData uses RequestsForPods (scheduler.go:471-491): a pod whose sidecar + init container push it past a
    small instance type must get the bigger one, and the NodeClaim's requests are the ceiled ones + pods: 1."""
    its = fake.default_instance_types()
    pool = NodePool(name="default")
    pod = Pod(name="p", uid=1, requests={"cpu": "500m"}, overhead={"cpu": "250m"},
              init_containers=[Container({"cpu": "1"}, {}, True), Container({"cpu": "2"}, {}, False)])
    assert effective_requests(pod) == {"cpu": 3250}  # max(0.5 + 1, 2 + 1) + 0.25
    s = Scheduler([pool], {"default": its}, backend=oracle_lib.solve)
    r = s.solve([pod])
    (claim,) = r.new_node_claims
    assert claim.requests["cpu"] >= 3250 and claim.requests["pods"] == 1
    by = {it.name: it for it in its}
    assert all(quantity_units("cpu", by[n].capacity["cpu"]) >= 3250 for n in claim.instance_type_options)
    plain = Scheduler([pool], {"default": its}, backend=oracle_lib.solve).solve([Pod(name="q", uid=2, requests={"cpu": "500m"})])
    assert len(plain.new_node_claims[0].instance_type_options) > len(claim.instance_type_options)
