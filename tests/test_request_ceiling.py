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


def test_template_keys_intern_the_same_classes():
    """Pod.template (an owner's pod-template key) lets the encoder intern a spec once per key: same problem, bit for bit."""
    import random
    import numpy as np
    from karpenter_b200.scheduler import Scheduler
    from tests import fuzz
    rng = random.Random(3)
    pl = fuzz.pods(rng, 3000)
    pools, its = fuzz.node_pools(rng), fuzz.instance_types(rng)
    s = Scheduler(pools, {p.name: its for p in pools}, backend=lambda p: None)
    plain = s.encode(pl).problem
    keys = {}
    for p in pl:
        p.template = keys.setdefault(repr((p.requests, p.labels, p.namespace, p.node_selector, p.node_affinity_required, p.tolerations,
                                           p.topology_spread_constraints, p.pod_affinity, p.pod_anti_affinity)), len(keys))
    keyed = s.encode(pl).problem
    assert plain.n_classes == keyed.n_classes
    for k in ("pod_class", "class_requests", "class_reqset", "pod_uid_lo"):
        assert np.array_equal(plain.get(k), keyed.get(k)), k
