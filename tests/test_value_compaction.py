"""Keys with more than 64 values (node.kubernetes.io/instance-type: one value per instance type) are encoded by merging
the values no pod / NodePool / offering mentions into one OTHER value (encode.py).  Checked on the oracle (CPU) and on the
CUDA path (GPU): selecting instance types by name works, and compaction never changes a result."""
import numpy as np
import pytest

from karpenter_b200 import kwok
from karpenter_b200.model import (CAPACITY_TYPE_LABEL, INSTANCE_TYPE_LABEL, ZONE_LABEL, NodePool, NodeSelectorRequirement,
                                  Pod)
from karpenter_b200.scheduler import Scheduler
from tests import oracle_lib

BACKENDS = [pytest.param("oracle", id="oracle"), pytest.param("gpu", id="gpu", marks=pytest.mark.gpu)]


def _solve(which, pools, its, pods, max_values=64):
    s = Scheduler(pools, {p.name: its for p in pools}, backend=oracle_lib.solve if which == "oracle" else None)
    orig = s._builder

    def builder():
        b = orig()
        b.max_values_per_key = max_values
        return b
    s._builder = builder
    try:
        return s.solve(pods)
    finally:
        s.close()


@pytest.mark.parametrize("which", BACKENDS)
def test_select_instance_type_by_name(which):
    its = kwok.aws_instance_types(600)  # 300+ distinct names: far beyond a 64-bit value mask
    names = sorted({it.name for it in its})
    pool = NodePool(name="default", requirements=[NodeSelectorRequirement(CAPACITY_TYPE_LABEL, "In", ("on-demand",))])
    want, avoid = names[10], names[20]
    pods = ([Pod(name=f"a{i}", uid=i + 1, requests={"cpu": "100m"}, node_selector={INSTANCE_TYPE_LABEL: want}) for i in range(5)] +
            [Pod(name=f"b{i}", uid=100 + i, requests={"cpu": "100m"},
                 node_affinity_required=[[NodeSelectorRequirement(INSTANCE_TYPE_LABEL, "NotIn", (avoid, want))]]) for i in range(5)] +
            [Pod(name="c", uid=999, requests={"cpu": "100m"}, node_selector={INSTANCE_TYPE_LABEL: "no-such-type"})])
    r = _solve(which, [pool], its, pods)
    assert len(r.pod_errors) == 1 and id(pods[-1]) in r.pod_errors
    for c in r.new_node_claims:
        kinds = {p.name[0] for p in c.pods}
        assert len(kinds) == 1  # In{want} and NotIn{want} never share a NodeClaim
        if kinds == {"a"}:
            assert set(c.instance_type_options) == {want}
        else:
            assert want not in c.instance_type_options and avoid not in c.instance_type_options
            assert len(set(c.instance_type_options)) > 100


@pytest.mark.parametrize("which", BACKENDS)
def test_compaction_does_not_change_results(which):
    """Same problem with the zone key encoded in full and compacted to {mentioned zones} + OTHER."""
    its = kwok.aws_instance_types(200)
    pool = NodePool(name="default", requirements=[NodeSelectorRequirement(CAPACITY_TYPE_LABEL, "In", ("on-demand", "spot"))])
    rng = np.random.default_rng(3)
    pods = []
    for i in range(3000):
        sel = {}
        z = int(rng.integers(0, 3))
        if z < 2:
            sel[ZONE_LABEL] = kwok.AWS_ZONES[z]  # zones c and d are never mentioned
        term = [[NodeSelectorRequirement(ZONE_LABEL, "NotIn", (kwok.AWS_ZONES[0],))]] if rng.random() < 0.2 and not sel else []
        pods.append(Pod(name=f"p{i}", uid=int(rng.integers(1, 1 << 60)), requests={"cpu": f"{int(rng.integers(1, 16)) * 250}m", "memory": f"{int(rng.integers(1, 8))}Gi"},
                        node_selector=sel, node_affinity_required=term))
    full = _solve(which, [pool], its, pods, max_values=64)
    compact = _solve(which, [pool], its, pods, max_values=2)
    assert np.array_equal(full.raw["pod_target"], compact.raw["pod_target"])
    assert np.array_equal(full.raw["claim_its"], compact.raw["claim_its"])
    assert np.array_equal(full.raw["claim_requests"], compact.raw["claim_requests"])
    assert full.raw["n_claims"] > 3
