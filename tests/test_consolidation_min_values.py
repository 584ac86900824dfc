"""minValues inside the consolidation decision (Strict policy): RemoveInstanceTypeOptionsByPriceAndMinValues
(nodeclaim.go:309-318), the spot-to-spot cap max(15, minimum needed) (consolidation.go:296-312, types.go:301-337),
filterOutSameInstanceType's second check (multinodeconsolidation.go:189-226).  The reference's cases
(consolidation_test.go:1329-1760) restated on the oracle (CPU tier) and on the CUDA path, bit-identical to the oracle."""
import pytest

from karpenter_b200 import fake
from karpenter_b200.model import (ARCH_LABEL, CAPACITY_TYPE_LABEL, INSTANCE_TYPE_LABEL, ZONE_LABEL, NodePool, Offering)
from karpenter_b200.model import NodeSelectorRequirement
from tests.test_reference_scenarios import _node, consolidate, pods


def req(key, op, *values, min_values=None):
    return NodeSelectorRequirement(key, op, tuple(values), min_values=min_values)

BACKENDS = [pytest.param("oracle", id="oracle"), pytest.param("gpu", id="gpu", marks=pytest.mark.gpu)]


def _ladder(n, ct):
    """n instance types of one capacity type in one zone, 1 cpu / 1 Gi each, prices 0.10, 0.11, ... (cheapest first)"""
    out = []
    for i in range(n):
        res = {"cpu": "1", "memory": "1Gi"}
        off = [Offering([req(CAPACITY_TYPE_LABEL, "In", ct), req(ZONE_LABEL, "In", "test-zone-1")], 0.10 + 0.01 * i, True)]
        out.append(fake.new_instance_type(f"{ct}-{i:02d}", res, offerings=off))
    return out


def _pool(min_values, *extra):
    return NodePool(name="default", requirements=[req(INSTANCE_TYPE_LABEL, "Exists", min_values=min_values),
                                                  req(ARCH_LABEL, "In", "amd64", "arm64"), *extra], limits={"cpu": "2000"})


def _single(which, ct, n_types, node_index, min_values, **kw):
    its = _ladder(n_types, ct)
    n = _node("node-0", its[node_index], ct=ct, pod_list=pods(1, requests={"cpu": "100m"}))
    (cmd,) = consolidate(which, [n], [["node-0"]], its=its,
                         np_=_pool(min_values, req(CAPACITY_TYPE_LABEL, "In", "spot", "on-demand")), **kw)
    return cmd, [it.name for it in its]


@pytest.mark.parametrize("which", BACKENDS)
def test_spot_to_spot_sends_as_many_types_as_min_values_needs(which):  # consolidation_test.go:1329-1445
    cmd, names = _single(which, "spot", 18, 17, 16, spot_to_spot=True)
    assert cmd.decision == "replace"
    assert sorted(cmd.replacement_instance_types) == names[:16]        # not 15: minValues needs 16 distinct instance types


@pytest.mark.parametrize("which", BACKENDS)
def test_spot_to_spot_keeps_the_default_fifteen_when_min_values_needs_fewer(which):  # consolidation_test.go:1546-1660
    cmd, names = _single(which, "spot", 18, 17, 10, spot_to_spot=True)
    assert cmd.decision == "replace" and sorted(cmd.replacement_instance_types) == names[:15]


@pytest.mark.parametrize("which", BACKENDS)
@pytest.mark.parametrize("ct", ["on-demand", "spot"])
def test_price_filter_that_breaks_min_values_is_no_command(which, ct):  # consolidation_test.go:1664-1745
    # the node runs the 16th cheapest of 18 types: 15 cheaper options, minValues wants 16
    cmd, _ = _single(which, ct, 18, 15, 16, spot_to_spot=True)
    assert cmd.decision == "noop" and cmd.n_new_node_claims == 1
    # with 16 cheaper options the same node is replaceable
    cmd, names = _single(which, ct, 18, 16, 16, spot_to_spot=True)
    assert cmd.decision == "replace" and sorted(cmd.replacement_instance_types) == names[:16]


@pytest.mark.parametrize("which", BACKENDS)
def test_filter_out_same_instance_type_checks_min_values_again(which):  # consolidation_test.go:1448-1545
    od = lambda name, price: fake.new_instance_type(name, {"cpu": "5"}, offerings=[Offering(
        [req(CAPACITY_TYPE_LABEL, "In", "on-demand"), req(ZONE_LABEL, "In", "test-zone-1")], price, True)])
    current, other = od("current-on-demand", 0.5), od("other-on-demand", 0.4)
    its = [current, other]
    np_ = _pool(2, req(CAPACITY_TYPE_LABEL, "In", "on-demand"))
    p = pods(4, requests={"cpu": "2"})
    nodes = [_node("node-0", current, pod_list=p[0:1]), _node("node-1", current, pod_list=p[1:2]),
             _node("node-2", current, pod_list=p[2:4])]
    # nodes 0 and 1 together: one replacement with two options (both cheaper than 1.0 in total), but one of them is the type
    # being removed -- filterOutSameInstanceType keeps only what is cheaper than it, and one type is not minValues == 2
    (pair,) = consolidate(which, nodes, [["node-0", "node-1"]], its=its, np_=np_, filter_same_instance_type=True)
    assert pair.decision == "noop" and pair.n_new_node_claims == 1
    # without that filter the command stands, with both types
    (pair,) = consolidate(which, nodes, [["node-0", "node-1"]], its=its, np_=np_, filter_same_instance_type=False)
    assert pair.decision == "replace" and sorted(pair.replacement_instance_types) == ["current-on-demand", "other-on-demand"]


@pytest.mark.parametrize("which", BACKENDS)
def test_best_effort_policy_is_refused_not_approximated(which):
    from karpenter_b200 import _native
    with pytest.raises((RuntimeError, _native.SolverError)):
        _single(which, "spot", 18, 17, 16, spot_to_spot=True, min_values_policy="BestEffort")
