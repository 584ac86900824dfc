"""Host ports (pkg/scheduling/hostportusage.go:35-118; NodeClaim.CanAdd nodeclaim.go:120-124, ExistingNode.CanAdd
existingnode.go:76-82, daemon ports scheduler.go:794-811).  The HostPort.Matches known-answer cases of the reference
(hostportusage_test.go:41-105) run against the encoder's conflict matrix; the scheduling behaviour on the oracle (CPU tier)
and on the CUDA path, bit-identical to the oracle (GPU tier)."""
import numpy as np
import pytest

from karpenter_b200 import fake
from karpenter_b200.model import (CAPACITY_TYPE_LABEL, HOSTNAME_LABEL, NodePool, NodeSelectorRequirement, Pod, StateNode,
                                  host_ports_match)
from karpenter_b200.scheduler import Scheduler
from tests import oracle_lib

BACKENDS = [pytest.param("oracle", id="oracle"), pytest.param("gpu", id="gpu", marks=pytest.mark.gpu)]


def test_matches_kats():  # hostportusage_test.go:41-105
    e1 = ("10.0.0.0", 4443, "TCP")
    assert host_ports_match(e1, e1)                                                  # identical entries match
    for unspecified in ("0.0.0.0", "::", ""):                                       # one unspecified address: match
        assert host_ports_match(e1, (unspecified, 4443, "TCP")) and host_ports_match((unspecified, 4443, "TCP"), e1)
    assert not host_ports_match(e1, ("10.0.0.0", 4443, "SCTP"))                      # mismatched protocols
    assert not host_ports_match(e1, ("10.0.0.0", 443, "TCP"))                        # mismatched ports
    assert not host_ports_match(e1, ("10.0.0.1", 4443, "TCP"))                       # two different specified IPs
    assert host_ports_match(("", 80, ""), ("1.2.3.4", 80, "TCP"))                    # GetHostPorts defaults: 0.0.0.0 / TCP


def test_encoder_conflict_matrix():
    pool = NodePool(name="default")
    pods = [Pod(name="a", uid=1, host_ports=[("10.0.0.0", 4443, "TCP")]), Pod(name="b", uid=2, host_ports=[("", 4443, "TCP")]),
            Pod(name="c", uid=3, host_ports=[("10.0.0.1", 4443, "TCP")]), Pod(name="d", uid=4, host_ports=[("10.0.0.0", 4443, "UDP")])]
    enc = Scheduler([pool], {"default": fake.default_instance_types()}, backend=oracle_lib.solve).encode(pods)
    p = enc.problem
    assert p.n_hostports == 4
    conf = p.get("hostport_conflicts")
    cls = p.get("class_hostports")[p.get("pod_class")]
    bit = [int(c).bit_length() - 1 for c in cls]
    m = [[bool(int(conf[bit[i]]) >> bit[j] & 1) for j in range(4)] for i in range(4)]
    assert m == [[True, True, False, False], [True, True, True, False], [False, True, True, False], [False, False, False, True]]


def _solve(which, pods, **kw):
    pool = NodePool(name="default", requirements=[NodeSelectorRequirement(CAPACITY_TYPE_LABEL, "In", ("on-demand",))])
    def run(backend):
        s = Scheduler([pool], {"default": fake.default_instance_types()}, backend=backend, **kw)
        try:
            return s.solve(pods)
        finally:
            s.close()
    r = run(oracle_lib.solve)
    if which == "gpu":
        from tests.parity import assert_same
        g = run(None)
        assert_same(g.raw, r.raw, "host ports ")
        r = g
    return r


@pytest.mark.parametrize("which", BACKENDS)
def test_pods_with_the_same_host_port_need_their_own_nodes(which):  # provisioning/suite_test.go:955-975 (daemonset ports)
    pods = [Pod(name=f"p{i}", uid=i + 1, requests={"cpu": "100m"}, host_ports=[("", 8080, "TCP")]) for i in range(3)]
    r = _solve(which, pods + [Pod(name="q", uid=9, requests={"cpu": "100m"})])
    assert not r.pod_errors and len(r.new_node_claims) == 3          # the port-less pod shares a node with one of them
    assert sorted(len(c.pods) for c in r.new_node_claims) == [1, 1, 2]


@pytest.mark.parametrize("which", BACKENDS)
def test_different_ips_share_a_node_wildcard_does_not(which):
    a = Pod(name="a", uid=1, requests={"cpu": "100m"}, host_ports=[("10.0.0.1", 443, "TCP")])
    b = Pod(name="b", uid=2, requests={"cpu": "100m"}, host_ports=[("10.0.0.2", 443, "TCP")])
    c = Pod(name="c", uid=3, requests={"cpu": "100m"}, host_ports=[("0.0.0.0", 443, "TCP")])
    assert len(_solve(which, [a, b]).new_node_claims) == 1
    assert len(_solve(which, [a, b, c]).new_node_claims) == 2
    udp = Pod(name="u", uid=4, requests={"cpu": "100m"}, host_ports=[("0.0.0.0", 443, "UDP")])
    assert len(_solve(which, [a, udp]).new_node_claims) == 1


@pytest.mark.parametrize("which", BACKENDS)
def test_daemon_host_ports_block_the_nodepool(which):  # scheduler.go:794-811: the daemonset already owns the port on every node
    p = Pod(name="p", uid=1, requests={"cpu": "100m"}, host_ports=[("", 9100, "TCP")])
    r = _solve(which, [p], daemon_host_ports={"default": [("", 9100, "TCP")]})
    assert len(r.pod_errors) == 1 and not r.new_node_claims
    r = _solve(which, [p], daemon_host_ports={"default": [("", 9101, "TCP")]})
    assert not r.pod_errors and len(r.new_node_claims) == 1


@pytest.mark.parametrize("which", BACKENDS)
def test_existing_node_ports(which):  # existingnode.go:76-82,153
    it = fake.default_instance_types()[0]
    def node(name, ports):
        return StateNode(name=name, labels={HOSTNAME_LABEL: name}, available={"cpu": "4", "memory": "4Gi", "pods": 10},
                         capacity=dict(it.capacity), managed=False, host_ports=ports)
    pods = [Pod(name=f"p{i}", uid=i + 1, requests={"cpu": "100m"}, host_ports=[("", 8080, "TCP")]) for i in range(3)]
    r = _solve(which, pods, state_nodes=[node("n-busy", [("10.1.1.1", 8080, "TCP")]), node("n-free", [])])
    # n-busy already uses 8080 on one address: the wildcard pods conflict with it; n-free takes exactly one of them
    assert list(r.existing_nodes) == ["n-free"] and len(r.existing_nodes["n-free"]) == 1
    assert len(r.new_node_claims) == 2 and not r.pod_errors
