"""More of the reference's topology suite (pkg/controllers/provisioning/scheduling/topology_test.go), including the
multi-round cases: a first provisioning pass launches nodes, a later pass sees them as existing nodes with bound pods.

`Cluster` plays the test environment: `provision(pods)` solves, then "launches" every NodeClaim the way the fake cloud
provider does (cheapest instance type of the options, first compatible offering, fake/cloudprovider.go:113-190) and
binds the pods, so the next pass counts them through countDomains (topology.go:328-426).  `skew()` is ExpectSkew
(expectations.go): pods matching the constraint's selector per domain, over all nodes.

CPU tier: the oracle.  GPU tier: the CUDA path on the same passes, every pass bit-identical to the oracle.
"""
from collections import Counter

import pytest

from karpenter_b200 import fake
from karpenter_b200.model import (ARCH_LABEL, CAPACITY_TYPE_LABEL, HOSTNAME_LABEL, INSTANCE_TYPE_LABEL, NODEPOOL_LABEL,
                                  OS_LABEL, ZONE_LABEL, LabelSelector, NodePool, Pod, PodAffinityTerm, StateNode, Taint,
                                  Toleration, TopologySpreadConstraint, quantity_units)
from karpenter_b200.scheduler import Scheduler
from tests import oracle_lib
from tests.test_reference_scenarios import BACKENDS, PRICE, req

LABELS = {"test": "test"}
SEL = LabelSelector.of(LABELS)
RR = {"cpu": "1.1"}  # fills a 2-cpu small instance: the launched node cannot take a second such pod


class Cluster:
    def __init__(self, which, pools=None, its=None, daemon_overhead=None):
        self.which = which
        self.daemon_overhead = daemon_overhead
        self.pools = pools or [NodePool(name="default", requirements=[req(CAPACITY_TYPE_LABEL, "In", "spot", "on-demand", "reserved")],
                                        limits={"cpu": "2000"})]
        self.its = its or fake.default_instance_types()
        self.nodes = []
        self.uid = 0
        self.bound = {}  # id(pod) -> node name

    def pods(self, n, **kw):
        out = []
        for _ in range(n):
            self.uid += 1
            out.append(Pod(name=f"p{self.uid}", uid=self.uid, **kw))
        return out

    def add_node(self, name, labels, pod_list=(), it_name="default-instance-type", taints=(), available=None):
        """test.Node + bound pods: a node no NodePool manages.  It is still an existing node of the Solve (every
        cluster node that is not being deleted is, provisioner.go:320-330) and feeds countDomains."""
        it = [i for i in self.its if i.name == it_name][0]
        cap = dict(it.capacity)
        n = StateNode(name=name, labels={HOSTNAME_LABEL: name, **labels}, taints=list(taints),
                      available=dict(available or {"cpu": "0", "memory": "0", "pods": 0}), capacity=cap, managed=False,
                      running_pods=list(pod_list))
        self.nodes.append(n)
        for p in pod_list:
            self.bound[id(p)] = name
        return n

    def _solve(self, backend, pod_list):
        s = Scheduler(self.pools, {p.name: self.its for p in self.pools}, state_nodes=self.nodes, backend=backend,
                      daemon_overhead=self.daemon_overhead)
        try:
            return s.solve(pod_list)
        finally:
            s.close()

    def provision(self, pod_list):
        r = self._solve(oracle_lib.solve, pod_list)
        if self.which == "gpu":
            from tests.parity import assert_same
            g = self._solve(None, pod_list)
            assert_same(g.raw, r.raw, "pass ")
            r = g
        its = {i.name: i for i in self.its}
        for name, ps in r.existing_nodes.items():
            node = [n for n in self.nodes if n.name == name][0]
            node.running_pods = list(node.running_pods) + list(ps)
            for res in ("cpu", "memory", "pods"):
                used = sum(quantity_units(res, p.requests.get(res, 0)) if res != "pods" else 1 for p in ps)
                left = quantity_units(res, node.available.get(res, 0)) - used
                node.available[res] = f"{left}m" if res == "cpu" else left
            for p in ps:
                self.bound[id(p)] = name
        for c in r.new_node_claims:
            it = its[min(c.instance_type_options, key=lambda n: (PRICE.get(n, fake.price_from_resources(its[n].capacity)), n))]
            labels = {}
            for key, rq in c.requirements.items():
                if not rq["complement"] and len(rq["values"]) == 1:
                    labels[key] = rq["values"][0]
            for o in it.offerings:  # the first offering the claim's requirements admit
                want = {x.key: x.values[0] for x in o.requirements}
                ok = o.available
                for key, v in want.items():
                    rq = c.requirements.get(key)
                    if rq is not None and ((v in rq["values"]) == rq["complement"]):
                        ok = False
                if ok:
                    labels.update(want)
                    break
            name = f"node-{len(self.nodes):03d}"
            labels.update({HOSTNAME_LABEL: name, NODEPOOL_LABEL: c.nodepool, INSTANCE_TYPE_LABEL: it.name, OS_LABEL: "linux",
                           ARCH_LABEL: [x for x in it.requirements if x.key == ARCH_LABEL][0].values[0]})
            pool = [p for p in self.pools if p.name == c.nodepool][0]
            labels.update(pool.labels)
            avail = {}
            for res in ("cpu", "memory", "pods"):
                a = quantity_units(res, it.capacity[res]) - quantity_units(res, it.overhead.get(res, 0))
                a -= sum(quantity_units(res, p.requests.get(res, 0)) if res != "pods" else 1 for p in c.pods)
                avail[res] = f"{a}m" if res == "cpu" else a
            cap = dict(it.capacity)
            cap["nodes"] = 1
            self.nodes.append(StateNode(name=name, labels=labels, taints=list(pool.taints), available=avail, capacity=cap,
                                        nodepool=c.nodepool, instance_type=it.name, running_pods=list(c.pods)))
            for p in c.pods:
                self.bound[id(p)] = name
        self.last = r
        return r

    def scheduled(self, pod):
        return id(pod) in self.bound

    def node_of(self, pod):
        return [n for n in self.nodes if n.name == self.bound[id(pod)]][0]

    def skew(self, key, selector=SEL, namespace="default"):
        cnt = Counter()
        for n in self.nodes:
            if key not in n.labels:
                continue
            for p in n.running_pods:
                if p.namespace == namespace and _selects(selector, p.labels):
                    cnt[n.labels[key]] += 1
        return sorted(cnt.values())

    def delete(self, pod):
        for n in self.nodes:
            n.running_pods = [p for p in n.running_pods if p is not pod]
        self.bound.pop(id(pod), None)


def _selects(sel, labels):
    if sel is None:
        return False
    if any(labels.get(k) != v for k, v in sel.match_labels):
        return False
    for k, op, vs in sel.match_expressions:
        if op == "In" and labels.get(k) not in vs:
            return False
        if op == "NotIn" and labels.get(k) in vs:
            return False
        if op == "Exists" and k not in labels:
            return False
        if op == "DoesNotExist" and k in labels:
            return False
    return True


def zone_pool(*zones):
    return [req(CAPACITY_TYPE_LABEL, "In", "spot", "on-demand", "reserved"), req(ZONE_LABEL, "In", *zones)]


def spread(key, skew=1, sel=SEL, **kw):
    return [TopologySpreadConstraint(skew, key, sel, **kw)]


# ---- Zonal (topology_test.go:106-545) --------------------------------------------------------------------------------
@pytest.mark.parametrize("which", BACKENDS)
def test_existing_pod_in_a_zone_the_nodepool_lost(which):  # topology_test.go:232-265
    c = Cluster(which)
    c.provision(c.pods(1, labels=LABELS, requests=RR, node_selector={ZONE_LABEL: "test-zone-3"}))
    c.pools[0].requirements = zone_pool("test-zone-1", "test-zone-2")
    c.provision(c.pods(6, labels=LABELS, requests=RR, topology_spread_constraints=spread(ZONE_LABEL)))
    assert c.skew(ZONE_LABEL) == [1, 2, 2]


@pytest.mark.parametrize("which", BACKENDS)
def test_non_minimum_domain_if_it_is_all_there_is(which):  # topology_test.go:266-307
    c = Cluster(which)
    tsc = spread(ZONE_LABEL, 5)
    for zone, n in (("test-zone-1", 1), ("test-zone-2", 1), ("test-zone-3", 10)):
        c.pools[0].requirements = zone_pool(zone)
        c.provision(c.pods(n, labels=LABELS, requests=RR, topology_spread_constraints=tsc))
    assert c.skew(ZONE_LABEL) == [1, 1, 6]


@pytest.mark.parametrize("which", BACKENDS)
def test_only_minimum_domains_when_already_violating(which):  # topology_test.go:308-346
    c = Cluster(which)
    tsc = spread(ZONE_LABEL)
    first = c.pods(9, labels=LABELS, requests=RR, topology_spread_constraints=tsc)
    c.provision(first)
    assert c.skew(ZONE_LABEL) == [3, 3, 3]
    for p in first:
        if c.node_of(p).labels[ZONE_LABEL] != "test-zone-1":
            c.delete(p)
    assert c.skew(ZONE_LABEL) == [3]
    c.provision(c.pods(3, labels=LABELS, requests=RR, topology_spread_constraints=tsc))
    assert c.skew(ZONE_LABEL) == [1, 2, 3]


@pytest.mark.parametrize("which", BACKENDS)
def test_do_not_schedule_keeps_max_skew(which):  # topology_test.go:347-379
    c = Cluster(which)
    tsc = spread(ZONE_LABEL)
    c.pools[0].requirements = zone_pool("test-zone-1")
    c.provision(c.pods(1, labels=LABELS, requests=RR, topology_spread_constraints=tsc))
    c.pools[0].requirements = zone_pool("test-zone-2", "test-zone-3")
    c.provision(c.pods(10, labels=LABELS, requests=RR, topology_spread_constraints=tsc))
    assert c.skew(ZONE_LABEL) == [1, 2, 2]


@pytest.mark.parametrize("which", BACKENDS)
def test_only_matching_bound_pods_on_nodes_with_the_domain_count(which):  # topology_test.go:412-444
    c = Cluster(which)
    first = c.pods(1)                                   # no labels: ignored
    wrong_ns = c.pods(1, labels=LABELS, namespace="wrong")
    counted1 = c.pods(2, labels=LABELS)
    counted2 = c.pods(1, labels=LABELS)
    nodomain = c.pods(1, labels=LABELS)
    c.add_node("first", {ZONE_LABEL: "test-zone-1"}, first + wrong_ns + counted1)
    c.add_node("second", {ZONE_LABEL: "test-zone-2"}, counted2)
    c.add_node("third", {}, nodomain)                   # no zone label: its pod is ignored
    c.provision(c.pods(2, labels=LABELS, topology_spread_constraints=spread(ZONE_LABEL)))
    assert c.skew(ZONE_LABEL) == [1, 2, 2]


@pytest.mark.parametrize("which", BACKENDS)
def test_min_domains_not_reachable(which):  # topology_test.go:482-501: two zones, minDomains 3 -> global minimum stays 0
    c = Cluster(which)
    c.pools[0].requirements = zone_pool("test-zone-1", "test-zone-2")
    c.provision(c.pods(3, labels=LABELS, topology_spread_constraints=spread(ZONE_LABEL, min_domains=3)))
    assert c.skew(ZONE_LABEL) == [1, 1]


@pytest.mark.parametrize("which", BACKENDS)
@pytest.mark.parametrize("min_domains", [3, 2])
def test_min_domains_satisfied(which, min_domains):  # topology_test.go:502-543
    c = Cluster(which)
    c.provision(c.pods(11, labels=LABELS, topology_spread_constraints=spread(ZONE_LABEL, min_domains=min_domains)))
    assert c.skew(ZONE_LABEL) == [3, 4, 4]


# ---- Hostname (topology_test.go:544-651) -----------------------------------------------------------------------------
@pytest.mark.parametrize("which", BACKENDS)
def test_hostname_up_to_max_skew(which):  # topology_test.go:558-570: maxSkew 4 -> all four pods on one node
    c = Cluster(which)
    c.provision(c.pods(4, labels=LABELS, topology_spread_constraints=spread(HOSTNAME_LABEL, 4)))
    assert c.skew(HOSTNAME_LABEL) == [4]


@pytest.mark.parametrize("which", BACKENDS)
def test_two_deployments_with_hostname_spread(which):  # topology_test.go:571-606
    c = Cluster(which)
    a, b = {"app": "a"}, {"app": "b"}
    pl = c.pods(2, labels=a, topology_spread_constraints=spread(HOSTNAME_LABEL, 1, LabelSelector.of(a)))
    pl += c.pods(2, labels=b, topology_spread_constraints=spread(HOSTNAME_LABEL, 1, LabelSelector.of(b)))
    r = c.provision(pl)
    assert not r.pod_errors
    assert c.skew(HOSTNAME_LABEL, LabelSelector.of(a)) == [1, 1] and c.skew(HOSTNAME_LABEL, LabelSelector.of(b)) == [1, 1]
    assert len(c.nodes) == 2  # each node takes one pod of each deployment


# ---- Capacity type (topology_test.go:652-940) ------------------------------------------------------------------------
@pytest.mark.parametrize("which", BACKENDS)
def test_capacity_type_do_not_schedule(which):  # topology_test.go:681-714
    c = Cluster(which)
    tsc = spread(CAPACITY_TYPE_LABEL)
    c.pools[0].requirements = [req(CAPACITY_TYPE_LABEL, "In", "spot")]
    c.provision(c.pods(1, labels=LABELS, requests=RR, topology_spread_constraints=tsc))
    assert c.skew(CAPACITY_TYPE_LABEL) == [1]
    c.pools[0].requirements = [req(CAPACITY_TYPE_LABEL, "In", "on-demand")]
    c.provision(c.pods(5, labels=LABELS, requests=RR, topology_spread_constraints=tsc))
    assert c.skew(CAPACITY_TYPE_LABEL) == [1, 2]  # on-demand gets 2, the rest fail


@pytest.mark.parametrize("which", BACKENDS)
def test_capacity_type_schedule_anyway(which):  # topology_test.go:716-745
    c = Cluster(which)
    tsc = spread(CAPACITY_TYPE_LABEL, when_unsatisfiable="ScheduleAnyway")
    c.pools[0].requirements = [req(CAPACITY_TYPE_LABEL, "In", "spot")]
    c.provision(c.pods(1, labels=LABELS, requests=RR, topology_spread_constraints=tsc))
    c.pools[0].requirements = [req(CAPACITY_TYPE_LABEL, "In", "on-demand")]
    c.provision(c.pods(5, labels=LABELS, requests=RR, topology_spread_constraints=tsc))
    assert c.skew(CAPACITY_TYPE_LABEL) == [1, 5]


@pytest.mark.parametrize("which", BACKENDS)
def test_hostname_and_zonal_constraints_together(which):  # topology_test.go:941-980
    c = Cluster(which)
    tsc = spread(ZONE_LABEL) + spread(HOSTNAME_LABEL, 3)
    for n, zones, hosts in ((2, [1, 1], None), (3, [1, 2, 2], None), (5, [3, 3, 4], None), (11, [7, 7, 7], None)):
        c.provision(c.pods(n, labels=LABELS, topology_spread_constraints=tsc))
        assert c.skew(ZONE_LABEL) == zones
        assert all(v <= 3 for v in c.skew(HOSTNAME_LABEL))


# ---- Pod affinity / anti-affinity (topology_test.go:1925-2980) --------------------------------------------------------
AFF = {"security": "s2"}
AFF_SEL = LabelSelector.of(AFF)


def same_node(c, *pods_):
    return len({c.bound[id(p)] for p in pods_}) == 1


@pytest.mark.parametrize("which", BACKENDS)
def test_pod_affinity_hostname(which):  # topology_test.go:1936-1969
    c = Cluster(which)
    target = c.pods(1, labels=AFF)
    follower = c.pods(1, pod_affinity=[PodAffinityTerm(AFF_SEL, HOSTNAME_LABEL)])
    spreaders = c.pods(10, labels=LABELS, topology_spread_constraints=spread(HOSTNAME_LABEL))
    c.provision(spreaders + target + follower)
    assert same_node(c, target[0], follower[0])


@pytest.mark.parametrize("which", BACKENDS)
def test_self_pod_affinity_hostname(which):  # topology_test.go:2013-2036
    c = Cluster(which)
    pl = c.pods(3, labels=AFF, pod_affinity=[PodAffinityTerm(AFF_SEL, HOSTNAME_LABEL)])
    c.provision(pl)
    assert same_node(c, *pl)


@pytest.mark.parametrize("which", BACKENDS)
def test_self_pod_affinity_first_empty_domain_only(which):  # topology_test.go:2037-2078: 5 pods per node
    c = Cluster(which)
    pl = c.pods(10, labels=AFF, pod_affinity=[PodAffinityTerm(AFF_SEL, HOSTNAME_LABEL)])
    c.provision(pl)
    placed = [p for p in pl if c.scheduled(p)]
    assert len(placed) == 5 and same_node(c, *placed)
    # a later batch does not schedule either: the one domain that holds matching pods is full
    later = c.pods(10, labels=AFF, pod_affinity=[PodAffinityTerm(AFF_SEL, HOSTNAME_LABEL)])
    c.provision(later)
    assert not any(c.scheduled(p) for p in later)


@pytest.mark.parametrize("which", BACKENDS)
def test_self_pod_affinity_zone(which):  # topology_test.go:2123-2146
    c = Cluster(which)
    pl = c.pods(3, labels=AFF, pod_affinity=[PodAffinityTerm(AFF_SEL, ZONE_LABEL)])
    c.provision(pl)
    assert same_node(c, *pl)


@pytest.mark.parametrize("which", BACKENDS)
def test_self_pod_affinity_zone_with_constraint(which):  # topology_test.go:2147-2177
    c = Cluster(which)
    pl = c.pods(3, labels=AFF, pod_affinity=[PodAffinityTerm(AFF_SEL, ZONE_LABEL)],
                node_affinity_required=[[req(ZONE_LABEL, "In", "test-zone-3")]])
    c.provision(pl)
    assert same_node(c, *pl) and c.node_of(pl[0]).labels[ZONE_LABEL] == "test-zone-3"


@pytest.mark.parametrize("which", BACKENDS)
def test_anti_affinity_on_zone(which):  # topology_test.go:2319-2357
    c = Cluster(which)
    zoned = [c.pods(1, labels=AFF, requests={"cpu": "2"}, node_selector={ZONE_LABEL: z})[0]
             for z in ("test-zone-1", "test-zone-2", "test-zone-3")]
    anti = c.pods(1, pod_anti_affinity=[PodAffinityTerm(AFF_SEL, ZONE_LABEL)])
    c.provision(zoned + anti)
    assert all(c.scheduled(p) for p in zoned) and not c.scheduled(anti[0])


@pytest.mark.parametrize("which", BACKENDS)
def test_anti_affinity_on_zone_other_schedules_first(which):  # topology_test.go:2358-2379
    c = Cluster(which)
    target = c.pods(1, labels=AFF, requests={"cpu": "2"})
    anti = c.pods(1, pod_anti_affinity=[PodAffinityTerm(AFF_SEL, ZONE_LABEL)])
    c.provision(target + anti)
    assert c.scheduled(target[0]) and not c.scheduled(anti[0])  # nobody knows yet which zone the target lands in


@pytest.mark.parametrize("which", BACKENDS)
def test_inverse_anti_affinity_on_zone(which):  # topology_test.go:2463-2498
    c = Cluster(which)
    term = [PodAffinityTerm(AFF_SEL, ZONE_LABEL)]
    zoned = [c.pods(1, requests={"cpu": "2"}, pod_anti_affinity=term, node_selector={ZONE_LABEL: z})[0]
             for z in ("test-zone-1", "test-zone-2", "test-zone-3")]
    victim = c.pods(1, labels=AFF)
    c.provision(zoned + victim)
    assert all(c.scheduled(p) for p in zoned) and not c.scheduled(victim[0])


@pytest.mark.parametrize("which", BACKENDS)
def test_schroedinger_anti_affinity(which):  # topology_test.go:2499-2529
    c = Cluster(which)
    anywhere = c.pods(1, requests={"cpu": "2"}, pod_anti_affinity=[PodAffinityTerm(AFF_SEL, ZONE_LABEL)])
    victim = c.pods(1, labels=AFF)
    c.provision(anywhere + victim)
    assert c.scheduled(anywhere[0]) and not c.scheduled(victim[0])  # it could be in any zone
    c.provision(victim)  # the launched node committed to one zone
    assert c.scheduled(victim[0])
    assert c.node_of(victim[0]).labels[ZONE_LABEL] != c.node_of(anywhere[0]).labels[ZONE_LABEL]


@pytest.mark.parametrize("which", BACKENDS)
def test_inverse_anti_affinity_with_existing_nodes(which):  # topology_test.go:2530-2579
    c = Cluster(which)
    term = [PodAffinityTerm(AFF_SEL, ZONE_LABEL)]
    zoned = [c.pods(1, requests={"cpu": "2"}, pod_anti_affinity=term, node_selector={ZONE_LABEL: z})[0]
             for z in ("test-zone-1", "test-zone-2", "test-zone-3")]
    c.provision(zoned)
    victim = c.pods(1, labels=AFF)
    c.provision(victim)
    assert not c.scheduled(victim[0])


@pytest.mark.parametrize("which", BACKENDS)
def test_affinity_to_a_pod_that_does_not_exist(which):  # topology_test.go:2710-2726
    c = Cluster(which)
    pl = c.pods(10, pod_affinity=[PodAffinityTerm(AFF_SEL, ZONE_LABEL)])
    c.provision(pl)
    assert not any(c.scheduled(p) for p in pl)


@pytest.mark.parametrize("which", BACKENDS)
@pytest.mark.parametrize("order", [(0, 1, 2, 3), (3, 2, 1, 0), (2, 0, 3, 1), (1, 3, 0, 2)])
def test_chain_of_dependent_affinities(which, order):  # topology_test.go:2789-2823: db <- web <- cache <- ui, any queue order
    c = Cluster(which)
    lab = [dict(type=t, spread="spread") for t in ("db", "web", "cache", "ui")]
    specs = [dict(labels=lab[0])] + [dict(labels=lab[i], pod_affinity=[PodAffinityTerm(LabelSelector.of(lab[i - 1]), HOSTNAME_LABEL)])
                                     for i in (1, 2, 3)]
    pl = [None] * 4
    for slot in order:  # uids (the queue's tie-break, queue.go:97-107) in a different order each time
        pl[slot] = c.pods(1, **specs[slot])[0]
    c.provision(pl)
    assert all(c.scheduled(p) for p in pl) and same_node(c, *pl)


@pytest.mark.parametrize("which", BACKENDS)
def test_unsatisfiable_dependency_terminates(which):  # topology_test.go:2824-2839
    c = Cluster(which)
    pl = c.pods(1, labels={"type": "db"}, pod_affinity=[PodAffinityTerm(LabelSelector.of({"type": "web"}), HOSTNAME_LABEL)])
    c.provision(pl)
    assert not c.scheduled(pl[0])


@pytest.mark.parametrize("which", BACKENDS)
@pytest.mark.parametrize("namespaces,together", [((), False), (("other-ns",), True)])
def test_affinity_namespaces(which, namespaces, together):  # topology_test.go:2840-2916
    c = Cluster(which)
    target = c.pods(1, labels=AFF, namespace="other-ns")
    follower = c.pods(1, pod_affinity=[PodAffinityTerm(AFF_SEL, HOSTNAME_LABEL, namespaces)])
    spreaders = c.pods(10, labels=LABELS, topology_spread_constraints=spread(HOSTNAME_LABEL))
    c.provision(spreaders + target + follower)
    assert c.scheduled(target[0])
    if together:
        assert same_node(c, target[0], follower[0])
    else:
        assert not c.scheduled(follower[0])  # the term only looks at the follower's own namespace


# ---- node inclusion policies (topology_test.go:1196-1661) ------------------------------------------------------------
SPREAD_LABEL, TINY = "fake-label", {"cpu": "100m", "memory": "1Gi", "pods": 10}


def _pool(labels=None, requirements=(), taints=(), name="default"):
    return NodePool(name=name, requirements=[req(CAPACITY_TYPE_LABEL, "In", "spot", "on-demand", "reserved")] + list(requirements),
                    labels=dict(labels or {}), taints=list(taints), limits={"cpu": "2000"})


@pytest.mark.parametrize("which", BACKENDS)
@pytest.mark.parametrize("policy,expect", [("Ignore", [1]), ("Honor", [5])])
def test_node_taints_policy_with_tainted_nodes(which, policy, expect):  # topology_test.go:1196-1335
    c = Cluster(which, pools=[_pool({SPREAD_LABEL: "baz"})])
    taint = [Taint("taintname", "taintvalue", "NoSchedule")]
    c.add_node("n1", {SPREAD_LABEL: "foo"}, taints=taint, available=TINY)
    c.add_node("n2", {SPREAD_LABEL: "bar"}, taints=taint, available=TINY)
    c.provision(c.pods(5, labels=LABELS, requests={"cpu": "1"},
                       topology_spread_constraints=spread(SPREAD_LABEL, node_taints_policy=policy)))
    # Ignore: foo and bar are domains the pods cannot reach, so only one pod fits before the skew blocks baz;
    # Honor: the tainted nodes do not count as domains at all
    assert c.skew(SPREAD_LABEL) == expect


@pytest.mark.parametrize("which", BACKENDS)
@pytest.mark.parametrize("policy,expect", [("Ignore", [1]), ("Honor", [2])])
def test_node_taints_policy_domains_from_nodepools(which, policy, expect):  # topology_test.go:1336-1447
    pools = [_pool(requirements=[req(SPREAD_LABEL, "In", "foo")]),
             NodePool(name="tainted", requirements=[req(CAPACITY_TYPE_LABEL, "Exists"), req(SPREAD_LABEL, "In", "bar")],
                      taints=[Taint("taint-key", "taint-value", "NoSchedule")], limits={"cpu": "2000"})]
    c = Cluster(which, pools=pools)
    c.provision(c.pods(2, labels=LABELS, topology_spread_constraints=spread(SPREAD_LABEL, node_taints_policy=policy)))
    assert c.skew(SPREAD_LABEL) == expect


@pytest.mark.parametrize("which", BACKENDS)
@pytest.mark.parametrize("policy,expect", [("Ignore", [1]), ("Honor", [5])])
def test_node_affinity_policy(which, policy, expect):  # topology_test.go:1529-1661
    c = Cluster(which, pools=[_pool({SPREAD_LABEL: "baz", "selector": "value"})])
    c.add_node("n1", {SPREAD_LABEL: "foo", "selector": "mismatch"}, available=TINY)
    c.add_node("n2", {SPREAD_LABEL: "bar", "selector": "mismatch"}, available=TINY)
    c.provision(c.pods(5, labels=LABELS, requests={"cpu": "1"}, node_selector={"selector": "value"},
                       topology_spread_constraints=spread(SPREAD_LABEL, node_affinity_policy=policy)))
    assert c.skew(SPREAD_LABEL) == expect
