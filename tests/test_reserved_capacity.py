"""Reserved capacity: ReservationManager + offeringsToReserve + ReservedOfferingModeStrict + FinalizeScheduling
(reservationmanager.go:28-110, nodeclaim.go:240-307, scheduler.go:447-453,632-646).

Restated from the reference's "Reserved Instance Types" cases (pkg/controllers/provisioning/scheduling/suite_test.go:
4323-4900): the multi-pass harness launches every NodeClaim the way the fake cloud provider's Create does
(pkg/cloudprovider/fake/cloudprovider.go:113-190: cheapest compatible available offering, reserved offerings first, a
launch uses up one instance of the reservation and the offering turns unavailable at zero) and the next pass sees the
result.  CPU tier: the oracle; GPU tier: the CUDA path, bit-identical to the oracle pass by pass."""
import copy

import pytest

from karpenter_b200 import fake
from karpenter_b200.model import (ARCH_LABEL, CAPACITY_TYPE_LABEL, HOSTNAME_LABEL, INSTANCE_TYPE_LABEL, NODEPOOL_LABEL,
                                  OS_LABEL, RESERVATION_ID_LABEL, ZONE_LABEL, LabelSelector, NodePool, NodeSelectorRequirement,
                                  Offering, Pod, PodAffinityTerm, PreferredSchedulingTerm, StateNode, quantity_units)
from karpenter_b200.scheduler import Scheduler
from tests import oracle_lib

BACKENDS = [pytest.param("oracle", id="oracle"), pytest.param("gpu", id="gpu", marks=pytest.mark.gpu)]


def req(key, op, *values):
    return NodeSelectorRequirement(key, op, tuple(values))


def _it(name, cpu, mem_gi):
    return fake.new_instance_type(name, resources={"cpu": str(cpu), "memory": f"{mem_gi}Gi"})


def _reserve(it, rid, capacity=1, zone="test-zone-1"):
    """suite_test.go:4349-4363: a reserved offering at 1/100 000 of the price, and `reserved` among the type's capacity types."""
    for r in it.requirements:
        if r.key == CAPACITY_TYPE_LABEL and "reserved" not in r.values:
            object.__setattr__(r, "values", tuple(r.values) + ("reserved",))
    it.offerings = list(it.offerings) + [Offering(
        [req(CAPACITY_TYPE_LABEL, "In", "reserved"), req(ZONE_LABEL, "In", zone), req(RESERVATION_ID_LABEL, "In", rid)],
        fake.price_from_resources(it.capacity) / 100_000.0, True, reservation_capacity=capacity)]
    return it


def catalog():
    large, medium, small = _it("large-instance-type", 6, 6), _it("medium-instance-type", 3, 3), _it("small-instance-type", 2, 2)
    _reserve(medium, "r-medium-instance-type")
    _reserve(small, "r-small-instance-type")
    return [large, medium, small]


class World:
    """Pods, NodePools, per-pool catalogs and the nodes launched so far; one provision() == one scheduling loop."""

    def __init__(self, which, pools, catalogs):
        self.which, self.pools, self.catalogs = which, pools, catalogs
        self.nodes, self.uid, self.bound = [], 0, {}

    def pods(self, n, **kw):
        out = []
        for _ in range(n):
            self.uid += 1
            out.append(Pod(name=f"p{self.uid}", uid=self.uid, **kw))
        return out

    def _solve(self, backend, pod_list):
        s = Scheduler(self.pools, self.catalogs, state_nodes=self.nodes, backend=backend)
        try:
            return s.solve(pod_list)
        finally:
            s.close()

    @staticmethod
    def _admits(reqs, o):
        for r in o.requirements:
            rq = reqs.get(r.key)
            if rq is not None and ((r.values[0] in rq["values"]) == rq["complement"]):
                return False
        return True

    def provision(self, pod_list):
        r = self._solve(oracle_lib.solve, pod_list)
        if self.which == "gpu":
            from tests.parity import assert_same
            g = self._solve(None, pod_list)
            assert_same(g.raw, r.raw, "pass ")
            r = g
        launched = []
        for c in r.new_node_claims:
            its = {i.name: i for i in self.catalogs[c.nodepool]}
            # fake Create: instance types with an available compatible offering, cheapest such offering first
            opts = []
            for n in c.instance_type_options:
                ok = [o for o in its[n].offerings if o.available and self._admits(c.requirements, o)]
                if ok:
                    opts.append((min(o.price for o in ok), n, ok))
            assert opts, "created nodeclaim with no available offerings"
            _, name, offerings = min(opts, key=lambda t: t[0])
            it = its[name]
            offering = next((o for o in offerings if any(x.key == CAPACITY_TYPE_LABEL and x.values == ("reserved",) for x in o.requirements)), None)
            if offering is not None:  # a launch uses up one instance of the reservation
                offering.reservation_capacity -= 1
                if offering.reservation_capacity == 0:
                    offering.available = False
            else:
                offering = offerings[0]
            node_name = f"node-{len(self.nodes):03d}"
            labels = {x.key: x.values[0] for x in offering.requirements}
            labels.update({HOSTNAME_LABEL: node_name, NODEPOOL_LABEL: c.nodepool, INSTANCE_TYPE_LABEL: it.name, OS_LABEL: "linux",
                           ARCH_LABEL: [x for x in it.requirements if x.key == ARCH_LABEL][0].values[0]})
            avail = {}
            for res in ("cpu", "memory", "pods"):
                a = quantity_units(res, it.capacity[res]) - quantity_units(res, it.overhead.get(res, 0))
                a -= sum(quantity_units(res, p.requests.get(res, 0)) if res != "pods" else 1 for p in c.pods)
                avail[res] = f"{a}m" if res == "cpu" else a
            cap = dict(it.capacity)
            cap["nodes"] = 1
            node = StateNode(name=node_name, labels=labels, available=avail, capacity=cap, nodepool=c.nodepool,
                             instance_type=it.name, running_pods=list(c.pods))
            self.nodes.append(node)
            launched.append(node)
            for p in c.pods:
                self.bound[id(p)] = node
        for name, ps in r.existing_nodes.items():
            node = [n for n in self.nodes if n.name == name][0]
            for p in ps:
                self.bound[id(p)] = node
        self.last = r
        return launched

    def unbound(self, pod_list):
        return [p for p in pod_list if id(p) not in self.bound]


def _pool(name="default", weight=0):
    return NodePool(name=name, weight=weight, requirements=[req(CAPACITY_TYPE_LABEL, "In", "spot", "on-demand", "reserved")])


@pytest.mark.parametrize("which", BACKENDS)
def test_no_fallback_while_compatible_reserved_offerings_are_available(which):  # suite_test.go:4365-4419
    w = World(which, [_pool()], {"default": catalog()})
    pl = w.pods(3, requests={"cpu": "1800m"})  # fits small and medium, two do not fit medium
    nodes = w.provision(pl)
    assert len(nodes) == 1  # the one claim reserved BOTH offerings (it could launch into either); the rest must wait
    n = nodes[0]
    assert n.labels[RESERVATION_ID_LABEL] == "r-small-instance-type" and n.labels[CAPACITY_TYPE_LABEL] == "reserved"
    assert n.labels[INSTANCE_TYPE_LABEL] == "small-instance-type"
    errs = set(w.last.pod_errors.values())
    assert len(w.last.pod_errors) == 2 and all("reserved" in e for e in errs)
    pl = w.unbound(pl)
    (n,) = w.provision(pl)
    assert n.labels[RESERVATION_ID_LABEL] == "r-medium-instance-type" and n.labels[INSTANCE_TYPE_LABEL] == "medium-instance-type"
    pl = w.unbound(pl)
    (n,) = w.provision(pl)  # both reservations are used up (offerings unavailable): plain capacity now
    assert RESERVATION_ID_LABEL not in n.labels and n.labels[CAPACITY_TYPE_LABEL] != "reserved"
    assert n.labels[INSTANCE_TYPE_LABEL] == "small-instance-type"


def _anti_pods(w, n, it_name, pools=None):
    sel = LabelSelector.of({"app": "test"})
    out = []
    for i in range(n):
        terms = [req(INSTANCE_TYPE_LABEL, "In", it_name)]
        if pools:
            terms.append(req(NODEPOOL_LABEL, "In", pools[i]))
        out += w.pods(1, labels={"app": "test"}, pod_anti_affinity=[PodAffinityTerm(sel, HOSTNAME_LABEL)],
                      node_affinity_required=[terms])
    return out


@pytest.mark.parametrize("which", BACKENDS)
def test_reservation_shared_across_nodepools(which):  # suite_test.go:4420-4469
    its = catalog()
    w = World(which, [_pool("np-1"), _pool("np-2")], {"np-1": its, "np-2": its})
    pl = _anti_pods(w, 2, "small-instance-type", ["np-1", "np-2"])
    nodes = w.provision(pl)  # one reservation behind both pools: the second pod gets nothing (and no OD / spot fallback)
    assert len(nodes) == 1 and nodes[0].labels[RESERVATION_ID_LABEL] == "r-small-instance-type"
    (n,) = w.provision(w.unbound(pl))
    assert RESERVATION_ID_LABEL not in n.labels and n.labels[CAPACITY_TYPE_LABEL] != "reserved"
    assert n.labels[INSTANCE_TYPE_LABEL] == "small-instance-type"


@pytest.mark.parametrize("which", BACKENDS)
def test_distinct_reservations_for_the_same_instance_pool(which):  # suite_test.go:4470-4536
    distinct = _reserve(_it("small-instance-type", 2, 2), "r-distinct")
    w = World(which, [_pool("np-1"), _pool("np-2")], {"np-1": catalog(), "np-2": [distinct]})
    pl = _anti_pods(w, 2, "small-instance-type", ["np-1", "np-2"])
    nodes = w.provision(pl)  # two pools, two reservations: both pods go out at once
    assert len(nodes) == 2 and all(n.labels[CAPACITY_TYPE_LABEL] == "reserved" for n in nodes)
    assert {n.labels[RESERVATION_ID_LABEL] for n in nodes} == {"r-small-instance-type", "r-distinct"}


@pytest.mark.parametrize("which", BACKENDS)
def test_multiple_reservations_for_the_same_instance_pool(which):  # suite_test.go:4537-4621
    its = catalog()
    small = [i for i in its if i.name == "small-instance-type"][0]
    _reserve(small, "r-small-instance-type-2", capacity=2)
    w = World(which, [_pool()], {"default": its})
    pl = _anti_pods(w, 4, "small-instance-type")
    nodes = w.provision(pl)  # the largest compatible reservation has two instances: two claims, no more
    assert len(nodes) == 2 and all(n.labels[CAPACITY_TYPE_LABEL] == "reserved" for n in nodes)
    nodes = w.provision(w.unbound(pl))  # one instance is left somewhere; pessimistic reservations: one pod per loop
    assert len(nodes) == 1 and nodes[0].labels[CAPACITY_TYPE_LABEL] == "reserved"
    (n,) = w.provision(w.unbound(pl))
    assert RESERVATION_ID_LABEL not in n.labels and n.labels[CAPACITY_TYPE_LABEL] != "reserved"


@pytest.mark.parametrize("which", BACKENDS)
def test_no_fallback_to_a_lower_weight_nodepool(which):  # suite_test.go:4622-4697
    fallback = _reserve(_it("small-instance-type", 2, 2), "r-fallback")
    w = World(which, [_pool("np-primary", 100), _pool("np-fallback", 50)], {"np-primary": catalog(), "np-fallback": [fallback]})
    pl = _anti_pods(w, 2, "small-instance-type")
    nodes = w.provision(pl)
    # the second pod is compatible with a reserved offering of the heavier pool that is taken: it must NOT slide to
    # np-fallback (whose own reservation is free) in this loop
    assert len(nodes) == 1 and nodes[0].labels[NODEPOOL_LABEL] == "np-primary" and RESERVATION_ID_LABEL in nodes[0].labels
    (n,) = w.provision(w.unbound(pl))  # np-primary's offering is unavailable now: plain capacity of np-primary
    assert n.labels[NODEPOOL_LABEL] == "np-primary" and RESERVATION_ID_LABEL not in n.labels


@pytest.mark.parametrize("which", BACKENDS)
def test_reserved_offering_error_does_not_relax_preferences(which):  # suite_test.go:4698-4790
    np2_small = _reserve(_it("small-instance-type", 2, 2), "r-np-2")
    w = World(which, [_pool("np-1"), _pool("np-2")], {"np-1": catalog(), "np-2": [np2_small]})
    sel = LabelSelector.of({"app": "test"})
    pl = []
    for _ in range(2):  # both PREFER np-1; relaxed, np-2 (first by name order, descending) would take the second pod at once
        pl += w.pods(1, labels={"app": "test"}, pod_anti_affinity=[PodAffinityTerm(sel, HOSTNAME_LABEL)],
                     node_affinity_required=[[req(INSTANCE_TYPE_LABEL, "In", "small-instance-type")]],
                     node_affinity_preferred=[PreferredSchedulingTerm(1, (req(NODEPOOL_LABEL, "In", "np-1"),))])
    nodes = w.provision(pl)
    assert len(nodes) == 1 and nodes[0].labels[NODEPOOL_LABEL] == "np-1"
    assert len(w.last.pod_errors) == 1 and "reserved" in list(w.last.pod_errors.values())[0]


@pytest.mark.parametrize("which", BACKENDS)
def test_fallback_mode_never_fails_on_reservations(which):
    """ReservedOfferingModeFallback (scheduler.go:66-68): without DisableReservedCapacityFallback a claim that cannot
    reserve simply holds no reservation and launches on plain capacity."""
    from karpenter_b200.encode import ProblemBuilder
    w = World(which, [_pool()], {"default": catalog()})
    pl = w.pods(3, requests={"cpu": "1800m"})

    class Loose(Scheduler):
        def _builder(self):
            b = super()._builder()
            b.reserved_offering_strict = False
            return b
    def solve(backend):
        s = Loose(w.pools, w.catalogs, backend=backend)
        try:
            return s.solve(pl)
        finally:
            s.close()
    r = solve(oracle_lib.solve)
    if which == "gpu":
        from tests.parity import assert_same
        g = solve(None)
        assert_same(g.raw, r.raw, "fallback ")
        r = g
    # the three pods pack onto one large node; when the second pod joined, the small / medium reservations the claim
    # held were released again (nodeclaim.go:216-218) instead of failing the add
    assert not r.pod_errors and len(r.new_node_claims) == 1 and r.new_node_claims[0].instance_type_options == ["large-instance-type"]
    assert int(r.raw["claim_reservations"][0]) == 0 and RESERVATION_ID_LABEL not in r.new_node_claims[0].requirements
