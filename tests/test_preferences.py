"""Preference relaxation (Preferences.Relax, preferences.go:38-146; trySchedule, scheduler.go:438-469).

Restates the reference's "Preferential Fallback" cases (suite_test.go:1125-1245) and the soft topology cases of
topology_test.go (ScheduleAnyway spreads, preferred pod affinity / anti-affinity) with the outcomes the reference
asserts.  CPU tier: through the oracle.  GPU tier: through the CUDA path, bit-identical to the oracle.
"""
from collections import Counter

import pytest

from karpenter_b200 import encode, fake
from karpenter_b200.model import (CAPACITY_TYPE_LABEL, HOSTNAME_LABEL, INSTANCE_TYPE_LABEL, ZONE_LABEL, LabelSelector, NodePool,
                                  NodeSelectorRequirement, Pod, PodAffinityTerm, PreferredSchedulingTerm, Taint, Toleration,
                                  TopologySpreadConstraint, WeightedPodAffinityTerm)
from tests.test_reference_scenarios import BACKENDS, nodepool, pods, req, run, zone_of

LABELS = {"test": "test"}
SEL = LabelSelector.of(LABELS)


def pref(weight, *reqs):
    return PreferredSchedulingTerm(weight, tuple(reqs))


# ---- Preferences.Relax itself: order and stopping rule (preferences.go:38-57) ----------------------------------------
def test_relax_order():
    p = Pod(name="p", node_affinity_required=[[req(ZONE_LABEL, "In", "a")], [req(ZONE_LABEL, "In", "b")]],
            node_affinity_preferred=[pref(1, req(ZONE_LABEL, "In", "c")), pref(5, req(ZONE_LABEL, "In", "d"))],
            pod_affinity_preferred=[WeightedPodAffinityTerm(1, PodAffinityTerm(SEL, ZONE_LABEL))],
            pod_anti_affinity_preferred=[WeightedPodAffinityTerm(1, PodAffinityTerm(SEL, HOSTNAME_LABEL))],
            topology_spread_constraints=[TopologySpreadConstraint(1, ZONE_LABEL, SEL, "ScheduleAnyway"),
                                         TopologySpreadConstraint(1, HOSTNAME_LABEL, SEL)])
    steps = []
    while p is not None:
        steps.append((len(p.node_affinity_required), len(p.pod_affinity_preferred), len(p.pod_anti_affinity_preferred),
                      [t.weight for t in p.node_affinity_preferred], len(p.topology_spread_constraints),
                      len(p.tolerations)))
        p = encode.relax(p, True)
    assert steps == [(2, 1, 1, [1, 5], 2, 0), (1, 1, 1, [1, 5], 2, 0), (1, 0, 1, [1, 5], 2, 0), (1, 0, 0, [1, 5], 2, 0),
                     (1, 0, 0, [1], 2, 0),  # the heaviest (5) goes first
                     (1, 0, 0, [], 2, 0), (1, 0, 0, [], 1, 0), (1, 0, 0, [], 1, 1)]
    # without a PreferNoSchedule taint on any NodePool no toleration is added
    q = Pod(name="q")
    assert encode.relax(q, False) is None
    assert encode.relax(Pod(name="q", tolerations=[Toleration("", "Exists", "", "PreferNoSchedule")]), True) is None


# ---- Required (suite_test.go:1126-1165) ------------------------------------------------------------------------------
@pytest.mark.parametrize("which", BACKENDS)
def test_final_required_term_is_not_relaxed(which):  # suite_test.go:1127-1142
    np_ = nodepool(requirements=[req(ZONE_LABEL, "In", "test-zone-1"), req(INSTANCE_TYPE_LABEL, "In", "default-instance-type")])
    r = run(which, pods(1, node_affinity_required=[[req(ZONE_LABEL, "In", "invalid")]]), np_)
    assert not r.new_node_claims and len(r.pod_errors) == 1


@pytest.mark.parametrize("which", BACKENDS)
def test_required_terms_relax_in_order(which):  # suite_test.go:1143-1165: invalid, invalid, zone-1, (zone-2 never reached)
    terms = [[req(ZONE_LABEL, "In", "invalid")], [req(ZONE_LABEL, "In", "invalid")], [req(ZONE_LABEL, "In", "test-zone-1")],
             [req(ZONE_LABEL, "In", "test-zone-2")]]
    r = run(which, pods(1, node_affinity_required=terms))
    assert not r.pod_errors and zone_of(r.new_node_claims[0]) == "test-zone-1"


# ---- Preferred (suite_test.go:1166-1245) -----------------------------------------------------------------------------
@pytest.mark.parametrize("which", BACKENDS)
def test_all_preferred_terms_relax(which):  # suite_test.go:1167-1186
    prefs = [pref(1, req(ZONE_LABEL, "In", "invalid")), pref(1, req(INSTANCE_TYPE_LABEL, "In", "invalid"))]
    r = run(which, pods(1, node_affinity_preferred=prefs))
    assert not r.pod_errors and len(r.new_node_claims) == 1


@pytest.mark.parametrize("which", BACKENDS)
def test_relax_to_lighter_weights(which):  # suite_test.go:1187-1212: weight 100 impossible, 50 -> zone-2, 1 never reached
    np_ = nodepool(requirements=[req(ZONE_LABEL, "In", "test-zone-1", "test-zone-2")])
    prefs = [pref(100, req(INSTANCE_TYPE_LABEL, "In", "test-zone-3")), pref(50, req(ZONE_LABEL, "In", "test-zone-2")),
             pref(1, req(ZONE_LABEL, "In", "test-zone-1"))]
    r = run(which, pods(1, node_affinity_preferred=prefs), np_)
    assert not r.pod_errors and zone_of(r.new_node_claims[0]) == "test-zone-2"


@pytest.mark.parametrize("which", BACKENDS)
def test_preference_conflicting_with_requirement(which):  # suite_test.go:1213-1233
    r = run(which, pods(1, node_affinity_preferred=[pref(1, req(ZONE_LABEL, "NotIn", "test-zone-3"))],
                        node_affinity_required=[[req(ZONE_LABEL, "In", "test-zone-3")]]))
    assert not r.pod_errors and zone_of(r.new_node_claims[0]) == "test-zone-3"


@pytest.mark.parametrize("which", BACKENDS)
def test_conflicting_preference_requirements(which):  # suite_test.go:1234-1243: In invalid AND NotIn invalid in one term
    r = run(which, pods(1, node_affinity_preferred=[pref(1, req(ZONE_LABEL, "In", "invalid"), req(ZONE_LABEL, "NotIn", "invalid"))]))
    assert not r.pod_errors and len(r.new_node_claims) == 1


@pytest.mark.parametrize("which", BACKENDS)
def test_satisfiable_preference_is_honoured(which):  # suite_test.go:335-349 family: a preference that CAN hold, holds
    r = run(which, pods(3, node_affinity_preferred=[pref(1, req(ZONE_LABEL, "In", "test-zone-2"))]))
    assert not r.pod_errors and {zone_of(c) for c in r.new_node_claims} == {"test-zone-2"}


@pytest.mark.parametrize("which", BACKENDS)
def test_failed_pod_requeues_with_its_preferences(which):  # scheduler.go:405-421: the ORIGINAL pod goes back into the queue
    # the pod can never schedule (20 000 cpu); every requeue cycle starts again from the unrelaxed class, and ends
    np_ = nodepool()
    pl = pods(2, requests={"cpu": "20000"}, node_affinity_preferred=[pref(1, req(ZONE_LABEL, "In", "test-zone-2"))])
    pl += pods(3, uid0=10, requests={"cpu": "1"})
    r = run(which, pl, np_)
    assert len(r.pod_errors) == 2 and sum(len(c.pods) for c in r.new_node_claims) == 3


# ---- ScheduleAnyway spreads (topology_test.go:715-741, 1049-1085) ----------------------------------------------------
@pytest.mark.parametrize("which", BACKENDS)
def test_schedule_anyway_violates_max_skew(which):  # topology_test.go:715-741, second half: one pod already on spot
    from tests.test_reference_scenarios import _node
    tsc = [TopologySpreadConstraint(1, CAPACITY_TYPE_LABEL, SEL, "ScheduleAnyway")]
    spot_pod = Pod(name="old", uid=999, labels=LABELS, requests={"cpu": "1.1"}, topology_spread_constraints=tsc)
    its = fake.default_instance_types()
    n = _node("n1", [i for i in its if i.name == "default-instance-type"][0], ct="spot")
    n.running_pods = [spot_pod]
    n.available = {"cpu": "0", "memory": "0", "pods": 0}  # full: the new pods need new nodes
    np_ = nodepool()
    np_.requirements = [req(CAPACITY_TYPE_LABEL, "In", "on-demand")]
    from karpenter_b200.scheduler import Scheduler
    from tests import oracle_lib
    from tests.parity import assert_same
    pl = pods(5, labels=LABELS, requests={"cpu": "1.1"}, topology_spread_constraints=tsc)

    def go(backend):
        s = Scheduler([np_], {np_.name: its}, state_nodes=[n], backend=backend)
        try:
            return s.solve(pl)
        finally:
            s.close()
    r = go(oracle_lib.solve)
    if which == "gpu":
        g = go(None)
        assert_same(g.raw, r.raw, "schedule-anyway ")
        r = g
    # on-demand ends up with all 5 pods although spot holds a single one
    assert not r.pod_errors and sum(len(c.pods) for c in r.new_node_claims) == 5
    assert all(c.requirements[CAPACITY_TYPE_LABEL]["values"] == ["on-demand"] for c in r.new_node_claims)


@pytest.mark.parametrize("which", BACKENDS)
def test_do_not_schedule_zone_with_schedule_anyway_hostname(which):  # topology_test.go:1049-1085
    tscs = [TopologySpreadConstraint(1, ZONE_LABEL, SEL), TopologySpreadConstraint(1, HOSTNAME_LABEL, SEL, "ScheduleAnyway")]
    np_ = nodepool(requirements=[req(ZONE_LABEL, "In", "test-zone-1", "test-zone-2")])
    np_b = NodePool(name="b", requirements=[req(ZONE_LABEL, "In", "test-zone-3")], limits={"cpu": "0"})
    from karpenter_b200.scheduler import Scheduler
    from tests import oracle_lib
    from tests.parity import assert_same
    its = fake.default_instance_types()
    pl = pods(10, labels=LABELS, topology_spread_constraints=tscs)

    def go(backend):
        s = Scheduler([np_, np_b], {np_.name: its, np_b.name: its}, backend=backend)
        try:
            return s.solve(pl)
        finally:
            s.close()
    r = go(oracle_lib.solve)
    if which == "gpu":
        g = go(None)
        assert_same(g.raw, r.raw, "both-constraints ")
        r = g
    # one pod per zone; zone 3 is only reachable through the disabled NodePool, so the zonal skew blocks the rest
    zones = Counter(zone_of(c) for c in r.new_node_claims for _ in c.pods)
    assert sorted(zones.values()) == [1, 1] and len(r.pod_errors) == 8


@pytest.mark.parametrize("which", BACKENDS)
def test_preferred_node_affinity_does_not_restrict_spread_domains(which):  # topology_test.go:1831-1852
    tsc = [TopologySpreadConstraint(1, ZONE_LABEL, SEL)]
    r = run(which, pods(6, labels=LABELS, topology_spread_constraints=tsc,
                        node_affinity_preferred=[pref(1, req(ZONE_LABEL, "In", "test-zone-1", "test-zone-2"))]))
    zones = Counter(zone_of(c) for c in r.new_node_claims for _ in c.pods)
    assert not r.pod_errors and sorted(zones.values()) == [2, 2, 2]


# ---- preferred pod affinity / anti-affinity (topology_test.go:2230-2300, 2630-2665) ---------------------------------
@pytest.mark.parametrize("which", BACKENDS)
def test_preferred_pod_affinity_may_be_violated(which):  # topology_test.go:2230-2262
    tsc = [TopologySpreadConstraint(1, HOSTNAME_LABEL, SEL)]
    aff = Pod(name="aff", uid=500, pod_affinity_preferred=[
        WeightedPodAffinityTerm(50, PodAffinityTerm(LabelSelector.of({"security": "s2"}), HOSTNAME_LABEL))])
    r = run(which, pods(10, labels=LABELS, topology_spread_constraints=tsc) + [aff])
    assert not r.pod_errors


@pytest.mark.parametrize("which", BACKENDS)
def test_preferred_pod_anti_affinity_may_be_violated(which):  # topology_test.go:2263-2290
    anti = [WeightedPodAffinityTerm(50, PodAffinityTerm(SEL, HOSTNAME_LABEL))]
    # 10 pods that prefer to avoid each other but ALSO must all share one zone-1 node type... they still schedule
    r = run(which, pods(10, labels=LABELS, pod_anti_affinity_preferred=anti))
    assert not r.pod_errors


@pytest.mark.parametrize("which", BACKENDS)
def test_affinity_preference_conflicting_with_required_spread(which):  # topology_test.go:2630-2665
    tsc = [TopologySpreadConstraint(1, HOSTNAME_LABEL, SEL)]
    aff_labels = {"security": "s2"}
    pl = pods(3, labels=LABELS, topology_spread_constraints=tsc, pod_affinity_preferred=[
        WeightedPodAffinityTerm(50, PodAffinityTerm(LabelSelector.of(aff_labels), HOSTNAME_LABEL))])
    pl += pods(1, uid0=100, labels=aff_labels)
    r = run(which, pl)
    assert not r.pod_errors
    per_node = sorted(sum(1 for p in c.pods if p.labels == LABELS) for c in r.new_node_claims)
    assert [n for n in per_node if n] == [1, 1, 1]  # three nodes due to the required hostname spread


# ---- PreferNoSchedule (preferences.go:132-146, scheduler.go:132-142) ------------------------------------------------
@pytest.mark.parametrize("which", BACKENDS)
def test_prefer_no_schedule_taint_is_tolerated_after_relaxation(which):  # suite_test.go "PreferNoSchedule" family
    np_ = nodepool(taints=[Taint("soft", "x", "PreferNoSchedule")])
    r = run(which, pods(2, requests={"cpu": "1"}), np_)
    assert not r.pod_errors and sum(len(c.pods) for c in r.new_node_claims) == 2


@pytest.mark.parametrize("which", BACKENDS)
def test_no_schedule_taint_is_never_relaxed(which):
    np_ = nodepool(taints=[Taint("hard", "x", "NoSchedule"), Taint("soft", "x", "PreferNoSchedule")])
    r = run(which, pods(1, requests={"cpu": "1"}), np_)
    assert len(r.pod_errors) == 1


# ---- PreferencePolicy Ignore (scheduler.go:81-101, topology.go:431,471,481) ----------------------------------------
def test_preference_policy_ignore_drops_soft_constraints():
    b = encode.ProblemBuilder()
    b.preference_policy = "Ignore"
    p = Pod(name="p", labels=LABELS, node_affinity_preferred=[pref(1, req(ZONE_LABEL, "In", "test-zone-2"))],
            pod_anti_affinity_preferred=[WeightedPodAffinityTerm(1, PodAffinityTerm(SEL, HOSTNAME_LABEL))],
            topology_spread_constraints=[TopologySpreadConstraint(1, ZONE_LABEL, SEL, "ScheduleAnyway")])
    c = b.pod_class(p)
    assert b.class_rows[c]["tscs"] == [] and b.class_rows[c]["reqset"] == b.class_rows[c]["strict"]
    assert b.pod_class(Pod(name="q", labels=LABELS)) == c  # indistinguishable from the bare pod
    assert b.relax_chains() == [-1]


# ---- BenchmarkRespectPreferences / BenchmarkIgnorePreferences (scheduling_benchmark_test.go:104-109, 217-256, 379-427) ---
def preference_benchmark(n_pods, policy, backend):
    from karpenter_b200.scheduler import Scheduler
    np_ = NodePool(name="default", requirements=[req(CAPACITY_TYPE_LABEL, "In", "spot", "on-demand", "reserved")],
                   limits={"cpu": "10000000", "memory": "10000000Gi"})
    nginx = LabelSelector.of({"app": "nginx"})
    pl = pods(n_pods, labels={"app": "nginx"}, requests={"cpu": "500m", "memory": "512Mi"},
              node_affinity_preferred=[pref(1, req(ZONE_LABEL, "In", "test-zone-1"))],  # satisfiable
              pod_anti_affinity_preferred=[WeightedPodAffinityTerm(10, PodAffinityTerm(nginx, ZONE_LABEL)),      # not satisfiable
                                           WeightedPodAffinityTerm(1, PodAffinityTerm(nginx, HOSTNAME_LABEL))])  # satisfiable
    s = Scheduler([np_], {np_.name: fake.instance_types(400)}, backend=backend, preference_policy=policy)
    try:
        return s.solve(pl)
    finally:
        s.close()


@pytest.mark.parametrize("policy", ["Respect", "Ignore"])
@pytest.mark.parametrize("which", BACKENDS)
def test_preference_benchmark_pods(which, policy):
    from tests import oracle_lib
    from tests.parity import assert_same
    n = 400 if which == "oracle" else 4000
    r = preference_benchmark(n, policy, oracle_lib.solve)
    if which == "gpu":
        g = preference_benchmark(n, policy, None)
        assert_same(g.raw, r.raw, f"preference benchmark {policy} ")
        r = g
    assert not r.pod_errors  # the benchmark fails on any PodError (scheduling_benchmark_test.go:177-179)
    if policy == "Respect":
        # the zonal anti-affinity preference holds for the first pod only (one zone is preferred), the hostname one always:
        # one pod per node, all in the preferred zone
        assert len(r.new_node_claims) == n and all(zone_of(c) == "test-zone-1" for c in r.new_node_claims)
    else:
        assert len(r.new_node_claims) < n // 10


# ---- Volume topology requirements (provisioning/suite_test.go:1852-2260; nodeclaim.go:136-176) ----------------------
@pytest.mark.parametrize("which", BACKENDS)
def test_volume_zone_is_respected(which):  # provisioning/suite_test.go:2036-2046: PV bound in test-zone-3
    r = run(which, pods(1, volume_requirements=[[req(ZONE_LABEL, "In", "test-zone-3")]]))
    assert not r.pod_errors and zone_of(r.new_node_claims[0]) == "test-zone-3"


@pytest.mark.parametrize("which", BACKENDS)
def test_volume_zone_incompatible_with_pod(which):  # provisioning/suite_test.go:2182-2194
    r = run(which, pods(1, volume_requirements=[[req(ZONE_LABEL, "In", "test-zone-3")]],
                        node_affinity_required=[[req(ZONE_LABEL, "In", "test-zone-1")]]))
    assert len(r.pod_errors) == 1


@pytest.mark.parametrize("which", BACKENDS)
def test_storage_class_zones(which):  # provisioning/suite_test.go:1968-1980: storage class allows zones 2 and 3
    r = run(which, pods(1, volume_requirements=[[req(ZONE_LABEL, "In", "test-zone-2", "test-zone-3")]]))
    z = r.new_node_claims[0].requirements[ZONE_LABEL]
    assert not r.pod_errors and sorted(z["values"]) == ["test-zone-2", "test-zone-3"] and not z["complement"]


@pytest.mark.parametrize("which", BACKENDS)
def test_volume_zone_survives_relaxation(which):  # provisioning/suite_test.go:2218-2258
    terms = [[req("example.com/label", "In", "unsupported")], [req(CAPACITY_TYPE_LABEL, "In", "on-demand")]]
    r = run(which, pods(1, volume_requirements=[[req(ZONE_LABEL, "In", "test-zone-3")]], node_affinity_required=terms))
    assert not r.pod_errors and zone_of(r.new_node_claims[0]) == "test-zone-3"


@pytest.mark.parametrize("which", BACKENDS)
def test_volume_zone_does_not_restrict_spread_domains(which):  # nodeclaim.go:160-176: the spread still sees three zones
    tsc = [TopologySpreadConstraint(1, ZONE_LABEL, SEL)]
    # all pods mount a zone-1 volume: the first lands in zone 1; the skew over zones 1, 2, 3 (0 pods in 2 and 3) then
    # blocks the others, which a pod whose OWN requirement were zone 1 would not be
    r = run(which, pods(3, labels=LABELS, topology_spread_constraints=tsc,
                        volume_requirements=[[req(ZONE_LABEL, "In", "test-zone-1")]]))
    assert len(r.pod_errors) == 2 and zone_of(r.new_node_claims[0]) == "test-zone-1"
    r = run(which, pods(3, labels=LABELS, topology_spread_constraints=tsc, node_selector={ZONE_LABEL: "test-zone-1"}))
    assert not r.pod_errors


# ---- topology groups born mid-solve (topology.go:162-194, topologygroup.go:186-202, topologynodefilter.go:30-64) ------
def _two_pools(which, pod_list):
    """pool a: untainted, room for a few pods only; pool b: PreferNoSchedule taint (pods reach it after relaxation)"""
    from karpenter_b200.scheduler import Scheduler
    from tests import oracle_lib
    from tests.parity import assert_same
    a = NodePool(name="a", weight=10, requirements=[req(CAPACITY_TYPE_LABEL, "In", "on-demand")], limits={"cpu": "8"})
    b = NodePool(name="b", requirements=[req(CAPACITY_TYPE_LABEL, "In", "on-demand")],
                 taints=[Taint("soft", "x", "PreferNoSchedule")], limits={"cpu": "2000"})
    its = fake.default_instance_types()

    def go(backend):
        s = Scheduler([a, b], {"a": its, "b": its}, backend=backend)
        try:
            return s.solve(pod_list)
        finally:
            s.close()
    r = go(oracle_lib.solve)
    if which == "gpu":
        g = go(None)
        assert_same(g.raw, r.raw, "born mid-solve ")
        r = g
    return r


@pytest.mark.parametrize("which", BACKENDS)
def test_zone_spread_group_born_by_toleration_relaxation(which):
    # the toleration is part of the spread's node filter, so the relaxed pods get a FRESH zonal spread group that never
    # saw the pods placed through pool a: both generations are balanced on their own
    tsc = [TopologySpreadConstraint(1, ZONE_LABEL, SEL)]
    r = _two_pools(which, pods(12, labels=LABELS, requests={"cpu": "1500m"}, topology_spread_constraints=tsc))
    assert not r.pod_errors
    by_pool = {}
    for c in r.new_node_claims:
        by_pool.setdefault(c.nodepool, Counter())[zone_of(c)] += len(c.pods)
    assert set(by_pool) == {"a", "b"}
    for pool, zones in by_pool.items():
        assert max(zones.values()) - min(zones.values()) <= 1 or len(zones) < 3, (pool, zones)
    raw = r.raw
    assert raw["n_groups"] == 2  # the NewTopology group and the one born when the first pod was relaxed


@pytest.mark.parametrize("which", BACKENDS)
def test_hostname_spread_group_born_by_toleration_relaxation(which):
    # the relaxed pods count hostnames in their own, fresh group: NodeClaims opened earlier look empty to it
    tsc = [TopologySpreadConstraint(2, HOSTNAME_LABEL, SEL)]
    r = _two_pools(which, pods(10, labels=LABELS, requests={"cpu": "1500m"}, topology_spread_constraints=tsc))
    assert not r.pod_errors
    assert all(len(c.pods) <= 2 for c in r.new_node_claims)
    assert {c.nodepool for c in r.new_node_claims} == {"a", "b"}


@pytest.mark.parametrize("which", BACKENDS)
def test_spread_group_born_by_dropping_a_required_term(which):
    # two alternative node-affinity terms under a zonal spread: the terms are part of the spread's node filter, so the pod
    # relaxed to its second alternative gets its own spread group
    tsc = [TopologySpreadConstraint(1, ZONE_LABEL, SEL)]
    terms = [[req(ZONE_LABEL, "In", "invalid")], [req(ZONE_LABEL, "In", "test-zone-2", "test-zone-3")]]
    r = run(which, pods(5, labels=LABELS, topology_spread_constraints=tsc, node_affinity_required=terms))
    zones = Counter(zone_of(c) for c in r.new_node_claims for _ in c.pods)
    assert not r.pod_errors and sorted(zones.items()) in ([("test-zone-2", 2), ("test-zone-3", 3)],
                                                          [("test-zone-2", 3), ("test-zone-3", 2)])
    assert r.raw["n_groups"] == 2
