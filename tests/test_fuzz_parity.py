"""Differential fuzzing: random scheduling problems (tests/fuzz.py) solved by the CUDA path and by the oracle must agree
bit for bit.  The CPU tier only checks that the generator produces well-formed, varied problems the oracle accepts."""
import collections

import numpy as np
import pytest

from karpenter_b200 import _native
from karpenter_b200.scheduler import Scheduler
from tests import fuzz, oracle_lib
from tests.parity import assert_same

import os

SEEDS = list(range(400))
CAP = int(os.environ.get("KP_FUZZ_SEEDS", "0"))  # > 0: only that many seeds per test (runs under compute-sanitizer)


def encode(seed):
    pools, per_pool, nodes, pl = fuzz.problem(seed)
    s = Scheduler(pools, per_pool, nodes, claim_order="go" if seed % 3 else "stable")
    return s.encode(pl)


def test_generator_is_wellformed_and_varied():
    stats = collections.Counter()
    for seed in SEEDS:
        enc = encode(seed)
        try:
            res = oracle_lib.solve(enc.problem)
        except RuntimeError:
            stats["rejected"] += 1
            continue
        t = res["pod_target"]
        stats["solved"] += 1
        stats["on_nodes"] += int((t >= 0).sum() > 0)
        stats["claims"] += int(res["n_claims"] > 0)
        stats["unsched"] += int((t == -1).sum() > 0)
        stats["multi_claim"] += int(res["n_claims"] > 3)
        stats["groups"] += int(res["n_groups"] > 0)
    assert stats["solved"] >= 300, stats
    for k in ("on_nodes", "claims", "unsched", "multi_claim", "groups"):
        assert stats[k] >= 10, stats


def encode_soft(seed):
    """the same problems with soft constraints sprinkled in (fuzz.soften); every fifth ignores preferences"""
    pools, per_pool, nodes, pl = fuzz.problem(seed, n_pods=[5, 20, 60, 150][seed % 4])
    fuzz.soften(seed, pools, pl)
    s = Scheduler(pools, per_pool, nodes, claim_order="go" if seed % 3 else "stable",
                  preference_policy="Ignore" if seed % 5 == 0 else "Respect",
                  min_values_policy="BestEffort" if seed % 7 == 0 else "Strict")
    return s.encode(pl)


def test_soft_generator_relaxes():
    stats = collections.Counter()
    for seed in range(150):
        enc = encode_soft(seed)
        nxt = enc.problem.get("class_relax_next")
        stats["chains"] += int((nxt >= 0).any())
        try:
            res = oracle_lib.solve(enc.problem)
        except RuntimeError:
            stats["rejected"] += 1
            continue
        stats["solved"] += 1
        stats["unsched"] += int((res["pod_target"] == -1).sum() > 0)
    assert stats["solved"] >= 100 and stats["chains"] >= 80, stats


@pytest.mark.gpu
@pytest.mark.parametrize("soft", [False, True], ids=["hard", "soft"])
def test_fuzz_parity_gpu(soft):
    h = _native.Handle()
    bad, ran = [], 0
    try:
        for seed in list(range(300) if soft else SEEDS)[:CAP or None]:
            enc = encode_soft(seed) if soft else encode(seed)
            try:
                orc = oracle_lib.solve(enc.problem)
            except RuntimeError:
                continue  # unsupported feature combination: both sides refuse (checked below)
            try:
                gpu = h.solve(enc.problem)
            except _native.SolverError as e:
                bad.append((seed, f"gpu refused: {e}"))
                continue
            ran += 1
            try:
                assert_same(gpu, orc, f"seed {seed} ")
            except AssertionError as e:
                bad.append((seed, str(e)[:200]))
    finally:
        h.close()
    assert not bad, bad[:10]
    assert ran >= (CAP * 3 // 4 if CAP else 300), ran


def encode_reserved(seed):
    """the soft problems with capacity reservations on some instance types (reserved offerings of capacity 1-3, some shared
    between types); strict mode on two seeds of three, fallback on the third"""
    import random
    from karpenter_b200.model import CAPACITY_TYPE_LABEL, RESERVATION_ID_LABEL, ZONE_LABEL, NodeSelectorRequirement, Offering
    pools, per_pool, nodes, pl = fuzz.problem(seed, n_pods=[5, 20, 60, 150][seed % 4])
    fuzz.soften(seed, pools, pl)
    rng = random.Random(31_000 + seed)
    its = per_pool[pools[0].name]
    ids = [f"r-{i}" for i in range(rng.randint(1, 5))]
    for it in its:
        if rng.random() < 0.6:
            for r in it.requirements:
                if r.key == CAPACITY_TYPE_LABEL and "reserved" not in r.values:
                    object.__setattr__(r, "values", tuple(r.values) + ("reserved",))
            for _ in range(rng.randint(1, 2)):
                it.offerings = list(it.offerings) + [Offering(
                    [NodeSelectorRequirement(CAPACITY_TYPE_LABEL, "In", ("reserved",)),
                     NodeSelectorRequirement(ZONE_LABEL, "In", (rng.choice(fuzz.ZONES),)),
                     NodeSelectorRequirement(RESERVATION_ID_LABEL, "In", (rng.choice(ids),))],
                    0.0001, rng.random() < 0.9, reservation_capacity=rng.randint(1, 3))]
    for p in pools:  # let the pools launch reserved capacity
        for i, r in enumerate(p.requirements):
            if r.key == CAPACITY_TYPE_LABEL and r.operator == "In" and "reserved" not in r.values:
                p.requirements[i] = NodeSelectorRequirement(CAPACITY_TYPE_LABEL, "In", tuple(r.values) + ("reserved",))

    class S(Scheduler):
        def _builder(self):
            b = super()._builder()
            b.reserved_offering_strict = seed % 3 != 2
            return b
    return S(pools, per_pool, nodes, claim_order="go" if seed % 3 else "stable").encode(pl)


def test_reserved_generator_reserves():
    stats = collections.Counter()
    for seed in range(120):
        enc = encode_reserved(seed)
        res = oracle_lib.solve(enc.problem)
        stats["held"] += int(res["claim_reservations"].any())
        stats["reserved_errors"] += int((res["pod_error"] == 3).any())
        stats["multi"] += int(res["n_claims"] > 2)
    assert stats["held"] >= 40 and stats["reserved_errors"] >= 10 and stats["multi"] >= 40, stats


@pytest.mark.gpu
def test_fuzz_reserved_capacity_parity_gpu():
    h = _native.Handle()
    bad = []
    try:
        for seed in range(CAP or 250):
            enc = encode_reserved(seed)
            orc = oracle_lib.solve(enc.problem)
            gpu = h.solve(enc.problem)
            try:
                assert_same(gpu, orc, f"seed {seed} ")
            except AssertionError as e:
                bad.append((seed, str(e)[:200]))
    finally:
        h.close()
    assert not bad, bad[:10]


def encode_ports(seed):
    """the soft problems with host ports on a third of the pods (a handful of ports / addresses / protocols, wildcard
    addresses included), ports already in use on some existing nodes and a daemon port on one NodePool"""
    import random
    pools, per_pool, nodes, pl = fuzz.problem(seed, n_pods=[5, 20, 60, 150][seed % 4])
    fuzz.soften(seed, pools, pl)
    rng = random.Random(47_000 + seed)
    entries = [(ip, port, proto) for ip in ("", "0.0.0.0", "10.0.0.1", "10.0.0.2") for port in (80, 443, 8080)
               for proto in ("TCP", "UDP")]
    rng.shuffle(entries)
    entries = entries[:rng.randint(2, 12)]
    for p in pl:
        if rng.random() < 0.35:
            p.host_ports = rng.sample(entries, rng.randint(1, min(3, len(entries))))
    for n in nodes:
        if rng.random() < 0.5:
            n.host_ports = rng.sample(entries, rng.randint(1, min(2, len(entries))))
    daemon = {pools[rng.randrange(len(pools))].name: [rng.choice(entries)]} if rng.random() < 0.4 else {}
    return Scheduler(pools, per_pool, nodes, claim_order="go" if seed % 3 else "stable", daemon_host_ports=daemon).encode(pl)


def test_ports_generator_conflicts():
    stats = collections.Counter()
    for seed in range(120):
        enc = encode_ports(seed)
        plain = encode_soft(seed)
        res, ref = oracle_lib.solve(enc.problem), oracle_lib.solve(plain.problem)
        stats["ports"] += int(enc.problem.n_hostports > 1)
        stats["changed"] += int(res["n_claims"] != ref["n_claims"] or not np.array_equal(res["pod_target"], ref["pod_target"]))
        stats["more_claims"] += int(res["n_claims"] > ref["n_claims"])
    assert stats["ports"] >= 100 and stats["changed"] >= 40 and stats["more_claims"] >= 20, stats


@pytest.mark.gpu
def test_fuzz_host_ports_parity_gpu():
    h = _native.Handle()
    bad = []
    try:
        for seed in range(CAP or 250):
            enc = encode_ports(seed)
            orc = oracle_lib.solve(enc.problem)
            gpu = h.solve(enc.problem)
            try:
                assert_same(gpu, orc, f"seed {seed} ")
            except AssertionError as e:
                bad.append((seed, str(e)[:200]))
    finally:
        h.close()
    assert not bad, bad[:10]


def encode_replicas(seed):
    """Few deployments with many replicas each (600 - 3 000 pods) on small instance types, so that the queue holds long runs of
    identical pods and the solve opens 50+ NodeClaims: the solver's cohort commits (kp_wsolve.cuh) against the pod-by-pod
    reference.  Every deployment gets its own CPU request (the queue then keeps its pods together)."""
    import random
    from karpenter_b200.model import quantity_units
    rng = random.Random(91_000 + seed)
    pools, per_pool, nodes, pl = fuzz.problem(seed, n_pods=[600, 1500, 3000][seed % 3], with_nodes=seed % 4 == 0)
    if seed % 5 == 0:
        fuzz.soften(seed, pools, pl)
    if seed % 4 == 1:  # topology-free seeds: the solver's lean build
        for p in pl:
            p.topology_spread_constraints, p.pod_affinity, p.pod_anti_affinity = [], [], []
            p.pod_affinity_preferred, p.pod_anti_affinity_preferred = [], []
    shapes = {}
    for p in pl:
        key = repr((p.requests, p.labels, p.namespace, p.node_selector, p.node_affinity_required, p.tolerations,
                    p.topology_spread_constraints, p.pod_affinity, p.pod_anti_affinity, p.node_affinity_preferred,
                    p.pod_affinity_preferred, p.pod_anti_affinity_preferred))
        i = shapes.setdefault(key, len(shapes))
        if seed % 7:  # one seed of seven keeps the interleaved queue of the plain generator
            p.requests = dict(p.requests, cpu=f"{quantity_units('cpu', p.requests['cpu']) + i}m")
        p.creation_timestamp = 0
    if seed % 2:  # small machines only: many NodeClaims
        for name, its in per_pool.items():
            small = sorted(its, key=lambda it: quantity_units("cpu", it.capacity["cpu"]))
            per_pool[name] = small[:max(2, len(small) // 3)]
    return Scheduler(pools, per_pool, nodes, claim_order="go" if seed % 3 else "stable").encode(pl)


def test_replica_generator_has_runs_and_claims():
    stats = collections.Counter()
    for seed in range(24):
        enc = encode_replicas(seed)
        res = oracle_lib.solve(enc.problem)
        stats["claims50"] += int(res["n_claims"] >= 50)
        stats["claims13_49"] += int(13 <= res["n_claims"] < 50)
        stats["groups"] += int(res["n_groups"] > 0)
        stats["lean"] += int(res["n_groups"] == 0)
        stats["placed"] += int((res["pod_target"] != -1).sum() > 300)
    assert stats["claims50"] >= 6 and stats["groups"] >= 6 and stats["lean"] >= 4 and stats["placed"] >= 8, stats


@pytest.mark.gpu
def test_fuzz_replica_cohorts_parity_gpu():
    h = _native.Handle()
    bad = []
    try:
        for seed in range(CAP or 120):
            enc = encode_replicas(seed)
            orc = oracle_lib.solve(enc.problem)
            gpu = h.solve(enc.problem)
            try:
                assert_same(gpu, orc, f"seed {seed} ")
            except AssertionError as e:
                bad.append((seed, str(e)[:200]))
    finally:
        h.close()
    assert not bad, bad[:10]


def encode_volumes(seed):
    """the soft problems with 2 - 3 volume-topology alternatives (zones, sometimes with a capacity type) on a third of the pods"""
    import random
    from karpenter_b200.model import CAPACITY_TYPE_LABEL, ZONE_LABEL, NodeSelectorRequirement
    pools, per_pool, nodes, pl = fuzz.problem(seed, n_pods=[5, 20, 60, 150][seed % 4])
    fuzz.soften(seed, pools, pl)
    rng = random.Random(53_000 + seed)
    shapes = {}
    for p in pl:
        key = id(p.requests), tuple(sorted(p.labels.items())), p.namespace
        if key not in shapes:
            alts = []
            if rng.random() < 0.35:
                for _ in range(rng.randint(2, 3)):
                    alt = [NodeSelectorRequirement(ZONE_LABEL, "In", tuple(rng.sample(fuzz.ZONES + ["no-such-zone"], rng.randint(1, 2))))]
                    if rng.random() < 0.3:
                        alt.append(NodeSelectorRequirement(CAPACITY_TYPE_LABEL, "In", (rng.choice(fuzz.CTS),)))
                    alts.append(alt)
            shapes[key] = alts
        if shapes[key]:
            p.volume_requirements = shapes[key]
    return Scheduler(pools, per_pool, nodes, claim_order="go" if seed % 3 else "stable").encode(pl)


def test_volume_generator_uses_later_alternatives():
    stats = collections.Counter()
    for seed in range(120):
        enc = encode_volumes(seed)
        nxt = enc.problem.get("class_vol_next")
        if nxt is None:
            continue
        stats["chains"] += 1
        res = oracle_lib.solve(enc.problem)
        one = enc.problem.get("class_vol_next").copy()
        enc.problem.set("class_vol_next", np.full_like(one, -1))   # only the first alternative
        first_only = oracle_lib.solve(enc.problem)
        enc.problem.set("class_vol_next", one)
        stats["later_alternative_matters"] += int(not np.array_equal(res["pod_target"], first_only["pod_target"]))
    assert stats["chains"] >= 60 and stats["later_alternative_matters"] >= 15, stats


@pytest.mark.gpu
def test_fuzz_volume_alternatives_parity_gpu():
    h = _native.Handle()
    bad = []
    try:
        for seed in range(CAP or 250):
            enc = encode_volumes(seed)
            orc = oracle_lib.solve(enc.problem)
            gpu = h.solve(enc.problem)
            try:
                assert_same(gpu, orc, f"seed {seed} ")
            except AssertionError as e:
                bad.append((seed, str(e)[:200]))
    finally:
        h.close()
    assert not bad, bad[:10]


def consolidation_case(seed):
    """A random small cluster (topology-free pods bound to nodes) and random candidate sets of 1-3 nodes."""
    import random
    from karpenter_b200.disruption import Consolidation
    rng = random.Random(10_000 + seed)
    its = fuzz.instance_types(rng)
    pools = fuzz.node_pools(rng)
    per_pool = {p.name: its for p in pools}
    nodes = fuzz.state_nodes(rng, its, pools, rng.randint(3, 14), [])
    for n in nodes:
        n.running_pods = []
        k = rng.randint(0, 4)
        cand = fuzz.pods(rng, 12)
        if seed % 4:  # three seeds of four stay topology-free (one warp per candidate set); the fourth takes the general path
            cand = [p for p in cand if not (p.topology_spread_constraints or p.pod_affinity or p.pod_anti_affinity)]
        n.pods = cand[:k]
    if seed % 3 == 0:  # soft constraints on the evicted pods: the simulation relaxes them like the provisioner does
        fuzz.soften(seed, pools, [p for n in nodes for p in n.pods])
    if seed % 5 == 1:  # host ports on some of the bound pods and on the nodes themselves
        ents = [(ip, port, "TCP") for ip in ("", "10.0.0.1", "10.0.0.2") for port in (80, 443)]
        for n in nodes:
            if rng.random() < 0.4:
                n.host_ports = [rng.choice(ents)]
            for p in n.pods:
                if rng.random() < 0.5:
                    p.host_ports = [rng.choice(ents)]
    names = [n.name for n in nodes]
    sets = [rng.sample(names, rng.randint(1, min(3, len(names)))) for _ in range(rng.randint(1, 12))]
    return pools, per_pool, nodes, sets, rng.random() < 0.5  # ... and whether spot-to-spot consolidation is enabled


def consolidation_extras(seed):
    """Pending pods and pods of deleting nodes that every simulation of the case also schedules (helpers.go:65-91):
    every other seed has some."""
    import random
    rng = random.Random(77_000 + seed)
    if seed % 2 == 0:
        return {}
    extra = fuzz.pods(rng, 8, uid0=50_000)
    if seed % 4 != 0 and seed % 4 != 3:  # keep the topology-free seeds topology-free
        pass
    if seed % 4:
        extra = [p for p in extra if not (p.topology_spread_constraints or p.pod_affinity or p.pod_anti_affinity)]
    k = rng.randint(0, len(extra))
    return dict(pending_pods=extra[:k], deleting_node_pods=extra[k:])


@pytest.mark.gpu
def test_fuzz_consolidation_parity_gpu():
    from karpenter_b200.disruption import Consolidation
    bad, ran = [], 0
    for seed in range(CAP or 200):
        pools, per_pool, nodes, sets, s2s = consolidation_case(seed)
        kw = dict(spot_to_spot=s2s, filter_same_instance_type=seed % 2 == 1, price_order=seed % 5 == 0,
                  **consolidation_extras(seed))
        orc = Consolidation(pools, per_pool, nodes, backend=oracle_lib.consolidate, **kw)
        try:
            orc.compute(sets)
        except RuntimeError:
            continue
        gpu = Consolidation(pools, per_pool, nodes, **kw)
        try:
            gpu.compute(sets)
        except _native.SolverError as e:
            if e.code == 5 and "minValues" in str(e):  # kp_consolidate refuses NodePools with minValues (so does the oracle)
                continue
            bad.append((seed, str(e)))
            continue
        finally:
            gpu.close()
        ran += 1
        from karpenter_b200 import _abi
        for k in _abi.CONSOL_PARITY_KEYS + (["repl_order_off", "repl_order"] if kw["price_order"] else []):
            if not np.array_equal(gpu.raw[k], orc.raw[k]):
                bad.append((seed, k, gpu.raw[k].tolist()[:8], orc.raw[k].tolist()[:8]))
                break
    assert ran >= (CAP // 2 if CAP else 120)
    assert not bad, bad[:6]


def test_consolidation_generator_on_oracle():
    from karpenter_b200.disruption import Consolidation
    decisions = collections.Counter()
    for seed in range(60):
        pools, per_pool, nodes, sets, s2s = consolidation_case(seed)
        try:
            for c in Consolidation(pools, per_pool, nodes, spot_to_spot=s2s, backend=oracle_lib.consolidate,
                                   **consolidation_extras(seed)).compute(sets):
                decisions[c.decision] += 1
        except RuntimeError:
            decisions["rejected"] += 1
    assert decisions["delete"] >= 5 and decisions["noop"] >= 5 and decisions["replace"] >= 5, decisions
