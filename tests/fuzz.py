"""Seeded random scheduling problems for differential testing (CUDA path vs oracle).

Covers what the fixed workloads C1-C5 do not: several NodePools with weights / limits / taints / labels, requirement
operators NotIn / Exists / DoesNotExist / Gt / Lt on pods and NodePools, custom (non well-known) label keys, pod
affinity and anti-affinity on zone and hostname, spread with minDomains / maxSkew 2 / taint + affinity policies,
namespaces, existing nodes (tainted, uninitialized, partially used) and pods already running on them.
"""
from __future__ import annotations

import random
from typing import List

from karpenter_b200 import fake
from karpenter_b200.model import (ARCH_LABEL, CAPACITY_TYPE_LABEL, HOSTNAME_LABEL, INSTANCE_TYPE_LABEL, NODEPOOL_LABEL,
                                  OS_LABEL, ZONE_LABEL, LabelSelector, NodePool, NodeSelectorRequirement, Offering, Pod,
                                  PodAffinityTerm, PreferredSchedulingTerm, StateNode, Taint, Toleration,
                                  TopologySpreadConstraint, WeightedPodAffinityTerm, quantity_units)

ZONES = ["test-zone-1", "test-zone-2", "test-zone-3"]
CTS = ["spot", "on-demand"]
TEAM = "example.com/team"  # a custom (not well-known) label


def _req(key, op, *values):
    return NodeSelectorRequirement(key, op, tuple(values))


def instance_types(rng: random.Random):
    out = []
    for i in range(rng.randint(3, 9)):
        cpu = rng.choice([1, 2, 4, 8, 16, 32])
        mem = cpu * rng.choice([1, 2, 4])
        offs = []
        for z in ZONES:
            for ct in CTS:
                if rng.random() < 0.75:
                    price = round(0.1 * cpu + 0.01 * mem + rng.random() * 0.05, 4) * (0.7 if ct == "spot" else 1.0)
                    offs.append(Offering([_req(CAPACITY_TYPE_LABEL, "In", ct), _req(ZONE_LABEL, "In", z)], price,
                                         rng.random() < 0.9))
        if not any(o.available for o in offs):
            offs.append(Offering([_req(CAPACITY_TYPE_LABEL, "In", "on-demand"), _req(ZONE_LABEL, "In", ZONES[0])], 1.0, True))
        out.append(fake.new_instance_type(f"it-{i}-{cpu}x", {"cpu": str(cpu), "memory": f"{mem}Gi",
                                                            "pods": str(rng.choice([3, 5, 10, 30]))},
                                          architecture=rng.choice(["amd64", "amd64", "arm64"]), offerings=offs))
    return out


def node_pools(rng: random.Random):
    pools = []
    for i in range(rng.randint(1, 3)):
        reqs = [_req(CAPACITY_TYPE_LABEL, rng.choice(["Exists", "In"]), *(CTS if rng.random() < 0.7 else ["on-demand"]))]
        if reqs[0].operator == "Exists":
            reqs[0] = _req(CAPACITY_TYPE_LABEL, "Exists")
        if rng.random() < 0.4:
            reqs.append(_req(ZONE_LABEL, rng.choice(["In", "NotIn"]), *rng.sample(ZONES, rng.randint(1, 2))))
        if rng.random() < 0.25:
            reqs.append(_req(fake.INTEGER_INSTANCE_LABEL, rng.choice(["Gt", "Lt"]), str(rng.choice([2, 4, 8]))))
        labels = {}
        if rng.random() < 0.5:
            labels[TEAM] = rng.choice(["a", "b"])
        taints = [Taint("dedicated", f"p{i}", "NoSchedule")] if rng.random() < 0.3 else []
        limits = {"cpu": str(rng.choice([8, 40, 200]))} if rng.random() < 0.3 else {}
        pools.append(NodePool(name=f"pool-{i}", weight=rng.choice([0, 0, 10, 50]), requirements=reqs, labels=labels,
                              taints=taints, limits=limits))
    return pools


def _selector(rng):
    app = rng.choice(["a", "b", "c"])
    if rng.random() < 0.8:
        return LabelSelector.of({"app": app})
    return LabelSelector.of(None, [("app", rng.choice(["In", "NotIn", "Exists"]), (app,) if rng.random() < 0.8 else ())])


def pods(rng: random.Random, n: int, uid0=1) -> List[Pod]:
    out = []
    shapes = []
    for _ in range(rng.randint(2, 8)):  # a handful of deployments: identical pods share a class
        kw = dict(requests={"cpu": rng.choice(["100m", "500m", "1", "2", "3500m"]),
                            "memory": rng.choice(["128Mi", "1Gi", "3Gi"])},
                  labels={"app": rng.choice(["a", "b", "c"])}, namespace=rng.choice(["default", "default", "other"]))
        sel = {}
        if rng.random() < 0.3:
            sel[ZONE_LABEL] = rng.choice(ZONES)
        if rng.random() < 0.2:
            sel[ARCH_LABEL] = rng.choice(["amd64", "arm64"])
        if rng.random() < 0.15:
            sel[TEAM] = rng.choice(["a", "b", "c"])
        kw["node_selector"] = sel
        if rng.random() < 0.35:
            term = []
            for _ in range(rng.randint(1, 2)):
                k = rng.choice([ZONE_LABEL, CAPACITY_TYPE_LABEL, TEAM, fake.INTEGER_INSTANCE_LABEL, fake.LABEL_INSTANCE_SIZE])
                if k == fake.INTEGER_INSTANCE_LABEL:
                    term.append(_req(k, rng.choice(["Gt", "Lt"]), str(rng.choice([1, 4, 8, 16]))))
                elif k == ZONE_LABEL:
                    term.append(_req(k, rng.choice(["In", "NotIn"]), *rng.sample(ZONES, rng.randint(1, 2))))
                elif k == CAPACITY_TYPE_LABEL:
                    term.append(_req(k, rng.choice(["In", "NotIn"]), rng.choice(CTS)))
                elif k == TEAM:
                    op = rng.choice(["In", "NotIn", "Exists", "DoesNotExist"])
                    term.append(_req(k, op, *([rng.choice(["a", "b"])] if op in ("In", "NotIn") else [])))
                else:
                    term.append(_req(k, rng.choice(["In", "NotIn"]), rng.choice(["small", "large"])))
            kw["node_affinity_required"] = [term]
        tols = []
        r = rng.random()
        if r < 0.3:
            tols.append(Toleration("dedicated", "Exists", "", ""))
        elif r < 0.45:
            tols.append(Toleration("dedicated", "Equal", rng.choice(["p0", "p1"]), "NoSchedule"))
        kw["tolerations"] = tols
        r = rng.random()
        if r < 0.3:
            kw["topology_spread_constraints"] = [TopologySpreadConstraint(
                rng.choice([1, 1, 2]), rng.choice([ZONE_LABEL, ZONE_LABEL, HOSTNAME_LABEL, CAPACITY_TYPE_LABEL]), _selector(rng),
                min_domains=rng.choice([None, None, 2, 3]), node_taints_policy=rng.choice([None, "Honor", "Ignore"]),
                node_affinity_policy=rng.choice([None, "Honor", "Ignore"]))]
        elif r < 0.45:
            kw["pod_anti_affinity"] = [PodAffinityTerm(_selector(rng), rng.choice([HOSTNAME_LABEL, HOSTNAME_LABEL, ZONE_LABEL]))]
        elif r < 0.55:
            kw["pod_affinity"] = [PodAffinityTerm(_selector(rng), rng.choice([HOSTNAME_LABEL, ZONE_LABEL]))]
        shapes.append(kw)
    for i in range(n):
        kw = rng.choice(shapes)
        out.append(Pod(name=f"p{uid0 + i}", uid=rng.getrandbits(100), creation_timestamp=rng.choice([0, 0, 5]), **kw))
    return out


def state_nodes(rng: random.Random, its, pools, n: int, running: List[Pod]):
    out = []
    for i in range(n):
        it = rng.choice(its)
        pool = rng.choice(pools)
        zone = rng.choice(ZONES)
        arch = [x for x in it.requirements if x.key == ARCH_LABEL][0].values[0]
        labels = {HOSTNAME_LABEL: f"node-{i:03d}", ZONE_LABEL: zone, CAPACITY_TYPE_LABEL: rng.choice(CTS), OS_LABEL: "linux",
                  ARCH_LABEL: arch, NODEPOOL_LABEL: pool.name, INSTANCE_TYPE_LABEL: it.name}
        labels.update(pool.labels)
        frac = rng.choice([0.0, 0.3, 0.8, 1.0])
        avail = {}
        for r in ("cpu", "memory", "pods"):
            a = quantity_units(r, it.capacity[r]) - quantity_units(r, it.overhead.get(r, 0))
            v = int(a * frac)
            avail[r] = f"{v}m" if r == "cpu" else v
        cap = dict(it.capacity)
        cap["nodes"] = 1
        here = [p for p in running if rng.random() < 1.0 / max(n, 1)]
        out.append(StateNode(name=f"node-{i:03d}", labels=labels, taints=list(pool.taints) if rng.random() < 0.7 else [],
                             available=avail, capacity=cap, nodepool=pool.name, instance_type=it.name,
                             initialized=rng.random() < 0.85, running_pods=here))
    return out


def problem(seed: int, n_pods=None, with_nodes=True):
    """(node_pools, instance_types per pool, state_nodes, pods) of one random scheduling problem."""
    rng = random.Random(seed)
    its = instance_types(rng)
    pools = node_pools(rng)
    per_pool = {p.name: (its if rng.random() < 0.7 else rng.sample(its, max(1, len(its) // 2))) for p in pools}
    n = n_pods if n_pods is not None else rng.choice([5, 20, 60, 150, 400])
    pl = pods(rng, n)
    nodes = []
    if with_nodes and rng.random() < 0.6:
        nodes = state_nodes(rng, its, pools, rng.randint(1, 12), pods(rng, rng.randint(0, 10), uid0=10_000))
    return pools, per_pool, nodes, pl


def soften(seed: int, pools, pl: List[Pod]):
    """Sprinkle soft constraints over a problem (own random stream, so `problem(seed)` itself is unchanged): preferred
    node affinity with mixed weights, ScheduleAnyway spreads, preferred pod (anti-)affinity, alternative required
    node-affinity terms, PreferNoSchedule taints on NodePools.  Pods of one deployment stay identical."""
    rng = random.Random(77_000 + seed)
    for pool in pools:
        if rng.random() < 0.25:
            pool.taints = list(pool.taints) + [Taint("soft", "x", "PreferNoSchedule")]
        if rng.random() < 0.25:  # minValues: keep a NodeClaim flexible across instance types / sizes / architectures
            k = rng.choice([INSTANCE_TYPE_LABEL, INSTANCE_TYPE_LABEL, fake.LABEL_INSTANCE_SIZE, ARCH_LABEL])
            pool.requirements = list(pool.requirements) + [
                NodeSelectorRequirement(k, "Exists", (), rng.choice([2, 3, 5]) if k == INSTANCE_TYPE_LABEL else 2)]
    shapes = {}
    for p in pl:
        shapes.setdefault(id(p.tolerations), []).append(p)  # pods(): the kwargs of one deployment share their lists
    for group in shapes.values():
        kw = {}
        if rng.random() < 0.4:
            terms = []
            for _ in range(rng.randint(1, 3)):
                k = rng.choice([ZONE_LABEL, ZONE_LABEL, CAPACITY_TYPE_LABEL, fake.LABEL_INSTANCE_SIZE])
                v = {ZONE_LABEL: ZONES + ["nowhere"], CAPACITY_TYPE_LABEL: CTS, fake.LABEL_INSTANCE_SIZE: ["small", "large", "huge"]}[k]
                terms.append(PreferredSchedulingTerm(rng.choice([1, 1, 10, 50, 100]),
                                                     (_req(k, rng.choice(["In", "In", "NotIn"]), rng.choice(v)),)))
            kw["node_affinity_preferred"] = terms
        tscs = list(group[0].topology_spread_constraints)
        r = rng.random()
        if tscs and r < 0.4:
            t = tscs[0]
            tscs[0] = TopologySpreadConstraint(t.max_skew, t.topology_key, t.label_selector, "ScheduleAnyway", t.min_domains,
                                               t.node_taints_policy, t.node_affinity_policy)
            kw["topology_spread_constraints"] = tscs
        elif r < 0.55:
            tscs.append(TopologySpreadConstraint(1, rng.choice([ZONE_LABEL, HOSTNAME_LABEL, CAPACITY_TYPE_LABEL]), _selector(rng),
                                                 "ScheduleAnyway"))
            kw["topology_spread_constraints"] = tscs
        if rng.random() < 0.25:
            kw["pod_anti_affinity_preferred"] = [
                WeightedPodAffinityTerm(rng.choice([1, 50]), PodAffinityTerm(_selector(rng), rng.choice([HOSTNAME_LABEL, ZONE_LABEL])))
                for _ in range(rng.randint(1, 2))]
        if rng.random() < 0.2:
            kw["pod_affinity_preferred"] = [
                WeightedPodAffinityTerm(rng.choice([1, 50]), PodAffinityTerm(_selector(rng), rng.choice([HOSTNAME_LABEL, ZONE_LABEL])))]
        # alternatives: the first term is tried first, dropped when the pod does not fit.  (Under a spread the relaxation
        # changes the spread's node filter, so the reference creates a fresh topology group mid-solve.)
        if rng.random() < 0.2:
            first = list(group[0].node_affinity_required[0]) if group[0].node_affinity_required else []
            alt = [_req(ZONE_LABEL, "In", rng.choice(ZONES + ["nowhere"]))]
            kw["node_affinity_required"] = [[_req(ZONE_LABEL, "In", rng.choice(["nowhere", ZONES[0]]))] + first, alt + first]
        if rng.random() < 0.15:  # a volume bound in one zone, or a storage class spanning two
            kw["volume_requirements"] = [[_req(ZONE_LABEL, "In", *rng.sample(ZONES, rng.choice([1, 1, 2])))]]
        for p in group:
            for k, v in kw.items():
                setattr(p, k, v)
    return pools, pl
