"""Interning of the object model (model.py) into the flat ``kp_problem`` of include/karpsolve.h.

This is the caller-side step the cgo shim performs inside ``Provisioner.NewScheduler`` in the reference
(pkg/controllers/provisioning/provisioner.go:238-307): strings -> integer ids, corev1 objects -> CSR tables.
It applies only representation changes that the reference's own constructors define:

* ``NewRequirementWithFlexibility`` operator canonicalisation (pkg/scheduling/requirement.go:48-102) and label
  normalisation (pkg/apis/v1/labels.go:117-123);
* ``NewPodRequirements`` / ``NewStrictPodRequirements`` (pkg/scheduling/requirements.go:90-110): node selector +
  first required node-affinity term (preferred terms are a "next" row);
* ``MakeTopologyNodeFilter`` requirement alternatives (topologynodefilter.go:38-64);
* ``OrderByWeight`` (pkg/utils/nodepool/nodepool.go:161-171) and ``sortExistingNodes`` (scheduler.go:738-751);
* key pruning: label keys that only instance types / node labels mention can never take part in a decision
  (``Intersects`` looks at shared keys only, requirements.go:237-274), so they are dropped from the universe.

Label values are interned in lexicographic order per key, which makes "ascending id" the canonical replacement for
Go's random map iteration order.
"""
from __future__ import annotations

from dataclasses import replace
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _abi
from .model import (CAPACITY_TYPE_LABEL, HOSTNAME_LABEL, NODEPOOL_LABEL, NORMALIZED_LABELS, RESERVATION_ID_LABEL, WELL_KNOWN_LABELS, InstanceType, LabelSelector,
                    NodePool, NodeSelectorRequirement, Pod, StateNode, Taint, Toleration, effective_requests, host_port_key,
                    host_ports_match, quantity_units)

EFFECTS = {"": 0, "NoSchedule": 1, "PreferNoSchedule": 2, "NoExecute": 3}
TOL_OPS = {"": 0, "Equal": 0, "Exists": 1, "Lt": 2, "Gt": 3}
SEL_OPS = {"In": 0, "NotIn": 1, "Exists": 2, "DoesNotExist": 3}
INT64_MAX = (1 << 63) - 1


def go_atoi(s: str) -> Optional[int]:
    """strconv.Atoi: optional sign + decimal digits, must fit in int64."""
    t = s[1:] if s[:1] in "+-" else s
    if not t or not t.isdigit() or not t.isascii():
        return None
    v = int(s)
    if v < -(1 << 63) or v > INT64_MAX:
        return None
    return v


def canonical_requirement(r: NodeSelectorRequirement):
    """NewRequirementWithFlexibility (requirement.go:48-102) -> (key, complement, values, gte, lte, min_values)."""
    key = NORMALIZED_LABELS.get(r.key, r.key)
    op = r.operator
    mv = r.min_values
    if op == "In":
        return (key, False, tuple(sorted(set(r.values))), None, None, mv)
    complement = op not in ("In", "DoesNotExist")
    values = tuple(sorted(set(r.values))) if op == "NotIn" else ()
    gte = lte = None
    if op == "Gt":
        v = int(r.values[0])
        if v == INT64_MAX:
            return (key, False, (), None, None, None)  # Gt MaxInt matches nothing: DoesNotExist
        gte = v + 1
    elif op == "Lt":
        lte = int(r.values[0]) - 1
    elif op == "Gte":
        gte = int(r.values[0])
    elif op == "Lte":
        lte = int(r.values[0])
    return (key, complement, values, gte, lte, mv)


def label_requirements(labels: Dict[str, str]):
    """NewLabelRequirements (requirements.go:64-70)."""
    return [canonical_requirement(NodeSelectorRequirement(k, "In", (v,))) for k, v in sorted(labels.items())]


def pod_requirements(pod: Pod, strict: bool = False):
    """NewPodRequirements / NewStrictPodRequirements (requirements.go:90-110): node selector, the heaviest preferred
    node-affinity term (not when strict), the first required term.  (The reference picks the heaviest term with an
    unstable sort; up to 12 terms that is the first of the heaviest, which is what this does.)"""
    reqs = label_requirements(pod.node_selector)
    if not strict and pod.node_affinity_preferred:
        heaviest = sorted(pod.node_affinity_preferred, key=lambda t: -t.weight)[0]
        reqs += [canonical_requirement(r) for r in heaviest.match_expressions]
    if pod.node_affinity_required:
        reqs += [canonical_requirement(r) for r in pod.node_affinity_required[0]]
    return reqs


PREFER_NO_SCHEDULE_TOLERATION = Toleration("", "Exists", "", "PreferNoSchedule")


def _tolerates_all_prefer_no_schedule(t: Toleration) -> bool:
    """corev1.Toleration.MatchToleration against {Operator: Exists, Effect: PreferNoSchedule} (preferences.go:133-146)."""
    return t.key == "" and t.operator == "Exists" and t.effect == "PreferNoSchedule" and t.value == ""


def relax(pod: Pod, tolerate_prefer_no_schedule: bool) -> Optional[Pod]:
    """Preferences.Relax (preferences.go:38-57): the pod with ONE soft constraint dropped, None when nothing is left.

    Order: the first required node-affinity term while several remain (they are alternatives) :74-88, the heaviest
    preferred pod affinity term :102-115, the heaviest preferred pod anti-affinity term :117-130, the heaviest preferred
    node-affinity term :59-72, the first ScheduleAnyway spread (the last constraint takes its place) :90-100, and -- only
    if some NodePool carries a PreferNoSchedule taint -- a toleration for all such taints :132-146."""
    if len(pod.node_affinity_required) > 1:
        return replace(pod, node_affinity_required=list(pod.node_affinity_required[1:]))
    for f in ("pod_affinity_preferred", "pod_anti_affinity_preferred", "node_affinity_preferred"):
        terms = getattr(pod, f)
        if terms:
            return replace(pod, **{f: sorted(terms, key=lambda t: -t.weight)[1:]})  # sort.SliceStable, drop the head
    for i, t in enumerate(pod.topology_spread_constraints):
        if t.when_unsatisfiable == "ScheduleAnyway":
            tscs = list(pod.topology_spread_constraints)
            tscs[i] = tscs[-1]
            return replace(pod, topology_spread_constraints=tscs[:-1])
    if tolerate_prefer_no_schedule and not any(_tolerates_all_prefer_no_schedule(t) for t in pod.tolerations):
        return replace(pod, tolerations=list(pod.tolerations) + [PREFER_NO_SCHEDULE_TOLERATION])
    return None


def pod_filter_requirements(pod: Pod):
    """MakeTopologyNodeFilter (topologynodefilter.go:38-64): one alternative per required node-affinity term."""
    base = label_requirements(pod.node_selector)
    if not pod.node_affinity_required:
        return [base]
    return [base + [canonical_requirement(r) for r in term] for term in pod.node_affinity_required]


class _Csr:
    def __init__(self):
        self.off = [0]
        self.items: List[int] = []

    def add(self, row: Iterable[int]) -> int:
        self.items.extend(row)
        self.off.append(len(self.items))
        return len(self.off) - 2


class _Dedup:
    def __init__(self):
        self.ids: Dict[object, int] = {}
        self.rows: List[object] = []

    def get(self, key) -> int:
        i = self.ids.get(key)
        if i is None:
            i = len(self.rows)
            self.ids[key] = i
            self.rows.append(key)
        return i


class ProblemBuilder:
    """Two-phase builder: collect symbolic rows, then `build()` interns strings and emits numpy arrays."""

    def __init__(self, resources: Sequence[str] = ("cpu", "memory", "pods", "ephemeral-storage")):
        self.resources = list(resources)
        self.reqsets = _Dedup()      # tuple of canonical requirement tuples
        self.taintsets = _Dedup()    # tuple of Taint
        self.tolsets = _Dedup()      # tuple of Toleration
        self.labelsets = _Dedup()    # tuple of (k, v)
        self.selectors = _Dedup()    # LabelSelector
        self.nssets = _Dedup()       # tuple of namespace strings
        self.namespaces = _Dedup()
        self.classes = _Dedup()      # class key -> id
        self.class_rows: List[dict] = []
        self.its: List[dict] = []
        self.it_index: Dict[str, int] = {}
        self.templates: List[dict] = []
        self.nodes: List[dict] = []
        self.pods: List[Tuple[int, int, int]] = []  # (class, creation, uid)
        self._template_class: Dict[object, int] = {}
        self.running: List[Tuple[int, int]] = []
        self.extra_keys = set()
        self.claim_order_mode = 0
        self.pod_arrays = None
        self.max_values_per_key = 64  # width of the per-key value mask; wider keys go through value compaction
        self.preference_policy = "Respect"  # or "Ignore": scheduler.IgnorePreferences (scheduler.go:81-101)
        self.min_values_policy = "Strict"   # or "BestEffort": scheduler.MinValuesPolicy (scheduler.go:110-114)
        self.max_instance_types = 0         # > 0: Results.TruncateInstanceTypes inside the solve (provisioner.go:380 uses 600)
        self.reserved_capacity = True       # FeatureGates.ReservedCapacity (options.go:169-177, default on)
        self.reserved_offering_strict = True  # DisableReservedCapacityFallback: what the provisioner runs (provisioner.go:347)

    # ---- resources ----
    def res_index(self, name: str) -> int:
        if name not in self.resources:
            self.resources.append(name)
        return self.resources.index(name)

    def res_vector(self, d: Dict[str, object]):
        pairs = [(self.res_index(k), quantity_units(k, v)) for k, v in d.items()]
        return pairs

    # ---- symbolic rows ----
    def reqset(self, reqs) -> int:
        return self.reqsets.get(tuple(reqs))

    def taintset(self, taints: Sequence[Taint]) -> int:
        return self.taintsets.get(tuple(taints))

    def tolset(self, tols: Sequence[Toleration]) -> int:
        return self.tolsets.get(tuple(tols))

    def add_instance_type(self, it: InstanceType) -> int:
        row = dict(
            name=it.name, reqset=self.reqset([canonical_requirement(r) for r in it.requirements]),
            capacity=self.res_vector(it.capacity), overhead=self.res_vector(it.overhead),
            offerings=[(self.reqset([canonical_requirement(r) for r in o.requirements]), float(o.price),
                        bool(o.available), self._reservation(o)) for o in it.offerings])
        self.it_index.setdefault(it.name, len(self.its))  # names may repeat (the AWS catalog lists linux + windows rows)
        self.its.append(row)
        return len(self.its) - 1

    def _reservation(self, o):
        """(interned reservation id, capacity) of a reserved offering, (-1, 0) otherwise.  Offering.ReservationID() is the
        value of the reservation-id requirement (types.go:432-434)."""
        ct = [r for r in o.requirements if r.key == CAPACITY_TYPE_LABEL and r.operator == "In" and tuple(r.values) == ("reserved",)]
        if not ct or not self.reserved_capacity:
            return (-1, 0)
        rid = [r.values[0] for r in o.requirements if r.key == RESERVATION_ID_LABEL and r.operator == "In" and len(r.values) == 1]
        name = rid[0] if rid else ""
        if not hasattr(self, "reservation_ids"):
            self.reservation_ids = {}
        return (self.reservation_ids.setdefault(name, len(self.reservation_ids)), int(o.reservation_capacity))

    def _ports(self, host_ports) -> int:
        """bit set of interned <ip, port, protocol> entries"""
        if not hasattr(self, "hostports"):
            self.hostports = {}
        m = 0
        for hp in host_ports:
            m |= 1 << self.hostports.setdefault(host_port_key(hp), len(self.hostports))
        return m

    def add_nodepool(self, np_: NodePool, instance_types: Sequence[str], daemon: Optional[Dict[str, object]] = None,
                     daemon_host_ports=()):
        """NewNodeClaimTemplate (nodeclaimtemplate.go:57-79): requirements + labels + nodepool/nodeclass labels.
        daemon_host_ports: the host ports of the daemonset pods that would run on a node of the pool
        (getDaemonHostPortUsage, scheduler.go:794-811)."""
        reqs = [canonical_requirement(r) for r in np_.requirements]
        labels = dict(np_.labels)
        labels[NODEPOOL_LABEL] = np_.name
        labels[np_.node_class_label_key()] = np_.node_class  # v1.NodeClassLabelKey(nodeClassRef group / kind)
        reqs += label_requirements(labels)
        self.templates.append(dict(
            name=np_.name, weight=np_.weight, reqset=self.reqset(reqs), taintset=self.taintset(np_.taints),
            its=[n if isinstance(n, int) else self.it_index[n] for n in instance_types],
            daemon=self.res_vector(daemon or {}), ports=self._ports(daemon_host_ports),
            limits=self.res_vector(np_.limits)))

    def pod_class(self, pod: Pod) -> int:
        respect = self.preference_policy != "Ignore"
        tscs = []
        for t in pod.topology_spread_constraints:
            soft = t.when_unsatisfiable != "DoNotSchedule"
            if soft and not respect:
                continue  # PreferencePolicyIgnore (topology.go:431)
            sel = t.label_selector
            if sel is not None and t.match_label_keys:  # topology.go:434-442
                extra = tuple((k, "In", (pod.labels[k],)) for k in t.match_label_keys if k in pod.labels)
                sel = LabelSelector(sel.match_labels, sel.match_expressions + extra)
            tscs.append((0, t.topology_key, sel, (pod.namespace,), t.max_skew,
                         -1 if t.min_domains is None else t.min_domains, t.node_taints_policy == "Honor",
                         t.node_affinity_policy != "Ignore", soft))
        # required and -- unless preferences are ignored -- preferred terms alike (topology.go:460-499)
        for kind, hard, pref in ((1, pod.pod_affinity, pod.pod_affinity_preferred),
                                 (2, pod.pod_anti_affinity, pod.pod_anti_affinity_preferred)):
            for t, soft in [(t, False) for t in hard] + [(w.term, True) for w in (pref if respect else [])]:
                ns = tuple(t.namespaces) if t.namespaces else (pod.namespace,)  # topology.go:503-526
                tscs.append((kind, t.topology_key, t.label_selector, ns, 0, -1, False, False, soft))
        strict_reqs = pod_requirements(pod, strict=True)
        reqs = pod_requirements(pod, strict=not respect)  # scheduler.go:471-491 updateCachedPodData
        vol_rest = ()
        if pod.volume_requirements:
            # CanAdd tries the volume alternatives in turn, each added to the NODE's requirements behind the pod's own and
            # kept out of the strict requirements the topology sees (nodeclaim.go:136-176, existingnode.go:98-140): this
            # class carries the first one (requirements = pod AND volume, strict requirements = pod), class_vol_next chains
            # to the class of the same pod with the remaining alternatives
            reqs = reqs + [canonical_requirement(r) for r in pod.volume_requirements[0]]
            vol_rest = tuple(tuple(canonical_requirement(r) for r in alt) for alt in pod.volume_requirements[1:])
        # everything a scheduling decision or a later relaxation step can depend on
        eff = effective_requests(pod)  # resources.Ceiling: init containers, sidecars, overhead, pod-level resources
        key = (tuple(sorted(eff.items())), tuple(reqs), tuple(strict_reqs),
               tuple(pod.tolerations), pod.namespace, tuple(sorted(pod.labels.items())), tuple(tscs),
               tuple(tuple(a) for a in pod_filter_requirements(pod)), tuple(sorted(host_port_key(h) for h in pod.host_ports)),
               tuple(pod.node_affinity_preferred) if respect else (),
               tuple((w.weight for w in pod.pod_affinity_preferred)) if respect else (),
               tuple((w.weight for w in pod.pod_anti_affinity_preferred)) if respect else (), vol_rest)
        n_before = len(self.classes.rows)
        cid = self.classes.get(key)
        if cid == n_before:
            requests = {k: (f"{v}m" if k == "cpu" else v) for k, v in eff.items()}
            vec = self.res_vector(requests)
            vec.append((self.res_index("pods"), 1))  # RequestsForPods adds pods: 1 (resources.go:37)
            row = dict(pod=pod, requests=vec, ports=self._ports(pod.host_ports), reqset=self.reqset(reqs), strict=self.reqset(strict_reqs),
                       tolset=self.tolset(pod.tolerations), namespace=self.namespaces.get(pod.namespace),
                       labelset=self.labelsets.get(tuple(sorted(pod.labels.items()))),
                       filters=[self.reqset(a) for a in pod_filter_requirements(pod)] if tscs else [],
                       tscs=[])
            for (kind, tkey, sel, ns, skew, mind, tp, ap, soft) in tscs:
                for n in ns:
                    self.namespaces.get(n)
                row["tscs"].append(dict(type=kind, key=NORMALIZED_LABELS.get(tkey, tkey),
                                        selector=-1 if sel is None else self.selectors.get(sel),
                                        nsset=self.nssets.get(tuple(ns)), max_skew=skew, min_domains=mind,
                                        taint_policy=int(tp), affinity_policy=int(ap), preferred=int(soft)))
                self.extra_keys.add(NORMALIZED_LABELS.get(tkey, tkey))
            row["vol_next"] = -1
            self.class_rows.append(row)
            if vol_rest:
                row["vol_next"] = self.pod_class(replace(pod, volume_requirements=list(pod.volume_requirements[1:])))
        return cid

    def add_pod(self, pod: Pod) -> int:
        t = pod.template
        if t is None:
            cid = self.pod_class(pod)
        else:  # pods of one template are one class: intern the spec once
            cid = self._template_class.get(t)
            if cid is None:
                cid = self._template_class[t] = self.pod_class(pod)
        self.pods.append((cid, pod.creation_timestamp, pod.uid))
        return len(self.pods) - 1

    def set_pod_arrays(self, pod_class, creation, uid_hi, uid_lo):
        """Bulk path for the big synthetic configs: numpy arrays of per-pod rows (classes registered via pod_class())."""
        self.pod_arrays = (np.asarray(pod_class, np.int32), np.asarray(creation, np.int64),
                           np.asarray(uid_hi, np.uint64), np.asarray(uid_lo, np.uint64))

    def add_running(self, pod: Pod, node: int):
        """A pod already bound to cluster node `node` (index returned by add_node): only counted by the topology."""
        self.running.append((self.pod_class(pod), node))

    def add_node(self, n: StateNode) -> int:
        template_index = {t["name"]: i for i, t in enumerate(self.templates)}
        labels = {k: v for k, v in n.labels.items() if k != HOSTNAME_LABEL}
        row = dict(name=n.name, hostname=n.labels.get(HOSTNAME_LABEL, n.name), reqset=self.reqset(label_requirements(labels)),
                   taintset=self.taintset(n.taints), available=self.res_vector(n.available),
                   capacity=self.res_vector(n.capacity),
                   flags=(1 if n.schedulable else 0) | (2 if n.initialized else 0) | (4 if n.managed else 0),
                   template=template_index.get(n.nodepool, -1) if n.nodepool else -1,
                   instance_type=n.instance_type, labels=labels,
                   ports=self._ports(list(n.host_ports) + [h for p in n.running_pods for h in p.host_ports]),
                   pod_ports=[self._ports(p.host_ports) for p in n.pods])
        self.nodes.append(row)
        return len(self.nodes) - 1

    # ---- build ----
    def relax_chains(self) -> List[int]:
        """class_relax_next: the class of each class's pods after one EFFECTIVE Preferences.Relax step.  Steps that leave
        the class unchanged (a preference the policy already ignores) are skipped: retrying an identical pod fails
        identically.  Relaxed classes are appended while walking, so the loop also covers them."""
        # NewScheduler: tolerate PreferNoSchedule during relaxation iff some NodePool adds such a taint (scheduler.go:132-142)
        tolerate = any(t.effect == "PreferNoSchedule" for tm in self.templates for t in self.taintsets.rows[tm["taintset"]])
        nxt: List[int] = []
        c = 0
        while c < len(self.class_rows):
            pod, n = self.class_rows[c]["pod"], -1
            while True:
                pod = relax(pod, tolerate)
                if pod is None:
                    break
                n = self.pod_class(pod)
                if n != c:
                    break
                n = -1
            nxt.append(n)
            c += 1
        return nxt

    def build(self) -> "EncodedProblem":
        relax_next = self.relax_chains()  # first: it may add classes
        R = len(self.resources)
        # active keys
        active = set(self.extra_keys) | {HOSTNAME_LABEL}
        off_reqsets = {o[0] for it in self.its for o in it["offerings"]}
        seeds = {t["reqset"] for t in self.templates} | off_reqsets
        for c in self.class_rows:
            seeds |= {c["reqset"], c["strict"], *c["filters"]}
        for s in seeds:
            for r in self.reqsets.rows[s]:
                active.add(r[0])
        keys = sorted(active)
        key_id = {k: i for i, k in enumerate(keys)}
        values: Dict[str, set] = {k: set() for k in keys}
        for rows in self.reqsets.rows:
            for r in rows:
                if r[0] in values:
                    values[r[0]].update(r[2])
        for n in self.nodes:
            values[HOSTNAME_LABEL].add(n["hostname"])
        # Value compaction.  A key carries one 64-bit value mask, but keys like node.kubernetes.io/instance-type have one
        # value per instance type.  Values that no pod / NodePool / offering / topology-filter requirement MENTIONS are
        # indistinguishable to the algorithm (they only ever occur as In{v} on an instance type or a node label and are
        # compared with sets of mentioned values or complements of them), so they collapse into one OTHER value.  Not
        # applicable when the key is a topology key (domain counts are per value) or carries Gt / Lt bounds.
        self.OTHER = "\uffff<any unmentioned value>"
        value_map: Dict[str, Dict[str, str]] = {}
        topo_keys = {t["key"] for c in self.class_rows for t in c["tscs"]}
        for k in keys:
            if k == HOSTNAME_LABEL or len(values[k]) <= self.max_values_per_key:
                continue
            bounded = any(r[0] == k and (r[3] is not None or r[4] is not None) for rows in self.reqsets.rows for r in rows)
            if k in topo_keys or bounded:
                continue  # stays too wide: the library answers KP_ERR_CAPACITY
            mentioned = set()
            for sd in seeds:
                for r in self.reqsets.rows[sd]:
                    if r[0] == k:
                        mentioned.update(r[2])
            if len(mentioned) >= 64:
                continue
            value_map[k] = {v: (v if v in mentioned else self.OTHER) for v in values[k]}
            values[k] = mentioned | {self.OTHER}
        value_id = {k: {v: i for i, v in enumerate(sorted(vs))} for k, vs in values.items()}
        for k, m in value_map.items():
            ids = value_id[k]
            value_id[k] = {v: ids[m[v]] for v in m}
            value_id[k][self.OTHER] = ids[self.OTHER]
        key_value_off = [0]
        value_int, value_is_int = [], []
        for k in keys:
            for v in sorted(values[k]):
                a = go_atoi(v)
                value_int.append(a if a is not None else 0)
                value_is_int.append(0 if a is None else 1)
            key_value_off.append(len(value_int))
        P = _abi.Problem()
        P.set("n_keys", len(keys))
        P.set("key_flags", [(1 if k in WELL_KNOWN_LABELS else 0) | (2 if k == HOSTNAME_LABEL else 0) for k in keys])
        P.set("key_value_off", key_value_off)
        P.set("value_int", value_int)
        P.set("value_is_int", value_is_int)
        # requirement sets (pruned to active keys)
        rs_off, rq_key, rq_flags, rq_gte, rq_lte, rq_min, rv_off, rvals = [0], [], [], [], [], [], [0], []
        for rows in self.reqsets.rows:
            for (k, comp, vals, gte, lte, mv) in rows:
                if k not in key_id:
                    continue
                rq_key.append(key_id[k])
                rq_flags.append((1 if comp else 0) | (2 if gte is not None else 0) | (4 if lte is not None else 0) |
                                (8 if mv is not None else 0))
                rq_gte.append(gte or 0)
                rq_lte.append(lte or 0)
                rq_min.append(mv or 0)
                rvals.extend(sorted({value_id[k][v] for v in vals}))  # a set: compaction may merge values
                rv_off.append(len(rvals))
            rs_off.append(len(rq_key))
        P.set("n_reqsets", len(self.reqsets.rows))
        P.set("reqset_off", rs_off)
        P.set("n_reqs", len(rq_key))
        for name, arr in (("req_key", rq_key), ("req_flags", rq_flags), ("req_gte", rq_gte), ("req_lte", rq_lte),
                          ("req_min_values", rq_min), ("req_val_off", rv_off), ("req_vals", rvals)):
            P.set(name, arr)
        # resources
        P.set("n_resources", R)
        flags = []
        for r in self.resources:
            flags.append((1 if r == "cpu" else 0) | (2 if r == "memory" else 0) | (4 if r.startswith("hugepages-") else 0) |
                         (8 if r == "nodes" else 0))
        P.set("res_flags", flags)

        def dense(pairs_list):
            m = np.zeros((len(pairs_list), R), np.int64)
            present = np.zeros(len(pairs_list), np.uint32)
            for i, pairs in enumerate(pairs_list):
                for r, v in pairs:
                    m[i, r] += v
                    present[i] |= np.uint32(1 << r)
            return m, present

        # taints / tolerations
        tt = _Dedup()
        tt.get("")
        taints = _Dedup()
        ts_csr = _Csr()
        for rows in self.taintsets.rows:
            ts_csr.add(taints.get((tt.get(t.key), tt.get(t.value), EFFECTS[t.effect])) for t in rows)
        tols = _Dedup()
        tl_csr = _Csr()
        for rows in self.tolsets.rows:
            tl_csr.add(tols.get((tt.get(t.key), TOL_OPS[t.operator], tt.get(t.value), EFFECTS[t.effect])) for t in rows)
        tti = [go_atoi(s) for s in tt.rows]
        P.set("n_tt_strings", len(tt.rows))
        P.set("tt_int", [a or 0 for a in tti])
        P.set("tt_is_int", [0 if a is None else 1 for a in tti])
        P.set("n_taints", len(taints.rows))
        P.set("taint_key", [t[0] for t in taints.rows])
        P.set("taint_value", [t[1] for t in taints.rows])
        P.set("taint_effect", [t[2] for t in taints.rows])
        P.set("n_taintsets", len(self.taintsets.rows))
        P.set("taintset_off", ts_csr.off)
        P.set("taintset_ids", ts_csr.items)
        P.set("n_tolerations", len(tols.rows))
        P.set("tol_key", [t[0] for t in tols.rows])
        P.set("tol_op", [t[1] for t in tols.rows])
        P.set("tol_value", [t[2] for t in tols.rows])
        P.set("tol_effect", [t[3] for t in tols.rows])
        P.set("n_tolsets", len(self.tolsets.rows))
        P.set("tolset_off", tl_csr.off)
        P.set("tolset_ids", tl_csr.items)
        # instance types
        cap, capp = dense([it["capacity"] for it in self.its])
        ovh, _ = dense([it["overhead"] for it in self.its])
        P.set("n_its", len(self.its))
        P.set("it_reqset", [it["reqset"] for it in self.its])
        P.set("it_capacity", cap)
        P.set("it_cap_present", capp)
        P.set("it_overhead", ovh)
        oo, orq, opr, oav, orsv, orid, ocap = [0], [], [], [], [], [], []
        for it in self.its:
            for (rs, price, av, (rid, rcap)) in it["offerings"]:
                orq.append(rs)
                opr.append(price)
                oav.append(1 if av else 0)
                # Offering.CapacityType() == reserved (types.go:385-387) while the ReservedCapacity gate is on
                orsv.append(1 if rid >= 0 else 0)
                orid.append(rid)
                ocap.append(rcap)
            oo.append(len(orq))
        P.set("off_reserved", orsv)
        n_rsv = len(getattr(self, "reservation_ids", {}))
        if n_rsv:
            P.set("off_reservation_id", orid)
            P.set("off_reservation_capacity", ocap)
        P.set("n_reservations", n_rsv)
        P.set("reserved_offering_strict", 1 if self.reserved_offering_strict else 0)
        self._reservation_fields_pending = n_rsv
        P.set("it_off_off", oo)
        P.set("off_reqset", orq)
        P.set("off_price", opr)
        P.set("off_available", oav)
        # templates in OrderByWeight order: weight desc, name desc
        order = sorted(range(len(self.templates)), key=lambda i: (-self.templates[i]["weight"],
                                                                     _neg_str(self.templates[i]["name"])))
        tmpls = [self.templates[i] for i in order]
        tmpl_index = {t["name"]: i for i, t in enumerate(tmpls)}
        P.set("n_templates", len(tmpls))
        P.set("tmpl_reqset", [t["reqset"] for t in tmpls])
        P.set("tmpl_taintset", [t["taintset"] for t in tmpls])
        tio, tits = [0], []
        for t in tmpls:
            tits.extend(t["its"])
            tio.append(len(tits))
        P.set("tmpl_it_off", tio)
        P.set("tmpl_its", tits)
        dm, _ = dense([t["daemon"] for t in tmpls])
        lm, lp = dense([t["limits"] for t in tmpls])
        P.set("tmpl_daemon", dm)
        if getattr(self, "hostports", None):  # host ports (hostportusage.go:35-108): entries interned to bits, who Matches whom
            ents = sorted(self.hostports.items(), key=lambda kv: kv[1])
            if len(ents) > 64:
                raise ValueError("more than 64 distinct host ports in one Solve")
            P.set("n_hostports", len(ents))
            P.set("hostport_conflicts", np.array([sum(1 << j for b, j in ents if host_ports_match(a, b)) for a, _ in ents], np.uint64))
            P.set("tmpl_hostports", np.array([t["ports"] for t in tmpls], np.uint64))
        P.set("tmpl_limits", lm)
        P.set("tmpl_limit_present", lp)
        # labels / selectors / namespaces
        lab = _Dedup()
        ls_off, lk, lv = [0], [], []
        for rows in self.labelsets.rows:
            for (k, v) in rows:
                lk.append(lab.get(("k", k)))
                lv.append(lab.get(("v", v)))
            ls_off.append(len(lk))
        P.set("n_labelsets", len(self.labelsets.rows))
        P.set("labelset_off", ls_off)
        P.set("label_key", lk)
        P.set("label_val", lv)
        so, sk, sop, svo, sv = [0], [], [], [0], []
        for sel in self.selectors.rows:
            exprs = [(k, "In", (v,)) for k, v in sel.match_labels] + list(sel.match_expressions)
            for (k, op, vs) in exprs:
                sk.append(lab.get(("k", k)))
                sop.append(SEL_OPS[op])
                sv.extend(lab.get(("v", v)) for v in vs)
                svo.append(len(sv))
            so.append(len(sk))
        P.set("n_selectors", len(self.selectors.rows))
        P.set("selector_off", so)
        P.set("selx_key", sk)
        P.set("selx_op", sop)
        P.set("selx_val_off", svo)
        P.set("selx_vals", sv)
        ns_csr = _Csr()
        for rows in self.nssets.rows:
            ns_csr.add(self.namespaces.get(n) for n in rows)
        P.set("n_nssets", len(self.nssets.rows))
        P.set("nsset_off", ns_csr.off)
        P.set("nsset_ids", ns_csr.items)
        # classes
        creq, _ = dense([c["requests"] for c in self.class_rows])
        P.set("n_classes", len(self.class_rows))
        P.set("class_requests", creq)
        P.set("class_reqset", [c["reqset"] for c in self.class_rows])
        P.set("class_strict_reqset", [c["strict"] for c in self.class_rows])
        P.set("class_tolset", [c["tolset"] for c in self.class_rows])
        P.set("class_namespace", [c["namespace"] for c in self.class_rows])
        P.set("class_labelset", [c["labelset"] for c in self.class_rows])
        fo, fr, to = [0], [], [0]
        cols = {k: [] for k in ("type", "key", "selector", "nsset", "max_skew", "min_domains", "taint_policy",
                                "affinity_policy", "preferred")}
        for c in self.class_rows:
            fr.extend(c["filters"])
            fo.append(len(fr))
            for t in c["tscs"]:
                for k in cols:
                    cols[k].append(key_id[t["key"]] if k == "key" else t[k])
            to.append(len(cols["type"]))
        P.set("class_filter_off", fo)
        P.set("class_filter_reqsets", fr)
        P.set("class_tsc_off", to)
        P.set("class_relax_next", relax_next)
        if getattr(self, "hostports", None):
            P.set("class_hostports", np.array([c["ports"] for c in self.class_rows], np.uint64))
        if any(c.get("vol_next", -1) >= 0 for c in self.class_rows):
            P.set("class_vol_next", np.array([c.get("vol_next", -1) for c in self.class_rows], np.int32))
        for k, arr in cols.items():
            P.set("tsc_" + k, arr)
        # pods
        if self.pod_arrays is not None:
            pc, cr, hi, lo = self.pod_arrays
        else:
            pc = np.array([p[0] for p in self.pods], np.int32)
            cr = np.array([p[1] for p in self.pods], np.int64)
            hi = np.array([(p[2] >> 64) & 0xFFFFFFFFFFFFFFFF for p in self.pods], np.uint64)
            lo = np.array([p[2] & 0xFFFFFFFFFFFFFFFF for p in self.pods], np.uint64)
        P.set("n_pods", len(pc))
        P.set("pod_class", pc)
        P.set("pod_creation", cr)
        P.set("pod_uid_hi", hi)
        P.set("pod_uid_lo", lo)
        # nodes: sortExistingNodes == initialized first, then name (stable)
        norder = sorted(range(len(self.nodes)), key=lambda i: (0 if self.nodes[i]["flags"] & 2 else 1,
                                                                 self.nodes[i]["name"]))
        nodes = [self.nodes[i] for i in norder]
        node_pos = {old: new for new, old in enumerate(norder)}
        av, avp = dense([n["available"] for n in nodes])
        cp, _ = dense([n["capacity"] for n in nodes])
        P.set("n_nodes", len(nodes))
        P.set("node_flags", [n["flags"] for n in nodes])
        P.set("node_reqset", [n["reqset"] for n in nodes])
        P.set("node_hostname", [value_id[HOSTNAME_LABEL][n["hostname"]] for n in nodes])
        P.set("node_taintset", [n["taintset"] for n in nodes])
        P.set("node_available", av)
        P.set("node_avail_present", avp)
        P.set("node_capacity", cp)
        if getattr(self, "hostports", None):
            # a node's reschedulable pods (`pods`) stay bound unless the node is a consolidation candidate: their ports count
            def all_ports(n):
                m = n["ports"]
                for b in n["pod_ports"]:
                    m |= b
                return m
            P.set("node_hostports", np.array([all_ports(n) for n in nodes], np.uint64))
        P.set("node_template", [tmpl_index.get(self.templates[n["template"]]["name"], -1) if n["template"] >= 0 else -1
                                for n in nodes])
        P.set("n_running", len(self.running))
        P.set("run_class", [r[0] for r in self.running])
        P.set("run_node", [node_pos[r[1]] for r in self.running])
        # minValues tables (cloudprovider/types.go:301-337): per key with minValues, the RAW values of every instance type
        mv_keys = sorted({r[0] for t in self.templates for r in self.reqsets.rows[t["reqset"]] if r[5] is not None},
                         key=lambda k: key_id[k])
        mv_off, mv_vals = [0], []
        for k in mv_keys:
            ids: Dict[str, int] = {}
            for it in self.its:
                for r in self.reqsets.rows[it["reqset"]]:
                    if r[0] == k:
                        mv_vals.extend(ids.setdefault(v, len(ids)) for v in r[2])
                mv_off.append(len(mv_vals))
        P.set("n_minvalue_keys", len(mv_keys))
        P.set("minvalue_key", [key_id[k] for k in mv_keys])
        P.set("minvalue_it_off", mv_off)
        P.set("minvalue_it_vals", mv_vals)
        self.minvalue_keys = mv_keys
        P.set("min_values_best_effort", 1 if self.min_values_policy == "BestEffort" else 0)
        P.set("max_instance_types", int(self.max_instance_types))
        P.set("claim_order_mode", self.claim_order_mode)
        enc = EncodedProblem(P, keys, {k: sorted(values[k]) for k in keys}, list(self.resources),
                             [it["name"] for it in self.its], [t["name"] for t in tmpls], [n["name"] for n in nodes],
                             node_pos, nodes)
        ids = getattr(self, "reservation_ids", {})
        enc.reservation_names = [n for n, _ in sorted(ids.items(), key=lambda kv: kv[1])]
        if ids:  # FinalizeScheduling's pins (nodeclaim.go:291-307) are applied by the solver itself
            P.set("reservation_capacity_type_key", enc.key_id(CAPACITY_TYPE_LABEL))
            P.set("reservation_reserved_value", enc.value_id(CAPACITY_TYPE_LABEL, "reserved"))
            P.set("reservation_id_key", enc.key_id(RESERVATION_ID_LABEL))
            P.set("reservation_value", [enc.value_id(RESERVATION_ID_LABEL, n) for n in enc.reservation_names])
        else:
            P.set("reservation_capacity_type_key", -1)
            P.set("reservation_id_key", -1)
        # what the decoder needs to report minValues per NodeClaim (nodeclaim.go:186-191)
        enc.min_values_policy = self.min_values_policy
        enc.tmpl_min_values = []
        for t in tmpls:
            need: Dict[str, int] = {}
            for r in self.reqsets.rows[t["reqset"]]:
                if r[5] is not None:
                    need[r[0]] = max(need.get(r[0], 0), r[5])  # Requirements.Add keeps the larger (requirement.go:180)
            enc.tmpl_min_values.append(need)
        enc.it_min_value_sets = {k: [{v for r in self.reqsets.rows[it["reqset"]] if r[0] == k for v in r[2]} for it in self.its]
                                 for k in mv_keys}
        return enc


def _neg_str(s: str):
    # descending string order as a sort key
    return tuple(-ord(c) for c in s) + (1,)


class EncodedProblem:
    """A kp_problem plus the string tables needed to decode results."""

    def __init__(self, problem, keys, values, resources, it_names, tmpl_names, node_names, node_pos, node_rows):
        self.problem = problem
        self.keys = keys
        self.values = values
        self.resources = resources
        self.it_names = it_names
        self.tmpl_names = tmpl_names
        self.node_names = node_names
        self.node_pos = node_pos
        self.node_rows = node_rows

    def key_id(self, key: str) -> int:
        return self.keys.index(key) if key in self.keys else -1

    def value_id(self, key: str, value: str) -> int:
        vs = self.values.get(key, [])
        return vs.index(value) if value in vs else -1

    def decode_requirements(self, res: dict, claim: int) -> Dict[str, dict]:
        out = {}
        woff = 0
        for k, key in enumerate(self.keys):
            nv = len(self.values[key])
            words = 0 if key == HOSTNAME_LABEL else (nv + 63) // 64
            f = int(res["claim_req_flags"][claim, k])
            if f & _abi.KP_SLOT_PRESENT:
                vals = [self.values[key][v] for v in range(nv)
                        if int(res["claim_req_mask"][claim, woff + (v >> 6)]) >> (v & 63) & 1]
                out[key] = dict(complement=bool(f & 1), values=vals,
                                gte=int(res["claim_req_gte"][claim, k]) if f & 2 else None,
                                lte=int(res["claim_req_lte"][claim, k]) if f & 4 else None)
            woff += words
        return out

    def decode_reservations(self, res: dict, claim: int) -> List[str]:
        """NodeClaim.reservedOfferings as reservation ids (kp_result.claim_reservations)."""
        m = int(res["claim_reservations"][claim]) if "claim_reservations" in res and len(res["claim_reservations"]) else 0
        return [n for i, n in enumerate(getattr(self, "reservation_names", [])) if m >> i & 1]

    def decode_min_values(self, res: dict, claim: int) -> Dict[str, dict]:
        """minValues of the NodeClaim's requirements.  Strict: the NodePool's.  BestEffort: lowered to the number of distinct
        values the final instance-type options offer whenever that is less (nodeclaim.go:186-191, scheduler.go:658-667)."""
        need = self.tmpl_min_values[int(res["claim_template"][claim])]
        if not need:
            return {}
        out = {}
        its = None
        for key, n in need.items():
            mv = n
            if self.min_values_policy == "BestEffort":
                if its is None:
                    w = res["claim_its"][claim]
                    its = [t for t in range(len(self.it_names)) if int(w[t >> 6]) >> (t & 63) & 1]
                have = len(set().union(*[self.it_min_value_sets[key][t] for t in its])) if its else 0
                mv = min(n, have)
            out[key] = dict(min_values=mv, relaxed=mv < n)
        return out

    def decode_its(self, res: dict, claim: int) -> List[str]:
        row = res["claim_its"][claim]
        return [n for i, n in enumerate(self.it_names) if int(row[i >> 6]) >> (i & 63) & 1]
