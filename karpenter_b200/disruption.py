"""Host-side mirror of the reference's consolidation decision over the C ABI.

    consolidation.computeConsolidation(ctx, candidates...) (Command, error)    pkg/controllers/disruption/consolidation.go:136-229
    SimulateScheduling(ctx, kubeClient, cluster, provisioner, candidates...)   pkg/controllers/disruption/helpers.go:51-142
    Command.Decision() / Candidate                                             pkg/controllers/disruption/types.go:74-182

`Consolidation(...)` takes what the disruption controller holds (NodePools, their instance types, the cluster's
StateNodes with the reschedulable pods bound to them) and `compute(candidate_sets)` evaluates every candidate set --
each one an independent computeConsolidation call -- in a single `kp_consolidate`.  `backend` lets the tests run the
identical encode / decode path against the CPU oracle.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from . import _abi
from .encode import ProblemBuilder
from .model import CAPACITY_TYPE_LABEL, InstanceType, NodePool, StateNode

DECISIONS = {0: "noop", 1: "delete", 2: "replace"}


@dataclass
class Command:
    """disruption.Command as its callers read it: the decision and, for a replacement, the NodeClaim to launch
    (Command.Replacements, consolidation.go:206-229): NodePool, requests, requirements after the capacity-type pins and
    the launchable instance types (in OrderByPrice order when the price order was asked for)."""
    decision: str
    replacement_instance_types: List[str]
    n_new_node_claims: int
    n_unscheduled: int
    replacement_nodepool: Optional[str] = None
    replacement_requirements: Optional[Dict[str, dict]] = None
    replacement_requests: Optional[Dict[str, int]] = None


class Consolidation:
    def __init__(self, node_pools: Sequence[NodePool], instance_types: Dict[str, List[InstanceType]],
                 state_nodes: Sequence[StateNode], spot_to_spot: bool = False, backend: Optional[Callable] = None,
                 device: int = -1, preference_policy: str = "Respect", filter_same_instance_type: bool = False,
                 pending_pods: Sequence = (), deleting_node_pods: Sequence = (), price_order: bool = False,
                 solve_backend: Optional[Callable] = None, min_values_policy: str = "Strict"):
        self.node_pools = list(node_pools)
        self.min_values_policy = min_values_policy  # options.MinValuesPolicy; kp_consolidate serves Strict only
        self.instance_types = instance_types
        # sortExistingNodes order (scheduler.go:738-751): initialized first, then by name
        self.state_nodes = sorted(state_nodes, key=lambda n: (not n.initialized, n.name))
        self.spot_to_spot = spot_to_spot
        self.preference_policy = preference_policy  # SimulateScheduling forwards the policy (helpers.go:97-101)
        # MultiNodeConsolidation applies filterOutSameInstanceType to every Replace of >= 2 nodes (multinodeconsolidation.go:154-163)
        self.filter_same_instance_type = filter_same_instance_type
        # SimulateScheduling also schedules the cluster's pending pods and the reschedulable pods of nodes that are
        # already being deleted (helpers.go:65-91)
        self.pending_pods = list(pending_pods)
        self.deleting_node_pods = list(deleting_node_pods)
        self.price_order = price_order
        self._backend = backend
        self._solve_backend = solve_backend  # tests: the oracle's solve for Consolidation.simulate
        self._device = device
        self._handle = None

    def _encode(self, candidate_sets):
        b = ProblemBuilder()
        b.preference_policy = self.preference_policy
        b.min_values_policy = self.min_values_policy
        index: Dict[int, int] = {}
        by_name: Dict[str, int] = {}
        for np_ in self.node_pools:
            ids = []
            for it in self.instance_types.get(np_.name, []):
                if id(it) not in index:
                    index[id(it)] = b.add_instance_type(it)
                    by_name.setdefault(it.name, index[id(it)])
                ids.append(index[id(it)])
            b.add_nodepool(np_, ids)
        pos = {}
        for n in self.state_nodes:
            pos[n.name] = b.add_node(n)
        off = [0]
        for n in self.state_nodes:  # the cluster's pod table: rows grouped by the node they run on
            for p in n.pods:
                b.add_pod(p)
            off.append(off[-1] + len(n.pods))
            for p in n.running_pods:
                b.add_running(p, pos[n.name])
        kinds = []
        for p in self.pending_pods:  # the last rows of the pod table: the extra pods of every simulation
            b.add_pod(p)
            kinds.append(_abi.KP_EXTRA_PENDING)
        for p in self.deleting_node_pods:
            b.add_pod(p)
            kinds.append(_abi.KP_EXTRA_DELETING_NODE)
        enc = b.build()
        assert all(enc.node_pos[pos[n.name]] == i for i, n in enumerate(self.state_nodes))
        sub_off = np.concatenate([[0], np.cumsum([len(s) for s in candidate_sets])]).astype(np.int32)
        sub_nodes = np.array([pos[name] for s in candidate_sets for name in s], np.int32)
        consol = _abi.ConsolInput(
            node_pod_off=np.asarray(off, np.int32),
            node_it=np.array([by_name.get(n.instance_type, -1) if n.instance_type else -1 for n in self.state_nodes], np.int32),
            node_is_spot=np.array([n.labels.get(CAPACITY_TYPE_LABEL) == "spot" for n in self.state_nodes], np.uint8),
            n_subsets=len(candidate_sets), subset_off=sub_off, subset_nodes=sub_nodes,
            spot_to_spot_enabled=int(self.spot_to_spot), capacity_type_key=enc.key_id(CAPACITY_TYPE_LABEL),
            ct_reserved=enc.value_id(CAPACITY_TYPE_LABEL, "reserved"), ct_spot=enc.value_id(CAPACITY_TYPE_LABEL, "spot"),
            ct_on_demand=enc.value_id(CAPACITY_TYPE_LABEL, "on-demand"),
            filter_same_instance_type=int(self.filter_same_instance_type),
            n_extra_pods=len(kinds), extra_pod_kind=np.asarray(kinds, np.uint8) if kinds else None,
            export_price_order=int(self.price_order))
        return enc, consol

    def compute(self, candidate_sets: Sequence[Sequence[str]]) -> List[Command]:
        """One computeConsolidation per candidate set (lists of node names)."""
        enc, consol = self._encode([list(s) for s in candidate_sets])
        if self._backend is not None:
            res = self._backend(enc.problem, consol)
        else:
            from . import _native
            if self._handle is None:
                self._handle = _native.Handle(self._device)
            res = self._handle.consolidate(enc.problem, consol)
        self.raw = res
        out = []
        for s in range(len(candidate_sets)):
            row = res["replacement_its"][s]
            its = [n for i, n in enumerate(enc.it_names) if int(row[i >> 6]) >> (i & 63) & 1]
            if res.get("repl_order_off") is not None:
                o = res["repl_order_off"]
                its = [enc.it_names[i] for i in res["repl_order"][o[s]:o[s + 1]]]
            cmd = Command(DECISIONS[int(res["decision"][s])], its, int(res["n_new_claims"][s]), int(res["n_unscheduled"][s]))
            if cmd.decision == "replace" and "repl_template" in res:
                view = {"claim_req_flags": res["repl_req_flags"], "claim_req_gte": res["repl_req_gte"],
                        "claim_req_lte": res["repl_req_lte"], "claim_req_mask": res["repl_req_mask"]}
                cmd.replacement_nodepool = enc.tmpl_names[int(res["repl_template"][s])]
                cmd.replacement_requirements = enc.decode_requirements(view, s)
                cmd.replacement_requests = {r: int(v) for r, v in zip(enc.resources, res["repl_requests"][s])}
            out.append(cmd)
        return out

    def simulate(self, candidate_set: Sequence[str]) -> dict:
        """SimulateScheduling for ONE candidate set without the consolidation decision (helpers.go:51-142): number of new
        NodeClaims, unscheduled non-pending pods, and the instance types of the first new claim BEFORE any price filter
        -- what validation.validateCommand compares a command against (validation.go:296-356).  A derived provisioning
        problem (the candidates' pods + extras pending, the other nodes existing) solved by kp_solve / the oracle."""
        names = set(candidate_set)
        b = ProblemBuilder()
        b.preference_policy = self.preference_policy
        b.min_values_policy = self.min_values_policy
        b.max_instance_types = 600  # SimulateScheduling truncates its results (helpers.go:120, scheduling.MaxInstanceTypes)
        index, it_names = {}, []
        for np_ in self.node_pools:
            ids = []
            for it in self.instance_types.get(np_.name, []):
                if id(it) not in index:
                    index[id(it)] = b.add_instance_type(it)
                    it_names.append(it.name)
                ids.append(index[id(it)])
            b.add_nodepool(np_, ids)
        kinds = []
        for n in self.state_nodes:
            if n.name in names:
                for p in n.pods:
                    b.add_pod(p)
                    kinds.append(0)
                continue
            at = b.add_node(n)
            for p in list(n.pods) + list(n.running_pods):
                b.add_running(p, at)
        for p in self.pending_pods:
            b.add_pod(p)
            kinds.append(_abi.KP_EXTRA_PENDING)
        for p in self.deleting_node_pods:
            b.add_pod(p)
            kinds.append(_abi.KP_EXTRA_DELETING_NODE)
        enc = b.build()
        if self._solve_backend is not None:
            res = self._solve_backend(enc.problem)
        else:
            from . import _native
            if self._handle is None:
                self._handle = _native.Handle(self._device)
            res = self._handle.solve(enc.problem)
        kinds = np.asarray(kinds, np.uint8)
        tgt = res["pod_target"]
        flags = enc.problem.get("node_flags")
        uns = int(((tgt == -1) & (kinds != _abi.KP_EXTRA_PENDING)).sum())
        on_node = tgt >= 0
        if on_node.any():
            uninit = (flags[tgt[on_node]] & _abi.KP_NODE_INITIALIZED) == 0
            uns += int((uninit & (kinds[on_node] == 0)).sum())
        its = enc.decode_its(res, 0) if int(res["n_claims"]) > 0 else []
        return {"n_new_claims": int(res["n_claims"]), "n_unscheduled": uns, "instance_types": its}

    def close(self):
        if self._handle is not None:
            self._handle.close()
            self._handle = None


# ---- the disruption front-end: which candidate sets to simulate ---------------------------
def eviction_cost(pod) -> float:
    """EvictionCost (pkg/utils/disruption/disruption.go:48-70): 1.0 + deletion-cost / 2^27 + priority / 2^25, clamped to
    [-10, 10]; an unparsable deletion-cost annotation is ignored."""
    cost = 1.0
    if pod.deletion_cost is not None:
        try:
            cost += float(pod.deletion_cost) / 2.0 ** 27
        except ValueError:
            pass
    if pod.priority is not None:
        cost += float(pod.priority) / 2.0 ** 25
    return min(10.0, max(-10.0, cost))


def rescheduling_cost(pods) -> float:
    """ReschedulingCost (disruption.go:72-78): sum of the eviction costs, in slice order (float addition is not
    associative: the order is the reference's)."""
    cost = 0.0
    for p in pods:
        cost += eviction_cost(p)
    return cost


def lifetime_remaining(node: StateNode) -> float:
    """LifetimeRemaining (disruption.go:36-46): fraction of the NodeClaim's ExpireAfter still ahead, 1.0 without expiry."""
    if node.expire_after_s is None:
        return 1.0
    total = float(node.expire_after_s)
    return min(1.0, max(0.0, (total - float(node.age_s)) / total))


def disruption_cost(node: StateNode) -> float:
    """Candidate.DisruptionCost (disruption/types.go:131-135): ReschedulingCost over ALL pods of the node x lifetime."""
    return rescheduling_cost(list(node.pods) + list(node.running_pods)) * lifetime_remaining(node)


def sort_candidates(nodes: Sequence[StateNode]) -> List[StateNode]:
    """consolidation.sortCandidates (consolidation.go:126-131): sort.Slice by DisruptionCost -- Go's unstable pdqsort, so
    cost ties fall exactly as they do in the reference (the library's host-side port, kp_go_sort_f64)."""
    from . import _native
    costs = np.array([disruption_cost(n) for n in nodes], np.float64)
    return [nodes[i] for i in _native.go_sort_order(costs)]


def interweave_by_nodepool(sorted_nodes: Sequence[StateNode], previously_unseen: Sequence[str] = ()) -> List[StateNode]:
    """SingleNodeConsolidation.shuffleCandidates (singlenodeconsolidation.go:148-176): round-robin over NodePools, the
    ones a previous pass timed out on first.  The reference iterates a Go map there; canon: NodePool names ascending."""
    groups: Dict[str, List[StateNode]] = {}
    for n in sorted_nodes:
        groups.setdefault(n.nodepool, []).append(n)
    order = [p for p in previously_unseen] + sorted(p for p in groups if p not in set(previously_unseen))
    out = []
    for i in range(max((len(g) for g in groups.values()), default=0)):
        for p in order:
            if i < len(groups.get(p, ())):
                out.append(groups[p][i])
    return out


MAX_PARALLEL = 100  # multinodeconsolidation.go:87


class MultiNodeConsolidation:
    """MultiNodeConsolidation.ComputeCommands (multinodeconsolidation.go:52-113) + firstNConsolidationOption (:118-171).

    The reference simulates ~log2(100) prefixes one after the other; here EVERY prefix [0..mid], mid = 1 .. max, goes to
    the device in one kp_consolidate call (filterOutSameInstanceType applied on the device, :154-163) and the binary
    search then reads the table -- same decisions, same command."""

    def __init__(self, engine: Consolidation):
        self.engine = engine
        self.engine.filter_same_instance_type = True
        self.last_prefix_table: List[Command] = []

    def compute_command(self, candidates: Sequence[StateNode], budgets: Dict[str, int]):
        """Returns (Command or None, candidate names of the command, constrained_by_budgets)."""
        budgets = dict(budgets)
        cands = sort_candidates(list(candidates))
        disruptable, constrained = [], False
        for c in cands:
            if budgets.get(c.nodepool, 0) == 0:
                constrained = True
                continue
            if not c.pods:  # empty nodes belong to the emptiness method
                continue
            disruptable.append(c)
            budgets[c.nodepool] -= 1
        max_parallel = min(max(len(disruptable), 0), MAX_PARALLEL)
        cmd, names = self.first_n_consolidation_option(disruptable, max_parallel)
        return cmd, names, constrained

    def first_n_consolidation_option(self, candidates: Sequence[StateNode], max_n: int):
        if len(candidates) < 2:
            return None, []
        lo_, hi = 1, max_n
        if len(candidates) <= max_n:
            hi = len(candidates) - 1
        names = [c.name for c in candidates]
        table = self.engine.compute([names[:mid + 1] for mid in range(1, hi + 1)]) if hi >= 1 else []
        self.last_prefix_table = table
        last, last_names = None, []
        while lo_ <= hi:
            mid = (lo_ + hi) // 2
            cmd = table[mid - 1]
            # a Replace whose options were all filtered out comes back as no-op (kp_consol_input.filter_same_instance_type)
            if cmd.decision in ("delete", "replace"):
                last, last_names = cmd, names[:mid + 1]
                lo_ = mid + 1
            else:
                hi = mid - 1
        return last, last_names


class SingleNodeConsolidation:
    """SingleNodeConsolidation.ComputeCommands (singlenodeconsolidation.go:56-131): candidates by disruption cost,
    interwoven by NodePool; the first one (budget permitting, non-empty) whose computeConsolidation is not a no-op wins.
    All candidates are simulated in one kp_consolidate call."""

    def __init__(self, engine: Consolidation):
        self.engine = engine
        self.engine.filter_same_instance_type = False
        self.previously_unseen: List[str] = []

    def compute_command(self, candidates: Sequence[StateNode], budgets: Dict[str, int]):
        cands = interweave_by_nodepool(sort_candidates(list(candidates)), self.previously_unseen)
        eligible, constrained = [], False
        for c in cands:
            if budgets.get(c.nodepool, 0) == 0:
                constrained = True
                continue
            if not c.pods:
                continue
            eligible.append(c)
        table = self.engine.compute([[c.name] for c in eligible]) if eligible else []
        for c, cmd in zip(eligible, table):
            if cmd.decision != "noop":
                return cmd, [c.name], constrained
        return None, [], constrained


def validate_command(engine: Consolidation, cmd: Command, candidate_names: Sequence[str]) -> bool:
    """validation.validateCommand (validation.go:296-356): re-simulate the candidates on the engine's CURRENT cluster
    state; the command stands iff every non-pending pod still schedules and the simulation wants no NodeClaim (and the
    command has no replacement) or exactly one whose instance types include every type the command would launch.  The
    re-simulation applies no price filter (kp_consolidate's n_new_claims / the claim's own instance types are read from a
    plain solve of the derived problem)."""
    if not candidate_names:
        return False
    saved = engine.filter_same_instance_type
    engine.filter_same_instance_type = False
    try:
        sim = engine.simulate(list(candidate_names))
    finally:
        engine.filter_same_instance_type = saved
    if sim["n_unscheduled"] > 0:
        return False
    if sim["n_new_claims"] == 0:
        return cmd.decision == "delete"
    if sim["n_new_claims"] > 1 or cmd.decision != "replace":
        return False
    return set(cmd.replacement_instance_types) <= set(sim["instance_types"])
