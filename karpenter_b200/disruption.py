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
    """disruption.Command as its callers read it: the decision and, for a replacement, the launchable instance types."""
    decision: str
    replacement_instance_types: List[str]
    n_new_node_claims: int
    n_unscheduled: int


class Consolidation:
    def __init__(self, node_pools: Sequence[NodePool], instance_types: Dict[str, List[InstanceType]],
                 state_nodes: Sequence[StateNode], spot_to_spot: bool = False, backend: Optional[Callable] = None,
                 device: int = -1, preference_policy: str = "Respect", filter_same_instance_type: bool = False):
        self.node_pools = list(node_pools)
        self.instance_types = instance_types
        # sortExistingNodes order (scheduler.go:738-751): initialized first, then by name
        self.state_nodes = sorted(state_nodes, key=lambda n: (not n.initialized, n.name))
        self.spot_to_spot = spot_to_spot
        self.preference_policy = preference_policy  # SimulateScheduling forwards the policy (helpers.go:97-101)
        # MultiNodeConsolidation applies filterOutSameInstanceType to every Replace of >= 2 nodes (multinodeconsolidation.go:154-163)
        self.filter_same_instance_type = filter_same_instance_type
        self._backend = backend
        self._device = device
        self._handle = None

    def _encode(self, candidate_sets):
        b = ProblemBuilder()
        b.preference_policy = self.preference_policy
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
            filter_same_instance_type=int(self.filter_same_instance_type))
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
            out.append(Command(DECISIONS[int(res["decision"][s])], its, int(res["n_new_claims"][s]),
                               int(res["n_unscheduled"][s])))
        return out

    def close(self):
        if self._handle is not None:
            self._handle.close()
            self._handle = None
