"""Host-side mirror of the reference's Scheduler API over the C ABI.

    scheduling.NewScheduler(ctx, kubeClient, nodePools, cluster, stateNodes, topology, instanceTypes, daemonSetPods, ...)
                                                    pkg/controllers/provisioning/scheduling/scheduler.go:116-129
    (*Scheduler).Solve(ctx, pods) (Results, error)                                          scheduler.go:381
    disruption.SimulateScheduling / consolidation.computeConsolidation     pkg/controllers/disruption/helpers.go:51,
                                                                            consolidation.go:136

`Scheduler(...)` takes the same inputs (NodePools, per-NodePool instance types, StateNodes, daemon overhead), `solve(pods)`
returns `Results{new_node_claims, existing_nodes, pod_errors}`.  The solve itself runs in libkarpsolve.so on the GPU; this
module only interns strings (encode.py) and decodes the result arrays.  `backend` lets the tests run the identical
encode/decode path against the CPU oracle.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from . import _abi
from .encode import EncodedProblem, ProblemBuilder
from .model import (CAPACITY_TYPE_LABEL, InstanceType, NodeClaimResult, NodePool, Pod, Results, StateNode)

POD_ERRORS = {
    1: "nodepool requirements filtered out all available instance types",                      # scheduler.go:511
    2: "incompatible with every nodepool (taints, requirements, topology, resources or offerings)",  # scheduler.go:683
    3: "one or more instance types with compatible reserved offerings are available, but could not be reserved",  # nodeclaim.go:277
    4: "pod didn't schedule because NodePool couldn't meet minValues requirements",  # scheduler.go:371 (after truncation)
}


class Scheduler:
    def __init__(self, node_pools: Sequence[NodePool], instance_types: Dict[str, List[InstanceType]],
                 state_nodes: Sequence[StateNode] = (), daemon_overhead: Optional[Dict[str, dict]] = None,
                 daemon_host_ports: Optional[Dict[str, list]] = None,
                 claim_order: str = "go", backend: Optional[Callable] = None, device: int = -1,
                 preference_policy: str = "Respect", min_values_policy: str = "Strict", max_instance_types: int = 0):
        self.node_pools = list(node_pools)
        self.max_instance_types = max_instance_types  # > 0: Results.TruncateInstanceTypes (provisioner.go:380: 600) in the solve
        self.instance_types = instance_types
        self.state_nodes = list(state_nodes)
        self.daemon_overhead = daemon_overhead or {}
        self.daemon_host_ports = daemon_host_ports or {}  # NodePool name -> host ports of its daemonset pods
        self.claim_order = claim_order
        self.preference_policy = preference_policy  # "Ignore" == scheduler.IgnorePreferences (scheduler.go:81-101)
        self.min_values_policy = min_values_policy  # "BestEffort" == MinValuesPolicyBestEffort (scheduler.go:110-114)
        self._backend = backend
        self._device = device
        self._handle = None

    # -- encoding -------------------------------------------------------------------------------------------------
    def _builder(self) -> ProblemBuilder:
        b = ProblemBuilder()
        b.claim_order_mode = 1 if self.claim_order == "stable" else 0
        b.preference_policy = self.preference_policy
        b.min_values_policy = self.min_values_policy
        b.max_instance_types = self.max_instance_types
        index: Dict[int, int] = {}
        for np_ in self.node_pools:
            ids = []
            for it in self.instance_types.get(np_.name, []):
                if id(it) not in index:
                    index[id(it)] = b.add_instance_type(it)
                ids.append(index[id(it)])
            b.add_nodepool(np_, ids, self.daemon_overhead.get(np_.name), self.daemon_host_ports.get(np_.name, ()))
        self._it_index = index
        return b

    def encode(self, pods: Sequence[Pod]) -> EncodedProblem:
        b = self._builder()
        self._node_index = {}
        for n in self.state_nodes:
            self._node_index[n.name] = b.add_node(n)
        for i, n in enumerate(self.state_nodes):
            for p in n.running_pods:
                b.add_running(p, self._node_index[n.name])
        for p in pods:
            b.add_pod(p)
        return b.build()

    def _run(self, enc: EncodedProblem) -> dict:
        if self._backend is not None:
            return self._backend(enc.problem)
        from . import _native
        if self._handle is None:
            self._handle = _native.Handle(self._device)
        return self._handle.solve(enc.problem)

    # -- Solve ----------------------------------------------------------------------------------------------------
    def solve(self, pods: Sequence[Pod]) -> Results:
        pods = list(pods)
        enc = self.encode(pods)
        res = self._run(enc)
        claims = []
        target = res["pod_target"]
        by_claim: Dict[int, List[Pod]] = {}
        existing: Dict[str, List[Pod]] = {}
        errors: Dict[int, str] = {}
        for i, p in enumerate(pods):
            t = int(target[i])
            if t == _abi.KP_TARGET_UNSCHEDULED:
                errors[id(p)] = POD_ERRORS.get(int(res["pod_error"][i]), "unschedulable")
            elif t >= 0:
                existing.setdefault(enc.node_names[t], []).append(p)
            elif res["claim_dropped"][-2 - t]:  # TruncateInstanceTypes dropped the NodeClaim (scheduler.go:368-373)
                errors[id(p)] = POD_ERRORS[4]
            else:
                by_claim.setdefault(-2 - t, []).append(p)
        for k in range(res["n_claims"]):
            if res["claim_dropped"][k]:
                continue
            reqs = enc.decode_requirements(res, k)  # FinalizeScheduling's reservation pins are already in
            for key, mv in enc.decode_min_values(res, k).items():
                reqs.setdefault(key, dict(complement=True, values=[], gte=None, lte=None)).update(mv)
            claims.append(NodeClaimResult(
                nodepool=enc.tmpl_names[int(res["claim_template"][k])], pods=by_claim.get(k, []),
                instance_type_options=enc.decode_its(res, k), requirements=reqs,
                requests={r: int(v) for r, v in zip(enc.resources, res["claim_requests"][k])},
                rank=int(res["claim_rank"][k])))
        claims.sort(key=lambda c: c.rank)  # the order of Results.NewNodeClaims
        return Results(new_node_claims=claims, existing_nodes=existing, pod_errors=errors, raw=res)

    def close(self):
        if self._handle is not None:
            self._handle.close()
            self._handle = None
