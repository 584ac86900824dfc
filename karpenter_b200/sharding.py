"""NodePool sharding of a provisioning job across ranks.

The reference's Solve shards exactly when (i) every pod is compatible with exactly one NodePool and (ii) no topology
group's selected pods span two pools (scheduler.go:565,602; topology.go:53,58).  Then rank r owns the pods, the template
and the NodeClaims of its pools and runs an ordinary, independent Scheduler.Solve; the only exchange is one all-reduce
(sum, int32) of the topology-domain counter table so that every rank ends with the global counters the next
provisioning round starts from.  No collective sits on the data path.

Sharding is bit-exact against the un-sharded solve under the *stable* claim order (`claim_order_mode = 1`): a stable
sort of the whole NodeClaim list restricted to one pool's claims is the stable sort of that pool's claims alone.  Go's
unstable `sort.Slice` (mode 0) permutes ties depending on the entire slice, so there a shard is defined as "the
reference run on that pool's pods" -- each shard is still bit-identical to the reference solver on its own input.

This module is plain host logic (numpy + torch.distributed); it is exercised on CPU with the gloo backend
(tests/test_sharding_gloo.py) and on GPUs with NCCL by bench.py.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np


def pools_of_rank(n_pools: int, rank: int, world: int) -> List[int]:
    """Round-robin NodePool -> rank map (pool i lives on rank i mod world)."""
    return [i for i in range(n_pools) if i % world == rank]


def counter_layout(group_sizes_per_rank: Sequence[Sequence[int]]):
    """Global layout of the [G x D] domain-counter table: the per-rank tables concatenated in rank order.
    Returns (offset of every rank's slice, total length)."""
    offs, total = [], 0
    for sizes in group_sizes_per_rank:
        offs.append(total)
        total += int(sum(sizes))
    return offs, total


def instance_offsets(slots_mine: Sequence[int], rank: int, world: int, dist=None, device=None):
    """Layout of the library-resident global counter table (include/karpsolve.h kp_comm_set_counter_layout): every rank
    contributes the slot counts of ITS instances (kp_comm_counter_slots; a rank may hold several NodePool shards as one
    kp_upload_batch); the table is the concatenation rank by rank, instance by instance.  One tiny all-gather of the
    sizes.  Returns (offset of each of my instances, total slots)."""
    sizes = [list(map(int, slots_mine))]
    if world > 1 and dist is not None:
        import torch
        n = torch.tensor([len(slots_mine)], dtype=torch.int64, device=device)
        counts = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(counts, n)
        width = max(int(c.item()) for c in counts)
        mine = torch.zeros(max(width, 1), dtype=torch.int64, device=device)
        if slots_mine:
            mine[:len(slots_mine)] = torch.tensor(list(map(int, slots_mine)), dtype=torch.int64, device=device)
        rows = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(rows, mine)
        sizes = [[int(v) for v in rows[r][:int(counts[r].item())].cpu()] for r in range(world)]
    offs, total = counter_layout(sizes)
    base = offs[rank if world > 1 and dist is not None else 0]
    mine_off = [base + int(sum(sizes[rank if len(sizes) > 1 else 0][:i])) for i in range(len(slots_mine))]
    return mine_off, total


def allreduce_domain_counts(result: dict, rank: int, world: int, dist=None, device=None) -> np.ndarray:
    """The single collective of the sharded job: every rank contributes its shard's domain counters (zeros elsewhere)
    and receives the global table.  `result` is the dict of one shard's kp_solve (karpenter_b200/_abi.py)."""
    import torch
    local = np.ascontiguousarray(result["domain_counts"], dtype=np.int32)
    if world == 1 or dist is None:
        return local.copy()
    # slice sizes differ by rank: exchange them first (tiny all-gather), then one sum all-reduce of the table
    n = torch.tensor([local.size], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    offs, total = counter_layout([[s] for s in sizes])
    table = torch.zeros(max(total, 1), dtype=torch.int32, device=device)
    if local.size:
        table[offs[rank]:offs[rank] + local.size] = torch.from_numpy(local).to(table.device)
    dist.all_reduce(table, op=dist.ReduceOp.SUM)
    return table[:total].cpu().numpy()


def merge_shards(shards: Sequence[dict], pod_index: Sequence[np.ndarray], n_pods: int,
                 template_index: Sequence[Sequence[int]]) -> Dict[str, np.ndarray]:
    """Host-side concat of per-shard results into the global NodeClaim set (what a gather to rank 0 produces).
    pod_index[r][i] = global row of shard r's pod i; template_index[r][n] = global NodePool index of shard r's template n.
    NodeClaims are renumbered shard by shard (claim k of shard r -> base_r + k)."""
    target = np.full(n_pods, -1, np.int32)
    error = np.zeros(n_pods, np.uint8)
    tmpl, npods, reqs, its = [], [], [], []
    base = 0
    for r, res in enumerate(shards):
        t = res["pod_target"].copy()
        is_claim = t <= -2
        t[is_claim] -= base
        target[pod_index[r]] = t
        error[pod_index[r]] = res["pod_error"]
        tmpl.append(np.asarray(template_index[r], np.int32)[res["claim_template"]])
        npods.append(res["claim_npods"])
        reqs.append(res["claim_requests"])
        its.append(res["claim_its"])
        base += int(res["n_claims"])
    return {"pod_target": target, "pod_error": error, "n_claims": base,
            "claim_template": np.concatenate(tmpl) if tmpl else np.zeros(0, np.int32),
            "claim_npods": np.concatenate(npods) if npods else np.zeros(0, np.int32),
            "claim_requests": np.concatenate(reqs) if reqs else np.zeros((0, 0), np.int64),
            "claim_its": np.concatenate(its) if its else np.zeros((0, 0), np.uint64)}


# ---- consolidation: candidate sets shard trivially ----------------------------------------------
def shard_subsets(consol: dict, rank: int, world: int) -> dict:
    """Rank r's share of a kp_consol_input: the candidate sets s with s % world == r (each one is an independent
    computeConsolidation over read-only cluster state, helpers.go:55-59).  Tables are replicated, nothing is exchanged on
    the data path; `gather_decisions` puts the per-rank answers back into candidate-set order."""
    S = int(consol["n_subsets"])
    off, nodes = np.asarray(consol["subset_off"]), np.asarray(consol["subset_nodes"])
    mine = np.arange(rank, S, world)
    sizes = (off[1:] - off[:-1])[mine]
    sub_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    sub_nodes = (np.concatenate([nodes[off[i]:off[i + 1]] for i in mine]).astype(np.int32) if len(mine) else nodes[:0].astype(np.int32))
    return dict(consol, n_subsets=len(mine), subset_off=sub_off, subset_nodes=sub_nodes)


def gather_decisions(per_rank: Sequence[dict], n_subsets: int) -> Dict[str, np.ndarray]:
    """Interleave the shard results (rank r holds sets r, r + world, ...) back into candidate-set order."""
    world = len(per_rank)
    out: Dict[str, np.ndarray] = {}
    for k in ("decision", "n_new_claims", "n_unscheduled", "replacement_its"):
        first = np.asarray(per_rank[0][k])
        full = np.zeros((n_subsets,) + first.shape[1:], first.dtype)
        for r, res in enumerate(per_rank):
            full[r::world] = np.asarray(res[k])[:len(range(r, n_subsets, world))]
        out[k] = full
    return out
