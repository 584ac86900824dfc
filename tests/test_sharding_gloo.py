"""N > 1 host logic on CPU: NodePool sharding (karpenter_b200/sharding.py) with world_size 2 over gloo.

The solver behind each shard is the CPU oracle here (no GPU in this tier); the sharding / merge / all-reduce code is the
same that bench.py runs over NCCL.  Bit-exactness claim under test: with the stable claim order the
union of the per-NodePool shard solves IS the un-sharded solve.
"""
import os
import socket

import numpy as np
import pytest

from karpenter_b200 import sharding, workloads
from tests import oracle_lib

N_PODS, N_POOLS, N_ITS, REPL = 3000, 2, 120, 40


def _problem(pools_subset=None):
    enc = workloads.config_c5(n_pods=N_PODS, n_pools=N_POOLS, n_its=N_ITS, app_replicas=REPL, pools_subset=pools_subset)
    enc.problem.set("claim_order_mode", 1)
    return enc


def _pool_of_pod():
    half = N_PODS // 2
    return np.concatenate([np.arange(half) % N_POOLS, np.arange(N_PODS - half) % N_POOLS])


def _canonical_claims(res, pod_ids, tmpl_map):
    """{frozenset(global pod ids)} -> (template, requests, its): independent of claim numbering."""
    out = {}
    t = res["pod_target"]
    for k in range(int(res["n_claims"])):
        members = frozenset(int(pod_ids[i]) for i in np.nonzero(t == -2 - k)[0])
        out[members] = (int(tmpl_map[int(res["claim_template"][k])]), tuple(res["claim_requests"][k].tolist()),
                        tuple(res["claim_its"][k].tolist()))
    return out


def test_union_of_shards_is_the_unsharded_solve():
    pool = _pool_of_pod()
    enc = _problem()
    names = list(enc.tmpl_names)  # OrderByWeight: weight desc, then name desc (utils/nodepool/nodepool.go:161-171)
    union = oracle_lib.solve(enc.problem)
    want = _canonical_claims(union, np.arange(N_PODS), list(range(N_POOLS)))
    got = {}
    unsched = set()
    for r in range(N_POOLS):
        ids = np.nonzero(pool == r)[0]
        res = oracle_lib.solve(_problem([r]).problem)
        assert len(res["pod_target"]) == len(ids)
        got.update(_canonical_claims(res, ids, [names.index(f"pool-{r}")]))
        unsched |= {int(ids[i]) for i in np.nonzero(res["pod_target"] == -1)[0]}
    assert got == want
    assert unsched == {int(i) for i in np.nonzero(union["pod_target"] == -1)[0]}
    # counters: the multiset of per-group domain-count vectors agrees
    def groups(res):
        off = res["group_domain_off"]
        return sorted(tuple(res["domain_counts"][off[g]:off[g + 1]].tolist()) for g in range(int(res["n_groups"])))
    per_shard = []
    for r in range(N_POOLS):
        per_shard += groups(oracle_lib.solve(_problem([r]).problem))
    assert sorted(per_shard) == groups(union)


def _layout_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # rank 0 holds three NodePool shards (a kp_upload_batch of three), rank 1 one: the layout the library is given
        slots = [12, 0, 7] if rank == 0 else [5]
        offs, total = sharding.instance_offsets(slots, rank, world, dist)
        q.put((rank, offs, total))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_counter_table_layout_gloo_world2():
    """Host logic of the library-side collective (kp_comm_set_counter_layout): slices by rank, then by instance."""
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_layout_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert got == [(0, [0, 12, 12], 24), (1, [19], 24)]
    assert sharding.instance_offsets([3, 4], 0, 1) == ([0, 3], 7)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = sharding.pools_of_rank(N_POOLS, rank, world)
        res = oracle_lib.solve(_problem(mine).problem)
        table = sharding.allreduce_domain_counts(res, rank, world, dist)
        q.put((rank, table.tolist(), res["domain_counts"].tolist(), int(res["n_claims"])))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_allreduce_of_domain_counters_gloo_world2():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, table0, local0, _), (r1, table1, local1, _) = got
    assert table0 == table1 == local0 + local1          # every rank holds the global table = concat of the shards
    assert sum(table0) > 0


def test_merge_shards_renumbers_claims():
    pool = _pool_of_pod()
    shards, idx = [], []
    for r in range(N_POOLS):
        shards.append(oracle_lib.solve(_problem([r]).problem))
        idx.append(np.nonzero(pool == r)[0])
    merged = sharding.merge_shards(shards, idx, N_PODS, [[r] for r in range(N_POOLS)])
    assert merged["n_claims"] == sum(int(s["n_claims"]) for s in shards)
    k = -2 - merged["pod_target"][merged["pod_target"] <= -2]
    assert np.array_equal(np.bincount(k, minlength=merged["n_claims"]), merged["claim_npods"])
    assert np.array_equal(merged["claim_template"] , np.concatenate([np.full(int(s["n_claims"]), r, np.int32)
                                                                    for r, s in enumerate(shards)]))


def test_consolidation_subsets_shard_and_gather():
    """Candidate sets split round-robin over 3 ranks, each solved on its own, gathered back: identical to one call."""
    from karpenter_b200 import _abi, sharding, workloads
    from tests import oracle_lib
    enc, consol = workloads.config_c4(n_nodes=300, n_pods=6000, n_candidates=12)
    full = oracle_lib.consolidate(enc.problem, _abi.ConsolInput(**consol))
    S = consol["n_subsets"]
    assert S > 100
    parts = [oracle_lib.consolidate(enc.problem, _abi.ConsolInput(**sharding.shard_subsets(consol, r, 3))) for r in range(3)]
    assert sum(len(p["decision"]) for p in parts) == S
    got = sharding.gather_decisions(parts, S)
    for k in ("decision", "n_new_claims", "n_unscheduled", "replacement_its"):
        assert np.array_equal(got[k], full[k]), k
    assert len(set(full["decision"].tolist())) >= 2  # the sample is not all one answer
