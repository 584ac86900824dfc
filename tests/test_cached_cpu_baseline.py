"""oracle/orc_cached.cpp -- the CUDA solver's algorithm (failure bits, accepted-signature fast path, threshold bitmaps, scan
bounds, incremental Go sort) as scalar C++ on one host core, used by bench.py's cpu_baseline legs -- against the oracle:
same targets, errors, NodeClaims, order, requests and instance-type lists on the topology-free shapes it serves."""
import numpy as np
import pytest

from karpenter_b200 import _abi, workloads
from tests import oracle_lib

KEYS = oracle_lib.CACHED_KEYS
cached_solve = oracle_lib.cached_solve


def same(a, b, what):
    for k in KEYS:
        if isinstance(a[k], np.ndarray):
            assert a[k].shape == b[k].shape and np.array_equal(a[k], b[k]), f"{what}{k}"
        else:
            assert a[k] == b[k], f"{what}{k}"


@pytest.mark.parametrize("n_pods,order", [(13, 0), (300, 0), (3000, 0), (3000, 1), (30_000, 0)])
def test_c2_shapes_match_the_oracle(n_pods, order):
    enc = workloads.config_c2(n_pods=n_pods, n_its=500)
    enc.problem.set("claim_order_mode", order)
    got = cached_solve(enc.problem)
    assert got is not None
    same(got[0], oracle_lib.solve(enc.problem, threads=8), f"C2[{n_pods}] ")


@pytest.mark.parametrize("n_pods", [60, 1000])
def test_c1_shapes_match_the_oracle(n_pods):
    enc = workloads.config_c1(n_pods=n_pods)
    got = cached_solve(enc.problem)
    assert got is not None
    same(got[0], oracle_lib.solve(enc.problem), f"C1[{n_pods}] ")


def test_topology_free_fuzz_problems_match_the_oracle():
    """the fuzz generator's problems that happen to be in scope (no topology, bounds, nodes, limits, preferences)"""
    from tests.test_fuzz_parity import encode
    ran = 0
    for seed in range(400):
        enc = encode(seed)
        got = cached_solve(enc.problem)
        if got is None:
            continue
        try:
            ref = oracle_lib.solve(enc.problem)
        except RuntimeError:
            continue
        same(got[0], ref, f"seed {seed} ")
        ran += 1
    assert ran >= 3, ran


def test_stripped_fuzz_problems_match_the_oracle():
    """the fuzz generator's problems made topology-free: several NodePools with weights and taints, selectors, In / NotIn /
    Exists / DoesNotExist node affinity, tolerations, up to 400 pods -- and both claim-order modes"""
    from karpenter_b200.scheduler import Scheduler
    from tests import fuzz
    ran = 0
    for seed in range(300):
        pools, per_pool, _, pl = fuzz.problem(seed, with_nodes=False)
        for p in pl:
            p.topology_spread_constraints, p.pod_affinity, p.pod_anti_affinity = [], [], []
        for np_ in pools:
            np_.limits = {}
        enc = Scheduler(pools, per_pool, [], claim_order="go" if seed % 3 else "stable").encode(pl)
        got = cached_solve(enc.problem)
        if got is None:
            continue
        try:
            ref = oracle_lib.solve(enc.problem)
        except RuntimeError:
            continue
        same(got[0], ref, f"seed {seed} ")
        ran += 1
    assert ran >= 80, ran


@pytest.mark.parametrize("apps,replicas,order", [(3, 5, 0), (10, 30, 0), (40, 50, 0), (40, 50, 1), (30, 400, 0)])
def test_c3_shapes_match_the_oracle(apps, replicas, order):
    enc = workloads.config_c3(n_apps=apps, replicas=replicas, n_its=300)
    enc.problem.set("claim_order_mode", order)
    got = cached_solve(enc.problem)
    assert got is not None
    same(got[0], oracle_lib.solve(enc.problem, threads=4), f"C3[{apps}x{replicas}] ")


def test_topology_fuzz_problems_match_the_oracle():
    """the fuzz generator's problems with their topology constraints (spread with minDomains and policies, pod affinity and
    anti-affinity on hostname / zone / capacity type, namespaces), without existing nodes and NodePool limits"""
    from karpenter_b200.scheduler import Scheduler
    from tests import fuzz
    ran = topo = 0
    for seed in range(400):
        pools, per_pool, _, pl = fuzz.problem(seed, with_nodes=False)
        for np_ in pools:
            np_.limits = {}
        enc = Scheduler(pools, per_pool, [], claim_order="go" if seed % 3 else "stable").encode(pl)
        got = cached_solve(enc.problem)
        if got is None:
            continue
        try:
            ref = oracle_lib.solve(enc.problem)
        except RuntimeError:
            continue
        same(got[0], ref, f"seed {seed} ")
        ran += 1
        topo += int(ref["n_groups"] > 0)
    assert ran >= 100 and topo >= 40, (ran, topo)


@pytest.mark.parametrize("kw", [dict(), dict(n_nodes=50, n_pods=1500), dict(n_nodes=400, n_pods=2500, fill=0.9)])
def test_existing_nodes_match_the_oracle(kw):
    enc = workloads.config_existing(**kw)
    got = cached_solve(enc.problem)
    assert got is not None
    same(got[0], oracle_lib.solve(enc.problem), f"existing nodes {kw} ")


def test_fuzz_problems_with_existing_nodes_match_the_oracle():
    """the fuzz generator's problems as they are (existing nodes with running pods, taints, topology), NodePool limits cleared"""
    from karpenter_b200.scheduler import Scheduler
    from tests import fuzz
    ran = nodes = 0
    for seed in range(400):
        pools, per_pool, state_nodes, pl = fuzz.problem(seed)
        for np_ in pools:
            np_.limits = {}
        enc = Scheduler(pools, per_pool, state_nodes, claim_order="go" if seed % 3 else "stable").encode(pl)
        got = cached_solve(enc.problem)
        if got is None:
            continue
        try:
            ref = oracle_lib.solve(enc.problem)
        except RuntimeError:
            continue
        same(got[0], ref, f"seed {seed} ")
        ran += 1
        nodes += int(bool(state_nodes) and (ref["pod_target"] >= 0).any())
    assert ran >= 100 and nodes >= 30, (ran, nodes)


def test_c5_shards_match_the_oracle():
    """C5's constraint mix (C2 half + C3 half, pods pinned to their NodePool) at 1/64 size, shard by shard"""
    shards = workloads.config_c5_shards(160_000, 8, 1000, 1000, [[p] for p in range(8)])
    for i, enc in enumerate(shards[:3]):
        got = cached_solve(enc.problem)
        assert got is not None
        same(got[0], oracle_lib.solve(enc.problem, threads=8), f"C5 shard {i} ")


CONSOL_KEYS = ["decision", "n_new_claims", "n_unscheduled", "replacement_its"]


@pytest.mark.parametrize("kw", [dict(n_nodes=300, n_pods=1500, n_candidates=12, max_subset=3),
                                dict(n_nodes=200, n_pods=2500, n_candidates=10, max_subset=3),
                                dict(n_nodes=1000, n_pods=12000, n_candidates=14, max_subset=3),
                                dict(n_nodes=400, n_pods=4000, n_candidates=16, max_subset=2, spot_fraction=1.0, spot_to_spot=True),
                                dict(n_nodes=400, n_pods=4000, n_candidates=16, max_subset=2, spot_fraction=0.5)])
def test_c4_consolidation_matches_the_oracle(kw):
    enc, consol = workloads.config_c4(**kw)
    ci = _abi.ConsolInput(**consol)
    got = oracle_lib.cached_consolidate(enc.problem, ci)
    assert got is not None
    ref = oracle_lib.consolidate(enc.problem, ci, threads=8)
    for k in CONSOL_KEYS:
        assert np.array_equal(got[0][k], ref[k]), (k, np.argwhere(got[0][k] != ref[k])[:5].tolist())


def test_c4_uninitialized_nodes_and_unknown_instance_type_match_the_oracle():
    enc, consol = workloads.config_c4(n_nodes=300, n_pods=1500, n_candidates=12, max_subset=2)
    flags = enc.problem.get("node_flags").copy()
    flags[::7] &= ~np.uint8(2)  # KP_NODE_INITIALIZED
    enc.problem.set("node_flags", flags)
    node_it = consol["node_it"].copy()
    node_it[consol["subset_nodes"][0]] = -1
    consol["node_it"] = node_it
    ci = _abi.ConsolInput(**consol)
    got = oracle_lib.cached_consolidate(enc.problem, ci)
    ref = oracle_lib.consolidate(enc.problem, ci, threads=8)
    for k in CONSOL_KEYS:
        assert np.array_equal(got[0][k], ref[k]), k


def test_out_of_scope_shapes_are_refused():
    from tests.test_fuzz_parity import encode_reserved
    assert cached_solve(encode_reserved(1).problem) is None
