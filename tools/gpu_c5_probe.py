"""Probe: C5 NodePool shards on the CUDA path -- one shard alone (kp_solve) and all shards as one kp_solve_batch."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time

from karpenter_b200 import _native, workloads

n_pods = int(sys.argv[1]) if len(sys.argv) > 1 else 1_600_000
mode = sys.argv[2] if len(sys.argv) > 2 else "both"
t = time.time()
shards = workloads.config_c5_shards(n_pods=n_pods)
print(f"C5 {n_pods} pods: 8 shards generated in {time.time() - t:.1f}s", flush=True)
h = _native.Handle()
if mode in ("both", "single"):
    t = time.time()
    res = h.solve(shards[0].problem)
    dt = time.time() - t
    st = h.stats()
    n = int(shards[0].problem.n_pods)
    print(f"pool 0 alone: {n} pods, {res['n_claims']} claims, unsched {(res['pod_target'] == -1).sum()}, kernels {st['solve_ms']:.0f} ms "
          f"({st['solve_ms'] * 1000 / n:.2f} us/pod), e2e {dt:.1f}s prep {st['prep_ms']:.0f} ms upload {st['upload_ms']:.0f} ms", flush=True)
if mode in ("both", "batch"):
    t = time.time()
    outs = h.solve_batch([e.problem for e in shards])
    dt = time.time() - t
    st = h.stats()
    print(f"8 pools as one batch: kernels {st['solve_ms']:.0f} ms ({n_pods / st['solve_ms'] * 1000:.0f} pods/s), e2e {dt:.1f}s, "
          f"prep {st['prep_ms']:.0f} ms upload {st['upload_ms']:.0f} ms download {st['download_ms']:.0f} ms, "
          f"claims {sum(int(o['n_claims']) for o in outs)}", flush=True)
h.close()
