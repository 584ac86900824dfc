"""ctypes binding of the CPU oracle (oracle/liborc.so). TEST INFRASTRUCTURE: imported by tests/, smoke() and the
cpu_baseline legs of bench.py only."""
import ctypes as C
import os
import subprocess

import numpy as np

from karpenter_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "liborc.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.orc_solve.argtypes = [C.c_void_p, C.c_void_p]
        _LIB.orc_solve_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        _LIB.orc_consolidate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB.orc_consolidate_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _LIB.orc_feasibility.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    return _LIB


def solve(problem, threads: int = 1) -> dict:
    """threads > 1: candidates are evaluated by a worker pool like the reference's parallelizeUntil (same result)."""
    r = _abi.kp_result()
    rc = lib().orc_solve_mt(problem.ref(), C.byref(r), int(threads))
    if rc != 0:
        raise RuntimeError(f"orc_solve failed: {rc}")
    out = _abi.result_to_dict(r, problem.n_resources)
    lib().orc_result_free(C.byref(r))
    return out


def consolidate(problem, consol, threads: int = 1) -> dict:
    r = _abi.kp_consol_result()
    rc = lib().orc_consolidate_mt(problem.ref(), consol.ref(), C.byref(r), int(threads))
    if rc != 0:
        raise RuntimeError(f"orc_consolidate failed: {rc}")
    out = _abi.consol_result_to_dict(r)
    lib().orc_consol_result_free(C.byref(r))
    return out


def feasibility(problem) -> np.ndarray:
    itw = (problem.n_its + 63) // 64
    out = np.zeros((problem.n_classes, problem.n_templates, itw), np.uint64)
    w = C.c_int32()
    rc = lib().orc_feasibility(problem.ref(), out.ctypes.data, C.byref(w))
    assert rc == 0 and w.value == itw
    return out


# ---- oracle/orc_cached.cpp: the CUDA solver's algorithm on one host core (bench.py cpu_baseline legs, tests) -------------
_CACHED = None
CACHED_KEYS = ["pod_target", "pod_error", "n_claims", "claim_template", "claim_npods", "claim_rank", "claim_requests", "claim_its"]


def cached_lib():
    global _CACHED
    if _CACHED is None:
        path = os.path.join(ROOT, "oracle", "liborc_cached.so")
        if not os.path.exists(path):
            build()
        _CACHED = C.CDLL(path)
        _CACHED.orc_cached_solve.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        _CACHED.orc_cached_free.argtypes = [C.c_void_p]
    return _CACHED


def cached_solve(problem):
    """-> (result dict with CACHED_KEYS, solve ms on one core with the tables prepared, prep ms), or None when the shape is
    outside what orc_cached.cpp serves (KP_ERR_UNSUPPORTED)"""
    r = _abi.kp_result()
    prep = C.c_double()
    rc = cached_lib().orc_cached_solve(problem.ref(), C.byref(r), C.byref(prep))
    if rc == 5:
        return None
    if rc != 0:
        raise RuntimeError(f"orc_cached_solve failed: {rc}")
    Cn, R, W = r.n_claims, problem.n_resources, r.it_words
    out = {"pod_target": _abi.view(r.pod_target, r.n_pods, np.int32).copy(), "pod_error": _abi.view(r.pod_error, r.n_pods, np.uint8).copy(),
           "n_claims": Cn, "claim_template": _abi.view(r.claim_template, Cn, np.int32).copy(),
           "claim_npods": _abi.view(r.claim_npods, Cn, np.int32).copy(), "claim_rank": _abi.view(r.claim_rank, Cn, np.int32).copy(),
           "claim_requests": _abi.view(r.claim_requests, Cn * R, np.int64).reshape(Cn, R).copy(),
           "claim_its": _abi.view(r.claim_its, Cn * W, np.uint64).reshape(Cn, W).copy()}
    ms = r.solve_ms
    cached_lib().orc_cached_free(C.byref(r))
    return out, ms, prep.value


def cached_consolidate(problem, consol):
    """kp_consolidate's fast path on one host core (oracle/orc_cached.cpp) -> (dict with decision / replacement_its /
    n_new_claims / n_unscheduled, solve ms, prep ms), or None outside its scope"""
    lib_ = cached_lib()
    lib_.orc_cached_consolidate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
    lib_.orc_cached_consol_free.argtypes = [C.c_void_p]
    r = _abi.kp_consol_result()
    prep = C.c_double()
    rc = lib_.orc_cached_consolidate(problem.ref(), consol.ref(), C.byref(r), C.byref(prep))
    if rc == 5:
        return None
    if rc != 0:
        raise RuntimeError(f"orc_cached_consolidate failed: {rc}")
    S, W = r.n_subsets, r.it_words
    out = {"decision": _abi.view(r.decision, S, np.uint8).copy(),
           "replacement_its": _abi.view(r.replacement_its, S * W, np.uint64).reshape(S, W).copy(),
           "n_new_claims": _abi.view(r.n_new_claims, S, np.int32).copy(), "n_unscheduled": _abi.view(r.n_unscheduled, S, np.int32).copy()}
    ms = r.solve_ms
    lib_.orc_cached_consol_free(C.byref(r))
    return out, ms, prep.value
