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
