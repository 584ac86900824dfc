"""ctypes view of include/karpsolve.h.

The struct layouts are parsed from the header itself, so the Python side can never drift from the C ABI.
`Problem` keeps the numpy arrays backing a `kp_problem` alive and exposes them by field name.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from typing import Dict

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "karpsolve.h")

_SCALARS = {
    "int32_t": (C.c_int32, np.int32), "int64_t": (C.c_int64, np.int64), "uint8_t": (C.c_uint8, np.uint8),
    "uint32_t": (C.c_uint32, np.uint32), "uint64_t": (C.c_uint64, np.uint64), "double": (C.c_double, np.float64),
    "int": (C.c_int, np.int32),
}


def _parse_header():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    defines = {}
    for m in re.finditer(r"#define\s+(KP_\w+)\s+\(?(-?(?:0x)?[0-9a-fA-F]+)u?\)?\s*$", src, flags=re.M):
        defines[m.group(1)] = int(m.group(2), 0)
    structs = {}
    for m in re.finditer(r"typedef struct (\w+) \{(.*?)\} (\w+);", src, flags=re.S):
        fields = []
        for decl in m.group(2).split(";"):
            decl = decl.strip()
            if not decl:
                continue
            fm = re.match(r"(const\s+)?(\w+)\s*(\*?)\s*(.+)$", decl)
            base, ptr, names = fm.group(2), fm.group(3), fm.group(4)
            for name in names.split(","):
                name = name.strip()
                p = ptr
                if name.startswith("*"):
                    p, name = "*", name[1:].strip()
                fields.append((name, base, bool(p)))
        structs[m.group(3)] = fields
    return defines, structs


DEFINES, STRUCTS = _parse_header()
globals().update(DEFINES)


def _make_struct(name):
    flds = []
    for fname, base, is_ptr in STRUCTS[name]:
        if base == "void":
            ct = C.c_void_p
        elif is_ptr:
            ct = C.c_void_p  # typed access goes through numpy; keeps None == NULL simple
        else:
            ct = _SCALARS[base][0]
        flds.append((fname, ct))
    return type(name, (C.Structure,), {"_fields_": flds})


kp_problem = _make_struct("kp_problem")
kp_result = _make_struct("kp_result")
kp_consol_input = _make_struct("kp_consol_input")
kp_consol_result = _make_struct("kp_consol_result")
kp_stats = _make_struct("kp_stats")


def _field_dtype(struct, fname):
    for n, base, is_ptr in STRUCTS[struct]:
        if n == fname:
            return _SCALARS[base][1], is_ptr
    raise KeyError(fname)


class _Holder:
    """Owns numpy arrays and mirrors them into a ctypes struct."""
    STRUCT = ""
    CT = None

    def __init__(self, **fields):
        self.arrays: Dict[str, np.ndarray] = {}
        self.c = self.CT()
        for k, v in fields.items():
            self.set(k, v)

    def set(self, name, value):
        dt, is_ptr = _field_dtype(self.STRUCT, name)
        if is_ptr:
            if value is None:
                self.arrays.pop(name, None)
                setattr(self.c, name, None)
                return
            arr = np.ascontiguousarray(value, dtype=dt)
            if arr.size == 0:
                arr = np.zeros(1, dtype=dt)[:0].copy()
                keep = np.zeros(1, dtype=dt)  # never hand out a dangling pointer for empty arrays
                self.arrays[name + "__pad"] = keep
                self.arrays[name] = arr
                setattr(self.c, name, keep.ctypes.data)
                return
            self.arrays[name] = arr
            setattr(self.c, name, arr.ctypes.data)
        else:
            setattr(self.c, name, int(value) if dt != np.float64 else float(value))

    def get(self, name):
        dt, is_ptr = _field_dtype(self.STRUCT, name)
        if is_ptr:
            return self.arrays.get(name)
        return getattr(self.c, name)

    def __getattr__(self, name):
        if name in ("arrays", "c"):
            raise AttributeError(name)
        try:
            return self.get(name)
        except KeyError:
            raise AttributeError(name)

    def ref(self):
        return C.byref(self.c)

    def nbytes(self):
        return sum(a.nbytes for k, a in self.arrays.items() if not k.endswith("__pad"))


class Problem(_Holder):
    STRUCT = "kp_problem"
    CT = kp_problem


class ConsolInput(_Holder):
    STRUCT = "kp_consol_input"
    CT = kp_consol_input


def view(ptr, n, dtype):
    """numpy copy of a C array returned by the library."""
    if not ptr or n <= 0:
        return np.zeros(0, dtype=dtype)
    ct = np.ctypeslib.as_ctypes_type(dtype)
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,)).copy()


def result_to_dict(r: kp_result, n_resources: int) -> dict:
    C_ = r.n_claims
    out = {
        "pod_target": view(r.pod_target, r.n_pods, np.int32),
        "pod_error": view(r.pod_error, r.n_pods, np.uint8),
        "n_claims": C_,
        "claim_template": view(r.claim_template, C_, np.int32),
        "claim_npods": view(r.claim_npods, C_, np.int32),
        "claim_rank": view(r.claim_rank, C_, np.int32),
        "claim_requests": view(r.claim_requests, C_ * n_resources, np.int64).reshape(C_, n_resources),
        "it_words": r.it_words,
        "claim_its": view(r.claim_its, C_ * r.it_words, np.uint64).reshape(C_, max(r.it_words, 0)),
        "n_keys": r.n_keys,
        "mask_words": r.mask_words,
        "claim_req_flags": view(r.claim_req_flags, C_ * r.n_keys, np.uint8).reshape(C_, r.n_keys),
        "claim_req_gte": view(r.claim_req_gte, C_ * r.n_keys, np.int64).reshape(C_, r.n_keys),
        "claim_req_lte": view(r.claim_req_lte, C_ * r.n_keys, np.int64).reshape(C_, r.n_keys),
        "claim_req_mask": view(r.claim_req_mask, C_ * r.mask_words, np.uint64).reshape(C_, r.mask_words),
        "n_groups": r.n_groups,
        "group_domain_off": view(r.group_domain_off, r.n_groups + 1, np.int32),
        "domain_counts": view(r.domain_counts, r.n_domain_slots, np.int32),
        "n_existing_evals": r.n_existing_evals,
        "n_inflight_evals": r.n_inflight_evals,
        "n_template_evals": r.n_template_evals,
        "n_commits": r.n_commits,
        "solve_ms": r.solve_ms,
        "claim_reservations": view(r.claim_reservations, C_, np.uint64),
        "claim_dropped": view(r.claim_dropped, C_, np.uint8),
    }
    return out


def consol_result_to_dict(r: kp_consol_result) -> dict:
    S = r.n_subsets
    return {
        "decision": view(r.decision, S, np.uint8),
        "it_words": r.it_words,
        "replacement_its": view(r.replacement_its, S * r.it_words, np.uint64).reshape(S, r.it_words),
        "n_new_claims": view(r.n_new_claims, S, np.int32),
        "n_unscheduled": view(r.n_unscheduled, S, np.int32),
        "solve_ms": r.solve_ms,
        "n_keys": r.n_keys, "mask_words": r.mask_words,
        "repl_template": view(r.repl_template, S, np.int32),
        "repl_requests": view(r.repl_requests, S * r.n_resources, np.int64).reshape(S, max(r.n_resources, 0)),
        "repl_req_flags": view(r.repl_req_flags, S * r.n_keys, np.uint8).reshape(S, max(r.n_keys, 0)),
        "repl_req_gte": view(r.repl_req_gte, S * r.n_keys, np.int64).reshape(S, max(r.n_keys, 0)),
        "repl_req_lte": view(r.repl_req_lte, S * r.n_keys, np.int64).reshape(S, max(r.n_keys, 0)),
        "repl_req_mask": view(r.repl_req_mask, S * r.mask_words, np.uint64).reshape(S, max(r.mask_words, 0)),
        "repl_order_off": view(r.repl_order_off, S + 1, np.int32) if r.repl_order_off else None,
        "repl_order": (view(r.repl_order, int(view(r.repl_order_off, S + 1, np.int32)[-1]), np.int32)
                       if r.repl_order_off else None),
    }


# keys of a consolidation result that must be bit-identical between the CUDA path and the oracle
CONSOL_PARITY_KEYS = ["decision", "n_new_claims", "n_unscheduled", "replacement_its", "repl_template", "repl_requests",
                      "repl_req_flags", "repl_req_gte", "repl_req_lte", "repl_req_mask"]


# result keys that must be bit-identical between the CUDA path and the oracle
PARITY_KEYS = [
    "pod_target", "pod_error", "n_claims", "claim_template", "claim_npods", "claim_rank", "claim_requests",
    "claim_its", "claim_req_flags", "claim_req_gte", "claim_req_lte", "claim_req_mask", "group_domain_off",
    "domain_counts", "claim_reservations", "claim_dropped",
]
