"""C-ABI checks that need no GPU: libkarpsolve.so loads, exports every entry point include/karpsolve.h declares, the
ctypes mirrors of the structs have the C compiler's layout, and without a CUDA device the library fails loudly instead of
falling back to a CPU path."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

from karpenter_b200 import _abi, _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "karpsolve.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"^\s*(?:const\s+char\s*\*|int|void|int64_t|double)\s+(kp_\w+)\s*\(", src, flags=re.M)))


def test_header_declares_the_documented_entry_points():
    names = declared_functions()
    for must in ("kp_version", "kp_create", "kp_destroy", "kp_last_error", "kp_solve", "kp_result_free", "kp_upload",
                 "kp_solve_resident", "kp_consolidate", "kp_consol_result_free", "kp_feasibility", "kp_get_stats"):
        assert must in names
    assert sorted(_native.EXPORTS) == names  # the binding lists exactly what the header declares


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_native.LIB_PATH), "run __graft_entry__.build() first"
    lib = C.CDLL(_native.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} is declared in karpsolve.h but not exported"
    lib.kp_version.restype = C.c_int
    m = re.search(r"#define\s+KP_ABI_VERSION\s+(\d+)", open(HEADER).read())
    assert lib.kp_version() == int(m.group(1))


def test_ctypes_structs_match_the_c_layout():
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "karpsolve.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu\n", sizeof(kp_problem), sizeof(kp_result), sizeof(kp_consol_input),
         sizeof(kp_consol_result), sizeof(kp_stats));
  printf("%zu %zu %zu\n", offsetof(kp_problem, n_pods), offsetof(kp_result, n_existing_evals),
         offsetof(kp_consol_input, spot_to_spot_enabled));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as td:
        src, exe = os.path.join(td, "t.c"), os.path.join(td, "t")
        open(src, "w").write(prog)
        subprocess.check_call(["/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc", "-I",
                               os.path.join(ROOT, "include"), src, "-o", exe])
        out = subprocess.check_output([exe], text=True).split()
    sizes = [C.sizeof(t) for t in (_abi.kp_problem, _abi.kp_result, _abi.kp_consol_input, _abi.kp_consol_result,
                                   _abi.kp_stats)]
    assert [int(x) for x in out[:5]] == sizes
    offs = [_abi.kp_problem.n_pods.offset, _abi.kp_result.n_existing_evals.offset,
            _abi.kp_consol_input.spot_to_spot_enabled.offset]
    assert [int(x) for x in out[5:]] == offs


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(_native.SolverError) as e:
        _native.Handle()
    assert e.value.code == 3  # KP_ERR_CUDA
