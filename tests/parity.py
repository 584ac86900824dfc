"""Bit-exact comparison of a CUDA-path result with the oracle's (PARITY_KEYS of karpenter_b200/_abi.py)."""
import numpy as np

from karpenter_b200 import _abi


def assert_same(gpu: dict, orc: dict, what=""):
    for k in _abi.PARITY_KEYS:
        a, b = gpu[k], orc[k]
        if isinstance(a, np.ndarray):
            assert a.shape == b.shape, f"{what}{k}: shape {a.shape} != {b.shape}"
            if not np.array_equal(a, b):
                idx = np.argwhere(a != b)[:5]
                raise AssertionError(f"{what}{k}: {len(np.argwhere(a != b))} mismatches, first at {idx.tolist()}: "
                                     f"gpu={a[tuple(idx[0])]} oracle={b[tuple(idx[0])]}")
        else:
            assert a == b, f"{what}{k}: {a} != {b}"
