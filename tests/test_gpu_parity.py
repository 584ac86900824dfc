"""-m gpu: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

from karpenter_b200 import _native, workloads
from tests import oracle_lib
from tests.parity import assert_same

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def handle():
    h = _native.Handle()
    yield h
    h.close()


@pytest.mark.parametrize("n_pods", [1, 13, 60, 200, 1000])
def test_c1_parity(handle, n_pods):
    enc = workloads.config_c1(n_pods=n_pods)
    assert_same(handle.solve(enc.problem), oracle_lib.solve(enc.problem), f"C1[{n_pods}] ")


def test_c1_stable_order(handle):
    enc = workloads.config_c1(n_pods=500)
    enc.problem.set("claim_order_mode", 1)
    assert_same(handle.solve(enc.problem), oracle_lib.solve(enc.problem), "C1 stable ")


def test_feasibility_parity(handle):
    enc = workloads.config_c2(n_pods=2000, n_its=500)
    assert np.array_equal(handle.feasibility(enc.problem), oracle_lib.feasibility(enc.problem))


@pytest.mark.parametrize("n_pods", [300, 3000])
def test_c2_parity(handle, n_pods):
    enc = workloads.config_c2(n_pods=n_pods, n_its=500)
    assert_same(handle.solve(enc.problem), oracle_lib.solve(enc.problem), f"C2[{n_pods}] ")


@pytest.mark.parametrize("apps,replicas", [(3, 5), (10, 30), (40, 50)])
def test_c3_parity(handle, apps, replicas):
    enc = workloads.config_c3(n_apps=apps, replicas=replicas, n_its=300)
    assert_same(handle.solve(enc.problem), oracle_lib.solve(enc.problem), f"C3[{apps}x{replicas}] ")


def test_c2_full_size_parity(handle):
    """BASELINE configs[1] at full size: 100k pods x 500 instance types, bit-identical to the oracle."""
    enc = workloads.config_c2()
    assert_same(handle.solve(enc.problem), oracle_lib.solve(enc.problem), "C2[100k] ")


def test_c3_medium_parity(handle):
    enc = workloads.config_c3(n_apps=200, replicas=200, n_its=1000)
    assert_same(handle.solve(enc.problem), oracle_lib.solve(enc.problem), "C3[200x200] ")
