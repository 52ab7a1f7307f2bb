"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"` covers the oracle against the golden vectors, the host logic and the C-ABI
symbol table (no compute calls).  `-m gpu` holds the parity tests proper; they call the CUDA
path through the C ABI and compare with the oracle, the committed golden fixtures and, when
oracle/_ref/libref_gpu.so travelled to the box, the unmodified reference kernel.
"""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

GOLDEN = ROOT / "tests" / "golden"
POSE_FILE = GOLDEN / "rendezvous_1s_dt0_01s.txt"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def dsrt():
    import dsrt_b200
    return dsrt_b200


@pytest.fixture(scope="session")
def poses(dsrt):
    return dsrt.read_pose_file(POSE_FILE)


@pytest.fixture(scope="session")
def mesh10k(dsrt):
    return dsrt.synth_mesh(10000)


@pytest.fixture(scope="session")
def oracle_scene10k(mesh10k):
    import oracle
    tris, mats = mesh10k
    return oracle.Scene(tris, mats)


@pytest.fixture(scope="session")
def ref_golden():
    p = GOLDEN / "ref_gpu_10k.npz"
    if not p.exists():
        pytest.skip("tests/golden/ref_gpu_10k.npz not generated yet")
    return np.load(p)


def cam12_of(cam):
    return np.array(cam.origin.tuple() + cam.lower_left_corner.tuple() + cam.horizontal.tuple() + cam.vertical.tuple(),
                    np.float32)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)
