import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def rv():
    import rvpt_b200
    return rvpt_b200


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.load()
    return oracle


class PreparedScene:
    """A scene after the host-side steps of RVPT::initialize(): BVH built,
    triangles permuted (rvpt.cpp:84-86). The same bytes feed engine and oracle."""

    def __init__(self, rv, scene):
        self.scene = scene
        self.nodes, self.perm = rv.build_bvh(scene.triangles)
        self.triangles = np.ascontiguousarray(scene.triangles[self.perm])
        self.materials = scene.materials


@pytest.fixture(scope="session")
def builtin(rv):
    return PreparedScene(rv, rv.builtin_scene())


@pytest.fixture(scope="session")
def cornell(rv):
    return PreparedScene(rv, rv.cornell_scene())


# poses: the literal default of main.cpp (camera.h:48-49) and the pinned one of SURVEY §3.4
DEFAULT_POSE = (0.0, 0.0, 0.0)
PINNED_POSE = (0.0, 0.8, -2.5)
CORNELL_POSE = (0.0, 1.2, -3.4)
