import importlib
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `pytest -m gpu` under gpurun)")


@pytest.fixture(scope="session")
def yart():
    """The product package; builds libyart_b200.so in-tree if it is stale or missing."""
    build = importlib.import_module("yet-another-raytracer_b200.build")
    build.build()
    mod = importlib.import_module("yet-another-raytracer_b200")
    mod.load_library()
    return mod


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure)."""
    from oracle import orc as o
    o.build()
    o.lib()
    return o


@pytest.fixture(scope="session")
def assets(yart):
    return yart.assets_dir()


@pytest.fixture(scope="session")
def ctx(yart):
    """A GPU context.  GPU tests must run the CUDA path: no device -> hard failure, never a skip."""
    c = yart.Context(0)
    yield c
    c.close()


_mesh_cache = {}


@pytest.fixture(scope="session")
def mesh_scene(yart, orc, assets):
    """name -> (product TriangleMesh, python MeshScene description, oracle Scene) for cube/sycee/david."""

    def get(name):
        if name not in _mesh_cache:
            m = yart.TriangleMesh.from_obj(os.path.join(assets, name + ".obj"))
            ms = orc.MeshScene(m.positions(), m.normals(), m.uvs())
            _mesh_cache[name] = (m, ms, orc.Scene(ms))
        return _mesh_cache[name]

    return get
