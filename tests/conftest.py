import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pd():
    """The product's Python binding over the C ABI (builds the .so here if it is missing)."""
    mod = importlib.import_module("soft-body-simulation-cuda_b200")
    if not os.path.exists(mod.LIB_PATH):
        build = importlib.import_module("soft-body-simulation-cuda_b200.build")
        build.build()
    return mod


@pytest.fixture(scope="session")
def O():
    import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def assets(tmp_path_factory):
    """TetGen assets + a context.json in the reference's schema, regenerated from the fixture."""
    import meshes
    root = tmp_path_factory.mktemp("scene")
    return meshes.write_assets(str(root))
