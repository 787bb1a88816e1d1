import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (runs on the B200 box through the C ABI)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def cornell():
    from ohao_engine_b200 import scenes
    return scenes.cornell_box(), scenes.cornell_camera()


@pytest.fixture(scope="session")
def helmet_small():
    from ohao_engine_b200 import scenes
    return scenes.helmet_class(ntris=5000, tex_size=256, env_size=(256, 128)), scenes.helmet_camera()


@pytest.fixture(scope="session")
def synthetic_small():
    from ohao_engine_b200 import scenes
    return scenes.synthetic_2m(nblobs=24, tris_per_blob=800, env_size=(128, 64)), scenes.synthetic_camera()
