import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def pkg():
    import vslam_b200_loader
    return vslam_b200_loader.pkg


@pytest.fixture(scope="session")
def pattern():
    from oracle.orb_pattern import load_pattern
    return load_pattern()


@pytest.fixture(scope="session")
def gpu_ctx(pkg):
    """One context for the whole GPU session.  Fails loudly (no fallback) if the library or device is missing."""
    ctx = pkg.Context(device=0, max_images=16, max_width=1241, max_height=376, max_keypoints=8192)
    yield ctx
    ctx.close()


@pytest.fixture(scope="module")
def gpu_ctx_big(pkg):
    """A context sized for multi-chunk batches of the host-buffer frontend (80 images, 1280 keypoints each)."""
    ctx = pkg.Context(device=0, max_images=80, max_width=1241, max_height=376, max_keypoints=1280)
    yield ctx
    ctx.close()
