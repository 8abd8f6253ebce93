import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tbpkg  # noqa: E402,F401  registers the package as `trafficbotsv1_5_b200`

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_ops():
    return torch.load(os.path.join(GOLDEN, "ops.pt"), weights_only=False)


@pytest.fixture(scope="session")
def golden_rollout():
    return torch.load(os.path.join(GOLDEN, "rollout_small.pt"), weights_only=False)


@pytest.fixture(scope="session")
def golden_checks():
    return torch.load(os.path.join(GOLDEN, "checks_dense.pt"), weights_only=False)


@pytest.fixture(scope="session")
def golden_navi():
    return torch.load(os.path.join(GOLDEN, "navi_pred.pt"), weights_only=False)


@pytest.fixture(scope="session")
def golden_wosac():
    return torch.load(os.path.join(GOLDEN, "wosac_post.pt"), weights_only=False)
