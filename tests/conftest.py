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


@pytest.fixture(scope="session", params=["checks_dense.pt", "checks_redlight.pt"])
def golden_checks_any(request):
    """Both TrafficRuleChecker fixtures: the dense scene (collisions / road edge) and the crafted scene in which
    `run_red_light` (168 positives) and `passive` (836) fire (synth.make_rule_scene_batch)."""
    return torch.load(os.path.join(GOLDEN, request.param), weights_only=False)


@pytest.fixture(scope="session")
def golden_navi():
    return torch.load(os.path.join(GOLDEN, "navi_pred.pt"), weights_only=False)


@pytest.fixture(scope="session")
def golden_wosac():
    return torch.load(os.path.join(GOLDEN, "wosac_post.pt"), weights_only=False)


@pytest.fixture(scope="session")
def golden_womd():
    return torch.load(os.path.join(GOLDEN, "womd_post.pt"), weights_only=False)


def _match_womd_modes(trajs, scores, ref_trajs, ref_scores, rtol=1e-5):
    """The reference's top-k is `sorted=False`: match every kept mode to the reference's by its (bit-identical,
    gathered) trajectory, then compare the scores mode by mode. Shapes [n_sc,n_ag,k,n_out,3] / [n_sc,n_ag,k]."""
    assert trajs.shape == ref_trajs.shape and scores.shape == ref_scores.shape
    d = (trajs[:, :, :, None] - ref_trajs[:, :, None]).abs().flatten(4).amax(-1)  # [n_sc,n_ag,k,k]
    idx = d.argmin(-1)
    assert float(d.min(-1)[0].max()) == 0.0, "a kept trajectory is not among the reference's"
    assert all(len(set(r.tolist())) == r.numel() for r in idx.flatten(0, 1)), "modes must map one to one"
    ref = torch.gather(ref_scores, 2, idx)
    assert torch.allclose(scores, ref, rtol=rtol, atol=1e-7), float((scores - ref).abs().max())


@pytest.fixture(scope="session")
def match_womd_modes():
    return _match_womd_modes
