import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


GOLDEN_CASES = ["nce_single", "nce_2attn", "prior_additive", "prior_mult", "prior_event_given", "nce_pred4",
                "tower_additive", "mult_2layers", "hier_2x7", "hier_options", "switch_bce", "switch_asl_master"]


def load_golden(name):
    import torch
    return torch.load(os.path.join(ROOT, "tests", "golden", f"{name}.pt"), weights_only=False)
