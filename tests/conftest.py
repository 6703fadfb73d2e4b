import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run by `pytest -m gpu` on the GPU box")


@pytest.fixture(scope="session")
def golden_q():
    return torch.load(os.path.join(GOLDEN, "quantizer_ref.pt"), weights_only=False)


@pytest.fixture(scope="session")
def golden_c():
    return torch.load(os.path.join(GOLDEN, "codec_oracle.pt"), weights_only=False)


@pytest.fixture(scope="session")
def golden_w():
    """Outputs of the reference's OWN quantization / quant_int packages (oracle/make_golden.py::wrap_vectors)."""
    return torch.load(os.path.join(GOLDEN, "wrap_ref.pt"), weights_only=False)


@pytest.fixture(scope="session")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
