import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "chatterbox-tts_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def tiny_cfg():
    from cbx_b200.config import ModelConfig
    return ModelConfig.tiny()


def bf16_round(sd):
    """The checkpoint the GPU path sees: every tensor bf16-representable, so parity measures kernels, not weight quantisation."""
    import torch
    return {k: (v.to(torch.bfloat16).float() if v.is_floating_point() else v) for k, v in sd.items()}
