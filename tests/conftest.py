import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference/finetune"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) when no device is visible, so `-m "not gpu"` and a
    plain run both stay green in the CPU container."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def load_golden(name):
    import torch
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def rel_err(x, ref):
    """Relative Frobenius error ||x-ref|| / ||ref||, computed in fp64."""
    import torch
    x = x.detach().double().cpu()
    ref = ref.detach().double().cpu()
    return float((x - ref).norm() / ref.norm().clamp_min(1e-300))


def reference_modules():
    """Import the live reference when it is present (build container only)."""
    if not os.path.isdir(REFERENCE):
        return None
    sys.path.insert(0, REFERENCE)
    try:
        import losses as ref_losses
        import optimizers as ref_opt
    finally:
        sys.path.pop(0)
    return ref_losses, ref_opt
