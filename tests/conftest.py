import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = os.environ.get("L32_REFERENCE", "/root/reference")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    """npz fixture -> dict of torch tensors / python scalars; *_bits arrays are bf16 bit patterns."""
    out = {}
    with np.load(os.path.join(GOLDEN, name), allow_pickle=False) as z:
        for k in z.files:
            a = z[k]
            if k.endswith("_bits"):
                out[k[:-5]] = torch.from_numpy(a.view(np.int16).copy()).view(torch.bfloat16).float()
            elif a.dtype.kind in "US":
                out[k] = a.tolist() if a.ndim else str(a)
            elif a.ndim == 0:
                out[k] = a.item()
            else:
                out[k] = torch.from_numpy(a.copy())
    return out


@pytest.fixture(scope="session")
def golden():
    return load_golden


def have_reference():
    return os.path.isfile(os.path.join(REFERENCE, "Model", "model.py"))
