import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"
GOLDEN_CASES = ["tiny_f32", "tiny_f64", "enc_small_f32", "pad_small_f32", "dec_small_f32", "odd_dims_f64"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    z = np.load(GOLDEN / f"{name}.npz")
    d = {k: torch.from_numpy(z[k]) for k in z.files}
    d["shape_list"] = [tuple(int(x) for x in row) for row in z["shapes"]]
    return d


@pytest.fixture(scope="session")
def c_oracle():
    from oracle.msda_oracle import COracle

    return COracle()


def rel_err(a, b):
    """max |a-b| / max |b| — the relative error used for every tolerance in this suite."""
    a, b = a.double(), b.double()
    denom = b.abs().max().clamp_min(1e-30)
    return ((a - b).abs().max() / denom).item()
