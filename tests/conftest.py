import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The C-ABI library must exist for both suites (the CPU suite checks that it loads and exports the ABI)."""
    from floppity_deprecated_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.LIB_PATH


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, f"golden_{name}.npz"), allow_pickle=False))


def golden_oracle_net(name):
    """Re-create the oracle network of a golden case (stored weights when present, seeded init otherwise)."""
    sys.path.insert(0, GOLDEN)
    import make_golden
    from oracle import nsf_oracle as O
    g = load_golden(name)
    D, C, T, K, H, Bk, B, store_w = make_golden.CASES[name]
    if store_w:
        net = O.build_nsf(torch.tensor(g["theta_train"]), torch.tensor(g["x_train"]), hidden_features=H,
                          num_transforms=T, num_bins=K, num_blocks=Bk)
        sd = {k[3:]: torch.tensor(v) for k, v in g.items() if k.startswith("sd/")}
        net.load_state_dict(sd)
    else:
        net = make_golden.build_case(name)[0]
    net.eval()
    return net, g, make_golden.CASES[name]


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(1.0, np.abs(b))
