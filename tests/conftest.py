import ast
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))     # tests may import the oracle (test infrastructure)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["cfg1_small", "cfg2_small", "cfg2_noNG", "cfg4_ragged", "missing_mask", "cfg3_small"]


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
    d = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    d["lists"] = ast.literal_eval(str(d["lists"]))
    return d


def oracle_components(g):
    """Oracle kernel components carrying the golden case's constrained hyper-parameters (leaf tensors)."""
    import lvae_oracle as orc
    L = g["mu"].shape[1]
    k0, k1 = orc.parse_kernel_lists(L, **g["lists"], id_covariate=2)
    ls = torch.from_numpy(g["lengthscale"])
    os_ = torch.from_numpy(g["outputscale"])
    i_ls = 0
    for i_c, comp in enumerate(k0 + k1):
        comp.outputscale = os_[i_c].clone().requires_grad_(True)
        for k in sorted(comp.lengthscales):
            comp.lengthscales[k] = ls[i_ls].clone().requires_grad_(True)
            i_ls += 1
    return k0, k1


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    return request.param, load_golden(request.param)
