import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_small():
    blob = dict(np.load(os.path.join(GOLD, "generic_small.npz")))
    with open(os.path.join(GOLD, "generic_small.json")) as f:
        meta = json.load(f)
    return blob, meta


@pytest.fixture(scope="session")
def golden_sliding():
    return dict(np.load(os.path.join(GOLD, "sliding_small.npz")))


def build_small_net(meta, blob, dtype=None, device="cuda"):
    """Product Generic_UNet with the golden (reference-generated) parameters loaded."""
    import torch
    from torch import nn
    from multitalent_b200.network_architecture.generic_UNet import Generic_UNet, InitWeights_He
    pool, convk = meta["pool"], meta["convk"]
    net = Generic_UNet(1, meta["base"], 47, len(pool), 2, 2, nn.Conv3d, nn.InstanceNorm3d, {'eps': 1e-5, 'affine': True},
                       nn.Dropout3d, {'p': 0, 'inplace': True}, nn.LeakyReLU, {'negative_slope': 1e-2, 'inplace': True},
                       True, False, lambda x: x, InitWeights_He(1e-2), pool, convk, False, True, True,
                       native_dtype=dtype or torch.float32)
    sd = {k[len("param/"):]: torch.from_numpy(v) for k, v in blob.items() if k.startswith("param/")}
    net.load_state_dict(sd)
    net.inference_apply_nonlin = nn.Sigmoid()
    return net.to(device)
