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
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def weights_checksum(sd):
    return float(sum(v.double().abs().sum().item() for k, v in sd.items() if v.is_floating_point()))


def zero_dropout(module):
    for m in module.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if hasattr(m, "dropout_p"):
            m.dropout_p = 0.0


def seeded_model(kind, seed=1337):
    """Product modules built on the CPU under the reference's seed (identical initial weights)."""
    from chap_b200 import networks
    torch.manual_seed(seed)
    if kind == "dualdecoder2d":
        m = networks.DualDecoder(1, 4, {"decoder_type": "mcnet"})
    elif kind == "unet2d":
        m = networks.UNet(1, 4)
    elif kind == "dualdecoder3d":
        m = networks.DualDecoder3d(1, 2, normalization='batchnorm', has_dropout=False)
    elif kind == "vnet":
        m = networks.VNet(1, 2, normalization='batchnorm', has_dropout=False)
    else:
        raise ValueError(kind)
    zero_dropout(m)
    return m


def rel_err(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def max_err(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max())
