import json
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
    config.addinivalue_line("markers", "gpu_draft: first GPU run of the row-f1 draft kernels (pytest -m gpu_draft; never part of -m gpu, "
                                       "skipped without CUDA)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"])) if "meta" in z.files else {}
    return z, meta


def vis_cases():
    """{name: (maps, heat, ratios)} of tests/golden/vis_upsample.npz (Pillow's own outputs, oracle/make_golden_vis.py)."""
    z = np.load(os.path.join(GOLDEN, "vis_upsample.npz"), allow_pickle=False)
    meta = json.loads(bytes(z["meta"]).decode())
    return {n: (z[n + ".maps"], z[n + ".heat"], z[n + ".ratios"]) for n in meta["cases"]}, meta


def golden_names(kind):
    out = []
    for f in sorted(os.listdir(GOLDEN)):
        if f.endswith(".npz") and f not in ("pe_sine.npz", "preprocess_u8.npz", "vis_upsample.npz"):
            if f.startswith("train_"):                       # row f1 goldens (oracle/make_golden_train.py)
                if kind == "train":
                    out.append(f[:-4])
            elif kind != "train" and (kind == "head") == f.startswith("head_"):
                out.append(f[:-4])
    return out


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Build libscouter_b200.so once if it is absent (nvcc cross-compiles without a GPU)."""
    import shutil
    from scouter_b200 import _lib
    if shutil.which("nvcc") or not os.path.exists(_lib.LIB_PATH):
        _lib.build()          # no-op when the .so is newer than every source
    yield


def model_args(meta_args):
    from oracle.refshim import make_args
    return make_args(**meta_args)


def rel_err(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
