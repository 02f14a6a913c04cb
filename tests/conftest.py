import importlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    # the C-ABI library is a build artefact (git-ignored): compile it if this checkout does not have it yet
    lib = os.path.join(ROOT, "autostyle-tts_b200", "libavs.so")
    if not os.path.exists(lib):
        try:
            importlib.import_module("autostyle-tts_b200.build").build()
        except Exception as e:  # the tests that need it will fail loudly with the reason
            print(f"[conftest] could not build libavs.so: {e}")


def load_pkg(sub: str = ""):
    return importlib.import_module("autostyle-tts_b200" + (("." + sub) if sub else ""))


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()


@pytest.fixture(scope="session")
def f1():
    """The reference's shipped 130 x 6144 style database (golden fixture F1)."""
    X = np.load(os.path.join(GOLDEN, "f1_vectors_fp16.npy")).astype(np.float32)
    with open(os.path.join(GOLDEN, "f1_rows.json"), encoding="utf-8") as f:
        rows = json.load(f)
    kat = np.load(os.path.join(GOLDEN, "f1_kat.npz"))
    return {"X": X, "pks": np.asarray(rows["pks"], dtype=np.int64), "meta": rows["meta"], "kat": kat, "info": rows}


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
