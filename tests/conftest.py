import os
import sys
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

os.environ.setdefault("RELORA_B200_NO_WANDB", "1")
os.environ.setdefault("WANDB_MODE", "disabled")
os.environ.setdefault("TOKENIZERS_PARALLELISM", "false")

REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")
    config.addinivalue_line("markers", "multigpu: needs >= 2 CUDA devices")
    config.addinivalue_line("markers", "slow: long-running")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        have_cuda = torch.cuda.is_available()
        n_gpu = torch.cuda.device_count() if have_cuda else 0
    except Exception:
        have_cuda, n_gpu = False, 0
    skip_gpu = pytest.mark.skip(reason="no CUDA device")
    skip_multi = pytest.mark.skip(reason="needs >= 2 CUDA devices")
    for item in items:
        if "gpu" in item.keywords and not have_cuda:
            item.add_marker(skip_gpu)
        if "multigpu" in item.keywords and n_gpu < 2:
            item.add_marker(skip_multi)


@pytest.fixture(scope="session")
def reference_modules():
    """Import the upstream package (read-only mount) with a stub ``bitsandbytes``; skip if absent."""
    if not os.path.isdir(os.path.join(REFERENCE, "peft_pretraining")):
        pytest.skip("reference tree not mounted")
    if "bitsandbytes" not in sys.modules:
        bnb = types.ModuleType("bitsandbytes")
        bnb.nn = types.ModuleType("bitsandbytes.nn")
        bnb.functional = types.ModuleType("bitsandbytes.functional")
        sys.modules["bitsandbytes"] = bnb
        sys.modules["bitsandbytes.nn"] = bnb.nn
        sys.modules["bitsandbytes.functional"] = bnb.functional
    if REFERENCE not in sys.path:
        sys.path.append(REFERENCE)
    try:
        import importlib

        mods = types.SimpleNamespace(
            llama=importlib.import_module("peft_pretraining.modeling_llama"),
            relora=importlib.import_module("peft_pretraining.relora"),
            training_utils=importlib.import_module("peft_pretraining.training_utils"),
        )
    except Exception as e:  # version drift in transformers etc.
        pytest.skip(f"reference modules not importable here: {type(e).__name__}: {e}")
    return mods
