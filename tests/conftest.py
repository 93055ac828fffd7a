import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")
    # a fresh checkout has no built library (the .so is git-ignored): build it once, in-tree, so that the suite does not depend on
    # `__graft_entry__.build()` having been called first (nvcc cross-compiles sm_100a without a GPU, ~30 s)
    lib = os.path.join(ROOT, "llm_mixed_q_b200", "libbq_b200.so")
    if not os.path.exists(lib):
        import shutil
        import subprocess

        if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
            subprocess.run(["bash", os.path.join(ROOT, "llm_mixed_q_b200", "csrc", "build.sh")], check=True,
                           stdout=subprocess.DEVNULL)


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_quantizers():
    arrays = np.load(os.path.join(GOLD, "quantizers.npz"))
    with open(os.path.join(GOLD, "manifest.json")) as f:
        manifest = json.load(f)
    return arrays, manifest["cases"]


@pytest.fixture(scope="session")
def golden_consumers():
    arrays = np.load(os.path.join(GOLD, "consumers.npz"))
    with open(os.path.join(GOLD, "consumers.json")) as f:
        return arrays, json.load(f)


@pytest.fixture(scope="session")
def golden_configs():
    with open(os.path.join(GOLD, "configs.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_hashed():
    with open(os.path.join(GOLD, "hashed.json")) as f:
        return json.load(f)


def f32(bits: np.ndarray) -> torch.Tensor:
    """int32 bit patterns -> fp32 tensor (keeps NaN payloads and -0.0)."""
    return torch.from_numpy(np.ascontiguousarray(bits)).view(torch.float32)


def case_input(arrays, case) -> torch.Tensor:
    x = f32(arrays[case["x"]])
    if case.get("transpose"):
        x = x.transpose(*case["transpose"])
    return x


def case_kwargs(case):
    return {k: (None if v == "NA->None" else v) for k, v in case["kwargs"].items()}


def n_bits_diff(a: torch.Tensor, b: torch.Tensor) -> int:
    """number of elements whose BIT PATTERNS differ (so -0.0 != +0.0).  NaN payloads are not compared: x86 produces
    the negative "real indefinite" quiet NaN (0xffc00000) for inf*0, the GPU its canonical 0x7fffffff."""
    a, b = a.contiguous(), b.contiguous()
    diff = a.view(torch.int32) != b.view(torch.int32)
    both_nan = torch.isnan(a) & torch.isnan(b)
    return int((diff & ~both_nan).sum())


def bits_equal(a: torch.Tensor, b: torch.Tensor) -> bool:
    return a.shape == b.shape and n_bits_diff(a, b) == 0
