"""Test configuration.

`-m "not gpu"`: the oracle against the reference build / properties, the host logic, the
C-ABI export list, and the kernels' logic under tests/emu (the CUDA sources compiled as C++
against a coroutine emulator -- test infrastructure, not a product path).
`-m gpu`: the parity tests proper, through the nvcc-built C ABI on a real B200.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: larger sizes")


def _make(directory, target=None):
    cmd = ["make", "-s", "-C", directory]
    if target:
        cmd.append(target)
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def oracle():
    import oracle_binding

    return oracle_binding.load()


@pytest.fixture(scope="session")
def emu_lib():
    """The product sources under the kernel-logic emulator (CPU)."""
    from daliti_b200.binding import load_library

    d = os.path.join(ROOT, "tests", "emu")
    _make(d)
    return load_library(os.path.join(d, "libdaliti_emu.so"))


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library (nvcc, sm_100a).  Fails loudly when missing: no fallback."""
    import torch

    from daliti_b200.binding import load_library

    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return load_library()


def backend_params():
    """Every device-path test runs twice: on the kernel-logic emulator here (small sizes), and on the GPU box."""
    return [pytest.param("emu", id="emu"), pytest.param("gpu", id="gpu", marks=pytest.mark.gpu)]


@pytest.fixture(params=backend_params())
def dev(request):
    """(library, is_gpu) for the requested backend."""
    if request.param == "emu":
        return request.getfixturevalue("emu_lib"), False
    return request.getfixturevalue("gpu_lib"), True
