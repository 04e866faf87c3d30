"""C-ABI surface: every function include/*.h declares is exported by the product library (and by the
emulator build the CPU tests use), and the product path refuses to run without a CUDA device."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in ("daliti_b200.h", "daliti_b200_lio.h", "daliti_b200_nccl.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        for m in re.finditer(r"\b(dlt_[a-z0-9_]+)\s*\(", src):
            names.add(m.group(1))
    return names


def exported(path):
    out = subprocess.run(["nm", "-D", "--defined-only", path], check=True, capture_output=True, text=True).stdout
    return {l.split()[-1] for l in out.splitlines() if l.strip()}


def test_headers_declare_something():
    names = declared_symbols()
    assert {"dlt_create", "dlt_measure", "dlt_map_incremental", "dlt_lio_process_scan"} <= names
    assert len(names) >= 40


def test_product_library_exports_every_declared_symbol():
    lib = os.path.join(ROOT, "daliti_b200", "lib", "libdaliti_b200.so")
    if not os.path.exists(lib):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "daliti_b200", "csrc")], check=True)
    missing = declared_symbols() - exported(lib)
    assert not missing, missing


def test_emulator_build_exports_every_declared_symbol(emu_lib):
    missing = declared_symbols() - exported(os.path.join(ROOT, "tests", "emu", "libdaliti_emu.so"))
    assert not missing, missing


def test_product_library_has_sm100a_kernels():
    lib = os.path.join(ROOT, "daliti_b200", "lib", "libdaliti_b200.so")
    out = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_no_cpu_fallback_without_a_device():
    """on a box without a GPU the product library must fail loudly (DLT_E_NO_DEVICE), not compute on the CPU"""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from daliti_b200.binding import DltError, ScanToMap, load_library

    lib = load_library()
    with pytest.raises(DltError, match="no usable CUDA device"):
        ScanToMap(lib)


def test_product_does_not_link_or_reference_the_oracle():
    lib = os.path.join(ROOT, "daliti_b200", "lib", "libdaliti_b200.so")
    ldd = subprocess.run(["ldd", lib], capture_output=True, text=True).stdout
    assert "oracle" not in ldd and "ikd" not in ldd
    for dirpath, _, files in os.walk(os.path.join(ROOT, "daliti_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_binding" not in txt and "liboracle" not in txt and '#include "../../oracle' not in txt, f


def test_headers_are_plain_c():
    """the drop-in boundary is a C ABI: both headers must compile as C99 and as C++14 on their own"""
    import subprocess

    for h in ("daliti_b200.h", "daliti_b200_lio.h", "daliti_b200_nccl.h"):
        src = f'#include "{os.path.join(ROOT, "include", h)}"\n'
        for args in (["gcc", "-x", "c", "-std=c99", "-pedantic"], ["g++", "-x", "c++", "-std=c++14"]):
            r = subprocess.run(args + ["-fsyntax-only", "-Wall", "-Werror", "-"], input=src, text=True, capture_output=True)
            assert r.returncode == 0, (h, args[0], r.stderr)
