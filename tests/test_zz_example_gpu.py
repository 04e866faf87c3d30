"""The C++ integration example against the CUDA library on the B200.  (Sorted last: added after the round's GPU budget
was spent, so it has not run on a GPU yet; the same program passes against the emulator build in tests/test_example.py.)"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_cpp_example_on_the_gpu(gpu_lib, tmp_path):
    lib_dir = os.path.join(ROOT, "daliti_b200", "lib")
    exe = str(tmp_path / "replay")
    subprocess.run(["g++", "-std=c++14", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "laser_mapping_replay.cpp"),
                    "-L", lib_dir, "-ldaliti_b200", "-Wl,-rpath," + lib_dir, "-o", exe], check=True)
    r = subprocess.run([exe, "8"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "updates 7" in r.stdout
