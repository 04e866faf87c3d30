"""Host-side algebra (daliti_b200/csrc/host) against textbook restatements, bit for bit: the fixed-size / AVX2-cloned 24 x 24
solves, and the IMU covariance propagation run behind the first kernel launches instead of in front of them."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_algebra_bit_identical(emu_lib, tmp_path):
    emu_dir = os.path.join(ROOT, "tests", "emu")
    exe = str(tmp_path / "host_unit")
    inc = [os.path.join(ROOT, "daliti_b200", "csrc", "host"), os.path.join(ROOT, "daliti_b200", "csrc"), os.path.join(ROOT, "include"), emu_dir]
    cmd = ["g++", "-std=c++17", "-O2", "-DDLT_EMU", *["-I" + d for d in inc], os.path.join(ROOT, "tests", "host_unit.cpp"),
           "-L", emu_dir, "-ldaliti_emu", "-Wl,-rpath," + emu_dir, "-o", exe]
    subprocess.run(cmd, check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host_unit: 0 failure(s)" in r.stdout
