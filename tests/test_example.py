"""examples/laser_mapping_replay.cpp -- the reference-side binding of INTEGRATION.md as a C++ program -- builds against
the C headers alone and follows a synthetic trajectory.  Here: linked against the kernel-logic emulator build (CPU);
the CUDA build of the same program is produced by __graft_entry__.build() and run by tests/test_zz_example_gpu.py."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_example_against_the_emulator(emu_lib, tmp_path):
    emu_dir = os.path.join(ROOT, "tests", "emu")
    exe = str(tmp_path / "replay_emu")
    subprocess.run(["g++", "-std=c++14", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "laser_mapping_replay.cpp"), "-L", emu_dir, "-ldaliti_emu", "-Wl,-rpath," + emu_dir, "-o", exe],
                   check=True)
    r = subprocess.run([exe, "6"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("scan ")]
    assert len(lines) == 6 and "iters 0" in lines[0]  # the first scan builds the map (laserMapping.cpp:780-793)
    assert all("ekf_stop 0" in ln for ln in lines)
    assert "updates 5" in r.stdout
