"""Native (g++, no CUDA) randomized tests of the engine's host-only data structures: the HBM arena
allocator and the interval set behind range parts.  Runs on the CPU-only box."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_arena_and_interval_set_against_brute_force_models(tmp_path):
    exe = tmp_path / "test_host_util"
    src = os.path.join(ROOT, "tests", "native", "test_host_util.cc")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-o", str(exe), src], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "host_util ok" in out.stdout
