"""Native (g++, no CUDA) randomized tests of the engine's host-only data structures: the HBM arena
allocator and the interval set behind range parts.  Runs on the CPU-only box."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_arena_and_interval_set_against_brute_force_models(tmp_path):
    exe = tmp_path / "test_host_util"
    src = os.path.join(ROOT, "tests", "native", "test_host_util.cc")
    flags = ["-fsanitize=address,undefined", "-fno-sanitize-recover=all"]
    if subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-Wall", "-Wextra", *flags, "-o", str(exe), src],
                      capture_output=True).returncode != 0:          # no sanitizer runtime: plain build
        subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-o", str(exe), src], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "host_util ok" in out.stdout


def test_manifest_parser_under_asan_ubsan(tmp_path):
    """200 000 mutated / random / truncated inputs through dm_manifest_parse built with
    -fsanitize=address,undefined: no over-read, no UB, only DM_OK or DM_EINVAL."""
    import pytest
    exe = tmp_path / "fuzz_manifest"
    src = os.path.join(ROOT, "tests", "native", "fuzz_manifest.cc")
    build = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
                            "-o", str(exe), src], capture_output=True, text=True)
    if build.returncode != 0 and ("asan" in build.stderr.lower() or "sanitize" in build.stderr.lower()):
        pytest.skip("this g++ has no sanitizer runtime")
    assert build.returncode == 0, build.stderr[-2000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, (out.stdout + out.stderr)[-3000:]
    assert "manifest fuzz ok" in out.stdout
