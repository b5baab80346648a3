"""The C-ABI entered the way cgo enters it (tests/native/cgo_shape.c): plain C99, every call on a short-lived
foreign thread, consecutive calls on one stream from different threads, error text fetched by id on another
thread.  On the CPU box against the fake-runtime build of the engine (host logic, threading, header usable from
C); under `-m gpu` against the real libdemodel_b200.so on the device."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "native", "cgo_shape.c")


def _build_harness(exe, libdir, libname):
    cmd = ["gcc", "-std=c99", "-O1", "-g", "-Wall", "-Wextra", "-pthread", SRC, "-o", str(exe),
           "-L", str(libdir), f"-l:{libname}", f"-Wl,-rpath,{libdir}"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-3000:]


def test_cgo_shaped_entry_over_the_fake_runtime(tmp_path):
    cs = os.path.join(ROOT, "demodel_b200", "csrc")
    lib = tmp_path / "libdemodel_b200_fake.so"
    build = subprocess.run(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-pthread",
                            "-I", os.path.join(ROOT, "tests", "native", "fake_cuda"), "-o", str(lib), "-x", "c++",
                            *[os.path.join(cs, f) for f in ("engine_core.cu", "engine_api.cu", "engine_cache.cu", "proxy_driver.cc",
                                                            "manifest.cc", "gunzip.cc")],
                            os.path.join(ROOT, "tests", "native", "fake_cuda.cc")],      # (its "kernels" are the CPU oracle)
                           capture_output=True, text=True, timeout=600)
    assert build.returncode == 0, build.stderr[-3000:]
    exe = tmp_path / "cgo_shape"
    _build_harness(exe, tmp_path, lib.name)
    for args in ([], [str(tmp_path / "cas")]):
        out = subprocess.run([str(exe), *args], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0 and "CGO SHAPE OK" in out.stdout, (out.stdout + out.stderr)[-3000:]


@pytest.mark.gpu
def test_cgo_shaped_entry_on_the_gpu(tmp_path):
    libdir = os.path.join(ROOT, "demodel_b200")
    assert os.path.exists(os.path.join(libdir, "libdemodel_b200.so")), "build the engine first (no CPU fallback)"
    exe = tmp_path / "cgo_shape"
    _build_harness(exe, libdir, "libdemodel_b200.so")
    for args in ([], [str(tmp_path / "cas")]):
        out = subprocess.run([str(exe), *args], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0 and "CGO SHAPE OK" in out.stdout, (out.stdout + out.stderr)[-3000:]
