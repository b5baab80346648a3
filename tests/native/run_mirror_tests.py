"""TEST INFRASTRUCTURE: run the host-buffer tests of tests/test_gpu_parity.py against the engine built over the
fake CUDA runtime (tests/native/fake_cuda.cc), in a process of its own.

Purpose: the Python mirror (demodel_b200/*.py) and the LOGIC of the GPU tests are exercised on the CPU-only box
every round, so a mistake in a test or in the ctypes layer is found here and not by the one GPU run at round end.
This proves nothing about digest arithmetic (the fake's "kernels" are the CPU oracle) and is not a product path:
the library it loads is built by the test into a temporary directory and selected by patching the loader's path
inside this process only.

usage: run_mirror_tests.py <path to the fake libdemodel_b200.so>
"""
import inspect
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import demodel_b200._lib as _lib  # noqa: E402

_lib.LIB_PATH = sys.argv[1]

import demodel_b200  # noqa: E402
from tests import _oracle  # noqa: E402
import tests.test_gpu_parity as T  # noqa: E402
import tests.test_manifest as TM  # noqa: E402
import tests.test_gpu_hooks as TH  # noqa: E402

import tempfile  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from fake_device import _Torch  # noqa: E402


def main():
    oracle = _oracle.load()
    golden = os.path.join(ROOT, "tests", "golden")
    eng = demodel_b200.Engine(device=0, hbm_cas_bytes=512 << 20, ring_bytes=128 << 20, slab_bytes=1 << 20)
    ran, failed = 0, []
    def is_gpu(mod, fn):
        marks = list(getattr(fn, "pytestmark", []))
        mm = getattr(mod, "pytestmark", [])
        marks += mm if isinstance(mm, list) else [mm]
        return any(m.name == "gpu" for m in marks)

    cases = [(name, fn) for mod in (T, TM, TH) for name, fn in sorted(vars(mod).items())
             if name.startswith("test_") and callable(fn) and is_gpu(mod, fn)]
    for name, fn in cases:
        if name in ("test_baseline_sized_properties", "test_ollama_pull_at_the_fixture_sizes"):
            continue                                      # GiB-sized bodies through the scalar oracle: too slow for the CPU suite
        params = list(inspect.signature(fn).parameters)
        variants = [{}]
        for mark in getattr(fn, "pytestmark", []):
            if mark.name == "parametrize":
                arg, values = mark.args[0], mark.args[1]
                variants = [dict(v, **{arg: x}) for v in variants for x in values]
        for extra in variants:
            kw = dict(extra)
            for p in params:
                if p == "eng":
                    kw[p] = eng
                elif p == "oracle":
                    kw[p] = oracle
                elif p == "golden_dir":
                    kw[p] = golden
                elif p == "torch_cuda":
                    kw[p] = _Torch
                elif p == "tmp_path":
                    import pathlib
                    kw[p] = pathlib.Path(tempfile.mkdtemp(prefix="dm_mirror_"))
            t0 = time.time()
            try:
                fn(**kw)
                print(f"ok    {name}{extra or ''}  {time.time() - t0:.1f}s", flush=True)
            except BaseException as ex:
                if type(ex).__name__ == "Skipped":            # pytest.skip() inside the test: not applicable here
                    print(f"skip  {name}{extra or ''}: {ex}", flush=True)
                    continue
                if isinstance(ex, (KeyboardInterrupt, SystemExit)):
                    raise
                failed.append(name)
                print(f"FAIL  {name}{extra or ''}", flush=True)
                traceback.print_exc()
            ran += 1
    st = eng.stats()
    leak = not (st["open_streams"] == 0 and st["open_readers"] == 0 and st["ring_slabs_free"] == st["ring_slabs_total"])
    eng.close()
    # __graft_entry__.smoke() is what the driver runs first on the GPU box: its control flow, too, is worth a CPU pass
    import types
    import numpy as np
    from fake_device import _Dev
    stub = types.ModuleType("torch")
    stub.uint8 = np.uint8
    stub.cuda = types.SimpleNamespace(is_available=lambda: True)
    stub.zeros = lambda n, dtype=None, device=None: _Dev(np.zeros(n, dtype=np.uint8))
    had = sys.modules.get("torch")
    sys.modules["torch"] = stub
    try:
        import __graft_entry__
        __graft_entry__.smoke()
        print("ok    __graft_entry__.smoke() over the fake runtime", flush=True)
    except BaseException:
        failed.append("smoke")
        print("FAIL  __graft_entry__.smoke()", flush=True)
        traceback.print_exc()
    finally:
        if had is not None:
            sys.modules["torch"] = had
        else:
            del sys.modules["torch"]
    ran += 1
    print(f"MIRROR TESTS: {ran} ran, {len(failed)} failed, leak={int(leak)}")
    return 1 if failed or leak or ran < 10 else 0


if __name__ == "__main__":
    sys.exit(main())
