"""TEST INFRASTRUCTURE: run bench.py's whole control flow (value leg, e2e leg, hit serving, CPU arm, JSON line) on the
CPU-only box against the engine built over the fake CUDA runtime, with a numpy stand-in for the few torch calls
bench.py makes.  The numbers it prints mean nothing (the fake's 'kernels' are the CPU oracle); what is checked is
that the script runs to its JSON line with every key of the contract - a slip in bench.py would otherwise only
show at the one GPU run at round end.  Nothing here is reachable from the product or from bench.py itself.

usage: run_bench_over_fake.py <fake libdemodel_b200.so> [bench.py args...]
"""
import os
import runpy
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import numpy as np  # noqa: E402

import demodel_b200._lib as _lib  # noqa: E402

_lib.LIB_PATH = sys.argv[1]

from fake_device import _Dev  # noqa: E402

torch = types.ModuleType("torch")
torch.uint8, torch.float64 = np.uint8, np.float64
torch.empty = lambda n, dtype=None, device=None: _Dev(np.empty(n, dtype=np.uint8))
torch.zeros = lambda n, dtype=None, device=None: _Dev(np.zeros(n, dtype=np.uint8))
torch.tensor = lambda v, dtype=None, device=None: np.asarray(v, dtype=np.float64)
torch.device = lambda *a: None
torch.cuda = types.SimpleNamespace(is_available=lambda: True, set_device=lambda d: None, synchronize=lambda: None)
torch.distributed = types.ModuleType("torch.distributed")
sys.modules["torch"] = torch
sys.modules["torch.distributed"] = torch.distributed

sys.argv = [os.path.join(ROOT, "bench.py")] + sys.argv[2:]
runpy.run_path(sys.argv[0], run_name="__main__")
