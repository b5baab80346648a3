"""Candidate kernel orderings that are built and selectable (DM_KERNEL_VARIANT) but not the default.

They run in a process of their own and are recorded as xfail/xpass: a candidate can never turn the suite red
or disturb the CUDA context the parity tests use.  What gates is test_gpu_parity.py, on the defaults.
The arithmetic of every round form is proven on the host by tests/test_native_host.py.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="candidate round ordering (deep variant 4): recorded, not gating, until measured")
def test_short_chain_round_is_bit_exact_on_the_gpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "check_variant.py"), "9,4"],
                         capture_output=True, text=True, timeout=240)
    print(out.stdout[-3000:])
    assert out.returncode == 0 and "VARIANT OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
