"""Guard on the decode kernel's register behaviour (CPU only: nvcc + cuobjdump, ~1 minute).

The persistent kernel's phase functions are __noinline__ and share the register file through ptxas'
inter-procedural allocation.  When `stage_acts` spills part of its 16-load batch, the spill store waits for
the load and serialises the staging (+1.6 us per GEMM phase, +400 us per token: DESIGN.md "Registers").
With the setmaxnreg reallocation no phase function should spill at all; this test keeps it that way."""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None or shutil.which("cuobjdump") is None, reason="needs the CUDA toolkit")
def test_decode_phase_functions_do_not_spill():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "spill_report.py")], capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    rows = {}
    for line in out.stdout.splitlines():
        m = re.match(r"(\w+)\s+n_ins\s+(\d+)\s+maxR\s+(-?\d+)\s+STL\s+(\d+)\s+LDL\s+(\d+)", line)
        if m:
            rows[m.group(1)] = tuple(int(m.group(i)) for i in (2, 3, 4, 5))
    assert {"gemm_phase", "stage_acts", "attn_item", "kernel"} <= set(rows), out.stdout
    for fn in ("gemm_phase", "stage_acts", "attn_item", "attn_scores"):
        n_ins, max_r, stl, ldl = rows[fn]
        assert stl == 0 and ldl == 0, f"{fn} spills (STL {stl}, LDL {ldl}):\n{out.stdout}"
    # consumer code really uses the raised budget (the launch bound alone would cap it at 168)
    assert rows["stage_acts"][1] > 168 or rows["gemm_phase"][1] > 168
    # the producer warp stays inside what setmaxnreg.dec leaves it
    assert rows["producer_loop"][1] < 40
    assert rows["kernel"][2] <= 12, out.stdout
