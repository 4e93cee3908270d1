"""The differential fuzzers (scripts/emu_fuzz*.py: emulated kernels / restatement / per-move driver against
the compiled reference) stay runnable: a couple of cases each.  The campaigns themselves are summarised in
profiles/r2_fuzz_summary.md."""
import os
import subprocess
import sys

import pytest

from tests import oracles

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not oracles.have_ref(9), reason="compiled reference (oracle/_ref) not available")


def run(script, *args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", script), *args], capture_output=True, text=True,
                         timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "MISMATCH" not in out.stdout
    return out.stdout


def test_restatement_fuzz_runs():
    assert "all 8 cases equal" in run("emu_fuzz.py", "--impl", "port", "--seed", "3", "--cases", "8")
    assert "all 4 cases equal" in run("emu_fuzz.py", "--impl", "port", "--seed", "4", "--cases", "4", "--quant", "64")


def test_kernel_fuzz_runs():
    assert "all 2 cases equal" in run("emu_fuzz.py", "--seed", "3", "--cases", "2")


def test_stream_and_driver_fuzz_run():
    assert "mismatches 0" in run("emu_fuzz_streams.py", "--seed", "2", "--cases", "1")
    assert "mismatches 0" in run("emu_fuzz_driver.py", "--seed", "2", "--cases", "1")
