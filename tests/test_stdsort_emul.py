"""oracle/stdsort_emul.h and elf_b200/csrc/stdsort.cuh (the restated libstdc++ std::sort, which decides
the order of moves with bit-equal probabilities in MCTSActor::pi2response; the second is the product's
device code, compiled for the host here) against the real std::sort of this toolchain:
tests/cxx/stdsort_check.cc sorts the same (move, probability) pairs with both -- random arrays full of
duplicates, sizes 0..400, and adversarial inputs that drive the introsort into its heap-sort fallback."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_restated_std_sort_equals_the_real_one(tmp_path):
    exe = str(tmp_path / "stdsort_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cxx", "stdsort_check.cc")])
    for seed in (1, 2, 3):
        out = subprocess.run([exe, str(seed), "1500"], capture_output=True, text=True)
        assert out.returncode == 0, out.stdout + out.stderr
        words = out.stdout.split()
        assert words[0] == "ok" and int(words[1]) > 1500
        assert int(words[3]) > 0  # the killer inputs reached the depth limit: the heap sort is covered
