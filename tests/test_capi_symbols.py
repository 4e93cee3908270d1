"""The C-ABI library loads on a CPU-only box and exports every symbol include/elfb200.h declares;
without a CUDA device every compute entry point fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in sorted(os.listdir(os.path.join(ROOT, "include"))):
        if not h.endswith(".h"):
            continue
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(elfb200_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    from elf_b200 import lib

    L = lib.load_library()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), f"libelfb200.so does not export {s}"
    # and the Python binding table covers the header exactly
    assert sorted(lib.SIGNATURES) == syms


def test_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import elf_b200

    with pytest.raises(elf_b200.ElfB200Error):
        elf_b200.GoBatch(4, board_size=19)


def test_missing_library_raises(tmp_path):
    from elf_b200 import lib

    with pytest.raises(lib.ElfB200Error):
        lib.load_library(str(tmp_path / "nope.so"))


def test_bad_arguments_rejected():
    from elf_b200 import lib

    L = lib.load_library()
    ctx = ctypes.c_void_p()
    assert L.elfb200_create(13, 4, 0, ctypes.byref(ctx)) != 0  # only 9 and 19 are compiled
    assert b"board_size" in L.elfb200_last_error()
    assert L.elfb200_create(19, 0, 0, ctypes.byref(ctx)) != 0
    assert L.elfb200_step(None, None, None) != 0
