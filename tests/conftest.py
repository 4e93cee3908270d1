import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_cuda():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device: on a machine without one a plain `pytest` skips them instead of
    failing in elfb200_create (the product has no CPU fallback, so they cannot run).  ELFB200_TEST_EMU=1
    (below) redirects board/search tests to the SIMT emulator instead."""
    if os.environ.get("ELFB200_TEST_EMU") == "1" or _have_cuda():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200 box); none visible here")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_lib():
    from tests import oracles

    return oracles.load_oracle()


# Developer aid: ELFB200_TEST_EMU=1 runs tests written for the GPU on the SIMT emulator build of the
# kernel sources (tests/simt_emu) -- `ELFB200_TEST_EMU=1 pytest tests/test_gpu_board.py -m gpu -k ...`.
# Only the board batch and the search are redirected; tests that need a real network on the device
# still need the device.  Never active unless the variable is set; never used by the product.
if os.environ.get("ELFB200_TEST_EMU") == "1":
    import elf_b200
    from tests import emu as _emu

    def _emu_gobatch(num_games, board_size=19, device=0):
        return _emu.emu_batch(num_games, board_size)

    elf_b200.GoBatch = _emu_gobatch
    elf_b200.MctsBatch = _emu.EmuSearch
    import elf_b200.board
    import elf_b200.mcts
    import elf_b200.selfplay

    elf_b200.board.GoBatch = _emu_gobatch
    elf_b200.mcts.MctsBatch = _emu.EmuSearch
    elf_b200.selfplay.GoBatch = _emu_gobatch
    elf_b200.selfplay.MctsBatch = _emu.EmuSearch
