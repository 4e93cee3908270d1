"""`_elfgames_go`: the module the reference's `src_py/elfgames/go/game.py` imports as `go`
(src_cpp/elfgames/go/train/Pybind.cc:22-62), served by the B200 engine: `GameContext(co, opt)` builds
the engine the options describe (elf_b200.compat.game_context) for modes selfplay / online / train."""
import os

from elf_b200 import compat
from elf_b200.compat import ContextOptions, GameOptions, WinRateStats  # noqa: F401

BOARD_SIZE = int(os.environ.get("ELFB200_BOARD", "19"))   # BOARD9x9 is a compile-time switch in the reference
DEVICE = int(os.environ.get("ELFB200_DEVICE", "0"))
FACTORIES = None  # tests substitute engine constructors here (compat.game_context's `factories`)


class GameContext(compat.GameContext):
    """train/game_context.h:37-85 -- GameContext(ContextOptions, GameOptions)"""

    def __init__(self, co, opt):
        built = compat.game_context(co, opt, board_size=BOARD_SIZE, device=DEVICE, factories=FACTORIES)
        self.__dict__.update(built.__dict__)
        self._context_options, self._game_options = co, opt


class GameStats:
    """common/game_stats.h as seen from Python: getWinRateStats() (selfplay.py:161-166)"""

    def __init__(self, engine):
        self._engine = engine

    def getWinRateStats(self):
        return self._engine.win_stats()


class Client:
    """train/distri_client.h:318-331 -- reached through GameContext.getClient()"""


class Server:
    """train/distri_server.h -- the training server (ZMQ receive loop, model selection) is control
    plane and out of scope; GameContext.getServer() is not available on this engine."""

    def __init__(self, *a, **k):
        raise NotImplementedError("the reference's training server (distri_server.h) is out of scope of elf_b200")


class GoGameSelfPlay:
    """common/game_selfplay.h:41-56 -- reached through GameContext.getGame(i) in online mode
    (showBoard / getNextPlayer / getLastMove / getScore / getLastScore)"""
