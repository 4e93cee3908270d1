"""`_elfgames_go_inference`: the inference-only module (src_cpp/elfgames/go/inference/Pybind.cc:22-43:
GameContext{ctx, getParams, getGame, setRequest}) that `src_py/elfgames/go/game_inference.py` imports;
always the online engine."""
from _elfgames_go import BOARD_SIZE, DEVICE, ContextOptions, GameOptions  # noqa: F401
import _elfgames_go as _go
from elf_b200 import compat


class GameContext(compat.GameContext):
    def __init__(self, co, opt):
        opt.mode = "online"  # inference/game_context.h:31-66 only builds GoGameSelfPlay in online mode
        built = compat.game_context(co, opt, board_size=BOARD_SIZE, device=DEVICE, factories=_go.FACTORIES)
        self.__dict__.update(built.__dict__)

    def setRequest(self, black_ver, white_ver, resign_thres, num_threads=-1):
        return self.getClient().setRequest(black_ver, white_ver, resign_thres, num_threads)
