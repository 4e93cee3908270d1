"""Driver for tests/test_dropin_shim.py::test_unmodified_gtp_console_script_end_to_end (run as a subprocess).

The reference's UNMODIFIED scripts/elfgames/go/df_console.py (__main__) on elf_b200/shim: rlpytorch load_env,
its df_model3 network loaded from a save file, console_lib.GoConsoleGTP on the human_actor / actor_black
labels of the online mode, GTP commands from stdin.  Engine behind the shim: the kernel sources on the SIMT
emulator (no GPU here; test infrastructure)."""
import io
import os
import runpy
import sys

import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
work = sys.argv[1]
sys.path[:0] = [os.path.join(ROOT, "elf_b200", "shim"), os.path.join(REF, "src_py"),
                os.path.join(REF, "scripts", "elfgames", "go"), ROOT]
os.environ.update(game="elfgames.go.game", model="df_pred", model_file="elfgames.go.df_model3", ELFB200_BOARD="9")

from elf_b200 import lib as _l  # noqa: E402
from elf_b200.model import PolicyValueNet  # noqa: E402
from elf_b200.online import OnlineGame  # noqa: E402
from tests import emu as E  # noqa: E402

n = 9
torch.manual_seed(0)
sd = {("resnet.resnet." + k[len("resnet."):] if k.startswith("resnet.") else k): v
      for k, v in PolicyValueNet(n, num_block=1, dim=8).state_dict().items()}
model_file = os.path.join(work, "save-1.bin")
torch.save({"state_dict": sd, "step": 0, "options": {}}, model_file)

E.emu_lib()
import _elfgames_go as go  # noqa: E402


def make_online(**kw):
    mcts_fields = {f[0] for f in _l.MctsOptions._fields_}
    mo = {k: kw.pop(k) for k in list(kw) if k in mcts_fields}
    kw.pop("board_size", None)
    kw.pop("device", None)
    gb = E.emu_batch(1, n)
    return OnlineGame(gb, E.EmuSearch(gb, **mo), **kw)


go.FACTORIES = {"online": make_online}
go.BOARD_SIZE = n
sys.argv = ["df_console.py", "--mode", "online", "--keys_in_reply", "V", "rv", "--use_mcts", "--mcts_verbose_time",
            "--mcts_use_prior", "--mcts_persistent_tree", "--load", model_file, "--gpu", "-1", "--num_block", "1", "--dim", "8",
            "--mcts_threads", "1", "--mcts_rollout_per_thread", "16", "--mcts_rollout_per_batch", "4", "--resign_thres", "0.0",
            "--mcts_virtual_loss", "1", "--mcts_puct", "1.5", "--no_check_loaded_options", "--batchsize", "4"]
sys.stdin = io.StringIO("boardsize 9\nclear_board\ngenmove b\nplay w E5\ngenmove b\nshowboard\nfinal_score\nquit\n")
runpy.run_path(os.path.join(REF, "scripts", "elfgames", "go", "df_console.py"), run_name="__main__")
print("DROPIN-CONSOLE-OK")
