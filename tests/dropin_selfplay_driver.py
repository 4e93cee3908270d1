"""Driver for tests/test_dropin_shim.py::test_unmodified_selfplay_script_end_to_end (run as a subprocess).

Puts elf_b200/shim in front of the reference's Python tree and executes the reference's UNMODIFIED
scripts/elfgames/go/selfplay.py (__main__): rlpytorch option parsing, its own df_model3 network loaded
from a save file in the game_start callback, Evaluator.actor as the model callback, GC.run() until
--suicide_after_n_games.  No GPU here, so the engine behind the shim is the kernel sources on the SIMT
emulator (test infrastructure); on a B200 the same command line runs on libelfb200.so."""
import os
import runpy
import sys

import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
work = sys.argv[1]
match = len(sys.argv) > 2 and sys.argv[2] == "match"  # evaluation match: two model versions, actor_white
sys.path[:0] = [os.path.join(ROOT, "elf_b200", "shim"), os.path.join(REF, "src_py"),
                os.path.join(REF, "scripts", "elfgames", "go"), ROOT]
os.environ.update(game="elfgames.go.game", model="df_pred", model_file="elfgames.go.df_model3", ELFB200_BOARD="9", root=work)

from elf_b200.model import PolicyValueNet  # noqa: E402
from elf_b200.selfplay import SelfPlay  # noqa: E402
from tests import emu as E  # noqa: E402

n = 9
# a model file in the reference's format (rlpytorch/model_base.py:83-108): save-<version>.bin under $root
torch.manual_seed(0)
sd = {("resnet.resnet." + k[len("resnet."):] if k.startswith("resnet.") else k): v
      for k, v in PolicyValueNet(n, num_block=1, dim=8).state_dict().items()}
torch.save({"state_dict": sd, "step": 0, "options": {}}, os.path.join(work, "save-3.bin"))
if match:
    torch.manual_seed(1)
    sd4 = {("resnet.resnet." + k[len("resnet."):] if k.startswith("resnet.") else k): v
           for k, v in PolicyValueNet(n, num_block=1, dim=8).state_dict().items()}
    torch.save({"state_dict": sd4, "step": 0, "options": {}}, os.path.join(work, "save-4.bin"))

E.emu_lib()
import _elfgames_go as go  # noqa: E402


def make_selfplay(**kw):
    gb = E.emu_batch(kw["num_games"], n)
    keys = ("num_rollouts", "num_rollouts_per_batch", "virtual_loss", "persistent_tree", "use_prior", "c_puct",
            "unexplored_q_zero", "root_unexplored_q_zero", "root_epsilon", "root_alpha", "komi", "ply_pass_enabled")
    mo = {k: v for k, v in kw.items() if k in keys}
    rest = {k: v for k, v in kw.items() if k not in mo and k not in ("board_size", "device")}
    white = E.EmuSearch(gb, **mo) if match else None  # the second AI's tree (GoGameSelfPlay::_ai2)
    return SelfPlay(board=gb, search=E.EmuSearch(gb, **mo), search_white=white, board_size=n, **rest, **mo)


go.FACTORIES = {"selfplay": make_selfplay}
go.BOARD_SIZE = n
sys.argv = ["selfplay.py", "--mode", "selfplay", "--num_games", "2", "--batchsize", "8", "--mcts_threads", "1",
            "--mcts_rollout_per_thread", "8", "--mcts_rollout_per_batch", "4", "--use_mcts", "--use_mcts_ai2",
            "--mcts_use_prior", "--mcts_persistent_tree", "--mcts_puct", "1.5", "--mcts_virtual_loss", "1",
            "--policy_distri_cutoff", "4", "--resign_thres", "0.0", "--move_cutoff", "8", "--selfplay_timeout_usec", "10",
            "--num_block0", "1", "--dim0", "8", "--num_block1", "1", "--dim1", "8", "--keys_in_reply", "V", "rv",
            "--gpu", "-1", "--suicide_after_n_games", "2", "--no_check_loaded_options0", "--no_check_loaded_options1",
            "--eval_model_pair", "3,4" if match else "3,-1"]
runpy.run_path(os.path.join(REF, "scripts", "elfgames", "go", "selfplay.py"), run_name="__main__")
print("DROPIN-SELFPLAY-OK")
