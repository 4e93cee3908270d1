"""Real drop-in: the reference's UNMODIFIED Python layer on the shim modules.

`elf_b200/shim` (modules `_elf`, `_elfgames_go`, `_elfgames_go_inference`) goes on sys.path in place
of the reference's compiled pybind modules; then the reference's own `src_py/elf` package,
`src_py/elfgames/go/game.py` (`Loader.initialize()`: option parsing -> ContextOptions / GameOptions ->
`go.GameContext(co, opt)` -> `GCWrapper`) and the `GC.start(); GC.run()...; GC.stop()` pump run without
a single edit.  On this GPU-less machine the engine behind the shim is the kernel sources on the SIMT
emulator (tests/simt_emu, test infrastructure); tests/test_gpu_compat.py drives the same shim classes
on the device, where the reference tree is not available."""
import os
import sys

import numpy as np
import pytest
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
SHIM = os.path.join(os.path.dirname(HERE), "elf_b200", "shim")

pytestmark = pytest.mark.timeout(900)


@pytest.fixture
def reference_python(monkeypatch):
    if not os.path.isdir(os.path.join(REF, "src_py", "elf")):
        pytest.skip("reference tree not present")
    for p in (os.path.join(REF, "scripts", "elfgames", "go"), os.path.join(REF, "src_py"), SHIM):
        monkeypatch.syspath_prepend(p)
    drop = [m for m in sys.modules if m.split(".")[0] in ("elf", "_elf", "_elfgames_go", "_elfgames_go_inference",
                                                          "elfgames", "server_addrs")]
    for m in drop:
        monkeypatch.delitem(sys.modules, m)
    yield
    for m in [m for m in sys.modules if m.split(".")[0] in ("elf", "_elf", "_elfgames_go", "_elfgames_go_inference",
                                                            "elfgames", "server_addrs")]:
        sys.modules.pop(m, None)


def test_option_spec_and_map_behave_like_the_reference(reference_python):
    """the reference's own PyOptionSpec / PyOptionMap subclasses on the shim's _options classes"""
    from elf.options import PyOptionSpec

    spec = PyOptionSpec()
    assert spec.addIntOption("num_games", "number of games", 1024)
    assert not spec.addIntOption("num_games", "again", 1)  # emplace: existing name wins
    spec.addBoolOption("verbose", "chatty", False)
    spec.addBoolOption("use_prior", "on by default", True)
    spec.addStrListOption("list_files", "files", [])
    spec.addFloatOption("puct", "required float")
    om = spec.parse(["--puct", "1.5", "--no_use_prior", "--list_files", "a", "b"])
    assert om.get("num_games") == 1024 and om.get("puct") == 1.5 and om.get("use_prior") is False
    assert om.get("verbose") is False and om.get("list_files") == ["a", "b"]
    with pytest.raises(RuntimeError, match="has not been set"):
        om.get("nope")
    other = PyOptionSpec()
    other.addIntOption("num_games", "other default", 7)
    other.addIntOption("batchsize", "bs", 128)
    spec.merge(other)
    assert spec.parse(["--puct", "1"]).get("num_games") == 1024  # merge keeps the existing option
    spec2 = spec.clone()
    spec2.addPrefixSuffixToOptionNames("", "0")
    assert "batchsize0" in spec2.getOptionNames() and "batchsize" in spec.getOptionNames()


def test_logging_surface(reference_python):
    import elf.logging as L

    assert L.LoggerLevel.from_str("warning") == L.LoggerLevel.warn
    assert L.LoggerLevel.from_str("nope") == L.LoggerLevel.invalid
    lg = L.getIndexedLogger("elfb200-test-", "")
    lg.info("hello")
    assert lg.name().startswith("elfb200-test-") and L.get(lg.name()) is lg
    L.set_level(L.LoggerLevel.err)
    assert not lg.should_log(L.LoggerLevel.info) and lg.should_log(L.LoggerLevel.critical)
    L.set_level(L.LoggerLevel.info)


@pytest.mark.parametrize("policy_only", [False, True])
def test_unmodified_game_py_selfplay_on_the_shim(reference_python, monkeypatch, policy_only):
    """scripts/elfgames/go/selfplay.py's skeleton with the reference's game.py untouched:
    Loader(option_map).initialize() -> reg_callback(actor_black / actor_white / game_start / game_end)
    -> GC.start(); setRequest; GC.run() x N; GC.stop().  The callback replies like Evaluator.actor
    (trainer.py:73-115): pi, V, a, rv."""
    from tests import emu as E
    from tests import oracles

    try:
        E.emu_lib()
    except Exception as e:
        pytest.skip(f"SIMT emulator build unavailable: {e}")
    monkeypatch.setenv("ELFB200_BOARD", "9")
    import _elfgames_go as go
    from elf_b200.selfplay import SelfPlay

    n = 9

    def make_selfplay(**kw):
        gb = E.emu_batch(kw["num_games"], n)
        mo = {k: v for k, v in kw.items() if k in ("num_rollouts", "num_rollouts_per_batch", "virtual_loss", "persistent_tree",
                                                    "use_prior", "c_puct", "unexplored_q_zero", "root_unexplored_q_zero",
                                                    "root_epsilon", "root_alpha", "komi", "ply_pass_enabled")}
        rest = {k: v for k, v in kw.items() if k not in mo and k not in ("board_size", "device")}
        return SelfPlay(board=gb, search=E.EmuSearch(gb, **mo), board_size=n, **rest, **mo)

    monkeypatch.setattr(go, "FACTORIES", {"selfplay": make_selfplay})
    monkeypatch.setattr(go, "BOARD_SIZE", n)
    from elfgames.go.game import Loader  # the reference's file, unmodified

    spec = Loader.get_option_spec()
    argv = ["--mode", "selfplay", "--num_games", "3", "--batchsize", "8", "--mcts_threads", "1",
            "--mcts_rollout_per_thread", "12", "--mcts_rollout_per_batch", "4", "--use_mcts", "--use_mcts_ai2",
            "--mcts_use_prior", "--mcts_persistent_tree", "--mcts_puct", "1.5", "--mcts_virtual_loss", "1",
            "--policy_distri_cutoff", "4", "--resign_thres", "0.0", "--move_cutoff", "12", "--selfplay_timeout_usec", "10",
            "--no_parameter_print"]
    if policy_only:
        argv += ["--white_use_policy_network_only"]
    option_map = spec.parse(argv)
    loader = Loader(option_map)
    GC = loader.initialize()  # -> go.ContextOptions / go.GameOptions / go.GameContext(co, opt) -> GCWrapper
    assert GC.params["num_action"] == n * n + 1 and GC.params["num_planes"] == 18
    seen = {"actor_black": 0, "actor_white": 0, "game_start": 0, "game_end": 0}

    def actor(name):
        def cb(batch):
            s = batch["s"]
            assert s.shape[1:] == (18, n, n) and batch.batchsize <= 8
            seen[name] += batch.batchsize
            bs = s.shape[0]
            pi, v = oracles.feature_net(s.numpy(), n * n + 1)
            return dict(pi=torch.from_numpy(pi), V=torch.from_numpy(v), a=torch.from_numpy(pi).argmax(1),
                        rv=torch.zeros(bs, dtype=torch.int64))
        return cb

    GC.reg_callback("actor_black", actor("actor_black"))
    GC.reg_callback("actor_white", actor("actor_white"))

    def game_start(batch):
        seen["game_start"] += 1
        assert int(batch["black_ver"][0]) == 5

    def game_end(batch):
        seen["game_end"] += 1
        wr = batch.GC.getClient().getGameStats().getWinRateStats()
        assert wr.total_games == wr.black_wins + wr.white_wins

    assert GC.reg_callback_if_exists("game_start", game_start) and GC.reg_callback_if_exists("game_end", game_end)
    GC.start()
    GC.GC.getClient().setRequest(5, -1, 0.0, -1)
    for _ in range(400):
        GC.run()
        if seen["game_end"] >= 3:
            break
    GC.stop()
    assert seen["game_start"] == 1 and seen["game_end"] >= 3
    assert seen["actor_black"] > 0
    wr = GC.GC.getClient().getGameStats().getWinRateStats()
    assert wr.total_games >= 3


@pytest.mark.parametrize("kind", ["selfplay", "match"])
def test_unmodified_selfplay_script_end_to_end(tmp_path, kind):
    """(kind "match": --eval_model_pair 3,4 -- an evaluation match, two model versions loaded in game_start,
    the second AI served through the actor_white label)
    the reference's own scripts/elfgames/go/selfplay.py, unmodified, as __main__: rlpytorch's load_env
    (option parsing through elf.options on the shim's _options), its df_model3 network loaded from
    save-<ver>.bin in the game_start callback after getClient().setRequest, Evaluator.actor as the
    model callback, GC.run() until --suicide_after_n_games, GC.stop().  Subprocess: the script owns
    sys.argv and global logger state (tests/dropin_selfplay_driver.py)."""
    import subprocess

    if not os.path.isdir(os.path.join(REF, "src_py", "rlpytorch")):
        pytest.skip("reference tree not present")
    from tests import emu as E

    try:
        E.emu_lib()
    except Exception as e:
        pytest.skip(f"SIMT emulator build unavailable: {e}")
    r = subprocess.run([sys.executable, os.path.join(HERE, "dropin_selfplay_driver.py"), str(tmp_path), kind],
                       capture_output=True, text=True, timeout=600, cwd=os.path.dirname(HERE))
    out = r.stdout + r.stderr
    assert r.returncode == 0 and "DROPIN-SELFPLAY-OK" in out, out[-3000:]
    assert "In game start" in out and "Finished loading model" in out and "#suicide_after_n_games: 2, total_games: 2" in out
    assert out.count("Finished loading model") == (2 if kind == "match" else 1)


def test_unmodified_gtp_console_script_end_to_end(tmp_path):
    """the reference's own scripts/elfgames/go/df_console.py, unmodified, as __main__ (online mode): GTP
    commands on stdin -- genmove (MCTS on the engine with the reference's network through Evaluator.actor),
    play, showboard (GameContext.getGame(0).showBoard()), final_score, quit (tests/dropin_console_driver.py)"""
    import re
    import subprocess

    if not os.path.isdir(os.path.join(REF, "src_py", "rlpytorch")):
        pytest.skip("reference tree not present")
    from tests import emu as E

    try:
        E.emu_lib()
    except Exception as e:
        pytest.skip(f"SIMT emulator build unavailable: {e}")
    r = subprocess.run([sys.executable, os.path.join(HERE, "dropin_console_driver.py"), str(tmp_path)],
                       capture_output=True, text=True, timeout=600, cwd=os.path.dirname(HERE))
    out = r.stdout + r.stderr
    assert r.returncode == 0 and "DROPIN-CONSOLE-OK" in out, out[-3000:]
    moves = re.findall(r"^= ([A-HJ][1-9])\s*$", out, re.M)
    assert len(moves) == 2, out[-2000:]          # two genmove answers
    assert "Last move: " + moves[1] in out and "nextPlayer: White" in out  # showboard after B, W(E5), B
    assert re.search(r"^ 5 .*O", out, re.M)      # the human's stone at E5
