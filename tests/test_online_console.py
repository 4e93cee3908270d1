"""Online mode: OnlineGame state machine, GTP console, board picture, and the
UNMODIFIED reference console (scripts/elfgames/go/console_lib.py) + GCWrapper running on
elf_b200.compat.OnlineEngine.

CPU-only: the board is the C restatement and the "search" is a one-wave stub (root evaluation,
arg-max of the network policy over legal moves), injected through the duck-typed board/search
arguments of OnlineGame -- the host logic under test is the same code the GPU engine runs
(tests/test_zz_gpu_online.py drives it with the real GoBatch/MctsBatch)."""
import builtins
import importlib.util
import io
import os

import numpy as np
import pytest
import torch

from elf_b200 import compat, console, online, sgf
from tests import oracles


class StubBoard:
    """GoBatch interface for one game over oracle/go_oracle.c"""

    num_games = 1

    def __init__(self, n, lib):
        self.board_size, self.lib = n, lib
        self.o = oracles.Oracle(n, lib)

    def forward(self, actions):
        return np.array([self.o.forward(int(actions[0]))])

    def info(self):
        return np.asarray(self.o.info(), np.int32)[None]

    def stones(self):
        return self.o.stones()[None]

    def evaluate(self, komi):
        return np.array([self.o.evaluate(komi)], np.float32)

    def features(self, d4=None):
        return self.o.features(0)[None].astype(np.float32)

    def reset(self, mask):
        self.o = oracles.Oracle(self.board_size, self.lib)

    def synchronize(self):
        pass


class StubSearch:
    """one wave, one leaf (the root): move = arg-max of the replied policy over legal moves; the
    resign rule is the device one (k_choose): side-to-move value < -1 + thres and ply >= 50"""

    waves_per_move = 1

    def __init__(self, board):
        self.b = board
        self.advanced, self.resets, self.evals = [], 0, 0

    def begin_move(self, active):
        self.pi = self.v = None

    def select(self):
        return torch.from_numpy(self.b.features())

    def expand_backup(self, pi, v):
        self.pi, self.v = pi[0].numpy().copy(), float(v[0])
        self.evals += 1

    def choose(self, cutoff, thres, never_resign, seed):
        n = self.b.board_size
        info = self.b.info()[0]
        side = self.v if info[1] == 1 else -self.v
        if side < -1.0 + thres and info[0] >= 50:
            return np.array([-1], np.int32), np.array([self.v], np.float32)
        legal = np.append(self.b.o.legal().astype(bool), True)
        p = np.where(legal, self.pi, -1.0)
        return np.array([int(p.argmax())], np.int32), np.array([self.v], np.float32)

    def advance(self, a):
        self.advanced.append(int(a[0]))

    def reset(self, mask):
        self.resets += 1


def policy_actor(prefer, value=0.0, n=9):
    """network stub: puts its mass on the first still-available action of `prefer`"""
    calls = []

    def actor(batch):
        s = batch["s"]
        assert s.shape[1:] == (18, n, n)
        k = s.shape[0]
        pi = torch.full((k, n * n + 1), 1e-4)
        for i, a in enumerate(prefer):
            pi[:, a] = 1.0 - 0.01 * i
        calls.append(k)
        return {"pi": pi, "V": torch.full((k,), float(value))}

    actor.calls = calls
    return actor


def make_game(oracle_lib, n=9, **kw):
    b = StubBoard(n, oracle_lib)
    return online.OnlineGame(b, StubSearch(b), **kw)


# ---- coordinates and the board picture, pinned on the compiled reference ------------------------
def test_gtp_vertices():
    assert online.move2xy("A1") == (0, 0) and online.move2xy("j9") == (8, 8) and online.move2xy("H8") == (7, 7)
    assert online.move2xy("pass") == (-1, -1) and online.xy2move(-1, -1) == "pass"
    assert online.xy2move(8, 0) == "J1" and online.xy2move(7, 18) == "H19" and online.xy2move(18, 18) == "T19"
    for n in (9, 19):
        for a in range(n * n):
            assert online.vertex2action(online.action2vertex(a, n), n) == a
    with pytest.raises(ValueError):
        online.vertex2action("K10", 9)


@pytest.mark.parametrize("n", [9, 19])
def test_vertex_and_board_picture_match_reference(n):
    if not oracles.have_ref(n):
        pytest.skip("oracle/_ref not built")
    for a in range(n * n + 1):
        assert online.action2vertex(a, n) == oracles.ref_vertex_str(n, a)
    rng = np.random.default_rng(n)
    r = oracles.Ref(n)
    checked = 0
    for t in range(160):
        i = r.info()
        pic = online.show_board(r.stones(), n, int(i[4]), int(i[2]), int(i[3]), int(i[1]))
        assert pic == oracles.ref_show_board(r), f"ply {t}\n{pic}\n---\n{oracles.ref_show_board(r)}"
        checked += 1
        legal = np.flatnonzero(r.legal())
        a = n * n if (len(legal) == 0 or t in (7, 30)) else int(rng.choice(legal))
        if not r.forward(a) or r.terminated():
            break
    assert checked > 100 and (i[2] + i[3]) > 0  # captures happened, the counters were on the picture


# ---- OnlineGame -----------------------------------------------------------------------------------
def test_human_branch(oracle_lib):
    g = make_game(oracle_lib)
    assert g.getNextPlayer() == "B" and g.getLastScore() == 0.0
    assert g.human(online.vertex2action("E5", 9)) == online.MOVED
    assert g.getNextPlayer() == "W" and g.getLastMove() == "E5"
    assert g.human(online.vertex2action("E5", 9)) == online.INVALID  # occupied
    assert g.human(12345) == online.INVALID
    assert g.human(online.SA_SKIP) == online.SKIP and g.getNextPlayer() == "W"  # nothing moved
    assert g.search.advanced == [online.vertex2action("E5", 9)]
    assert g.human(online.SA_PASS) == online.MOVED and g.getLastMove() == "PASS"
    # second pass ends the game: value = evaluate(komi) of the final position, then restart
    assert g.human(81) == online.FINISHED
    assert g.finished == [(pytest.approx(81 - 7.5), 4, "two_pass")]  # one black stone owns the whole board
    assert g.getLastScore() == pytest.approx(73.5) and g.seq == 1 and g.search.resets == 1
    assert g.info()[0] == 1 and g.getLastMove() == "PASS"  # lastMove() of a fresh game = last move of the finished one
    # clear on a fresh game does nothing; after a move it finishes the game
    assert g.human(online.SA_CLEAR) == online.CLEARED and g.seq == 1
    g.human(0)
    assert g.human(online.SA_CLEAR) == online.CLEARED and g.seq == 2 and g.finished[-1][2] == "clear"
    # resign: the side to move loses
    g.human(0)
    assert g.human(online.SA_RESIGN) == online.RESIGNED
    assert g.getLastScore() == 1.0 and g.getLastMove() == "RESIGN"  # white to move resigned
    assert g.human(online.SA_RESIGN) == online.RESIGNED and g.getLastScore() == -1.0


def test_ai_branch(oracle_lib):
    g = make_game(oracle_lib, resign_thres=0.1)
    e5, c3 = online.vertex2action("E5", 9), online.vertex2action("C3", 9)
    actor = policy_actor([e5, c3])
    assert g.genmove(actor) == e5 and g.genmove(actor) == c3  # e5 taken -> next preference
    assert actor.calls == [1, 1] and g.search.advanced == [e5, c3] and g.info()[0] == 3
    # resign needs ply >= 50
    losing = policy_actor([e5, c3, 0, 1, 2], value=-0.99)
    assert g.genmove(losing) == 0  # black to move, value says black loses, but ply < 50
    while g.info()[0] < 50 or g.getNextPlayer() != "W":
        assert g.human(int(np.flatnonzero(g.board.o.legal())[0])) == online.MOVED
    assert g.info()[0] >= 50 and g.getNextPlayer() == "W" and g.seq == 0
    assert g.genmove(losing) != online.SA_RESIGN  # white to move: -0.99 is good for white
    assert g.genmove(losing) == online.SA_RESIGN and g.finished[-1][2] == "resign"
    assert g.getLastScore() == -1.0 and g.getLastMove() == "RESIGN"


def test_following_pass(oracle_lib):
    e5 = online.vertex2action("E5", 9)
    for follow, v, expect_pass in ((True, 0.95, True), (True, 0.5, False), (False, 0.95, False)):
        g = make_game(oracle_lib, following_pass=follow)
        g.human(e5)  # black stone: black is ahead on the board
        g.human(81)  # white passes
        a = g.genmove(policy_actor([0, 1], value=v))
        assert (a == 81) == expect_pass
        if expect_pass:
            assert g.finished[-1][2] == "two_pass" and g.getLastScore() == pytest.approx(73.5)


def test_move_cutoff_and_terminal_position(oracle_lib):
    g = make_game(oracle_lib, move_cutoff=4)
    actor = policy_actor(list(range(30)))
    assert [g.genmove(actor) for _ in range(3)] == [0, 1, 2]
    assert g.seq == 1 and g.finished[-1][1:] == (4, "max_step")  # ply reached move_cutoff -> finish_game(FR_MAX_STEP)


def test_sgf_preload_and_follow(oracle_lib, tmp_path):
    from tests.test_sgf import SGF_MAKE

    f = tmp_path / "game.sgf"
    f.write_text(SGF_MAKE)
    rec = sgf.Sgf.loads(SGF_MAKE, 9)
    g = make_game(oracle_lib, preload_sgf=str(f), preload_sgf_move_to=10)
    assert g.info()[0] == 11 and g.search.advanced == rec.actions()[:10]
    actor = policy_actor([80, 79, 78])
    # the search runs, but the move comes from the record (game_selfplay.cc:387-400)
    got = [g.genmove(actor) for _ in range(52)]
    assert got == rec.actions()[10:62] and len(actor.calls) == 52
    assert g.genmove(actor) is None and g.finished[-1][2] == "max_step"  # record exhausted
    # finish_game -> _state_ext.restart(): an EMPTY board; the preload is not repeated and the SGF
    # iterator stays exhausted (both belong to GoGameSelfPlay::restart, game_selfplay.cc:202-219)
    assert g.info()[0] == 1 and g._sgf_pos == len(rec.actions())
    with pytest.raises(RuntimeError):
        bad = tmp_path / "bad.sgf"
        bad.write_text("(;SZ[9];B[aa];W[aa])")
        make_game(oracle_lib, preload_sgf=str(bad), preload_sgf_move_to=5)


# ---- GTP ------------------------------------------------------------------------------------------
def test_gtp_session(oracle_lib):
    g = make_game(oracle_lib)
    e5, d4 = online.vertex2action("E5", 9), online.vertex2action("D4", 9)
    c = console.GtpConsole(g, policy_actor([e5, d4]))
    out = io.StringIO()
    script = "\n".join([
        "protocol_version", "name", "version", "boardsize 9", "boardsize 19", "komi 7.5", "komi 6.5", "clear_board",
        "# a comment", "", "7 play B E5", "play B D4", "play W E5", "genmove b", "genmove W", "known_command genmove",
        "known_command undo", "frobnicate", "play B", "showboard", "play b pass", "play w pass", "final_score", "quit",
        "name"]) + "\n"
    c.run(io.StringIO(script), out)
    r = out.getvalue().split("\n\n")
    assert r[0] == "= 2" and r[1] == "= DF2" and r[2] == "= 1.0" and r[3] == "="
    assert r[4] == "? We only support 9x9 board for now" and r[5] == "=" and r[6].startswith("? We only support 7.5")
    assert r[7] == "=" and r[8] == "=7"
    assert r[9].startswith("? Specified next player B is not the same as the next player W")
    assert r[10] == "? illegal move"
    assert r[11].startswith("? Specified next player b")
    assert r[12] == "= D4"  # E5 is taken, the stub's second preference
    assert r[13] == "= true" and r[14] == "= false" and r[15] == "? unknown command" and r[16].startswith("? Invalid command")
    assert " X)" not in r[17] and " O)" in r[17] and "Last move: D4, nextPlayer: Black" in r[17]
    assert r[18] == "=" and r[19] == "=" and r[20] == "= W+7.5"  # 1 black, 1 white stone, komi 7.5
    assert r[21] == "=" and len(r) == 23 and c.exit  # nothing answered after quit
    assert sorted(c.commands) == sorted(
        ["protocol_version", "name", "version", "list_commands", "known_command", "boardsize", "komi", "clear_board", "play",
         "genmove", "showboard", "final_score", "quit", "exit"])


# ---- the reference's own console on the compat surface ----------------------------------------------
REF = "/root/reference"


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
def test_reference_gtp_console_runs_on_online_engine(oracle_lib, monkeypatch, capsys):
    ref_utils = _load(REF + "/src_py/elf/utils_elf.py", "ref_utils_elf_online")
    ref_console = _load(REF + "/scripts/elfgames/go/console_lib.py", "ref_console_lib")
    g = make_game(oracle_lib)
    eng = compat.OnlineEngine(g)
    GC = compat.GameContext(eng, batchsize=4)
    # the reference's desc for mode == "online" (src_py/elfgames/go/game.py:363-374)
    desc = {"human_actor": dict(input=["s"], reply=["pi", "a", "V"], batchsize=1),
            "actor_black": dict(input=["s"], reply=["pi", "V", "a", "rv"], timeout_usec=10, batchsize=4)}
    gcw = ref_utils.GCWrapper(GC, 4, desc, num_recv=2, gpu=None, use_numpy=False, params=GC.getParams())
    e5, d4 = online.vertex2action("E5", 9), online.vertex2action("D4", 9)
    net = policy_actor([e5, d4])

    class Evaluator:  # what df_console.py passes in (rlpytorch Evaluator): only .actor is used
        def actor(self, batch):
            r = net(batch)
            k = batch["s"].shape[0]
            return dict(pi=r["pi"], V=r["V"], a=torch.zeros(k, dtype=torch.int64), rv=torch.zeros(k, dtype=torch.int64))

    con = ref_console.GoConsoleGTP(gcw, Evaluator())
    gcw.reg_callback_if_exists("actor_black", con.actor)
    gcw.reg_callback_if_exists("human_actor", lambda batch: con.prompt("", batch))
    script = iter(["name", "boardsize 9", "play B E5", "genmove B", "genmove W", "showboard", "final_score",
                   "play B pass", "play W pass", "final_score", "clear_board", "quit"])
    monkeypatch.setattr(builtins, "input", lambda prompt="": next(script))
    gcw.start()
    GC.getClient().setRequest(0, -1, 0.05, -1)  # df_console.py:71-72
    for _ in range(40):
        gcw.run()
        if con.exit:
            break
    gcw.stop()
    assert con.exit and g.resign_thres == pytest.approx(0.05)
    out = capsys.readouterr().out
    replies = [ln for ln in out.split("\n") if ln.startswith(("=", "?"))]
    assert replies[0] == "= DF2"
    assert replies[3].startswith("? Specified next player B")  # genmove B while white is to move
    assert "= D4" in replies  # the AI's move, reported at the next prompt via getLastMove()
    assert "Last move: D4, nextPlayer: Black" in out and "O)" in out  # showboard through GC.getGame(0)
    assert "= W+7.5" in replies  # final_score after the two passes
    assert g.finished[0][2] == "two_pass" and eng.win_stats().total_games == len(g.finished)
    assert net.calls == [1]  # exactly one network round trip (the stub search has one leaf)
