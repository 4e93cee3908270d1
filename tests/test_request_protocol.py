"""Server requests (MsgRequest) on the self-play engine: model-version switching, waiting, async
updates, thread limits and player_swap -- GoGameSelfPlay::OnReceive / restart
(common/game_selfplay.cc:159-270) and DispatcherCallback (common/dispatcher_callback.h:27-99) for a
batch of games -- plus the notifications the reference's selfplay.py sees (game_start once per
actionable request, game_end once per finished game).

CPU-only: boards are the C restatement, the search is a one-wave stub (root evaluation, arg-max of
the policy over legal moves), injected through SelfPlay(board=, search=, search_white=)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from elf_b200 import compat
from elf_b200.selfplay import SelfPlay
from tests import oracles

N = 9
P1 = N * N + 1


class Boards:
    """GoBatch interface over G independent oracle games"""

    def __init__(self, G, lib, n=N):
        self.num_games, self.board_size, self.lib = G, n, lib
        self.o = [oracles.Oracle(n, lib) for _ in range(G)]
        self.resets = np.zeros(G, int)

    def forward(self, actions):
        return np.array([True if a < 0 else bool(o.forward(int(a))) for o, a in zip(self.o, actions)])

    def info(self):
        return np.stack([np.asarray(o.info(), np.int32) for o in self.o])

    def evaluate(self, komi):
        return np.array([o.evaluate(komi) for o in self.o], np.float32)

    def features(self, d4=None):
        return np.stack([o.features(0 if d4 is None else int(d4[g])) for g, o in enumerate(self.o)]).astype(np.float32)

    def reset(self, mask):
        for g in range(self.num_games):
            if mask is None or mask[g]:
                self.o[g] = oracles.Oracle(self.board_size, self.lib)
                self.resets[g] += 1

    def synchronize(self):
        pass


class Search:
    """MctsBatch interface: one wave whose leaves are the roots of the active games"""

    waves_per_move = 1

    def __init__(self, boards, name):
        self.b, self.name = boards, name
        G = boards.num_games
        self.pi = np.zeros((G, P1), np.float32)
        self.v = np.zeros(G, np.float32)
        self.searched = np.zeros(G, bool)
        self.resets = np.zeros(G, int)
        self.evals = 0

    def begin_move(self, active):
        G = self.b.num_games
        self.active = np.ones(G, bool) if active is None else np.asarray(active).astype(bool)
        self.searched[:] = False

    def select(self):
        self.ids = np.flatnonzero(self.active)
        return torch.from_numpy(self.b.features()[self.ids])

    def expand_backup(self, pi, v):
        if pi is None:
            return
        self.pi[self.ids] = pi[: len(self.ids)].numpy()
        self.v[self.ids] = v[: len(self.ids)].numpy()
        self.searched[self.ids] = True
        self.evals += len(self.ids)

    def search(self, actor, active=None):
        self.begin_move(active)
        s = self.select()
        if s.shape[0]:
            r = actor({"s": s})
            self.expand_backup(r["pi"], r["V"].reshape(-1))

    def choose(self, cutoff, thres, never_resign, seed):
        G = self.b.num_games
        acts = np.full(G, -2, np.int32)
        info = self.b.info()
        for g in np.flatnonzero(self.searched):
            side = self.v[g] if info[g, 1] == 1 else -self.v[g]
            if side < -1.0 + thres and info[g, 0] >= 50 and not (never_resign is not None and never_resign[g]):
                acts[g] = -1
                continue
            legal = np.append(self.b.o[g].legal().astype(bool), True)
            acts[g] = int(np.where(legal, self.pi[g], -1.0).argmax())
        return acts, self.v.copy()

    def advance(self, actions):
        pass

    def reset(self, mask):
        self.resets += np.ones_like(self.resets) if mask is None else np.asarray(mask).astype(int)

    def errors(self):
        return np.zeros(4, np.int32)


def net(tag, log):
    """network stub: prefers low action indices (games fill the board column by column and end by
    two passes / the move cutoff); records how many positions it saw under `tag`"""

    def actor(batch):
        k = batch["s"].shape[0]
        log.append((tag, k))
        pi = torch.linspace(1.0, 0.1, P1).repeat(k, 1)
        return {"pi": pi, "V": torch.zeros(k)}

    return actor


def make(oracle_lib, G=4, two=False, **kw):
    b = Boards(G, oracle_lib)
    log = []
    sp = SelfPlay(net("black", log), num_games=G, board_size=N, policy_distri_cutoff=0, never_resign_ratio=0.0,
                  actor_white=net("white", log) if two else None, board=b, search=Search(b, "ai"),
                  search_white=Search(b, "ai2") if two else None, **kw)
    return sp, b, log


def test_onreceive_table(oracle_lib):
    sp, b, log = make(oracle_lib)
    assert sp.step() == 4 and sp.step() == 4 and (b.info()[:, 0] == 3).all()  # no protocol: plays at once
    # first request = the reference's "was waiting" case: restart with the new model
    assert sp.set_request(5, -1, 0.1) == "update_model"
    assert (b.info()[:, 0] == 1).all() and (b.resets == 1).all() and (sp.mcts.resets == 1).all()
    assert sp.resign_thres == pytest.approx(0.1)
    sp.step()
    # same versions: thresholds only (black/white thresholds are averaged, go_state_ext.h:62-63)
    assert sp.set_request(5, -1, 0.2, white_resign_thres=0.4, never_resign_prob=0.25) == "update_request_only"
    assert sp.resign_thres == pytest.approx(0.3) and sp.never_resign_ratio == 0.25 and (b.info()[:, 0] == 2).all()
    # new version, synchronous: games restart
    assert sp.set_request(6, -1, 0.2) == "update_model" and (b.info()[:, 0] == 1).all() and (b.resets == 2).all()
    sp.step()
    # new version, async: the model changes under the running games
    assert sp.set_request(7, -1, 0.2, async_=True) == "update_model_async" and (b.info()[:, 0] == 2).all()
    assert sp.set_request(7, -1, 0.2, async_=True) == "update_request_only"
    # wait: nobody plays
    assert sp.set_request(-1, -1) == "only_wait" and sp.idle.all()
    assert sp.step() == 0 and (b.info()[:, 0] == 2).all()
    # leaving the wait state restarts even for the same version pair as before
    assert sp.set_request(7, -1, 0.2, async_=True) == "update_model" and sp.idle is None and (b.resets == 3).all()
    # thread limit: only the first two games play, the parked ones keep their position
    assert sp.set_request(7, -1, 0.2, async_=True, num_game_thread_used=2) == "update_request_only"
    assert sp.idle.tolist() == [False, False, True, True]
    assert sp.step() == 2 and b.info()[:, 0].tolist() == [2, 2, 1, 1]
    n_eval = sp.mcts.evals
    assert sp.step() == 2 and sp.mcts.evals == n_eval + 2  # parked games cost no network evaluations
    # parked games that are used again start afresh, running ones carry on
    assert sp.set_request(7, -1, 0.2, async_=True, num_game_thread_used=-1) == "update_request_only"
    assert sp.idle is None and b.resets.tolist() == [3, 3, 4, 4] and b.info()[:, 0].tolist() == [3, 3, 1, 1]
    assert sp.step() == 4


def test_match_routing_and_player_swap(oracle_lib):
    sp, b, log = make(oracle_lib, two=True)
    assert sp.set_request(10, 9, 0.0) == "update_model"
    sp.step()  # black to move everywhere: only _ai (actor_black, version 10) is asked
    assert log == [("black", 4)]
    sp.step()
    assert log[-1] == ("white", 4)
    b.o[0].forward(N * N)  # game 0: an extra pass flips the side to move
    del log[:]
    sp.step()
    assert sorted(log) == [("black", 3), ("white", 1)]
    # swap: restart (same versions, different swap) and the roles flip
    assert sp.set_request(10, 9, 0.0, player_swap=True) == "update_model" and sp.swap
    del log[:]
    sp.step()
    assert log == [("white", 4)]  # black to move is now played by the "white" model
    sp.step()
    assert log[-1] == ("black", 4)
    labels = [(lab, m.tolist()) for _, _, lab, m in sp.phases(b.info())]
    assert labels == [("actor_black", [0, 0, 0, 0]), ("actor_white", [1, 1, 1, 1])]  # black to move -> actor_white
    # back to self-play with one model: no swap without a second AI
    assert sp.set_request(11, -1, 0.0, player_swap=True) == "update_model" and not sp.swap


def test_games_finish_and_restart_under_protocol(oracle_lib):
    sp, b, log = make(oracle_lib, G=3, move_cutoff=6)
    sp.set_request(1, -1, 0.0, num_game_thread_used=2)
    for _ in range(12):
        sp.step()
    assert sp.games_finished == 4 and all(r[1:] == (6, "max_step") for r in sp.results)  # 2 games x 2 rounds
    assert b.resets.tolist() == [3, 3, 0]  # request restart + one restart per finished game; the parked game is not touched


# ---- through the pybind-compatible surface, with the reference's own GCWrapper -----------------------
REF_UTILS = "/root/reference/src_py/elf/utils_elf.py"


@pytest.mark.skipif(not os.path.exists(REF_UTILS), reason="reference tree not present (GPU box)")
def test_version_switch_through_compat_surface(oracle_lib):
    spec = importlib.util.spec_from_file_location("ref_utils_elf_req", REF_UTILS)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    sp, b, _ = make(oracle_lib, G=4, move_cutoff=8)
    sp.actor = None  # the callbacks own the networks
    eng = compat.SelfPlayEngine(sp)
    GC = compat.GameContext(eng, batchsize=3)
    desc = {  # src_py/elfgames/go/game.py:375-405
        "actor_black": dict(input=["s"], reply=["pi", "V", "a", "rv"], batchsize=3, timeout_usec=10),
        "actor_white": dict(input=["s"], reply=["pi", "V", "a", "rv"], batchsize=3, timeout_usec=10),
        "game_end": dict(batchsize=1),
        "game_start": dict(batchsize=1, input=["black_ver", "white_ver"], reply=None),
    }
    gcw = ref.GCWrapper(GC, 3, desc, num_recv=2, gpu=None, use_numpy=False, params=GC.getParams())
    log = []
    loaded = {"ver": None}

    def actor(batch):
        k = batch["s"].shape[0]
        log.append(("actor", loaded["ver"], k))
        return dict(pi=torch.linspace(1.0, 0.1, P1).repeat(k, 1), V=torch.zeros(k), a=torch.zeros(k, dtype=torch.int64),
                    rv=torch.zeros(k, dtype=torch.int64))

    def game_start(batch):  # scripts/elfgames/go/selfplay.py:138-156: load the models named by the request
        loaded["ver"] = int(batch["black_ver"][0])
        log.append(("start", loaded["ver"], int(batch["white_ver"][0])))

    def game_end(batch):
        log.append(("end", batch.GC.getClient().getGameStats().getWinRateStats().total_games))

    gcw.reg_callback("actor_black", actor)
    gcw.reg_callback("actor_white", actor)
    gcw.reg_callback_if_exists("game_start", game_start)
    gcw.reg_callback_if_exists("game_end", game_end)
    gcw.start()
    GC.getClient().setRequest(5, -1, 0.1, -1)  # scripts/elfgames/go/selfplay.py:186-187
    for _ in range(9):
        gcw.run()
    assert log[0] == ("start", 5, -1)  # the versions reach Python before any position of that model
    assert [e[:2] for e in log[1:]] == [("actor", 5)] * 8 and [e[2] for e in log[1:5]] == [3, 1, 3, 1]  # 4 leaves in chunks of <= 3
    assert sp.resign_thres == pytest.approx(0.1) and eng.replies == ["update_model"]
    GC.getClient().setRequest(6, -1, 0.1, -1)  # mid-move: takes effect at the move boundary
    n = len(log)
    while ("start", 6, -1) not in log:
        gcw.run()
    i = log.index(("start", 6, -1))
    assert all(e[:2] == ("actor", 5) for e in log[n:i])  # the move in flight finished on the old model
    assert (b.info()[:, 0] == 1).all()  # ... then every game restarted
    for _ in range(40):
        gcw.run()
    assert all(e[1] == 6 for e in log[i + 1:] if e[0] == "actor")
    ends = [e for e in log if e[0] == "end"]
    assert len(ends) == sp.games_finished >= 4 and ends[-1][1] == sp.games_finished
    assert [e for e in log if e[0] == "start"] == [("start", 5, -1), ("start", 6, -1)]  # no per-game game_start
    # every game parked: the pump has nothing to wait for
    GC.getClient().setRequest(-1, -1, 0.1, -1)
    with pytest.raises(RuntimeError, match="waiting for a request"):
        for _ in range(10):
            gcw.run()
    gcw.stop()


def test_records_name_every_model_a_game_saw(oracle_lib):
    class VisitSearch(Search):
        def results(self):
            return {"visits": np.where(self.pi > 0.5, 10, -1).astype(np.int32)}

    b = Boards(2, oracle_lib)
    sp = SelfPlay(net("black", []), num_games=2, board_size=N, policy_distri_cutoff=0, never_resign_ratio=0.0,
                  move_cutoff=6, record_games=True, board=b, search=VisitSearch(b, "ai"))
    sp.set_request(3, -1, 0.0)
    sp.step()
    sp.step()
    assert sp.set_request(4, -1, 0.0, async_=True) == "update_model_async"  # new model in the middle of the games
    while len(sp.records) < 4:
        sp.step()
    assert [r["result"]["using_models"] for r in sp.records] == [[3, 4], [3, 4], [4], [4]]
    assert [r["request"]["vers"]["black_ver"] for r in sp.records] == [4] * 4
    assert all(r["request"]["client_ctrl"]["async"] for r in sp.records)


@pytest.mark.skipif(not oracles.have_ref(9), reason="oracle/_ref not built")
def test_request_message_format_against_reference(oracle_lib):
    """the server's MsgRequest JSON: what record.request_json writes is what the reference's own
    MsgRequest::createFromJson reads (and writes back), and parse_request + set_request_msg consume
    what the reference writes"""
    import json

    from elf_b200 import record

    msg = record.request_json(7, 5, mcts_opt=dict(num_rollouts=1600, c_puct=0.85, virtual_loss=2, persistent_tree=1),
                              black_resign_thres=0.04, white_resign_thres=0.08, never_resign_prob=0.1, player_swap=True,
                              async_=True, num_game_thread_used=3)
    back = oracles.ref_request_roundtrip(json.dumps(msg))
    assert back is not None
    back = json.loads(back)
    assert back["vers"]["mcts_opt"]["num_rollouts_per_thread"] == 1600 and back["vers"]["mcts_opt"]["alg_opt"]["c_puct"] == pytest.approx(0.85)
    assert back["client_ctrl"]["player_swap"] is True and back["client_ctrl"]["async"] is True
    kw, opt = record.parse_request(json.dumps(back))  # the reference's own serialisation
    assert kw == dict(black_ver=7, white_ver=5, black_resign_thres=pytest.approx(0.04), white_resign_thres=pytest.approx(0.08),
                      never_resign_prob=pytest.approx(0.1), player_swap=True, async_=True, num_game_thread_used=3)
    assert opt["virtual_loss"] == 2
    # fields the reference treats as optional / mandatory
    m2 = json.loads(json.dumps(record.request_json(3)))
    del m2["client_ctrl"]["async"], m2["client_ctrl"]["player_swap"]  # a self-play request may omit both
    assert oracles.ref_request_roundtrip(json.dumps(m2)) is not None and record.parse_request(m2)[0]["player_swap"] is False
    m3 = json.loads(json.dumps(record.request_json(3, 2)))
    del m3["client_ctrl"]["player_swap"]  # an evaluation request may not
    assert oracles.ref_request_roundtrip(json.dumps(m3)) is None
    with pytest.raises(KeyError):
        record.parse_request(m3)
    # applied to the engine: versions, thresholds, the parked games, and mcts_opt into the records
    class VisitSearch(Search):
        def results(self):
            return {"visits": np.where(self.pi > 0.5, 10, -1).astype(np.int32)}

    b = Boards(3, oracle_lib)
    sp = SelfPlay(net("black", []), num_games=3, board_size=N, policy_distri_cutoff=0, never_resign_ratio=0.0,
                  move_cutoff=4, record_games=True, board=b, search=VisitSearch(b, "ai"))
    assert sp.set_request_msg(json.dumps(record.request_json(9, mcts_opt=dict(num_rollouts=400), black_resign_thres=0.02,
                                                             white_resign_thres=0.06, num_game_thread_used=2))) == "update_model"
    assert sp.resign_thres == pytest.approx(0.04) and sp.idle.tolist() == [False, False, True]
    while len(sp.records) < 2:
        sp.step()
    r = sp.records[0]
    assert r["request"]["vers"]["black_ver"] == 9 and r["request"]["vers"]["mcts_opt"]["num_rollouts_per_thread"] == 400
    assert r["request"]["client_ctrl"]["num_game_thread_used"] == 2 and r["result"]["using_models"] == [9]
    assert oracles.ref_record_roundtrip(json.dumps(r)) is not None


def test_num_games_per_thread_and_policy_for_all(oracle_lib):
    """a slot stops after num_games_per_thread games (GoStateExt::finished); with
    policy_distri_training_for_all every move's MCTS policy is recorded, not only those up to the cutoff"""
    class VisitSearch(Search):
        def results(self):
            return {"visits": np.where(self.pi > 0.5, 10, -1).astype(np.int32)}

    b = Boards(2, oracle_lib)
    sp = SelfPlay(net("black", []), num_games=2, board_size=N, policy_distri_cutoff=1, never_resign_ratio=0.0,
                  move_cutoff=5, record_games=True, policy_distri_training_for_all=True, num_games_per_thread=2,
                  board=b, search=VisitSearch(b, "ai"))
    moved = [sp.step() for _ in range(12)]
    assert moved[:8] == [2] * 8 and moved[8:] == [0] * 4  # 2 games x 4 moves per slot, then every slot has stopped
    assert sp.games_finished == 4 and sp.stopped.all() and sp.idle.all()
    assert all(len(r["result"]["policies"]) == 4 for r in sp.records)  # for_all: one policy per move
    with pytest.raises(RuntimeError, match="waiting"):  # nothing left for the pump either
        from elf_b200 import compat as _c
        eng = _c.SelfPlayEngine(sp)
        eng.next_label()
