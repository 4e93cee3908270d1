"""The pybind-compatible surface (elf_b200.compat) against the UNMODIFIED reference GCWrapper
(/root/reference/src_py/elf/utils_elf.py, loaded straight from the reference tree when present):
the reference's own allocation code (Allocator.spec2batches -> AnyP.field()/set()) and its
wait -> callback -> step pump must run on our Context/SharedMem/AnyP objects without change.
A stub engine stands in for the GPU search here (CPU-only box); the GPU test of the same surface
with the real engine is tests/test_gpu_compat.py."""
import importlib.util
import os

import numpy as np
import pytest
import torch

REF_UTILS = "/root/reference/src_py/elf/utils_elf.py"


class StubEngine:
    """emits two waves of 5 and 3 'leaves' per move with recognisable features, records replies"""

    board_size = 9
    num_action = 82

    def __init__(self):
        self.waves = [5, 3]
        self.w = 0
        self.off = 0
        self.cur = None
        self.replies = []
        self.started = False

    def start(self):
        self.started = True

    def stop(self):
        self.started = False

    def win_stats(self):
        from elf_b200 import compat

        return compat.WinRateStats(0, 0)

    def next_batch(self, max_n):
        if self.cur is None or self.off >= self.cur.shape[0]:
            n = self.waves[self.w % 2]
            self.cur = torch.arange(n, dtype=torch.float32).reshape(n, 1, 1, 1).expand(n, 18, 9, 9) + 100 * self.w
            self.w += 1
            self.off = 0
        k = min(max_n, self.cur.shape[0] - self.off)
        out = self.cur[self.off:self.off + k]
        self.off += k
        return k, out

    def reply(self, pi, v):
        self.replies.append((pi.clone(), v.clone()))

    events = None

    def poll_event(self):
        return self.events.pop(0) if self.events else None


def _load_reference_gcwrapper():
    if not os.path.exists(REF_UTILS):
        pytest.skip("reference tree not present")
    spec = importlib.util.spec_from_file_location("ref_utils_elf", REF_UTILS)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_reference_gcwrapper_runs_on_compat_surface(capsys):
    from elf_b200 import compat

    ref = _load_reference_gcwrapper()
    eng = StubEngine()
    GC = compat.GameContext(eng, batchsize=4)
    # the reference's own desc for selfplay's actor_black (src_py/elfgames/go/game.py:375-381)
    desc = {"actor_black": dict(input=["s"], reply=["pi", "V", "a", "rv"], batchsize=4, timeout_usec=10)}
    gcw = ref.GCWrapper(GC, 4, desc, num_recv=2, gpu=None, use_numpy=False, params=GC.getParams())
    seen = []

    def actor(batch):
        s = batch["s"]
        assert s.shape[1:] == (18, 9, 9)
        assert batch.batchsize == s.shape[0] and batch.max_batchsize == 4
        seen.append(s[:, 0, 0, 0].clone())
        n = s.shape[0]
        pi = torch.full((n, 82), 1.0 / 82) + s[:, 0, 0, 0].reshape(-1, 1)
        return dict(pi=pi, V=s[:, 0, 0, 0] * 0.5, a=torch.zeros(n, dtype=torch.int64), rv=torch.zeros(n, dtype=torch.int64))

    gcw.reg_callback("actor_black", actor)
    gcw.start()
    for _ in range(4):  # wave of 5 -> chunks 4 + 1, wave of 3 -> 3, next wave of 5 -> 4
        gcw.run()
    gcw.stop()
    assert [t.tolist() for t in seen] == [[0, 1, 2, 3], [4], [100, 101, 102], [200, 201, 202, 203]]
    assert len(eng.replies) == 4
    np.testing.assert_allclose(eng.replies[1][1].numpy(), [2.0])
    np.testing.assert_allclose(eng.replies[2][0][:, 0].numpy(), np.array([100, 101, 102]) + 1 / 82, rtol=1e-6)
    p = GC.getParams()
    assert p["num_action"] == 82 and p["num_planes"] == 18 and p["ACTION_PASS"] == -99


def test_compat_objects_have_the_pybind_names():
    from elf_b200 import compat

    eng = StubEngine()
    ctx = compat.GameContext(eng, 8).ctx()
    o = ctx.createSharedMemOptions("actor_black", 8)
    o.setTimeout(10)
    sm = ctx.allocateSharedMem(o, ["s", "pi", "V", "a", "rv"])
    assert sm.getSharedMemOptions().label() == "actor_black" and sm.getSharedMemOptions().batchsize() == 8
    assert sm.getSharedMemOptions().idx() == 0
    f = sm["s"].field()
    assert (f.name(), f.type_name(), f.sz().vec()) == ("s", "float", [8, 18, 9, 9])
    assert sm["a"].field().type_name() == "int64_t" and sm["pi"].field().sz().vec() == [8, 82]
    with pytest.raises(KeyError):
        ctx.allocateSharedMem(o, ["offline_a_not_supported"])
    with pytest.raises(RuntimeError):
        ctx.wait()  # before start()
    assert isinstance(ctx.version(), str)


def test_game_start_and_game_end_labels_reach_their_callbacks():
    """selfplay-mode desc (game.py:375-405): actor_black/actor_white + game_start/game_end with
    batchsize 1; notifications are delivered through the same wait/step pump."""
    from elf_b200 import compat

    ref = _load_reference_gcwrapper()
    eng = StubEngine()
    eng.events = [("game_start", {"black_ver": 7, "white_ver": -1}), ("game_end", {})]
    GC = compat.GameContext(eng, batchsize=4)
    desc = {
        "actor_black": dict(input=["s"], reply=["pi", "V", "a", "rv"], batchsize=4, timeout_usec=10),
        "actor_white": dict(input=["s"], reply=["pi", "V", "a", "rv"], batchsize=4, timeout_usec=10),
        "game_end": dict(batchsize=1),
        "game_start": dict(batchsize=1, input=["black_ver", "white_ver"], reply=None),
    }
    gcw = ref.GCWrapper(GC, 4, desc, num_recv=2, gpu=None, use_numpy=False, params=GC.getParams())
    log = []
    n_actor = [0]

    def actor(batch):
        n = batch["s"].shape[0]
        n_actor[0] += 1
        return dict(pi=torch.zeros(n, 82), V=torch.zeros(n), a=torch.zeros(n, dtype=torch.int64), rv=torch.zeros(n, dtype=torch.int64))

    gcw.reg_callback("actor_black", actor)
    gcw.reg_callback("actor_white", actor)
    assert gcw.reg_callback_if_exists("game_start", lambda b: log.append(("start", int(b["black_ver"][0]), int(b["white_ver"][0]))))
    assert gcw.reg_callback_if_exists("game_end", lambda b: log.append(("end", b.GC.getClient().getGameStats().getWinRateStats().total_games)))
    gcw.start()
    for _ in range(3):
        gcw.run()
    gcw.stop()
    assert log == [("start", 7, -1), ("end", 0)] and n_actor[0] == 1


def test_game_context_from_reference_options():
    """go.GameContext(co, opt): the reference's option structs (same field names and defaults) are
    turned into the engine's constructor arguments"""
    from elf_b200 import compat

    co, opt = compat.ContextOptions(), compat.GameOptions()
    assert (co.mcts_options.num_threads, co.mcts_options.num_rollouts_per_thread, co.mcts_options.alg_opt.c_puct) == (16, 100, 5.0)
    assert (opt.policy_distri_cutoff, opt.resign_thres, opt.komi, opt.white_puct, opt.preload_sgf_move_to) == (20, 0.05, 7.5, -1.0, -1)
    with pytest.raises(AttributeError):
        compat.GameOptions(no_such_field=1)
    # what src_py/elf/context_utils.py:89-112 and src_py/elfgames/go/game.py:264-345 set for start_selfplay.sh
    co.num_games, co.batchsize = 32, 128
    m = co.mcts_options
    m.num_threads, m.num_rollouts_per_thread, m.num_rollouts_per_batch, m.virtual_loss, m.persistent_tree = 8, 100, 8, 1, True
    m.alg_opt.c_puct, m.alg_opt.unexplored_q_zero = 1.5, True
    opt.mode, opt.policy_distri_cutoff, opt.resign_thres, opt.move_cutoff, opt.seed = "selfplay", 30, 0.01, 200, 5
    opt.white_puct, opt.white_mcts_rollout_per_thread = 0.85, 50
    got = {}

    class FakeSelfPlay:
        def __init__(self, **kw):
            got.update(kw)
            self.N, self.G, self.resign_thres, self.policy_only = kw["board_size"], kw["num_games"], kw["resign_thres"], {}

    GC = compat.game_context(co, opt, board_size=9, factories={"selfplay": FakeSelfPlay})
    assert isinstance(GC, compat.GameContext) and GC.getParams()["board_size"] == 9
    assert got["num_games"] == 32 and got["num_rollouts"] == 800 and got["num_rollouts_per_batch"] == 8
    assert got["c_puct"] == 1.5 and got["virtual_loss"] == 1 and got["persistent_tree"] == 1 and got["unexplored_q_zero"] == 1
    assert got["policy_distri_cutoff"] == 30 and got["resign_thres"] == 0.01 and got["move_cutoff"] == 200 and got["seed"] == 5
    assert got["never_resign_ratio"] == 0.1 and got["komi"] == 7.5 and got["actor"] is None
    assert got["white_mcts_opts"] == {"c_puct": 0.85, "num_rollouts": 400}  # only what differs for the second AI
    assert got["rng"] == "reference"  # GameOptions::seed != 0: the reference's games are reproducible, ours are the same games
    # online and train modes
    opt.mode, opt.preload_sgf, opt.preload_sgf_move_to, opt.following_pass = "online", "g.sgf", 12, True

    class FakeGame:
        N, resign_thres = 9, 0.0

    def fake_online(**kw):
        got.clear()
        got.update(kw)
        return FakeGame()

    GC = compat.game_context(co, opt, board_size=9, factories={"online": fake_online})
    assert isinstance(GC._engine, compat.OnlineEngine)
    assert (got["preload_sgf"], got["preload_sgf_move_to"], got["following_pass"], got["num_rollouts"]) == ("g.sgf", 12, True, 800)
    opt.mode, opt.num_future_actions = "train", 3

    class FakeReplay:
        def __init__(self, **kw):
            got.clear()
            got.update(kw)
            self.N, self.K, self.B = kw["board_size"], kw["num_future_actions"], kw["num_states"]

    GC = compat.game_context(co, opt, board_size=9, factories={"replay": FakeReplay})
    assert isinstance(GC._engine, compat.TrainEngine) and got["num_states"] == 128 and got["num_future_actions"] == 3
    assert GC.ctx().allocateSharedMem(GC.ctx().createSharedMemOptions("train", 128), ["offline_a"])["offline_a"].field().sz().vec() == [128, 3]
    opt.mode = "bogus"
    with pytest.raises(ValueError, match="Unknown mode"):
        compat.game_context(co, opt)
    # TSOptions::pick_method: unknown names fail like TreeSearchT::chooseAction, the two unused ones loudly
    co.mcts_options.pick_method = "best_guess"
    with pytest.raises(ValueError, match="MCTS Pick method unknown! best_guess"):
        compat.game_context(co, opt)
    co.mcts_options.pick_method = "strongest_prior"
    with pytest.raises(NotImplementedError):
        compat.game_context(co, opt)
    co.mcts_options.pick_method = "most_visited"
    # policy-only colours go through the same pump (tests/test_dropin_shim.py plays such games)
    opt.mode, opt.white_use_policy_network_only = "selfplay", True
    GC = compat.game_context(co, opt, factories={"selfplay": FakeSelfPlay})
    assert isinstance(GC._engine, compat.SelfPlayEngine) and got["white_use_policy_network_only"] is True
    assert got["rng"] == "reference"  # GameOptions::seed is set: policy-only colours run on the reference streams too
    opt.white_use_policy_network_only, opt.seed = False, 0
    got.clear()
    compat.game_context(co, opt, factories={"selfplay": FakeSelfPlay})
    assert "rng" not in got  # seed 0: the reference seeds from the clock
