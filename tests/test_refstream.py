"""The reference's random streams (include/elfb200_refstream.h, elf_b200/refstream.py) against the
COMPILED reference (oracle/_ref): container order of the root edges, root Dirichlet noise, the D4
code of every evaluated leaf, sampled moves and the never-resign draw -- whole self-play games, move
for move.  The search kernels run on the SIMT emulator here (tests/test_gpu_mcts.py has the device
twin)."""
import numpy as np
import pytest
import torch

from elf_b200.refstream import RefStream
from tests import emu, oracles

pytestmark = pytest.mark.skipif(not oracles.have_ref(9), reason="compiled reference (oracle/_ref) not available")


def _quantised(pi, quant):
    """quant > 0: probabilities on a grid of 1/quant -- bit-equal values everywhere, like the replies of a
    half-precision network (equal fp16 logits)"""
    return pi if not quant else (np.floor(pi * np.float32(quant)) / np.float32(quant)).astype(np.float32)


def plane_actor(n, quant=0):
    def actor(batch):
        pi, v = oracles.feature_net(batch["s"].float().numpy(), n * n + 1)
        return {"pi": torch.from_numpy(_quantised(pi, quant)), "V": torch.from_numpy(v)}

    return actor


def ref_net(n, quant=0):
    def net(feats, hashes):
        pi, v = oracles.feature_net(feats, n * n + 1)
        return _quantised(pi, quant), v

    return net


@pytest.mark.parametrize("n", [9, 19])
def test_edge_order_is_the_reference_containers(n):
    if not oracles.have_ref(n):
        pytest.skip("compiled reference not available")
    st = oracles.Ref(n)
    rng = np.random.default_rng(5)
    for ply in range(30 if n == 9 else 8):
        ref = oracles.RefMcts(n, num_rollouts=24, num_rollouts_per_batch=4, callback=ref_net(n))
        r = ref.act(st)
        order = ref.last_order()
        has = np.flatnonzero(r["visits"] >= 0)
        assert sorted(order.tolist()) == has.tolist()
        pri = r["prior"][has]
        assert len(np.unique(pri)) == len(pri)  # distinct priors: the insertion order is well defined
        inserted = has[np.argsort(-pri, kind="stable")]  # pi2response: descending prior
        ours = inserted[RefStream.edge_order(n, inserted)]
        np.testing.assert_array_equal(ours, order)
        legal = np.flatnonzero(st.legal())
        st.forward(int(rng.choice(legal)) if len(legal) else n * n)


def play_reference_game(n, seed, opts, eps, alpha, flip, cutoff, thres, ratio, max_moves, quant=0):
    """one game thread of GoGameSelfPlay::act (game_selfplay.cc:272-430) composed from the reference's
    own pieces: init_ai's seed draw, MCTSAI_T::act, mcts_make_diverse_move, shouldResign, forward"""
    g = oracles.RefRng(n, seed)
    ref = oracles.RefMcts(n, callback=ref_net(n, quant), root_epsilon=eps, root_alpha=alpha, rotation_flip=flip,
                          seed=g.next(), **opts)
    rc = oracles.RefResign(n, thres, ratio)
    st = oracles.Ref(n)
    log = []
    for _ in range(max_moves):
        info = st.info()
        ply, nxt = int(info[0]), int(info[1])
        r = ref.act(st)
        a = r["best_action"]
        if ply <= cutoff:
            a = ref.sample(g)
        resign = rc.check(r["best_q"], nxt, g) and ply >= 50
        log.append({"visits": r["visits"].copy(), "prior": r["prior"].copy(), "action": int(a), "value": r["best_q"],
                    "resign": bool(resign), "best": int(r["best_action"])})
        if resign:
            break
        assert st.forward(int(a))
        if st.info()[9]:
            break
    return log


@pytest.mark.parametrize("n,eps,flip,moves,quant", [(9, 0.25, 1, 30, 0), (9, 0.0, 1, 16, 0), (9, 0.25, 0, 16, 0),
                                                    (19, 0.25, 1, 8, 0), (9, 0.25, 1, 24, 512), (9, 0.0, 1, 24, 64)])
def test_whole_games_move_for_move(n, eps, flip, moves, quant):
    """quant > 0: the network's probabilities collide bit for bit (what half precision does); the edges are
    then stored in the order std::sort leaves equal elements in (search option std_sort_ties), which decides
    the container order of every later tie-break"""
    if not oracles.have_ref(n):
        pytest.skip("compiled reference not available")
    G, seed = 3, 20240917
    opts = dict(num_rollouts=32, num_rollouts_per_batch=4, virtual_loss=1, persistent_tree=1, c_puct=1.5, komi=7.5)
    cutoff, thres, ratio, alpha = 12, 0.05, 0.1, 0.3
    seeds = [seed, seed, seed + 1]  # GameOptions::seed seeds every game thread alike; one more for variety
    logs = [play_reference_game(n, s, opts, eps, alpha, flip, cutoff, thres, ratio, moves, quant) for s in seeds]

    gb = emu.emu_batch(G, n)
    mc = emu.EmuSearch(gb, rotation_flip=flip, std_sort_ties=1, **opts)
    rs = RefStream(G, n, np.array(seeds, np.uint64))
    rs.init_actor(0)
    mc.attach_ref_stream(rs, 0, eps, alpha)
    actor = plane_actor(n, quant)
    alive = np.ones(G, bool)
    drawn = np.zeros(G, bool)
    never = np.zeros(G, bool)
    for t in range(moves):
        info = gb.info()
        mc.search(actor, active=alive.astype(np.uint8))
        res = mc.results()
        pri = mc.root_priors()
        cho = mc.ref_choose(sample=(info[:, 0] <= cutoff), mask=alive.astype(np.uint8))
        need = alive & ~drawn
        u = rs.game_uniform(need.astype(np.uint8))
        never[need] = u[need] < ratio
        drawn |= need
        side = np.where(info[:, 1] == 1, cho["value"], -cho["value"]).astype(np.float64)
        resign = alive & ~never & ~(side >= -1.0 + float(np.float32(thres))) & (info[:, 0] >= 50)
        acts = np.full(G, -2, np.int32)
        for g in np.flatnonzero(alive):
            L = logs[g][t]
            np.testing.assert_array_equal(res["visits"][g], L["visits"], err_msg=f"game {g} move {t}: visits")
            has = L["visits"] >= 0
            np.testing.assert_array_equal(pri[g][has], L["prior"][has], err_msg=f"game {g} move {t}: priors")
            assert cho["best_action"][g] == L["best"], (g, t)
            assert cho["action"][g] == L["action"], (g, t)
            # W is a float sum whose order the reference leaves to the addresses of its leaf nodes
            # (batch_rollouts walks an unordered_map keyed by Node*): equal up to the last bits
            assert np.isclose(cho["value"][g], L["value"], rtol=2e-6, atol=1e-7, equal_nan=True), (g, t)
            assert bool(resign[g]) == L["resign"]
            acts[g] = cho["action"][g]
            if t + 1 == len(logs[g]):
                alive[g] = False
        gb.forward(acts)
        mc.advance(acts)
        if not alive.any():
            break
    assert t >= min(20, moves - 1)
    # the two game threads with the same GameOptions::seed played the same game; the third did not
    assert [m["action"] for m in logs[0]] == [m["action"] for m in logs[1]]
    assert [m["action"] for m in logs[0]] != [m["action"] for m in logs[2]]


def test_selfplay_driver_on_the_reference_streams():
    """SelfPlay(rng="reference"): the whole per-move driver (search, choice, resign check, forward, tree
    advance, restart) replays the reference game threads' games, including the game after a restart"""
    from elf_b200.selfplay import SelfPlay

    n, G, moves = 9, 2, 26
    opts = dict(num_rollouts=24, num_rollouts_per_batch=4, virtual_loss=1, persistent_tree=1, c_puct=1.5, komi=7.5)
    eps, alpha, cutoff, thres, ratio, move_cutoff = 0.25, 0.3, 8, 0.05, 0.1, 11
    seeds = np.array([777, 778], np.uint64)

    # reference side: one generator per game thread lives across games; move_cutoff ends a game
    # (finish_game -> endGame resets the tree, _state_ext.restart() resets ResignCheck)
    expect = []
    for s in seeds:
        g = oracles.RefRng(n, int(s))
        ref = oracles.RefMcts(n, callback=ref_net(n), root_epsilon=eps, root_alpha=alpha, rotation_flip=1,
                              seed=g.next(), **opts)
        rc = oracles.RefResign(n, thres, ratio)
        st = oracles.Ref(n)
        played = []
        for _ in range(moves):
            ply = int(st.info()[0])
            r = ref.act(st)
            a = ref.sample(g) if ply <= cutoff else r["best_action"]
            rc.check(r["best_q"], int(st.info()[1]), g)
            assert st.forward(int(a))
            played.append(int(a))
            if st.info()[9] or int(st.info()[0]) >= move_cutoff:
                ref.end_game(st)
                st = oracles.Ref(n)
                rc.reset()
        expect.append(played)

    gb = emu.emu_batch(G, n)
    mc = emu.EmuSearch(gb, rotation_flip=1, std_sort_ties=1, **opts)
    sp = SelfPlay(plane_actor(n), num_games=G, board_size=n, board=gb, search=mc, rng="reference", seed=seeds,
                  policy_distri_cutoff=cutoff, resign_thres=thres, never_resign_ratio=ratio, move_cutoff=move_cutoff,
                  root_epsilon=eps, root_alpha=alpha, **opts)
    got = [[] for _ in range(G)]
    fwd = gb.forward

    def logged(acts):
        for g in range(G):
            got[g].append(int(acts[g]))
        return fwd(acts)

    gb.forward = logged
    for _ in range(moves):
        sp.step()
    assert got == expect
    assert sp.games_finished >= 2 * G


def test_game_context_with_a_seed_runs_the_reference_streams():
    """compat.game_context with GameOptions::seed != 0: the wait()/step() pump (begin_move by the engine,
    finish_move at the end of a move) consumes the streams exactly as SelfPlay.step() does -- same games"""
    from elf_b200 import compat, lib as _l
    from elf_b200.selfplay import SelfPlay

    n, G = 9, 2
    co, opt = compat.ContextOptions(), compat.GameOptions()
    co.num_games, co.batchsize = G, 8
    m = co.mcts_options
    m.num_threads, m.num_rollouts_per_thread, m.num_rollouts_per_batch, m.virtual_loss, m.persistent_tree = 1, 16, 4, 1, True
    m.alg_opt.c_puct, m.root_epsilon, m.root_alpha = 1.5, 0.25, 0.3
    opt.mode, opt.policy_distri_cutoff, opt.move_cutoff, opt.seed = "selfplay", 6, 9, 4242
    logs = []

    def factory(**kw):
        assert kw["rng"] == "reference"
        fields = {f[0] for f in _l.MctsOptions._fields_}
        gb = emu.emu_batch(kw["num_games"], kw["board_size"])
        mk = {k: v for k, v in kw.items() if k in fields and k not in ("seed", "root_epsilon", "root_alpha")}
        sp = SelfPlay(board=gb, search=emu.EmuSearch(gb, **mk), **kw)
        log, fwd = [], gb.forward

        def logged(acts):
            log.append([int(a) for a in acts])
            return fwd(acts)

        gb.forward = logged
        logs.append(log)
        return sp

    # (1) through the pump
    GC = compat.game_context(co, opt, board_size=n, factories={"selfplay": factory})
    ctx = GC.ctx()
    sm = ctx.allocateSharedMem(ctx.createSharedMemOptions("actor_black", 8), ["s", "pi", "V", "a", "rv"])
    bufs = {}
    for k in ("s", "pi", "V", "a", "rv"):
        f = sm[k].field()
        dt = {"float": torch.float32, "int64_t": torch.int64, "int32_t": torch.int32}[f.type_name()]
        bufs[k] = torch.zeros(*f.sz().vec(), dtype=dt)
        sm[k].set(bufs[k].data_ptr(), [i * bufs[k].element_size() for i in bufs[k].stride()])
    ctx.start()
    sp = GC._engine.sp
    it = 0
    while len(logs[0]) < 14 and it < 2000:
        s = ctx.wait()
        k = s.effective_batchsize()
        pi, v = oracles.feature_net(bufs["s"][:k].numpy(), n * n + 1)
        bufs["pi"][:k] = torch.from_numpy(pi)
        bufs["V"][:k] = torch.from_numpy(v)
        ctx.step()
        it += 1
    ctx.stop()
    assert len(logs[0]) >= 14 and sp.games_finished >= G
    # (2) the same options through SelfPlay.step()
    GC2 = compat.game_context(co, opt, board_size=n, factories={"selfplay": factory})
    sp2 = GC2._engine.sp
    sp2.actor = plane_actor(n)
    for _ in range(14):
        sp2.step()
    assert logs[0][:14] == logs[1][:14]
    assert all(a[0] == a[1] for a in logs[1])  # GameOptions::seed seeds every game thread alike: identical games


@pytest.mark.parametrize("n,R", [(9, 12), (9, 40), (19, 16)])
def test_device_tie_break_is_the_reference_container_order(n, R):
    """DEFAULT path (no reference stream attached): the most visited root edge reported by k_results and
    played by k_choose is the FIRST maximum in the reference's container order -- the kernels replay
    libstdc++'s hash-table insertions (a 19x19 root has ~355 edges: every bucket count 13..541 is crossed) -- checked against
    std::unordered_map itself (RefStream.edge_order) and, through it, the compiled reference"""
    G = 6
    gb = emu.emu_batch(G, n)
    rng = np.random.default_rng(n + R)
    os_ = [oracles.Oracle(n) for _ in range(G)]
    for _ in range(6):
        acts = np.array([int(rng.choice(np.flatnonzero(o.legal()))) for o in os_], np.int32)
        for o, a in zip(os_, acts):
            o.forward(int(a))
        assert gb.forward(acts).all()
    mc = emu.EmuSearch(gb, num_rollouts=R, num_rollouts_per_batch=4, rotation_flip=0)
    actor = plane_actor(n)
    tied = differs = 0
    for mv in range(3):
        res = mc.act(actor)
        e = mc.root_edges()
        a_dev, _ = mc.choose(0, 0.0, None, 0)  # cutoff 0: most visited
        for g in range(G):
            k = int(e["n_edges"][g])
            act, vis = e["actions"][g, :k], e["visits"][g, :k]
            order = RefStream.edge_order(n, act)
            top = vis.max()
            first = next(int(i) for i in order if vis[i] == top)
            assert res["best_action"][g] == act[first], (mv, g)
            assert a_dev[g] == act[first], (mv, g)
            if (vis == top).sum() > 1:
                tied += 1
                differs += first != int(np.flatnonzero(vis == top)[0])
        gb.forward(a_dev)
        mc.advance(a_dev)
    assert tied >= 3 and differs >= 1, (tied, differs)  # ties happen, and storage order would have chosen differently


@pytest.mark.parametrize("two_models", [False, True])
def test_policy_only_colour_on_the_reference_streams(two_models):
    """white_use_policy_network_only under rng="reference" (game_selfplay.cc:359-371): white moves by
    MCTSAI_T::actPolicyOnly -- no root noise, no sampled move, one D4 draw when its root still has to be
    evaluated, the largest prior in container order, and MCTSGoAI::getValue of THAT result as predicted
    value (W/N of the chosen edge once the shared tree's root has visits, NaN for an unvisited edge) --
    while black searches; moves and predicted values equal the compiled reference's game threads"""
    from elf_b200.selfplay import SelfPlay

    n, G, moves = 9, 2, 24
    opts = dict(num_rollouts=24, num_rollouts_per_batch=4, virtual_loss=1, persistent_tree=1, c_puct=1.5, komi=7.5)
    eps, alpha, cutoff, thres, ratio, move_cutoff = 0.25, 0.3, 6, 0.05, 0.1, 15
    seeds = np.array([991, 992], np.uint64)

    expect, expect_v = [], []
    for s in seeds:
        g = oracles.RefRng(n, int(s))
        mk = lambda: oracles.RefMcts(n, callback=ref_net(n), root_epsilon=eps, root_alpha=alpha, rotation_flip=1,
                                     seed=g.next(), **opts)
        ai = mk()  # init_ai(_ai) then init_ai(_ai2): the seeds come off the game generator in that order
        ai2 = mk() if two_models else None
        rc = oracles.RefResign(n, thres, ratio)
        st = oracles.Ref(n)
        played, values = [], []
        for _ in range(moves):
            ply, nxt = int(st.info()[0]), int(st.info()[1])
            cur = ai2 if (ai2 is not None and nxt == 2) else ai
            if nxt == 2:
                r = cur.act(st, policy_only=True)
                a = r["best_action"]
            else:
                r = cur.act(st)
                a = cur.sample(g) if ply <= cutoff else r["best_action"]
            rc.check(r["best_q"], nxt, g)
            assert st.forward(int(a))
            played.append(int(a))
            values.append(float(r["best_q"]))
            if st.info()[9] or int(st.info()[0]) >= move_cutoff:
                ai.end_game(st)
                if ai2 is not None:
                    ai2.end_game(st)
                st = oracles.Ref(n)
                rc.reset()
        expect.append(played)
        expect_v.append(values)

    gb = emu.emu_batch(G, n)
    mc = emu.EmuSearch(gb, rotation_flip=1, std_sort_ties=1, **opts)
    mc2 = emu.EmuSearch(gb, rotation_flip=1, std_sort_ties=1, **opts) if two_models else None
    sp = SelfPlay(plane_actor(n), actor_white=plane_actor(n) if two_models else None, num_games=G, board_size=n, board=gb,
                  search=mc, search_white=mc2, rng="reference", seed=seeds, policy_distri_cutoff=cutoff, resign_thres=thres,
                  never_resign_ratio=ratio, move_cutoff=move_cutoff, root_epsilon=eps, root_alpha=alpha,
                  white_use_policy_network_only=True, **opts)
    got, got_v = [[] for _ in range(G)], [[] for _ in range(G)]
    fin = sp.finish_move

    def logged(info, res=None, chosen=None):
        for g in range(G):
            got[g].append(int(chosen[0][g]))
            got_v[g].append(float(chosen[1][g]))
        return fin(info, res=res, chosen=chosen)

    sp.finish_move = logged
    for _ in range(moves):
        sp.step()
    assert got == expect
    # edge reward sums agree to the last bits only (see test_whole_games_move_for_move)
    np.testing.assert_allclose(np.array(got_v), np.array(expect_v), rtol=2e-6, atol=1e-7, equal_nan=True)
    assert sp.games_finished >= G


def test_nan_value_resigns_like_the_reference():
    """ResignCheck::check is written `if (value >= -1 + thres) return false; return true;`
    (game_utils.h:36-39): the NaN predicted value of an unvisited policy-only edge resigns"""
    from elf_b200.selfplay import SelfPlay

    rc = oracles.RefResign(9, 0.05, 0.0)
    g = oracles.RefRng(9, 1)
    want = [bool(rc.check(v, nxt, g)) for v, nxt in ((float("nan"), 1), (float("nan"), 2), (-0.99, 1), (0.99, 2), (0.0, 1))]
    sp = SelfPlay.__new__(SelfPlay)
    sp.never_resign, sp.resign_thres = np.zeros(5, bool), 0.05
    info = np.zeros((5, 12), np.int32)
    info[:, 0], info[:, 1] = 60, [1, 2, 1, 2, 1]
    acts = np.zeros(5, np.int32)
    sp._resign(info, acts, np.array([np.nan, np.nan, -0.99, 0.99, 0.0], np.float32), np.ones(5, bool))
    assert [a == -1 for a in acts] == want == [True, True, True, True, False]


def test_policy_only_colour_default_rng_matches_the_reference():
    """the same without random draws (no noise, no D4, nothing sampled): the device-side choice for the
    searched colour and the host-side policy-only choice, move for move and value for value"""
    from elf_b200.selfplay import SelfPlay

    n, G, moves = 9, 3, 20
    opts = dict(num_rollouts=24, num_rollouts_per_batch=4, virtual_loss=1, persistent_tree=1, c_puct=1.5, komi=7.5)
    openings = [[], [40], [20, 60]]
    expect, expect_v = [], []
    for op in openings:
        ai = oracles.RefMcts(n, callback=ref_net(n), rotation_flip=0, **opts)
        st = oracles.Ref(n)
        for a in op:
            assert st.forward(a)
        played, values = [], []
        for _ in range(moves):
            r = ai.act(st, policy_only=int(st.info()[1]) == 1)  # black moves by policy only here
            assert st.forward(int(r["best_action"]))
            played.append(int(r["best_action"]))
            values.append(float(r["best_q"]))
        expect.append(played)
        expect_v.append(values)
    gb = emu.emu_batch(G, n)
    for t in range(2):
        gb.forward(np.array([op[t] if t < len(op) else -2 for op in openings], np.int32))
    mc = emu.EmuSearch(gb, rotation_flip=0, **opts)
    sp = SelfPlay(plane_actor(n), num_games=G, board_size=n, board=gb, search=mc, policy_distri_cutoff=-1,
                  never_resign_ratio=1.0, black_use_policy_network_only=True, **opts)
    got, got_v = [[] for _ in range(G)], [[] for _ in range(G)]
    fin = sp.finish_move

    def logged(info, res=None, chosen=None):
        for g in range(G):
            got[g].append(int(chosen[0][g]))
            got_v[g].append(float(chosen[1][g]))
        return fin(info, res=res, chosen=chosen)

    sp.finish_move = logged
    for _ in range(moves):
        sp.step()
    assert got == expect
    np.testing.assert_allclose(np.array(got_v), np.array(expect_v), rtol=2e-6, atol=1e-7, equal_nan=True)
