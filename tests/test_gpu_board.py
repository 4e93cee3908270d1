"""GPU parity tests for the board path: CUDA (through the C ABI) vs the oracle, bit-exact."""
import numpy as np
import pytest

from tests import oracles

pytestmark = pytest.mark.gpu


def _gobatch(G, n):
    import elf_b200

    return elf_b200.GoBatch(G, board_size=n)


def _random_candidate(rng, o, avoid_eyes=True):
    """pick a move like the playout policy but with numpy randomness (covers other streams)"""
    n = o.n
    legal = o.legal()
    nxt = int(o.info()[1])
    cand = legal.copy()
    if avoid_eyes:
        cand &= 1 - o.true_eyes(nxt)
    idx = np.flatnonzero(cand)
    if len(idx) == 0:
        return n * n
    return int(rng.choice(idx))


@pytest.mark.parametrize("against", ["port", "reference"])
@pytest.mark.parametrize("n,G,plies", [(19, 64, 600), (9, 96, 200)])
def test_step_parity_every_ply(n, G, plies, against, oracle_lib):
    """GoState::forward for a batch: after EVERY ply compare hash, info words, stones, legal
    mask, and periodically score / eyes / features with the oracle -- the C restatement and the
    compiled UNMODIFIED reference (oracle/_ref) alike."""
    if against == "reference" and not oracles.have_ref(n):
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(1234 + n)
    gb = _gobatch(G, n)
    os_ = [oracles.Oracle(n, oracle_lib) if against == "port" else oracles.Ref(n) for _ in range(G)]
    for t in range(plies):
        acts = np.empty(G, np.int32)
        exp_ok = np.empty(G, bool)
        for g, o in enumerate(os_):
            r = rng.random()
            if o.terminated():
                a = n * n if r < 0.5 else int(rng.integers(0, n * n))
            elif r < 0.03:
                a = n * n  # pass
            elif r < 0.08:
                a = int(rng.integers(0, n * n))  # arbitrary, often illegal, point
            elif r < 0.10:
                a = -1  # untouched
            else:
                a = _random_candidate(rng, o, avoid_eyes=(r < 0.9))
            acts[g] = a
            exp_ok[g] = o.forward(a) if a >= 0 else False
        ok = gb.forward(acts)
        np.testing.assert_array_equal(ok, exp_ok, err_msg=f"ok flags at ply {t}")
        h = gb.getHashCode()
        info = gb.info()
        st = gb.stones()
        lg = gb.legal_mask()
        for g, o in enumerate(os_):
            assert int(h[g]) == o.hash(), f"hash g={g} t={t}"
            oi = o.info()
            oi[8] = 0  # ko_age is not part of the device state
            np.testing.assert_array_equal(info[g], oi, err_msg=f"info g={g} t={t}")
            np.testing.assert_array_equal(st[g], o.stones(), err_msg=f"stones g={g} t={t}")
            np.testing.assert_array_equal(lg[g, :-1], o.legal(), err_msg=f"legal g={g} t={t}")
            assert lg[g, -1] == 1
        if t % 37 == 5 or t == plies - 1:
            sc = gb.tt_score()
            ev = gb.evaluate(7.5)
            e0 = gb.true_eyes(0)
            e1 = gb.true_eyes(1)
            d4 = rng.integers(0, 8, G).astype(np.int32)
            ft = gb.features(d4)
            for g, o in enumerate(os_):
                assert sc[g] == o.tt_score()
                assert ev[g] == np.float32(o.evaluate(7.5))
                np.testing.assert_array_equal(e0[g], o.true_eyes(int(o.info()[1])))
                np.testing.assert_array_equal(e1[g], o.true_eyes(1))
                np.testing.assert_array_equal(ft[g], o.features(int(d4[g])), err_msg=f"features g={g} t={t}")
    gb.close()


@pytest.mark.parametrize("n,G,layout", [(19, 256, 0), (9, 300, 0), (19, 257, 1)])
def test_playout_matches_oracle(n, G, layout, oracle_lib):
    """whole device-resident playouts: per-game checksum (hash, captures, legal mask of every
    position), ply count, final score and hash identical to the oracle's.  layout 1 = k_playout2 (two
    board rows per lane, three games per warp; 257 games leave a partly filled last warp)."""
    gb = _gobatch(G, n)
    gb.set_playout_layout(layout)
    res = gb.playout(seed=77, first_game_id=1000)
    exp = oracles.oracle_playout_many(n, 77, 1000, G, lib=oracle_lib)
    np.testing.assert_array_equal(res["plies"], exp["plies"])
    np.testing.assert_array_equal(res["chk"], exp["chk"])
    np.testing.assert_array_equal(res["score"], exp["score"])
    assert res["total_plies"] == exp["total_plies"]
    gb.close()


def test_reset_mask_and_edge_cases(oracle_lib):
    n, G = 19, 5
    gb = _gobatch(G, n)
    # empty board: everything legal, hash 0, ply 1, black to move
    info = gb.info()
    assert (info[:, 0] == 1).all() and (info[:, 1] == 1).all()
    assert (gb.getHashCode() == 0).all()
    assert gb.legal_mask().all()
    assert (gb.features() [:, :16] == 0).all() and (gb.features()[:, 16] == 1).all()
    # out-of-range actions are rejected, game untouched
    ok = gb.forward(np.array([n * n + 1, 10**6, -5, 0, n * n], np.int32))
    assert ok.tolist() == [False, False, False, True, True]
    # playing on an occupied point is rejected
    ok = gb.forward(np.array([0, 0, 0, 0, 0], np.int32))
    assert ok.tolist() == [True, True, True, False, True]
    # two passes terminate; further moves are refused (go_state.cc:78)
    gb.forward(np.array([-1, -1, -1, -1, n * n], np.int32))
    gb.forward(np.array([-1, -1, -1, -1, n * n], np.int32))
    info = gb.info()
    assert info[4, 9] == 1 and info[4, 10] == 1
    ok = gb.forward(np.full(G, 5, np.int32))
    assert ok.tolist() == [True, True, True, True, False]
    # masked reset
    gb.reset(np.array([0, 0, 0, 0, 1], np.uint8))
    info = gb.info()
    assert info[4, 0] == 1 and info[4, 9] == 0 and info[0, 0] > 1
    gb.close()


# ---- committed golden fixtures (generated from the compiled reference) ---------------------
import json
import os

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


@pytest.mark.parametrize("n", [19, 9])
def test_golden_playouts(n):
    gold = _load(f"playouts_{n}.json")
    G = len(gold["games"])
    gb = _gobatch(G, n)
    res = gb.playout(seed=gold["seed"], first_game_id=0)
    for e in gold["games"]:
        g = e["game_id"]
        assert (int(res["plies"][g]), f"{int(res['chk'][g]):016x}", int(res["score"][g])) == (e["plies"], e["chk"], e["score"])
    gb.close()


@pytest.mark.parametrize("n", [19, 9])
def test_golden_traces_and_positions(n):
    """replay the reference's own move lists through the step API: hash and captures after every
    ply, and the full observable state at the fixture positions."""
    games = [e for e in _load(f"playouts_{n}.json")["games"] if "moves" in e]
    positions = _load(f"positions_{n}.json")["positions"]
    gb = _gobatch(len(games), n)
    T = max(e["plies"] for e in games)
    for t in range(T):
        acts = np.array([e["moves"][t] if t < e["plies"] else -1 for e in games], np.int32)
        ok = gb.forward(acts)
        h = gb.getHashCode()
        info = gb.info()
        for g, e in enumerate(games):
            if t < e["plies"]:
                assert ok[g]
                assert f"{int(h[g]):016x}" == e["hashes"][t]
                assert info[g, 2:4].tolist() == e["caps"][t]
        here = [p for p in positions if p["after_ply"] == t + 1]
        if here:
            st, lg, sc, ev = gb.stones(), gb.legal_mask(), gb.tt_score(), gb.evaluate(7.5)
            eb, ew = gb.true_eyes(1), gb.true_eyes(2)
            feats = {d4: gb.features(np.full(len(games), d4, np.int32)) for d4 in (0, 3, 5, 6)}
            for p in here:
                g = p["game_id"]
                if p["after_ply"] > games[g]["plies"]:
                    continue
                gi = info[g].tolist()
                pi = list(p["info"])
                pi[8] = 0
                assert gi == pi
                assert st[g].tolist() == p["stones"]
                assert lg[g, :-1].tolist() == p["legal"]
                assert np.flatnonzero(eb[g]).tolist() == p["eyes_black"]
                assert np.flatnonzero(ew[g]).tolist() == p["eyes_white"]
                assert sc[g] == p["tt_score"] and ev[g] == np.float32(p["evaluate_7_5"])
                for d4, ones in p["features_ones"].items():
                    assert np.flatnonzero(feats[int(d4)][g].reshape(-1)).tolist() == ones
    gb.close()


def test_full_size_playout_properties(oracle_lib):
    """BASELINE config 2 at full size (4096 games): size-independent properties + a sampled
    exact check against the oracle."""
    n, G = 19, 4096
    gb = _gobatch(G, n)
    a = gb.playout(seed=11, first_game_id=0)
    b = gb.playout(seed=11, first_game_id=0)
    for k in ("chk", "plies", "score", "hash"):
        np.testing.assert_array_equal(a[k], b[k])  # deterministic / idempotent
    assert (a["plies"] > 100).all() and (a["plies"] <= 2 * n * n).all()
    assert a["total_plies"] == int(a["plies"].sum())
    assert (np.abs(a["score"]) <= n * n).all()
    # different game ids give different games
    c = gb.playout(seed=11, first_game_id=G)
    assert (a["chk"] != c["chk"]).mean() > 0.999
    # shifting the id window reproduces the overlap exactly
    d = gb.playout(seed=11, first_game_id=100)
    np.testing.assert_array_equal(d["chk"][: G - 100], a["chk"][100:])
    # sampled exact parity
    idx = np.linspace(0, G - 1, 48).astype(int)
    for g in idx:
        t, chk, sc = oracles.oracle_playout(n, 11, int(g), lib=oracle_lib)
        assert (t, chk, sc) == (int(a["plies"][g]), int(a["chk"][g]), int(a["score"][g]))
    gb.close()


@pytest.mark.parametrize("n,G,budget,layout", [(19, 512, 700, 0), (9, 510, 400, 0), (19, 64, 37, 0), (19, 511, 700, 1)])
def test_stream_playout_matches_oracle(n, G, budget, layout, oracle_lib):
    """steady-state mode: every slot plays exactly `budget` plies across consecutive games; the
    per-slot fold of game checksums, the ply count and the number of games equal the oracle's."""
    gb = _gobatch(G, n)
    gb.set_playout_layout(layout)
    res = gb.playout_stream(seed=5, first_game_id=70, plies_per_slot=budget)
    assert (res["plies"] == budget).all() and res["total_plies"] == budget * G
    for slot in np.linspace(0, G - 1, 40).astype(int):
        t, acc, games = oracles.oracle_playout_stream(n, 5, 70, int(slot), G, budget, lib=oracle_lib)
        assert t == budget
        assert (int(res["chk"][slot]), int(res["games"][slot])) == (acc, games), f"slot {slot}"
    if budget > 2 * n * n // 2:
        assert (res["games"] >= 2).any()
    gb.close()


def test_config5_small_board_stress(oracle_lib):
    """BASELINE config 5: 9x9, 16384 concurrent games (three games per warp): deterministic,
    ply counts in range, sampled exact parity with the oracle, stream mode consistent."""
    n, G = 9, 16384
    gb = _gobatch(G, n)
    a = gb.playout(seed=9, first_game_id=0)
    b = gb.playout(seed=9, first_game_id=0)
    np.testing.assert_array_equal(a["chk"], b["chk"])
    assert (a["plies"] > 20).all() and (a["plies"] <= 2 * n * n).all()
    for g in np.linspace(0, G - 1, 64).astype(int):
        t, chk, sc = oracles.oracle_playout(n, 9, int(g), lib=oracle_lib)
        assert (t, chk, sc) == (int(a["plies"][g]), int(a["chk"][g]), int(a["score"][g]))
    s = gb.playout_stream(seed=9, first_game_id=0, plies_per_slot=300)
    assert (s["plies"] == 300).all() and (s["games"] >= 2).all()
    for slot in (0, 5, 16383):
        t, acc, games = oracles.oracle_playout_stream(n, 9, 0, slot, G, 300, lib=oracle_lib)
        assert (acc, games) == (int(s["chk"][slot]), int(s["games"][slot]))
    gb.close()


def _ref_playouts_all_threads(n, seed, first, count):
    """`count` playouts of the compiled reference (oracle/_ref: ref_playout drives the unmodified
    GoState), spread over the host threads (ctypes releases the GIL during the call)"""
    import ctypes
    import os
    from concurrent.futures import ThreadPoolExecutor

    L = oracles.load_ref(n)
    chk = np.zeros(count, np.uint64)
    plies = np.zeros(count, np.int32)
    score = np.zeros(count, np.int32)

    def work(lo, hi):
        c, s = ctypes.c_uint64(), ctypes.c_int32()
        for i in range(lo, hi):
            plies[i] = L.ref_playout(seed, first + i, 2 * n * n, None, None, None, ctypes.byref(c), ctypes.byref(s))
            chk[i], score[i] = c.value, s.value

    nt = max(1, min(64, len(os.sched_getaffinity(0))))
    step = -(-count // (nt * 8))
    with ThreadPoolExecutor(nt) as ex:
        list(ex.map(lambda lo: work(lo, min(lo + step, count)), range(0, count, step)))
    return chk, plies, score


@pytest.mark.parametrize("n,G,batches,layout", [(19, 4096, 3, 0), (9, 16384, 1, 0), (19, 4096, 3, 1)])
def test_ten_thousand_playouts_bit_exact_vs_compiled_reference(n, G, batches, layout):
    """north_star parity bar: >= 10k random playouts, GAME BY GAME against the reference C++ itself
    (not the restatement): the per-game checksum folds hash, both capture counts, side to move and the
    full legal mask of EVERY position, so equality means those were bit-identical at every ply; plus
    ply count and final Tromp-Taylor score.  19x19: 3 x 4096 = 12,288 games (BASELINE config 2's batch,
    three times); 9x9: 16,384 games (config 5's batch)."""
    if not oracles.have_ref(n):
        pytest.skip("oracle/_ref not built")
    gb = _gobatch(G, n)
    gb.set_playout_layout(layout)  # 1: the two-rows-per-lane kernel (k_playout2)
    total = 0
    for b in range(batches):
        first = 5_000_000 + (b + 3 * layout) * G
        res = gb.playout(seed=2024, first_game_id=first)
        chk, plies, score = _ref_playouts_all_threads(n, 2024, first, G)
        bad = np.flatnonzero((res["chk"] != chk) | (res["plies"] != plies) | (res["score"] != score))
        assert bad.size == 0, f"{bad.size} of {G} games differ from the reference, first: game id {first + int(bad[0])}"
        total += int(plies.sum())
    print(f"{batches * G} playouts ({total} positions) bit-exact vs the compiled reference, {n}x{n}")
    gb.close()


@pytest.mark.parametrize("n,G", [(19, 37), (9, 50)])
def test_feature_formats_and_store_modes(n, G):
    """the feature writer's fast formats and both ways a staged tile leaves shared memory (bulk/TMA
    store, vector stores) produce the planes of elfb200_features (oracle-checked above): float32
    NCHW incl. an odd batch and an 8-byte-aligned destination, binary16 / bfloat16 NHWC with 24 or 32
    channels (zeros above plane 17)"""
    import torch

    from elf_b200 import lib as L

    gb = _gobatch(G, n)
    rng = np.random.default_rng(n)
    os_ = [oracles.Oracle(n) for _ in range(G)]
    for t in range(70):
        acts = np.array([_random_candidate(rng, o) for o in os_], np.int32)
        for o, a in zip(os_, acts):
            o.forward(int(a))
        gb.forward(acts)
    d4 = rng.integers(0, 8, G).astype(np.int32)
    want = gb.features(d4)  # host copy of the float32 planes
    for g in (0, G // 2, G - 1):
        np.testing.assert_array_equal(want[g], os_[g].features(int(d4[g])))
    dev = torch.device("cuda", gb.device)
    stream = torch.cuda.ExternalStream(gb.stream, device=dev)
    d4_dev = torch.from_numpy(d4).to(dev)
    for mode in (1, 0):
        gb.set_feature_store(mode)
        out = torch.full((G + 1, 18, n, n), -7.0, device=dev)
        gb.features_dev(out.data_ptr(), d4_dev.data_ptr())
        gb.synchronize()
        np.testing.assert_array_equal(out[:G].cpu().numpy(), want)
        assert (out[G] == -7.0).all()  # odd batch: the tail tile must not spill
        out.fill_(-7.0)
        gb.features_dev(out[1:].data_ptr(), d4_dev.data_ptr())  # destination only 8-byte aligned
        gb.synchronize()
        np.testing.assert_array_equal(out[1:].cpu().numpy(), want)
        assert (out[0] == -7.0).all()
        for fmt, dt in ((L.FEAT_F16_NHWC, torch.float16), (L.FEAT_BF16_NHWC, torch.bfloat16)):
            for cpad in (24, 32):
                o16 = torch.full((G + 1, n, n, cpad), 5.0, dtype=dt, device=dev)
                gb.features_dev(o16.data_ptr(), d4_dev.data_ptr(), fmt, cpad)
                gb.synchronize()
                got = o16[:G].float().permute(0, 3, 1, 2).cpu().numpy()
                np.testing.assert_array_equal(got[:, :18], want, err_msg=f"mode {mode} fmt {fmt} cpad {cpad}")
                assert (got[:, 18:] == 0).all() and (o16[G] == 5.0).all()
    del stream
    gb.close()


@pytest.mark.parametrize("n,G", [(9, 40), (19, 24)])
def test_darkforest_features_vs_compiled_reference(n, G):
    """row f4 remainder: BoardFeature::extract (the 25 DarkForest planes, GameOptions::use_df_feature) on the
    device == the compiled reference's, plane by plane, under random D4 codes, along games with captures and
    kos; the placement plies behind the history planes survive elfb200_replay"""
    if not oracles.have_ref(n):
        pytest.skip("oracle/_ref not built")
    gb = _gobatch(G, n)
    refs = [oracles.Ref(n) for _ in range(G)]
    rng = np.random.default_rng(100 + n)
    lists = [[] for _ in range(G)]
    plies = 75 if n == 9 else 260
    for t in range(plies):
        acts = np.empty(G, np.int32)
        for g, r in enumerate(refs):
            idx = np.flatnonzero(r.legal() & (1 - r.true_eyes(int(r.info()[1]))))
            acts[g] = int(rng.choice(idx)) if len(idx) else n * n
            assert r.forward(acts[g])
            lists[g].append(int(acts[g]))
        assert gb.forward(acts).all()
        if t % 17 == 3 or t >= plies - 3:
            d4 = rng.integers(0, 8, G).astype(np.int32)
            got = gb.features_df(d4)
            for g, r in enumerate(refs):
                want = r.features_df(int(d4[g]))
                for pl in range(25):
                    np.testing.assert_array_equal(got[g, pl], want[pl], err_msg=f"plane {pl} game {g} ply {t} d4 {d4[g]}")
    assert sum(int(r.info()[2] + r.info()[3]) for r in refs) > G  # plenty of captures
    gb2 = _gobatch(G, n)
    gb2.replay(lists)
    np.testing.assert_array_equal(gb2.features_df(), gb.features_df())
    gb.close()
    gb2.close()
