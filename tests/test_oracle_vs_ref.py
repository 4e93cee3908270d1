"""Pin the C restatement (oracle/go_oracle.c) against the compiled UNMODIFIED reference
(oracle/_ref, built by oracle/Makefile from /root/reference).  Skipped when oracle/_ref is
absent (e.g. a checkout without the reference tree)."""
import numpy as np
import pytest

from tests import oracles


def _need_ref(n):
    if not oracles.have_ref(n):
        pytest.skip("oracle/_ref not built (needs /root/reference)")


@pytest.mark.parametrize("n,games,plies", [(19, 6, 760), (9, 12, 180)])
def test_every_ply_equal(n, games, plies, oracle_lib):
    _need_ref(n)
    rng = np.random.default_rng(99 + n)
    for g in range(games):
        o = oracles.Oracle(n, oracle_lib)
        r = oracles.Ref(n)
        for t in range(plies):
            u = rng.random()
            legal = r.legal()
            nxt = int(r.info()[1])
            if r.terminated():
                a = int(rng.integers(0, n * n + 1))
            elif u < 0.03:
                a = n * n
            elif u < 0.08:
                a = int(rng.integers(0, n * n))
            else:
                cand = legal & (1 - r.true_eyes(nxt)) if u < 0.9 else legal
                idx = np.flatnonzero(cand)
                a = int(rng.choice(idx)) if len(idx) else n * n
            assert o.forward(a) == r.forward(a), f"forward g={g} t={t} a={a}"
            assert o.hash() == r.hash(), f"hash g={g} t={t}"
            np.testing.assert_array_equal(o.info(), r.info(), err_msg=f"info g={g} t={t}")
            np.testing.assert_array_equal(o.stones(), r.stones())
            np.testing.assert_array_equal(o.legal(), r.legal(), err_msg=f"legal g={g} t={t}")
            if t % 7 == 0:
                for pl in (1, 2):
                    np.testing.assert_array_equal(o.true_eyes(pl), r.true_eyes(pl))
                assert o.tt_score() == r.tt_score()
                assert o.evaluate(7.5) == r.evaluate(7.5)
                d4 = int(rng.integers(0, 8))
                np.testing.assert_array_equal(o.features(d4), r.features(d4), err_msg=f"features d4={d4}")
                st = r.stones()
                for a2 in np.flatnonzero(st)[:6]:
                    assert oracles.oracle_group(o, int(a2)) == oracles.ref_group(r, int(a2))
                assert oracles.oracle_num_groups(o) == oracles.ref_num_groups(r)


@pytest.mark.parametrize("n,count", [(19, 40), (9, 200)])
def test_playout_checksums_equal(n, count, oracle_lib):
    _need_ref(n)
    exp = oracles.oracle_playout_many(n, 5, 300, count, lib=oracle_lib)
    for i in range(count):
        t, chk, sc = oracles.ref_playout(n, 5, 300 + i)
        assert (t, chk, sc) == (int(exp["plies"][i]), int(exp["chk"][i]), int(exp["score"][i])), f"game {i}"


@pytest.mark.parametrize("n", [19, 9])
def test_d4_action_maps_equal(n, oracle_lib):
    _need_ref(n)
    R = oracles.load_ref(n)
    for d4 in range(8):
        for a in list(range(0, n * n + 1, 7)) + [n * n, n * n - 1]:
            assert oracle_lib.go_d4_action2action(n, d4, a) == R.ref_d4_action2action(d4, a)
