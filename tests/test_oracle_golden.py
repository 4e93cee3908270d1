"""The oracle (C restatement) against the committed golden fixtures generated from the compiled
reference by scripts/gen_golden.py.  Runs anywhere (no reference tree, no GPU)."""
import json
import os

import numpy as np
import pytest

from tests import oracles

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


@pytest.mark.parametrize("n", [19, 9])
def test_playout_fixtures(n, oracle_lib):
    gold = load(f"playouts_{n}.json")
    for e in gold["games"]:
        if "moves" in e:
            t, chk, sc, moves, hashes, caps = oracles.oracle_playout(n, gold["seed"], e["game_id"], trace=True, lib=oracle_lib)
            assert moves.tolist() == e["moves"]
            assert [f"{int(h):016x}" for h in hashes] == e["hashes"]
            assert caps.tolist() == e["caps"]
        else:
            t, chk, sc = oracles.oracle_playout(n, gold["seed"], e["game_id"], lib=oracle_lib)
        assert (t, f"{chk:016x}", sc) == (e["plies"], e["chk"], e["score"])


@pytest.mark.parametrize("n", [19, 9])
def test_position_fixtures(n, oracle_lib):
    gold = load(f"positions_{n}.json")
    games = {e["game_id"]: e for e in load(f"playouts_{n}.json")["games"] if "moves" in e}
    for pos in gold["positions"]:
        o = oracles.Oracle(n, oracle_lib)
        for a in games[pos["game_id"]]["moves"][: pos["after_ply"]]:
            assert o.forward(a)
        assert f"{o.hash():016x}" == pos["hash"]
        assert o.info().tolist() == pos["info"]
        assert o.stones().tolist() == pos["stones"]
        assert o.legal().tolist() == pos["legal"]
        assert np.flatnonzero(o.true_eyes(1)).tolist() == pos["eyes_black"]
        assert np.flatnonzero(o.true_eyes(2)).tolist() == pos["eyes_white"]
        assert o.tt_score() == pos["tt_score"]
        assert o.evaluate(7.5) == pos["evaluate_7_5"]
        for d4, ones in pos["features_ones"].items():
            assert np.flatnonzero(o.features(int(d4)).reshape(-1)).tolist() == ones
