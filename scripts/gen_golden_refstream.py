#!/usr/bin/env python
"""Generate tests/golden/refstream_games.json from the compiled UNMODIFIED reference (oracle/_ref):
self-play games of reference game threads with GameOptions::seed set -- root Dirichlet noise, one D4
code per evaluated leaf, moves sampled while ply <= policy_distri_cutoff, the never-resign draw,
game ends by move_cutoff -- composed from the reference's own pieces in the order
GoGameSelfPlay::act uses them (as tests/test_refstream.py does live).  The network is
tests/oracles.feature_net (a fixed function of the planes).  The fixture lets the stream tests run
where the compiled reference is not available.

Run in the CPU container after `make -C oracle ref`."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import oracles  # noqa: E402

CASES = [
    dict(n=9, seeds=[777, 778], moves=26, cutoff=8, move_cutoff=11, eps=0.25, alpha=0.3, flip=1, thres=0.05, ratio=0.1,
         opts=dict(num_rollouts=24, num_rollouts_per_batch=4, virtual_loss=1, persistent_tree=1, c_puct=1.5, komi=7.5)),
    dict(n=19, seeds=[31337], moves=6, cutoff=30, move_cutoff=-1, eps=0.25, alpha=0.03, flip=1, thres=0.05, ratio=0.1,
         opts=dict(num_rollouts=16, num_rollouts_per_batch=8, virtual_loss=1, persistent_tree=1, c_puct=1.5, komi=7.5)),
]


def play(c, seed):
    n = c["n"]
    net = lambda feats, hashes: oracles.feature_net(feats, n * n + 1)  # noqa: E731
    g = oracles.RefRng(n, seed)
    ref = oracles.RefMcts(n, callback=net, root_epsilon=c["eps"], root_alpha=c["alpha"], rotation_flip=c["flip"],
                          seed=g.next(), **c["opts"])
    rc = oracles.RefResign(n, c["thres"], c["ratio"])
    st = oracles.Ref(n)
    played, tops = [], []
    for _ in range(c["moves"]):
        ply = int(st.info()[0])
        r = ref.act(st)
        a = ref.sample(g) if ply <= c["cutoff"] else r["best_action"]
        rc.check(r["best_q"], int(st.info()[1]), g)
        assert st.forward(int(a))
        played.append(int(a))
        tops.append([int(r["best_action"]), int(r["visits"].max()), int(r["total_visits"])])
        if st.info()[9] or (c["move_cutoff"] > 0 and int(st.info()[0]) >= c["move_cutoff"]):
            ref.end_game(st)
            st = oracles.Ref(n)
            rc.reset()
    return {"seed": seed, "actions": played, "best_maxvisits_total": tops}


def main():
    out = []
    for c in CASES:
        out.append({**c, "games": [play(c, s) for s in c["seeds"]]})
    path = os.path.join(ROOT, "tests", "golden", "refstream_games.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
