"""Online mode with the real engine (GoBatch of one game + MctsBatch): the moves `genmove` plays
must be the ones the search restatement (oracle/mcts_oracle.c) chooses on the same positions.

The host logic is also covered on CPU (tests/test_online_console.py); this is the G = 1 combination
(human moves interleaved with searches) on the device (first green hardware run: round 1's final
GPU test pass)."""
import numpy as np
import pytest

from tests import oracles

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def test_online_game_matches_search_restatement(oracle_lib):
    import torch

    from elf_b200 import console, online

    n = 9
    opts = dict(num_rollouts=96, num_rollouts_per_batch=8, c_puct=1.5, virtual_loss=1, persistent_tree=1)  # rotation_flip off below: the fake net is keyed by hash
    g = online.OnlineGame.create(board_size=n, rotation_flip=0, **opts)

    def actor(batch):
        h, _, _ = g.search.leaf_info()
        pi, v = oracles.fakenet(h, n * n + 1)
        assert batch["s"].shape[0] == len(h)
        return {"pi": torch.from_numpy(pi).to(g.search.device), "V": torch.from_numpy(v).to(g.search.device)}

    o = oracles.Oracle(n, oracle_lib)
    om = oracles.OracleMcts(n, lib=oracle_lib, **opts)
    c = console.GtpConsole(g, actor)
    rng = np.random.default_rng(3)
    for t in range(12):
        who = "b" if int(o.info()[1]) == 1 else "w"
        if t % 3 == 2:  # an operator move in between: the tree has to follow it
            a = int(rng.choice(np.flatnonzero(o.legal())))
            assert c.execute(f"play {who} {online.action2vertex(a, n)}") == "=\n\n"
        else:
            a = om.act(o)["best_action"]
            assert c.execute(f"genmove {who}") == f"= {online.action2vertex(a, n)}\n\n"
        assert o.forward(a)
        assert int(g.board.getHashCode()[0]) == o.hash()
    assert "Last move" in c.execute("showboard") and g.seq == 0
    assert g.search.errors().sum() == 0
