"""GPU parity of the batched tree search against the UNMODIFIED reference search
(elf::ai::tree_search::MCTSAI_T + the reference Go actor logic, through oracle/_ref) with one
search thread, fixed rollouts per batch, rotation_flip off and the deterministic fake net
(oracle/fakenet.h): root visit counts must agree within +-1 per edge (BASELINE.json north_star)."""
import numpy as np
import pytest

from tests import oracles

pytestmark = pytest.mark.gpu


def _fake_actor(mcts, n):
    import torch

    def actor(batch):
        h, _, _ = mcts.leaf_info()
        pi, v = oracles.fakenet(h, n * n + 1)
        return {"pi": torch.from_numpy(pi).to(mcts.device), "V": torch.from_numpy(v).to(mcts.device)}

    return actor


def _run_parity(n, G, moves, opts, open_plies, tol=1):
    import elf_b200

    if not oracles.have_ref(n):
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(5 + n)
    gb = elf_b200.GoBatch(G, board_size=n)
    refs = [oracles.Ref(n) for _ in range(G)]
    # distinct openings
    for t in range(open_plies):
        acts = np.empty(G, np.int32)
        for g, r in enumerate(refs):
            idx = np.flatnonzero(r.legal())
            acts[g] = int(rng.choice(idx))
            assert r.forward(acts[g])
        assert gb.forward(acts).all()
    mc = elf_b200.MctsBatch(gb, rotation_flip=0, **opts)
    rms = [oracles.RefMcts(n, **opts) for _ in range(G)]
    actor = _fake_actor(mc, n)
    worst = 0
    for mv in range(moves):
        res = mc.act(actor)
        acts = np.empty(G, np.int32)
        for g in range(G):
            rr = rms[g].act(refs[g])
            gv, rv = res["visits"][g], rr["visits"]
            assert ((gv >= 0) == (rv >= 0)).all(), f"edge sets differ: move {mv} game {g}"
            d = np.abs(gv - rv)[rv >= 0]
            worst = max(worst, int(d.max()))
            assert d.max() <= tol, f"visits differ by {d.max()} at move {mv} game {g}: gpu {gv[rv>=0][d.argmax()]} ref {rv[rv>=0][d.argmax()]}"
            assert res["total_visits"][g] == rr["total_visits"]
            assert abs(res["root_value"][g] - rr["root_value"]) < 1e-6
            acts[g] = rr["best_action"]
            assert refs[g].forward(acts[g])
        assert gb.forward(acts).all()
        mc.advance(acts)
    assert (mc.errors() == 0).all(), mc.errors()
    # same number of network evaluations as the reference (no extra / missing expansions)
    assert mc.eval_count() == sum(r.num_evals() for r in rms)
    mc.close()
    gb.close()
    return worst


def test_mcts_parity_19_persistent():
    _run_parity(19, G=6, moves=6, open_plies=6,
                opts=dict(num_rollouts=200, num_rollouts_per_batch=8, virtual_loss=1, persistent_tree=1, c_puct=1.5))


def test_mcts_parity_19_fresh_tree_batch1():
    _run_parity(19, G=4, moves=3, open_plies=30,
                opts=dict(num_rollouts=128, num_rollouts_per_batch=1, virtual_loss=0, persistent_tree=0, c_puct=0.85))


def test_mcts_parity_9_endgame():
    # 9x9 late in the game: terminal leaves, passes, pass suppression, superko inside the tree
    _run_parity(9, G=8, moves=12, open_plies=60,
                opts=dict(num_rollouts=160, num_rollouts_per_batch=4, virtual_loss=2, persistent_tree=1, c_puct=1.5))
