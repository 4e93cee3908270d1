"""Differential fuzz of the reference-stream mode (root Dirichlet noise, one D4 code per evaluated leaf,
sampled moves, never-resign draw, container-order ties) on the emulated search kernels against game
threads composed from the COMPILED reference's own pieces (tests/test_refstream.py::play_reference_game).
Test infrastructure; needs oracle/_ref, no GPU.  `python scripts/emu_fuzz_streams.py --seed 1 --cases 20`"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from elf_b200.refstream import RefStream  # noqa: E402
from tests import emu, oracles  # noqa: E402
from tests.test_refstream import plane_actor, play_reference_game  # noqa: E402


def one_case(rng, n, case):
    opts = dict(num_rollouts=int(rng.integers(4, 60 if n == 9 else 24)), num_rollouts_per_batch=int(rng.integers(1, 9)),
                virtual_loss=int(rng.integers(0, 4)), persistent_tree=int(rng.random() < 0.8),
                c_puct=float(rng.choice([0.5, 1.5, 5.0])), komi=float(rng.choice([5.5, 7.5])),
                ply_pass_enabled=int(rng.choice([0, 30])), unexplored_q_zero=int(rng.integers(0, 2)),
                root_unexplored_q_zero=int(rng.integers(0, 2)))
    eps = float(rng.choice([0.0, 0.25, 0.5]))
    alpha = float(rng.choice([0.03, 0.3, 1.0]))
    flip = int(rng.integers(0, 2))
    cutoff = int(rng.choice([-1, 6, 30, 400]))
    thres, ratio = float(rng.choice([0.05, 0.3, 0.9])), float(rng.choice([0.0, 0.1, 1.0]))
    moves = int(rng.integers(6, 70 if n == 9 else 10))
    G = int(rng.integers(1, 4))
    seeds = [int(x) for x in rng.integers(1, 2**31, G)]
    tag = f"case {case}: n={n} G={G} moves={moves} eps={eps} alpha={alpha} flip={flip} cutoff={cutoff} thres={thres} ratio={ratio} {opts}"
    logs = [play_reference_game(n, s, opts, eps, alpha, flip, cutoff, thres, ratio, moves) for s in seeds]
    emu.emu_lib().simt_emu_set_order(case % 3)
    gb = emu.emu_batch(G, n)
    mc = emu.EmuSearch(gb, rotation_flip=flip, **opts)
    rs = RefStream(G, n, np.array(seeds, np.uint64))
    rs.init_actor(0)
    mc.attach_ref_stream(rs, 0, eps, alpha)
    actor = plane_actor(n)
    alive, drawn, never = np.ones(G, bool), np.zeros(G, bool), np.zeros(G, bool)
    for t in range(moves):
        info = gb.info()
        mc.search(actor, active=alive.astype(np.uint8))
        if mc.errors()[3]:
            print("ok (pruned)", tag, flush=True)
            return True
        res = mc.results()
        cho = mc.ref_choose(sample=(info[:, 0] <= cutoff), mask=alive.astype(np.uint8))
        need = alive & ~drawn
        u = rs.game_uniform(need.astype(np.uint8))
        never[need] = u[need] < float(np.float32(ratio))
        drawn |= need
        side = np.where(info[:, 1] == 1, cho["value"], -cho["value"]).astype(np.float64)
        resign = alive & ~never & ~(side >= -1.0 + float(np.float32(thres))) & (info[:, 0] >= 50)
        acts = np.full(G, -2, np.int32)
        for g in np.flatnonzero(alive):
            L = logs[g][t]
            bad = None
            if not np.array_equal(res["visits"][g], L["visits"]):
                bad = "visits"
            elif cho["best_action"][g] != L["best"] or cho["action"][g] != L["action"]:
                bad = f"move {cho['best_action'][g]}/{cho['action'][g]} vs {L['best']}/{L['action']}"
            elif not np.isclose(cho["value"][g], L["value"], rtol=2e-6, atol=1e-7, equal_nan=True):
                bad = f"value {cho['value'][g]} vs {L['value']}"
            elif bool(resign[g]) != L["resign"]:
                bad = "resign"
            if bad:
                print("MISMATCH", bad, "game", g, "move", t, "seed", seeds[g], tag, flush=True)
                return False
            acts[g] = cho["action"][g]
            if t + 1 == len(logs[g]):
                alive[g] = False
        gb.forward(acts)
        mc.advance(acts)
        if not alive.any():
            break
    print("ok", tag, flush=True)
    return True


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cases", type=int, default=10)
    ap.add_argument("--board", type=int, default=9)
    ap.add_argument("--only", type=int, default=-1, help="run just this case index")
    a = ap.parse_args()
    if not oracles.have_ref(a.board):
        sys.exit("oracle/_ref is not built")
    bad = 0
    for c in (range(a.cases) if a.only < 0 else [a.only]):
        bad += not one_case(np.random.default_rng([a.seed, c]), a.board, c)
    print("cases", a.cases, "mismatches", bad)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
