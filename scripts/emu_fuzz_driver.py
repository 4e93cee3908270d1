"""Differential fuzz of the per-move DRIVER (elf_b200.selfplay.SelfPlay, rng="reference") on the emulated
kernels against game threads composed from the COMPILED reference's pieces the way GoGameSelfPlay::act
chains them (game_selfplay.cc:272-430): search or policy-only move, sampled move while ply <= cutoff,
shouldResign (never-resign draw at a game's first move, resign only from ply 50), forward, game end on two
passes / superko / ply cap / move_cutoff, finish_game (tree reset, ResignCheck reset) and the next game on
the same generators.  Random option sets; the sequence of moves (-1 = resign) must be identical.
Test infrastructure; needs oracle/_ref, no GPU.  `python scripts/emu_fuzz_driver.py --seed 1 --cases 20`"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import emu, oracles  # noqa: E402
from tests.test_refstream import plane_actor, ref_net  # noqa: E402


def reference_thread(n, seed, opts, eps, alpha, flip, cutoff, thres, ratio, move_cutoff, po_colour, quant, moves, two):
    g = oracles.RefRng(n, int(seed))
    mk = lambda: oracles.RefMcts(n, callback=ref_net(n, quant), root_epsilon=eps, root_alpha=alpha, rotation_flip=flip,
                                 seed=g.next(), **opts)
    ai = mk()  # init_ai(_ai), then init_ai(_ai2): the seeds come off the game generator in that order
    ai2 = mk() if two else None
    rc = oracles.RefResign(n, thres, ratio)
    st = oracles.Ref(n)
    played = []

    def finish():
        nonlocal st
        ai.end_game(st)
        if ai2 is not None:
            ai2.end_game(st)
        st = oracles.Ref(n)
        rc.reset()

    for _ in range(moves):
        ply, nxt = int(st.info()[0]), int(st.info()[1])
        ref = ai2 if (ai2 is not None and nxt == 2) else ai  # game_selfplay.cc:366-367
        if nxt == po_colour:
            r = ref.act(st, policy_only=True)
            a = r["best_action"]
        else:
            r = ref.act(st)
            a = ref.sample(g) if ply <= cutoff else r["best_action"]
        if rc.check(r["best_q"], nxt, g) and ply >= 50:
            played.append(-1)
            finish()
            continue
        assert st.forward(int(a))
        played.append(int(a))
        if st.info()[9] or (move_cutoff > 0 and int(st.info()[0]) >= move_cutoff):
            finish()
    return played


def one_case(rng, n, case):
    from elf_b200.selfplay import SelfPlay

    opts = dict(num_rollouts=int(rng.integers(4, 40 if n == 9 else 16)), num_rollouts_per_batch=int(rng.integers(1, 9)),
                virtual_loss=int(rng.integers(0, 3)), persistent_tree=int(rng.random() < 0.8),
                c_puct=float(rng.choice([0.5, 1.5, 5.0])), komi=float(rng.choice([5.5, 7.5])),
                ply_pass_enabled=int(rng.choice([0, 20, 60])))
    eps, alpha = float(rng.choice([0.0, 0.25])), float(rng.choice([0.03, 0.3]))
    flip = int(rng.integers(0, 2))
    cutoff = int(rng.choice([-1, 4, 20, 400]))
    thres, ratio = float(rng.choice([0.05, 0.5, 0.95, 1.5])), float(rng.choice([0.0, 0.3, 1.0]))
    move_cutoff = int(rng.choice([-1, 7, 30, 70]))
    po_colour = int(rng.choice([0, 0, 1, 2]))  # 0: both colours search
    quant = int(rng.choice([0, 0, 256]))
    moves = int(rng.integers(10, 140 if n == 9 else 14))
    G = int(rng.integers(1, 4))
    seeds = np.array([int(x) for x in rng.integers(1, 2**31, G)], np.uint64)
    two = bool(rng.random() < 0.3)  # an evaluation match: a second AI with its own tree plays white
    tag = (f"case {case}: n={n} G={G} moves={moves} eps={eps} alpha={alpha} flip={flip} cutoff={cutoff} thres={thres} "
           f"ratio={ratio} move_cutoff={move_cutoff} policy_only={po_colour} quant={quant} two_models={two} {opts}")
    expect = [reference_thread(n, s, opts, eps, alpha, flip, cutoff, thres, ratio, move_cutoff, po_colour, quant, moves, two)
              for s in seeds]
    emu.emu_lib().simt_emu_set_order(case % 3)
    gb = emu.emu_batch(G, n)
    mc = emu.EmuSearch(gb, rotation_flip=flip, std_sort_ties=int(quant > 0), **opts)
    mc2 = emu.EmuSearch(gb, rotation_flip=flip, std_sort_ties=int(quant > 0), **opts) if two else None
    sp = SelfPlay(plane_actor(n, quant), actor_white=plane_actor(n, quant) if two else None, search_white=mc2,
                  num_games=G, board_size=n, board=gb, search=mc, rng="reference", seed=seeds,
                  policy_distri_cutoff=cutoff, resign_thres=thres, never_resign_ratio=ratio, move_cutoff=move_cutoff,
                  root_epsilon=eps, root_alpha=alpha, black_use_policy_network_only=po_colour == 1,
                  white_use_policy_network_only=po_colour == 2, **opts)
    got = [[] for _ in range(G)]
    fwd = gb.forward

    def logged(acts):
        for g in range(G):
            got[g].append(int(acts[g]))
        return fwd(acts)

    gb.forward = logged
    for t in range(moves):
        sp.step()
        if mc.errors()[3] or (mc2 is not None and mc2.errors()[3]):
            print("ok (pruned)", tag, flush=True)
            return True
    if got != expect:
        for g in range(G):
            if got[g] != expect[g]:
                k = next(i for i, (x, y) in enumerate(zip(got[g], expect[g])) if x != y)
                print("MISMATCH game", g, "seed", int(seeds[g]), "at move", k, "ours", got[g][k], "ref", expect[g][k], tag, flush=True)
                break
        return False
    print("ok", f"games finished {sp.games_finished}", tag, flush=True)
    return True


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cases", type=int, default=10)
    ap.add_argument("--board", type=int, default=9)
    ap.add_argument("--only", type=int, default=-1)
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
