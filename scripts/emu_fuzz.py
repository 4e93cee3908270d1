"""Differential fuzz of the SEARCH KERNELS (built for the SIMT emulator, tests/simt_emu) against the
COMPILED reference search (oracle/_ref): random option sets, random openings of any length, several
moves on a persistent tree, terminated games left inactive.  Test infrastructure: needs no GPU and is
not part of the product.  `python scripts/emu_fuzz.py --seed 1 --cases 40 [--board 19]`; prints one
line per case and exits 1 at the first difference (the case is reproducible from seed + index)."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import emu, oracles  # noqa: E402


QUANT = 0  # --quant Q: network probabilities on a grid of 1/Q (bit-equal values, as half precision gives) + std_sort_ties


def net(n):
    def f(feats, hashes):
        pi, v = oracles.fakenet(hashes, n * n + 1)
        if QUANT:
            pi = (np.floor(pi * np.float32(QUANT)) / np.float32(QUANT)).astype(np.float32)
        return pi, v

    return f


def fake_actor(search, n):
    f = net(n)

    def actor(batch):
        h, _, _ = search.leaf_info()
        pi, v = f(None, h)
        return {"pi": torch.from_numpy(pi), "V": torch.from_numpy(v)}

    return actor


def one_case(rng, n, case, replay=None, dump=None, verbose=False):
    if replay is not None:
        return run_case(n, case, replay["opts"], replay["G"], replay["opening"], replay["moves"], None, verbose)
    opts = dict(
        num_rollouts=int(rng.integers(4, 200 if n == 9 else 70)), num_rollouts_per_batch=int(rng.integers(1, 17)),
        virtual_loss=int(rng.integers(0, 6)), persistent_tree=int(rng.random() < 0.7),
        c_puct=float(rng.choice([0.3, 0.85, 1.5, 2.5, 5.0])), unexplored_q_zero=int(rng.integers(0, 2)),
        root_unexplored_q_zero=int(rng.integers(0, 2)), ply_pass_enabled=int(rng.choice([0, 20, 60, 300])),
        remove_pass_if_dangerous=int(rng.integers(0, 2)), komi=float(rng.choice([0.5, 5.5, 6.5, 7.5])),
        use_prior=int(rng.random() > 0.15))
    G = int(rng.integers(1, 5))
    open_plies = int(rng.integers(0, 2 * n * n if rng.random() < 0.3 else n * n))
    moves = int(rng.integers(2, 9 if n == 9 else 4))
    # the opening is played on scratch states first so that a failing case can be dumped and replayed
    scratch = [oracles.Ref(n) for _ in range(G)]
    opening = []
    for _ in range(open_plies):
        acts = np.full(G, -2, np.int32)
        for g, o in enumerate(scratch):
            if o.info()[9]:
                continue
            idx = np.flatnonzero(o.legal())
            # mostly stones, sometimes a pass (two in a row end the game: terminal roots are skipped below)
            acts[g] = int(rng.choice(idx)) if len(idx) and rng.random() > 0.03 else n * n
            assert o.forward(int(acts[g]))
        opening.append([int(a) for a in acts])
    return run_case(n, case, opts, G, opening, moves, dump, verbose)


class PortSearch:
    """the C restatement (oracle/mcts_oracle.c) behind the few calls run_case makes: --impl port fuzzes
    the restatement against the reference without the emulator (much faster)"""

    def __init__(self, n, G, opts):
        self.os = [oracles.Oracle(n) for _ in range(G)]
        self.ms = [oracles.OracleMcts(n, callback=net(n) if QUANT else None, std_sort_ties=int(QUANT > 0), **opts) for _ in range(G)]
        self.n, self.G = n, G

    def forward(self, acts):
        return np.array([a == -2 or o.forward(int(a)) for o, a in zip(self.os, acts)])

    def act(self, actor, active):
        P1 = self.n * self.n + 1
        res = {"visits": np.full((self.G, P1), -1, np.int32), "total_visits": np.zeros(self.G, np.int32),
               "best_action": np.zeros(self.G, np.int32), "root_value": np.zeros(self.G, np.float32)}
        self.pri = np.zeros((self.G, P1), np.float32)
        for g in np.flatnonzero(active):
            w = self.ms[g].act(self.os[g])
            res["visits"][g], res["total_visits"][g] = w["visits"], w["total_visits"]
            res["best_action"][g], res["root_value"][g] = w["best_action"], w["root_value"]
            self.pri[g] = w["prior"]
        return res

    def root_priors(self):
        return self.pri

    def advance(self, acts):
        pass

    def errors(self):
        return np.zeros(4, np.int32)


IMPL = "emu"


def run_case(n, case, opts, G, opening, moves, dump, verbose):
    if IMPL == "port":
        mc = PortSearch(n, G, opts)
        gb = mc
    else:
        emu.emu_lib().simt_emu_set_order(case % 3)
        gb = emu.emu_batch(G, n)
        mc = emu.EmuSearch(gb, rotation_flip=0, std_sort_ties=int(QUANT > 0), **opts)
    refs = [oracles.Ref(n) for _ in range(G)]
    rms = [oracles.RefMcts(n, callback=net(n) if QUANT else None, **opts) for _ in range(G)]
    for row in opening:
        acts = np.array(row, np.int32)
        for g, o in enumerate(refs):
            if acts[g] != -2:
                assert o.forward(int(acts[g]))
        gb.forward(acts)
    actor = fake_actor(mc, n) if IMPL == "emu" else None
    tag = f"seed-case {case}: n={n} G={G} open={len(opening)} moves={moves} {opts}"

    def fail():
        if dump:
            import json

            json.dump({"n": n, "case": case, "opts": opts, "G": G, "opening": opening, "moves": moves}, open(dump, "w"))
            print("dumped", dump, flush=True)
        return False
    for mv in range(moves):
        live = np.array([not o.info()[9] for o in refs])
        if not live.any():
            break
        res = mc.act(actor, active=live.astype(np.uint8))
        pri = mc.root_priors()
        if verbose:
            print("move", mv, "errors", mc.errors(), flush=True)
        if mc.errors()[3]:
            # the bounded node pool recycled subtrees of the persistent tree (deviation (ii)):
            # from here on the statistics legitimately differ from an unbounded tree's
            print("ok (pruned at move %d)" % mv, tag, flush=True)
            return True
        acts = np.full(G, -2, np.int32)
        for g in np.flatnonzero(live):
            w = rms[g].act(refs[g])
            if not (np.array_equal(res["visits"][g], w["visits"]) and res["total_visits"][g] == w["total_visits"]
                    and res["best_action"][g] == w["best_action"] and res["root_value"][g] == np.float32(w["root_value"])
                    and np.array_equal(pri[g][w["visits"] >= 0], w["prior"][w["visits"] >= 0])):
                print("MISMATCH", tag, "move", mv, "game", g, flush=True)
                print(" ours ", res["best_action"][g], res["total_visits"][g], res["root_value"][g], flush=True)
                print(" ref  ", w["best_action"], w["total_visits"], w["root_value"], flush=True)
                d = np.flatnonzero(res["visits"][g] != w["visits"])
                print(" visits differ at", d[:10], res["visits"][g][d[:10]], w["visits"][d[:10]], flush=True)
                return fail()
            acts[g] = w["best_action"]
            assert refs[g].forward(int(acts[g]))
        ok = gb.forward(acts)
        assert ok[live].all()
        mc.advance(acts)
    e = mc.errors()
    if verbose:
        print("errors (root mismatch, pool overflow, depth cut, prunes):", e, flush=True)
    if e[0] or e[1] or e[2]:
        print("ERRORS", e, tag, flush=True)
        return fail()
    print("ok", tag, flush=True)
    mc.close() if hasattr(mc, "close") else None
    return True


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cases", type=int, default=20)
    ap.add_argument("--board", type=int, default=9)
    ap.add_argument("--dump-dir", default="/tmp", help="where a failing case is written (JSON)")
    ap.add_argument("--replay", help="run one dumped case again")
    ap.add_argument("--keep-going", action="store_true")
    ap.add_argument("--impl", choices=["emu", "port"], default="emu")
    ap.add_argument("--quant", type=int, default=0, help="quantise the fake net's probabilities to 1/Q and search with std_sort_ties")
    a = ap.parse_args()
    global IMPL, QUANT
    IMPL, QUANT = a.impl, a.quant
    if a.replay:
        import json

        r = json.load(open(a.replay))
        sys.exit(0 if one_case(None, r["n"], r["case"], replay=r, verbose=True) else 1)
    if not oracles.have_ref(a.board):
        sys.exit("oracle/_ref is not built")
    rng = np.random.default_rng(a.seed)
    for c in range(a.cases):
        if not one_case(rng, a.board, c, dump=os.path.join(a.dump_dir, f"emu_fuzz_fail_b{a.board}_s{a.seed}_c{c}.json")):
            if not a.keep_going:
                sys.exit(1)
    print("all", a.cases, "cases equal")


if __name__ == "__main__":
    main()
