"""Warp collectives per game-ply / per rollout of every kernel, counted by the SIMT emulator.

    python tests/simt_emu/collective_profile.py > profiles/r1_collectives.md

A static property of the kernel sources on a given input (no GPU, no timing): how many SHFL / REDUX /
VOTE / MATCH / __syncwarp operations a warp executes.  These are the instructions whose latency
chains the board kernels are made of, so the table says where a rewrite can save them.
TEST INFRASTRUCTURE ONLY."""
import collections
import ctypes
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from tests import emu as E, oracles  # noqa: E402


def counters(reset=True):
    L = E.emu_lib()
    buf = ctypes.create_string_buffer(1 << 16)
    L.simt_emu_counters.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
    k = L.simt_emu_counters(buf, 1 << 16, int(reset))
    out = collections.defaultdict(dict)
    for line in buf.value.decode().splitlines():
        kern, op, n = line.split()
        out[kern][op] = int(n)
    return out


def table(title, c, unit, units):
    ops = sorted({o for d in c.values() for o in d})
    if not ops:
        print(f"\n### {title}\n\n(no warp collectives: one thread per output cell)")
        return
    print(f"\n### {title}\n")
    print("| kernel | " + " | ".join(ops) + f" | total | per {unit} |")
    print("|---|" + "---|" * (len(ops) + 2))
    for kern in sorted(c):
        tot = sum(c[kern].values())
        print(f"| `{kern}` | " + " | ".join(str(c[kern].get(o, 0)) for o in ops) + f" | {tot} | {tot / units:.1f} |")


def main():
    print("# Warp collectives executed by the kernels (SIMT emulator count, `tests/simt_emu/collective_profile.py`)")
    print("\nCounts are per WARP-wide operation.  19x19 runs one game per warp, 9x9 three.")
    for n, G in ((19, 4), (9, 6)):
        counters()
        gb = E.emu_batch(G, n)
        counters()
        r = gb.playout(77, first_game_id=0)
        plies = int(r["total_plies"])
        warps = G if n == 19 else (G + 2) // 3
        table(f"k_playout, {n}x{n}, {G} games to the end ({plies} plies; per game-ply of ONE warp = total / ({plies} / {G // warps} games per warp))",
              counters(), "game-ply (all games)", plies)
        o = [oracles.Oracle(n) for _ in range(G)]
        rng = np.random.default_rng(n)
        steps = 60
        for _ in range(steps):
            acts = np.array([int(rng.choice(np.flatnonzero(x.legal()))) for x in o], np.int32)
            gb.forward(acts)
            for x, a in zip(o, acts):
                x.forward(int(a))
        table(f"k_step, {n}x{n}, {G} games x {steps} plies", counters(), "game-ply", G * steps)
        gb.features(None)
        gb.legal_mask()
        table(f"k_features / k_export, {n}x{n}, {G} positions", counters(), "position", G)
        R, B = 64, 8
        mc = E.EmuSearch(gb, rotation_flip=0, num_rollouts=R, num_rollouts_per_batch=B)

        def actor(batch):
            h, _, _ = mc.leaf_info()
            pi, v = oracles.fakenet(h, n * n + 1)
            return {"pi": torch.from_numpy(pi), "V": torch.from_numpy(v)}

        counters()
        mc.act(actor)
        table(f"search, {n}x{n}, {G} games x {R} rollouts (batch {B}) from a {steps}-ply position", counters(), "rollout",
              G * R)


if __name__ == "__main__":
    main()
