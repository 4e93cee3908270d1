"""SelfPlay(rng="reference") against tests/golden/refstream_games.json -- games of the compiled
reference's game threads (scripts/gen_golden_refstream.py) -- so that the stream parity stays checked
where oracle/_ref is not available.  Kernels on the SIMT emulator; the device twin is in
tests/test_gpu_mcts.py."""
import json
import os

import numpy as np
import pytest
import torch

from tests import emu, oracles

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refstream_games.json")


def run_case(c, make_board, make_search):
    from elf_b200.selfplay import SelfPlay

    n, G = c["n"], len(c["games"])

    def actor(batch):
        pi, v = oracles.feature_net(batch["s"].float().cpu().numpy(), n * n + 1)
        dev = batch["s"].device
        return {"pi": torch.from_numpy(pi).to(dev), "V": torch.from_numpy(v).to(dev)}

    gb = make_board(G, n)
    mc = make_search(gb, rotation_flip=c["flip"], **c["opts"])
    sp = SelfPlay(actor, num_games=G, board_size=n, board=gb, search=mc, rng="reference",
                  seed=np.array(c["seeds"], np.uint64), policy_distri_cutoff=c["cutoff"], resign_thres=c["thres"],
                  never_resign_ratio=c["ratio"], move_cutoff=c["move_cutoff"], root_epsilon=c["eps"],
                  root_alpha=c["alpha"], **c["opts"])
    got = [[] for _ in range(G)]
    fwd = gb.forward

    def logged(acts):
        for g in range(G):
            got[g].append(int(acts[g]))
        return fwd(acts)

    gb.forward = logged
    for t in range(c["moves"]):
        sp.step()
    return got


@pytest.mark.parametrize("case", range(2))
def test_reference_stream_games_equal_the_golden_reference_games(case):
    c = json.load(open(GOLD))[case]
    got = run_case(c, emu.emu_batch, emu.EmuSearch)
    for g, game in enumerate(c["games"]):
        assert got[g] == game["actions"], f"game {g} (seed {game['seed']})"
