"""Memcheck for the kernels without a GPU: run the emulator build under AddressSanitizer.

    python tests/simt_emu/asan_run.py            (re-executes itself with libasan preloaded)

Every "device" buffer of the C ABI is a heap allocation in the emulator build, so an out-of-bounds
or use-after-free access by any kernel (or by the host side of the library) is reported by ASan
with the kernel source line.  Exercised: playouts (both modes), elfb200_replay with random --
including refused -- moves, the export and feature kernels, and the search with a node pool small
enough to overflow (tree drops), random D4, the device-side move choice and tree advance, and a
search without the prior term (every descent step goes through the container-order tie-break) and one
with std_sort_ties on replies full of equal probabilities (the sequential std::sort restatement).
TEST INFRASTRUCTURE ONLY."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def main():
    if os.environ.get("SIMT_EMU_ASAN_CHILD") != "1":
        asan = subprocess.check_output(["gcc", "-print-file-name=libasan.so"], text=True).strip()
        env = dict(os.environ, LD_PRELOAD=os.path.realpath(asan), SIMT_EMU_ASAN_CHILD="1",
                   ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0")
        sys.exit(subprocess.call([sys.executable, os.path.abspath(__file__)], env=env))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import numpy as np
    import torch

    import build
    from elf_b200 import lib as _l
    from tests import emu as E, oracles

    E._lib = _l.load_library(build.build(asan=True))
    for n, G in ((9, 7), (19, 3)):
        gb = E.emu_batch(G, n)
        r = gb.playout(1234, first_game_id=50)
        want = oracles.oracle_playout_many(n, 1234, 50, G)
        assert (r["chk"] == want["chk"]).all()
        gb.playout_stream(99, first_game_id=7, plies_per_slot=200)
        rng = np.random.default_rng(n)
        gb.replay([[int(a) for a in rng.integers(0, n * n + 1, int(rng.integers(0, 2 * n * n)))] for _ in range(G)])
        gb.features(np.arange(G, dtype=np.int32) % 8)
        gb.legal_mask(), gb.info(), gb.tt_score(), gb.true_eyes(0), gb.stones()
        gb2 = E.emu_batch(G, n)
        mc = E.EmuSearch(gb2, rotation_flip=1, num_rollouts=40, num_rollouts_per_batch=8, persistent_tree=1,
                         nodes_per_game=48, root_epsilon=0.25, root_alpha=0.3)

        def actor(batch):
            h, _, _ = mc.leaf_info()
            pi, v = oracles.fakenet(h, n * n + 1)
            return {"pi": torch.from_numpy(pi), "V": torch.from_numpy(v)}

        for mv in range(5):
            mc.act(actor)
            a, _ = mc.choose(3, 0.05, None, mv)
            gb2.forward(a)
            mc.advance(a)
        mc.root_priors()
        # exact PUCT ties at every step (no prior term): uct_tie_break's full rescans, tie bit words and
        # hash-table scratch, full-scan nodes
        gb3 = E.emu_batch(G, n)
        mt = E.EmuSearch(gb3, rotation_flip=0, num_rollouts=48, num_rollouts_per_batch=6, persistent_tree=1, use_prior=0)

        def actor_t(batch):
            h, _, _ = mt.leaf_info()
            pi, v = oracles.fakenet(h, n * n + 1)
            return {"pi": torch.from_numpy(pi), "V": torch.from_numpy(v)}

        for mv in range(3):
            mt.act(actor_t)
            a, _ = mt.choose(-1, 0.05, None, mv)
            gb3.forward(a)
            mt.advance(a)
        mt.close(), gb3.close()
        # replies full of bit-equal probabilities with std_sort_ties: k_expand<N, true>'s sequential
        # std::sort restatement (shared-memory key / order / move arrays, explicit partition stack)
        gb4 = E.emu_batch(G, n)
        ms = E.EmuSearch(gb4, rotation_flip=1, num_rollouts=32, num_rollouts_per_batch=8, persistent_tree=1, std_sort_ties=1)

        def actor_q(batch):
            h, _, _ = ms.leaf_info()
            pi, v = oracles.fakenet(h, n * n + 1)
            return {"pi": torch.from_numpy((np.floor(pi * 16) / 16).astype(np.float32)), "V": torch.from_numpy(v)}

        for mv in range(3):
            ms.act(actor_q)
            a, _ = ms.choose(-1, 0.05, None, mv)
            gb4.forward(a)
            ms.advance(a)
        ms.close(), gb4.close()
        print(f"{n}x{n}: no ASan report; tree prunes {int(mc.errors()[3])}, pool overflows {int(mc.errors()[1])}", flush=True)
        mc.close(), gb2.close(), gb.close()
    print("ASAN RUN CLEAN")


if __name__ == "__main__":
    main()
