#!/usr/bin/env python
"""One visit to every kernel of libelfb200.so at BASELINE's batch (4096 games, 19x19, 800 rollouts per
move in waves of 8) in a realistic state, for `ncu --set full` (north_star: each kernel ships with an
ncu capture).  The search state is built unprofiled (one full 800-rollout move with a table-lookup
net, tree advance, 60 waves into the second move); then cudaProfilerStart() brackets two waves and one
launch of everything else.  Run as

  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_ \\
      -o gpurun_out/r2_tour python scripts/kernel_tour.py

scripts/ncu_summary.py turns the report into profiles/r2_kernels.md + profiles/traffic.json."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import elf_b200  # noqa: E402
from elf_b200 import lib as L  # noqa: E402

G, N, R, B = int(os.environ.get("TOUR_GAMES", 4096)), 19, 800, 8
dev = torch.device("cuda", 0)
torch.manual_seed(0)
gb = elf_b200.GoBatch(G, board_size=N)
rng = np.random.default_rng(0)
for _ in range(30):  # opening: random legal moves (k_step / k_export, unprofiled)
    lg = gb.legal_mask()[:, :-1].astype(np.float64) + 1e-9
    c = (lg / lg.sum(1, keepdims=True)).cumsum(1)
    a = np.minimum((c < rng.random((G, 1))).sum(1), N * N - 1).astype(np.int32)
    gb.forward(a)
mc = elf_b200.MctsBatch(gb, num_rollouts=R, num_rollouts_per_batch=B, rotation_flip=1, seed=1)
P1 = N * N + 1
table = torch.rand(4096, P1, device=dev).softmax(1)
vals = torch.rand(4096, device=dev) * 2 - 1


def actor(batch):
    x = batch["s"] if "s" in batch else batch["s_nhwc"]
    idx = (torch.arange(x.shape[0], device=dev) * 7919) % 4096
    return {"pi": table[idx].contiguous(), "V": vals[idx].contiguous()}


mc.search(actor)                      # move 1: 100 waves
a, _ = mc.choose(0, 0.0)
gb.forward(a)
mc.advance(a)
mc.begin_move()
for _ in range(60):                   # 60 waves into move 2: persistent tree + 480 new rollouts per game
    mc.wave(actor)
gb.synchronize()
torch.cuda.synchronize()

torch.cuda.profiler.start()
mc.wave(actor)                        # k_select, k_leaf_features (float32 NCHW, bulk store), k_expand, k_backup
mc.set_feature_format("f16")
mc.wave(actor)                        # k_leaf_features fp16 NHWC
gb.set_feature_store(1)
mc.wave(actor)                        # k_leaf_features fp16 NHWC, staged tile + bulk (TMA) store (for comparison)
gb.set_feature_store(0)
mc.set_feature_format("f32")
mc.results()                          # k_results
a, _ = mc.choose(0, 0.0)              # k_choose
mc.root_priors()                      # k_root_priors
gb.forward(a)                         # k_step
mc.advance(a)                         # k_advance
mc.begin_move()                       # k_begin (persistent roots)
out32 = torch.empty((G, 18, N, N), dtype=torch.float32, device=dev)
out16 = torch.empty((G, N, N, 24), dtype=torch.float16, device=dev)
gb.features_dev(out32.data_ptr())                         # k_features float32
gb.features_dev(out16.data_ptr(), None, L.FEAT_F16_NHWC, 24)  # k_features fp16 NHWC
df = gb.features_df()                  # k_features_df (DarkForest 25 planes)
gb.legal_mask()                       # k_export
gb.tt_score()
moves = [list(rng.integers(0, N * N, 120)) for _ in range(256)]
gb2 = elf_b200.GoBatch(256, board_size=N)
gb2.replay(moves)                     # k_replay
m = np.zeros(G, np.uint8)
m[::2] = 1
mc.reset(m)                           # k_tree_reset
gb.reset(m)                           # k_reset
gb.playout_stream_launch(20260922, 0, 512)  # k_playout, the bench's configs[1] step
gb.set_playout_layout(1)
gb.playout_stream_launch(20260922, 0, 512)  # k_playout2 (two rows per lane)
gb.set_playout_layout(0)
gb.synchronize()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("tour done", mc.errors())
