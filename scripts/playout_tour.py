#!/usr/bin/env python
"""k_playout in both lane layouts at 4096 and 16384 games (one bench step each) for ncu."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import elf_b200  # noqa: E402

for G in (4096, 16384):
    gb = elf_b200.GoBatch(G, board_size=19)
    for layout in (0, 1):
        gb.set_playout_layout(layout)
        gb.playout_stream_launch(20260922, 0, 512)  # warm
        gb.synchronize()
        torch.cuda.profiler.start()
        gb.playout_stream_launch(20260922, 0, 512)
        gb.synchronize()
        torch.cuda.profiler.stop()
    gb.close()
print("done")
