#!/usr/bin/env python
"""CUDA-event micro-benchmarks of single kernels (outside ncu): the feature writer in every format /
store mode at 32768 positions, and the playout kernel in both lane layouts at 4096 and 16384 games.
L2 is flushed between repetitions (256 MiB write).  Prints one JSON object."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import elf_b200  # noqa: E402
from elf_b200 import lib as L  # noqa: E402

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0


def timed(gb, fn, reps=20):
    st = torch.cuda.ExternalStream(gb.stream, device=dev)
    fn()
    gb.synchronize()
    ms = []
    for r in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            flush.fill_(r & 0xFF)
            a.record(st)
        fn()
        with torch.cuda.stream(st):
            b.record(st)
        gb.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms)), float(np.min(ms))


out = {"peak_hbm_GBps": PEAK}
# ---- feature writer --------------------------------------------------------------------------------
G, N = 32768, 19
gb = elf_b200.GoBatch(G, board_size=N)
rng = np.random.default_rng(0)
for _ in range(24):
    lg = gb.legal_mask()[:, :-1].astype(np.float32) + 1e-6
    c = (lg / lg.sum(1, keepdims=True)).cumsum(1)
    a = np.minimum((c < rng.random((G, 1), dtype=np.float32)).sum(1), N * N - 1).astype(np.int32)
    gb.forward(a)
d4 = torch.from_numpy(rng.integers(0, 8, G).astype(np.int32)).to(dev)
o32 = torch.empty((G, 18, N, N), dtype=torch.float32, device=dev)
o16 = torch.empty((G, N, N, 24), dtype=torch.float16, device=dev)
feat = {}
for name, fn, byts in (
        ("f32_nchw", lambda: gb.features_dev(o32.data_ptr(), d4.data_ptr()), 26792),
        ("f16_nhwc_direct_stores", lambda: gb.features_dev(o16.data_ptr(), d4.data_ptr(), L.FEAT_F16_NHWC, 24), 800 + 17328)):
    med, best = timed(gb, fn)
    feat[name] = {"ms_median": med, "ms_min": best, "GBps": G * byts / med / 1e6, "frac_of_peak": G * byts / med / 1e6 / PEAK,
                  "bytes_per_position": byts}
gb.set_feature_store(1)
med, best = timed(gb, lambda: gb.features_dev(o16.data_ptr(), d4.data_ptr(), L.FEAT_F16_NHWC, 24))
feat["f16_nhwc_staged_bulk_store"] = {"ms_median": med, "ms_min": best, "GBps": G * 18128 / med / 1e6, "frac_of_peak": G * 18128 / med / 1e6 / PEAK}
gb.set_feature_store(0)
out["k_features_32768_positions"] = feat
gb.close()
del o32, o16
# ---- playout kernel, both lane layouts ---------------------------------------------------------------
po = {}
for G in (4096, 16384):
    gb = elf_b200.GoBatch(G, board_size=N)
    for layout in (0, 1):
        gb.set_playout_layout(layout)
        med, best = timed(gb, lambda: gb.playout_stream_launch(20260922, 0, 512), reps=8)
        plies = gb.playout_results()["total_plies"]
        po[f"G{G}_layout{layout}"] = {"ms_median": med, "moves_per_s": plies / med * 1e3, "plies": plies}
    gb.close()
out["k_playout_stream_512"] = po
print(json.dumps(out, indent=1))
