#!/usr/bin/env python
"""Where do the microseconds of a host-driven 4096-game step go?  (a) k_step alone, device-timed over
back-to-back launches; (b) launch + kernel + stream synchronise with device-resident actions;
(c) elfb200_step with host buffers (mapped window + completion flag)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import elf_b200  # noqa: E402

G, N = 4096, 19
gb = elf_b200.GoBatch(G, board_size=N)
gold = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "playouts_19.json")))
mv = next(e["moves"] for e in gold["games"] if "moves" in e)
dev = torch.device("cuda", 0)
st = torch.cuda.ExternalStream(gb.stream, device=dev)
acts_dev = [torch.full((G,), a, dtype=torch.int32, device=dev) for a in mv]
ok_dev = torch.empty(G, dtype=torch.uint8, device=dev)
out = {}
# (a) device time of k_step, back to back
gb.reset()
gb.synchronize()
a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a_.record(st)
for t in acts_dev:
    gb.forward_dev(t.data_ptr(), ok_dev.data_ptr())
b_.record(st)
gb.synchronize()
out["k_step_us_device_back_to_back"] = 1e3 * a_.elapsed_time(b_) / len(mv)
# (b) launch + kernel + synchronise, device-resident actions
gb.reset()
t0 = time.perf_counter()
for t in acts_dev:
    gb.forward_dev(t.data_ptr(), ok_dev.data_ptr())
    gb.synchronize()
out["launch_kernel_sync_us"] = 1e6 * (time.perf_counter() - t0) / len(mv)
# (c) host buffers through elfb200_step
gb.reset()
acts = np.empty(G, np.int32)
t0 = time.perf_counter()
for a in mv:
    acts.fill(a)
    ok = gb.forward(acts)
out["elfb200_step_host_us"] = 1e6 * (time.perf_counter() - t0) / len(mv)
out["all_accepted"] = bool(ok.all())
# python-only overhead of the same loop with the library call stubbed out is not measurable here; numpy fill alone:
t0 = time.perf_counter()
for a in mv:
    acts.fill(a)
out["numpy_fill_us"] = 1e6 * (time.perf_counter() - t0) / len(mv)
print(json.dumps(out))
