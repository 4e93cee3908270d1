#!/usr/bin/env python
"""Probe first-call latency and steady-state throughput of the policy/value net variants
(PyTorch/cuDNN plumbing, not our kernels): picks the NN settings for the config-3 bench."""
import sys
import time

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from elf_b200.model import PolicyValueNet  # noqa: E402


def probe(name, dtype, channels_last, batch, blocks=20, dim=256, iters=5):
    torch.manual_seed(0)
    m = PolicyValueNet(19, num_block=blocks, dim=dim).cuda().eval()
    if channels_last:
        m = m.to(memory_format=torch.channels_last)
    if dtype == "half_weights":
        m = m.half()
    x = torch.rand(batch, 18, 19, 19, device="cuda")
    if channels_last:
        x = x.contiguous(memory_format=torch.channels_last)

    def fwd():
        with torch.no_grad():
            if dtype == "half_weights":
                return m(x.half())
            with torch.autocast("cuda", dtype=dtype, enabled=dtype is not None):
                return m(x)

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fwd()
    torch.cuda.synchronize()
    first = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(iters):
        fwd()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / iters
    flops = 2 * batch * 361 * (18 * dim * 9 + blocks * 2 * dim * dim * 9)
    print(f"{name:34s} batch {batch:5d} first {first:7.2f}s steady {dt*1e3:8.2f} ms  {batch/dt:9.0f} pos/s  {flops/dt/1e12:6.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    cfgs = {
        "bf16_cl": (torch.bfloat16, True), "bf16_nchw": (torch.bfloat16, False),
        "fp16_cl": (torch.float16, True), "fp16w_cl": ("half_weights", True), "tf32_nchw": (None, False),
    }
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    for k, (dt, cl) in cfgs.items():
        if which in ("all", k):
            for b in (256, 2048):
                probe(k, dt, cl, b)
