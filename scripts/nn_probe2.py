#!/usr/bin/env python
"""Probe round 2: how fast can the PyTorch/cuDNN plumbing evaluate the 20x256 policy/value net at
NN batch 256?  Variants: eager module (round 1), cudnn.benchmark, BN folded into the convolutions
with cuDNN's fused conv+bias(+add)+ReLU calls, and the same under a CUDA graph.  Not our kernels:
this picks the settings of elf_b200.model.FusedActor."""
import sys
import time

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from elf_b200.model import Actor, PolicyValueNet  # noqa: E402


def timeit(fn, iters=10):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def report(name, batch, ms, blocks=20, dim=256):
    flops = 2 * batch * 361 * (18 * dim * 9 + blocks * 2 * dim * dim * 9)
    print(f"{name:44s} batch {batch:5d} {ms:8.3f} ms {batch / ms * 1e3:9.0f} pos/s {flops / ms / 1e9:7.1f} TFLOP/s", flush=True)


def main():
    torch.manual_seed(0)
    dev = torch.device("cuda")
    model = PolicyValueNet(19, num_block=20, dim=256).to(dev).eval()
    for bench in (False, True):
        torch.backends.cudnn.benchmark = bench
        act = Actor(PolicyValueNet(19, num_block=20, dim=256).to(dev), batchsize=256)
        for b in (256, 1024):
            x = torch.rand(b, 18, 19, 19, device=dev)
            report(f"eager fp16 CL (round 1) benchmark={bench}", b, timeit(lambda: act({"s": x})))
    from elf_b200.model import FusedActor

    for dtype in (torch.float16, torch.bfloat16):
        for graph in (False, True):
            try:
                fa = FusedActor(model, batchsize=256, dtype=dtype, cuda_graph=graph)
            except Exception as e:
                print("FusedActor failed:", dtype, graph, repr(e), flush=True)
                continue
            for b in (256, 2048):
                x = torch.rand(b, 18, 19, 19, device=dev)
                report(f"fused {dtype} graph={graph}", b, timeit(lambda: fa({"s": x})))
            # accuracy against the fp32 module
            x = (torch.rand(256, 18, 19, 19, device=dev) > 0.7).float()
            with torch.no_grad():
                ref = model(x)
            out = fa({"s": x})
            print(f"   max |pi - pi_fp32| = {(out['pi'] - ref['pi']).abs().max().item():.3e}  "
                  f"max |V - V_fp32| = {(out['V'].reshape(-1) - ref['V'].reshape(-1)).abs().max().item():.3e}", flush=True)
    # bigger NN batch for reference (not BASELINE's config): how far is batch 256 from the asymptote
    for nb in (512, 1024):
        fa = FusedActor(model, batchsize=nb, dtype=torch.float16, cuda_graph=True)
        x = torch.rand(4096, 18, 19, 19, device=dev)
        report(f"fused fp16 graph NN batch {nb}", 4096, timeit(lambda: fa({"s": x}), 3))
    # fp16 NHWC input handed in directly (what k_leaf_features' fast mode will write)
    fa = FusedActor(model, batchsize=256, dtype=torch.float16, cuda_graph=True)
    xh = torch.rand(4096, 19, 19, fa.cpad, device=dev).half()
    report("fused fp16 graph, NHWC fp16 input", 4096, timeit(lambda: fa({"s_nhwc": xh}), 3))
    # two streams, two graphs: do concurrent batch-256 chunks overlap better?
    fa2 = FusedActor(model, batchsize=256, dtype=torch.float16, cuda_graph=True)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def two():
        s1.wait_stream(torch.cuda.current_stream())
        s2.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s1):
            fa({"s_nhwc": xh[:2048]})
        with torch.cuda.stream(s2):
            fa2({"s_nhwc": xh[2048:]})
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)

    report("fused fp16 graph, 2 streams x 2048", 4096, timeit(two, 3))


if __name__ == "__main__":
    main()
