"""FusedActor (BatchNorm folded into the convolutions, fused heads) computes what the module computes.
CPU: float32 through the plain-op branch; the cuDNN fused calls and the CUDA graph are covered by
tests/test_gpu_mcts.py::test_fused_actor_matches_module on the GPU."""
import numpy as np
import torch

from elf_b200.model import FusedActor, PolicyValueNet


def randomise_bn(model):
    for mod in model.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.3)
            mod.running_var.uniform_(0.5, 2.0)
            mod.weight.data.uniform_(0.5, 1.5)
            mod.bias.data.normal_(0, 0.3)


def test_fused_actor_equals_module_fp32():
    torch.manual_seed(3)
    m = PolicyValueNet(9, num_block=3, dim=16).eval()
    randomise_bn(m)
    fa = FusedActor(m, batchsize=4, dtype=torch.float32, cuda_graph=False)
    x = (torch.rand(10, 18, 9, 9) > 0.6).float()
    with torch.no_grad():
        ref = m(x)
    out = fa({"s": x})
    assert out["pi"].shape == (10, 82) and out["V"].shape == (10,)
    np.testing.assert_allclose(out["pi"].numpy(), ref["pi"].numpy(), atol=2e-6)
    np.testing.assert_allclose(out["V"].numpy(), ref["V"].reshape(-1).numpy(), atol=2e-6)
    # the NHWC, channel-padded input of the search's fast feature mode gives the same answer
    xn = torch.zeros(10, 9, 9, fa.cpad)
    xn[..., :18] = x.permute(0, 2, 3, 1)
    out2 = fa({"s_nhwc": xn})
    np.testing.assert_allclose(out2["pi"].numpy(), out["pi"].numpy(), atol=1e-6)
    np.testing.assert_allclose(out2["V"].numpy(), out["V"].numpy(), atol=1e-6)
