"""The policy/value net twin: same parameter count and state_dict layout as the reference's
Model_PolicyValue (df_model3.py:113-306) up to the documented key remapping."""
import torch

from elf_b200.model import PolicyValueNet, load_reference_state_dict


def test_shapes_and_parameter_count():
    net = PolicyValueNet(19, num_block=20, dim=256)
    n = sum(p.numel() for p in net.parameters())
    assert 23_900_000 < n < 24_100_000  # "~24 M params" 
    out = net.eval()(torch.zeros(2, 18, 19, 19))
    assert out["pi"].shape == (2, 362) and out["V"].shape == (2, 1) and out["logpi"].shape == (2, 362)
    assert torch.allclose(out["pi"].sum(1), torch.ones(2), atol=1e-5)
    assert (out["V"].abs() <= 1).all()


def test_reference_key_layout_loads():
    torch.manual_seed(0)
    a = PolicyValueNet(9, num_block=2, dim=8)
    # what the reference would have saved: tower one level deeper, DataParallel prefix, wrapped dict
    ref_sd = {}
    for k, v in a.state_dict().items():
        if k.startswith("resnet."):
            k = "resnet.resnet." + k[len("resnet."):]
        ref_sd["module." + k if k.startswith("init_conv") else k] = v.clone()
    b = PolicyValueNet(9, num_block=2, dim=8)
    missing, unexpected = load_reference_state_dict(b, {"state_dict": ref_sd, "stats": {}})
    assert not missing and not unexpected
    x = torch.rand(3, 18, 9, 9)
    ya, yb = a.eval()(x), b.eval()(x)
    assert torch.equal(ya["pi"], yb["pi"]) and torch.equal(ya["V"], yb["V"])
