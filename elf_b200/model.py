"""Policy/value ResNet with the architecture of the reference's ``Model_PolicyValue``
(``src_py/elfgames/go/df_model3.py:113-306``): 3x3 conv + BN + ReLU stem, ``num_block`` residual
blocks (conv-BN-ReLU, conv-BN, add, ReLU; ``df_model3.py:20-84``), policy head 1x1 conv (2 ch) + BN
+ ReLU -> Linear(2d, d+1) -> log-softmax, value head 1x1 conv (1 ch) + BN + ReLU -> Linear(d, 256)
-> ReLU -> Linear(256, 1) -> tanh.  In the drop-in deployment the reference's own module is used
unchanged; this twin exists because the GPU box has no copy of the reference tree, and for
random-init benchmarking (BASELINE config 3: 20 blocks x 256 channels).

This is PyTorch plumbing (cuDNN convolutions), not part of the hand-written hot path.
"""
import torch
import torch.nn as nn


def _conv(cin, cout, k, relu=True):
    layers = [nn.Conv2d(cin, cout, k, padding=k // 2), nn.BatchNorm2d(cout, momentum=0.1, eps=1e-5)]
    if relu:
        layers.append(nn.ReLU())
    return nn.Sequential(*layers)


class Block(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.conv_lower = _conv(dim, dim, 3)
        self.conv_upper = _conv(dim, dim, 3, relu=False)
        self.relu = nn.ReLU()

    def forward(self, s):
        return self.relu(self.conv_upper(self.conv_lower(s)) + s)


class PolicyValueNet(nn.Module):
    def __init__(self, board_size=19, num_planes=18, num_block=20, dim=256):
        super().__init__()
        self.board_size = board_size
        d = board_size * board_size
        self.init_conv = _conv(num_planes, dim, 3)
        self.resnet = nn.Sequential(*[Block(dim) for _ in range(num_block)])
        self.pi_final_conv = _conv(dim, 2, 1)
        self.value_final_conv = _conv(dim, 1, 1)
        self.pi_linear = nn.Linear(2 * d, d + 1)
        self.value_linear1 = nn.Linear(d, 256)
        self.value_linear2 = nn.Linear(256, 1)

    def forward(self, x):
        s = x["s"] if isinstance(x, dict) else x
        d = self.board_size * self.board_size
        s = self.resnet(self.init_conv(s))
        pi = self.pi_linear(self.pi_final_conv(s).reshape(-1, 2 * d))
        logpi = torch.log_softmax(pi.float(), dim=1)
        v = torch.relu(self.value_linear1(self.value_final_conv(s).reshape(-1, d)))
        v = torch.tanh(self.value_linear2(v).float())
        return dict(logpi=logpi, pi=logpi.exp(), V=v)


class Actor:
    """The model-interface callback (``Evaluator.actor``, rlpytorch/trainer/trainer.py:73-115) for
    inference: ``actor(batch) -> {"pi", "V"}`` on the GPU.  Evaluates in chunks of ``batchsize``
    positions (the reference's NN batch; BASELINE config 3 uses 256).  ``dtype`` float16/bfloat16
    converts the weights once (channels_last, tensor-core convolutions through cuDNN); float32
    keeps them as they are.  (Measured on B200, 20x256 net: fp16 weights + channels_last reach
    ~32 k positions/s already at batch 256; bf16 autocast on the 18-channel stem was pathologically
    slow to start, see scripts/nn_probe.py.)"""

    def __init__(self, model, batchsize=256, dtype=torch.float16, channels_last=True):
        self.model = model.eval()
        self.batchsize = batchsize
        self.dtype = dtype
        self.channels_last = channels_last
        if dtype != torch.float32:
            self.model = self.model.to(dtype)
        if channels_last:
            self.model = self.model.to(memory_format=torch.channels_last)
        self.num_batches = 0
        self.num_positions = 0

    @torch.no_grad()
    def __call__(self, batch):
        s = batch["s"]
        n = s.shape[0]
        pis, vs = [], []
        for i in range(0, n, self.batchsize):
            x = s[i:i + self.batchsize].to(self.dtype)
            if self.channels_last:
                x = x.contiguous(memory_format=torch.channels_last)
            out = self.model(x)
            pis.append(out["pi"].float())
            vs.append(out["V"].float().reshape(-1))
            self.num_batches += 1
        self.num_positions += n
        return {"pi": torch.cat(pis) if len(pis) > 1 else pis[0], "V": torch.cat(vs) if len(vs) > 1 else vs[0]}


def load_reference_state_dict(model, state_dict):
    """Load weights saved from the reference's ``Model_PolicyValue`` (``df_model3.py``): its tower lives
    one level deeper (``resnet.resnet.<i>...`` because ``GoResNet`` wraps the ``nn.Sequential``) and
    DataParallel / DDP wrappers prefix ``module.``; everything else has the same names here.
    Returns (missing, unexpected) like ``load_state_dict(strict=False)``."""
    sd = state_dict.get("state_dict", state_dict)
    out = {}
    for k, v in sd.items():
        k = k.replace("module.", "")
        if k.startswith("resnet.resnet."):
            k = "resnet." + k[len("resnet.resnet."):]
        out[k] = v
    return model.load_state_dict(out, strict=False)


def broadcast_weights(model, src=0):
    """NCCL broadcast of the frozen weights from rank ``src`` (the only collective of the path)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src=src)
