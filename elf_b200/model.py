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


def fold_batchnorm(conv, bn):
    """(weight, bias) of the convolution that equals ``bn(conv(x))`` in eval mode (float32)."""
    w = conv.weight.detach().float()
    b = conv.bias.detach().float() if conv.bias is not None else torch.zeros(w.shape[0], device=w.device)
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    return w * scale.reshape(-1, 1, 1, 1), (b - bn.running_mean.detach().float()) * scale + bn.bias.detach().float()


class FusedActor:
    """Inference engine for the policy/value net behind the same callback contract as ``Actor``
    (``Evaluator.actor``, rlpytorch/trainer/trainer.py:73-115): ``actor(batch) -> {"pi", "V"}``.

    Still PyTorch/cuDNN plumbing, but arranged for a frozen net: BatchNorm (eval) is folded into the
    convolution weights once, every tower layer is ONE cuDNN call with the bias, residual add and
    ReLU in its epilogue (``cudnn_convolution_relu`` / ``cudnn_convolution_add_relu``), the two 1x1
    head convolutions read the tower output once, and a full NN batch (``batchsize`` positions,
    BASELINE config 3: 256) is replayed as a CUDA graph.  Takes a module with the attribute names of
    ``PolicyValueNet`` / the reference's ``Model_PolicyValue`` (after ``load_reference_state_dict``).

    Input: ``batch["s"]`` float32 ``[n,18,N,N]`` (the GoFeature contract) or ``batch["s_nhwc"]``
    ``dtype`` ``[n,N,N,cpad]`` as written by the search's fast feature mode
    (``elfb200_mcts_select_nhwc16``); channels 18..cpad-1 are zero padding so that the stem is a
    tensor-core convolution (cuDNN wants channel counts in multiples of 8 for 16-bit types)."""

    def __init__(self, model, batchsize=256, dtype=torch.float16, cuda_graph=True, cpad=24):
        self.batchsize, self.dtype, self.cpad = int(batchsize), dtype, int(cpad)
        self.N = model.board_size
        self.device = next(model.parameters()).device
        self.num_batches = self.num_positions = 0
        cl = torch.channels_last

        def prep(w, b):
            return w.to(dtype).contiguous(memory_format=cl), b.to(dtype).contiguous()

        w, b = fold_batchnorm(model.init_conv[0], model.init_conv[1])
        wp = torch.zeros(w.shape[0], self.cpad, 3, 3, device=w.device)
        wp[:, : w.shape[1]] = w
        self.stem = prep(wp, b)
        self.blocks = []
        for blk in model.resnet:
            self.blocks.append((prep(*fold_batchnorm(blk.conv_lower[0], blk.conv_lower[1])),
                                prep(*fold_batchnorm(blk.conv_upper[0], blk.conv_upper[1]))))
        wpi, bpi = fold_batchnorm(model.pi_final_conv[0], model.pi_final_conv[1])
        wv, bv = fold_batchnorm(model.value_final_conv[0], model.value_final_conv[1])
        wh = torch.zeros(8, wpi.shape[1], 1, 1, device=w.device)  # 2 policy + 1 value channels, padded to 8
        bh = torch.zeros(8, device=w.device)
        wh[:2], wh[2:3], bh[:2], bh[2:3] = wpi, wv, bpi, bv
        self.head = prep(wh, bh)
        self.pi_linear = (model.pi_linear.weight.detach().to(dtype), model.pi_linear.bias.detach().to(dtype))
        self.v1 = (model.value_linear1.weight.detach().to(dtype), model.value_linear1.bias.detach().to(dtype))
        self.v2 = (model.value_linear2.weight.detach().to(dtype), model.value_linear2.bias.detach().to(dtype))
        self.graph = None
        if cuda_graph and self.device.type == "cuda":
            self._capture()

    @staticmethod
    def _conv_relu(x, wb, pad, z=None):
        """relu(conv(x) + bias [+ z]) as one cuDNN call; plain ops for CPU tensors (tests of the folding)"""
        w, b = wb
        if x.is_cuda:
            if z is None:
                return torch.cudnn_convolution_relu(x, w, b, (1, 1), (pad, pad), (1, 1), 1)
            return torch.cudnn_convolution_add_relu(x, w, z, 1.0, b, (1, 1), (pad, pad), (1, 1), 1)
        y = torch.nn.functional.conv2d(x, w, b, padding=pad)
        return torch.relu(y if z is None else y + z)

    def _tower(self, x):
        x = self._conv_relu(x, self.stem, 1)
        for lower, upper in self.blocks:
            x = self._conv_relu(self._conv_relu(x, lower, 1), upper, 1, z=x)
        return x

    def _forward(self, x):
        """x: dtype [n,cpad,N,N] channels_last -> (pi float32 [n,N*N+1], V float32 [n])"""
        d = self.N * self.N
        h = self._conv_relu(self._tower(x), self.head, 0)
        n = h.shape[0]
        pi = torch.nn.functional.linear(h[:, :2].reshape(n, 2 * d), *self.pi_linear)
        v = torch.relu(torch.nn.functional.linear(h[:, 2].reshape(n, d), *self.v1))
        v = torch.tanh(torch.nn.functional.linear(v, *self.v2).float()).reshape(-1)
        return torch.softmax(pi.float(), dim=1), v

    def _capture(self):
        B, N = self.batchsize, self.N
        self.x_static = torch.zeros(B, self.cpad, N, N, dtype=self.dtype, device=self.device).contiguous(
            memory_format=torch.channels_last)
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(3):  # warm-up outside the capture: cuDNN plans, workspace
                self._forward(self.x_static)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.pi_static, self.v_static = self._forward(self.x_static)

    def _as_input(self, batch, i, j):
        """positions [i,j) as a dtype [m,cpad,N,N] channels_last view or copy"""
        if "s_nhwc" in batch:
            return batch["s_nhwc"][i:j].permute(0, 3, 1, 2)  # [m,N,N,cpad] memory == channels_last
        s = batch["s"][i:j]
        x = torch.zeros(s.shape[0], self.cpad, self.N, self.N, dtype=self.dtype, device=s.device).contiguous(
            memory_format=torch.channels_last)
        x[:, : s.shape[1]] = s
        return x

    @torch.no_grad()
    def __call__(self, batch):
        src = batch["s_nhwc"] if "s_nhwc" in batch else batch["s"]
        n, B = src.shape[0], self.batchsize
        pi = torch.empty(n, self.N * self.N + 1, dtype=torch.float32, device=src.device)
        v = torch.empty(n, dtype=torch.float32, device=src.device)
        for i in range(0, n, B):
            j = min(i + B, n)
            if self.graph is not None and j - i == B:
                if "s_nhwc" in batch:
                    self.x_static.copy_(batch["s_nhwc"][i:j].permute(0, 3, 1, 2))
                else:
                    self.x_static[:, : src.shape[1]].copy_(src[i:j])  # float32 NCHW -> 16-bit NHWC in one pass
                self.graph.replay()
                pi[i:j].copy_(self.pi_static)
                v[i:j].copy_(self.v_static)
            else:
                p, vv = self._forward(self._as_input(batch, i, j))
                pi[i:j], v[i:j] = p, vv
            self.num_batches += 1
        self.num_positions += n
        return {"pi": pi, "V": v}


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
