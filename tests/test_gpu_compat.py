"""The pybind-compatible surface with the REAL GPU engine behind it.  The driver below issues the
same calls, in the same order, as the reference's Allocator.spec2batches + GCWrapper.run()
(src_py/elf/utils_elf.py:32-99,426-437); it is re-stated here because the reference tree does not
exist on the GPU box (tests/test_compat_surface.py runs the unmodified GCWrapper on CPU)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _alloc(p, device_mode):
    f = p.field()
    dt = {"float": torch.float32, "int64_t": torch.int64, "int32_t": torch.int32}[f.type_name()]
    if device_mode:
        v = torch.zeros(*f.sz().vec(), dtype=dt, device="cuda")
    else:
        v = torch.zeros(*f.sz().vec(), dtype=dt).pin_memory()
    p.set(v.data_ptr(), [i * v.element_size() for i in v.stride()])
    return v


@pytest.mark.parametrize("device_mode", [False, True], ids=["pinned_host", "cuda_fast_mode"])
def test_wait_step_pump_plays_games(device_mode):
    import elf_b200
    from elf_b200 import compat
    from elf_b200.model import Actor, PolicyValueNet

    torch.manual_seed(0)
    n, G, BS = 9, 16, 24
    net = PolicyValueNet(n, num_block=1, dim=16).cuda()
    nn_actor = Actor(net, batchsize=BS, dtype=torch.float32, channels_last=False)
    sp = elf_b200.selfplay.SelfPlay(None, num_games=G, board_size=n, policy_distri_cutoff=0, num_rollouts=16,
                                    num_rollouts_per_batch=4, move_cutoff=12, seed=3, rotation_flip=1)
    GC = compat.GameContext(compat.SelfPlayEngine(sp), batchsize=BS)
    ctx = GC.ctx()
    keys = ["s", "pi", "V", "a", "rv"]
    opts = ctx.createSharedMemOptions("actor_black", BS)
    opts.setTimeout(10)
    smems, bufs = [], {}
    for _ in range(2):  # num_recv = 2 (game.py:428)
        sm = ctx.allocateSharedMem(opts, keys)
        bufs[sm.getSharedMemOptions().idx()] = {k: _alloc(sm[k], device_mode) for k in keys}
        smems.append(sm)
    ctx.start()
    batches = 0
    while sp.games_finished < G and batches < 4000:
        sm = ctx.wait()
        k = sm.effective_batchsize()
        assert 0 < k <= BS
        b = bufs[sm.getSharedMemOptions().idx()]
        s = b["s"][:k]
        s_gpu = s if device_mode else s.cuda(non_blocking=True)
        # planes 16/17 are the side-to-move indicators: exactly one of them is all ones
        ind = s_gpu[:, 16:18].reshape(k, 2, -1)
        assert ((ind.sum(2) == n * n).sum(1) == 1).all()
        out = nn_actor({"s": s_gpu})
        b["pi"][:k].copy_(out["pi"])
        b["V"][:k].copy_(out["V"])
        torch.cuda.synchronize()
        ctx.step()
        batches += 1
    ctx.stop()
    assert sp.games_finished >= G and sp.moves_played >= 11 * G
    assert (sp.mcts.errors() == 0).all()
    assert GC.getClient().getGameStats().getWinRateStats().total_games == sp.games_finished
    sp.close()


def test_two_models_route_to_actor_black_and_actor_white():
    """evaluation match (GoGameSelfPlay::_ai2): black's and white's searches keep separate trees
    and their leaves arrive under the labels actor_black / actor_white."""
    import elf_b200
    from elf_b200 import compat
    from elf_b200.model import Actor, PolicyValueNet

    torch.manual_seed(1)
    n, G, BS = 9, 8, 32
    nets = {lab: Actor(PolicyValueNet(n, num_block=1, dim=8).cuda(), batchsize=BS, dtype=torch.float32, channels_last=False)
            for lab in ("actor_black", "actor_white")}
    sp = elf_b200.selfplay.SelfPlay(nets["actor_black"], actor_white=nets["actor_white"], num_games=G, board_size=n,
                                    policy_distri_cutoff=0, num_rollouts=16, num_rollouts_per_batch=4, move_cutoff=10,
                                    seed=2, rotation_flip=0)
    GC = compat.GameContext(compat.SelfPlayEngine(sp), batchsize=BS)
    ctx = GC.ctx()
    keys = ["s", "pi", "V", "a", "rv"]
    bufs, counts = {}, {"actor_black": 0, "actor_white": 0}
    for lab in ("actor_black", "actor_white"):
        o = ctx.createSharedMemOptions(lab, BS)
        for _ in range(2):
            sm = ctx.allocateSharedMem(o, keys)
            bufs[sm.getSharedMemOptions().idx()] = {k: _alloc(sm[k], True) for k in keys}
    ctx.start()
    it = 0
    while sp.games_finished < G and it < 3000:
        sm = ctx.wait()
        lab = sm.getSharedMemOptions().label()
        k = sm.effective_batchsize()
        b = bufs[sm.getSharedMemOptions().idx()]
        out = nets[lab]({"s": b["s"][:k]})
        b["pi"][:k].copy_(out["pi"])
        b["V"][:k].copy_(out["V"])
        torch.cuda.synchronize()
        ctx.step()
        counts[lab] += k
        it += 1
    ctx.stop()
    assert counts["actor_black"] > 0 and counts["actor_white"] > 0
    assert sp.games_finished >= G
    assert (sp.mcts.errors() == 0).all() and (sp.mcts2.errors() == 0).all()
    sp.close()
