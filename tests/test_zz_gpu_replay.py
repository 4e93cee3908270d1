"""elfb200_replay / ReplayBatch on the GPU board batch.

k_replay is k_step's body in a loop (same device functions); its logic is also covered on the SIMT
emulator (tests/test_emu_kernels.py::test_replay_kernel, three lane orders) and ReplayBatch's
arithmetic is pinned on the compiled reference (tests/test_replay_records.py).  First green
hardware run: round 1's final GPU test pass."""
import numpy as np
import pytest

from tests import oracles

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


@pytest.mark.parametrize("n,G", [(9, 100), (19, 64)])
def test_replay_equals_stepwise_forward(oracle_lib, n, G):
    import elf_b200

    rng = np.random.default_rng(n)
    lists, os_ = [], []
    for g in range(G):
        o = oracles.Oracle(n, oracle_lib)
        mv = []
        for _ in range(int(rng.integers(0, 70 if n == 9 else 250))):
            lg = np.flatnonzero(o.legal())
            a = int(rng.choice(lg)) if len(lg) and rng.random() > 0.04 else n * n
            if rng.random() < 0.05:
                a = int(rng.integers(n * n))
            mv.append(a)
            o.forward(a)
        lists.append(mv)
        os_.append(o)
    gb = elf_b200.GoBatch(G, board_size=n)
    gb.forward(np.full(G, 3, np.int32))
    gb.replay(lists)
    ref = elf_b200.GoBatch(G, board_size=n)
    for t in range(max(len(m) for m in lists)):
        ref.forward(np.array([m[t] if t < len(m) else -1 for m in lists], np.int32))
    assert [int(h) for h in gb.getHashCode()] == [o.hash() for o in os_]
    np.testing.assert_array_equal(gb.info(), ref.info())
    np.testing.assert_array_equal(gb.legal_mask(), ref.legal_mask())
    d4 = np.arange(G, dtype=np.int32) % 8
    np.testing.assert_array_equal(gb.features(d4), ref.features(d4))
    for _ in range(30):
        acts = np.array([int(rng.choice(np.flatnonzero(o.legal()))) if o.legal().any() and not o.terminated() else n * n
                         for o in os_], np.int32)
        ok = gb.forward(acts)
        for g, o in enumerate(os_):
            assert bool(ok[g]) == bool(o.forward(int(acts[g])))
        assert [int(h) for h in gb.getHashCode()] == [o.hash() for o in os_]


def test_replay_batch_equals_oracle_boards(oracle_lib):
    from elf_b200 import record, replay
    from tests.test_request_protocol import Boards

    n, B = 9, 16
    rng = np.random.default_rng(6)
    recs = []
    for i in range(4):
        o = oracles.Oracle(n, oracle_lib)
        r = record.GameRecorder(n, i, 8)
        for _ in range(25 + 7 * i):
            lg = np.flatnonzero(o.legal())
            a = int(rng.choice(lg)) if len(lg) else n * n
            row = np.full(n * n + 1, -1, np.int32)
            row[rng.choice(n * n + 1, 10, replace=False)] = rng.integers(1, 99, 10)
            r.on_move(int(o.info()[0]), a, row, 0.0)
            o.forward(a)
        recs.append(r.finish(1.0 if i % 2 else -1.0, False, model_ver=i))
    a_ = replay.ReplayBatch(B, board_size=n, num_future_actions=2, seed=1)
    b_ = replay.ReplayBatch(B, board_size=n, num_future_actions=2, seed=1, board=Boards(B, oracle_lib))
    a_.add_records(recs)
    b_.add_records(recs)
    for _ in range(3):
        x, y = a_.sample(), b_.sample()
        for k in x:
            np.testing.assert_array_equal(x[k], y[k])
    import torch

    dst = torch.full((B, 18, n, n), -1.0, device="cuda")
    z, y = a_.sample(s_out=dst), b_.sample()
    assert z["s"] is dst
    np.testing.assert_array_equal(dst.cpu().numpy(), y["s"])
    np.testing.assert_array_equal(z["offline_a"], y["offline_a"])
