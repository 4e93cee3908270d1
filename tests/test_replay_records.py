"""Self-play records and the training-sample path pinned on the compiled
reference (oracle/_ref, ref_offline_shim.cc):

* every record the engine writes is accepted by the reference's own ``Record::createFromJson``
  (whose ``JSON_LOAD`` throws on any missing field -- ``createBatchFromJson`` would silently drop
  the record) and comes back field for field;
* ``quantise_policy`` == ``MCTSPolicy::normalize`` + ``GoStateExt::addMCTSPolicy``;
* ``ReplayBatch.sample`` == ``GoStateExtOffline`` + the ``train`` extractors of ``GoFeature`` for
  every field, all 8 D4 codes, moves with and without a stored policy, 1 and 3 future actions.

CPU-only: records come from the self-play host loop running on oracle boards and a stub search
(tests/test_request_protocol.py), replay runs on the same oracle boards."""
import json

import numpy as np
import pytest
import torch

from elf_b200 import record, replay
from elf_b200.selfplay import SelfPlay
from tests import oracles
from tests.test_request_protocol import Boards, P1, Search

needs_ref9 = pytest.mark.skipif(not oracles.have_ref(9), reason="oracle/_ref not built")
needs_ref19 = pytest.mark.skipif(not oracles.have_ref(19), reason="oracle/_ref not built")


def same(a, b, path=""):
    if isinstance(a, dict):
        assert set(a) == set(b), (path, set(a) ^ set(b))
        for k in a:
            same(a[k], b[k], path + "/" + k)
    elif isinstance(a, list):
        assert len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            same(x, y, f"{path}[{i}]")
    elif isinstance(a, float) or isinstance(b, float):
        assert abs(a - b) <= 1e-6 * max(1.0, abs(a)), (path, a, b)  # the reference stores float32
    else:
        assert a == b, (path, a, b)


@pytest.fixture(scope="module")
def selfplay_records(oracle_lib):
    """records of 9x9 games from the real host loop (SelfPlay.finish_move + GameRecorder)"""
    G = 4
    b = Boards(G, oracle_lib)
    rng = np.random.default_rng(11)

    def actor(batch):
        k = batch["s"].shape[0]
        return {"pi": torch.from_numpy(rng.random((k, P1)).astype(np.float32)), "V": torch.from_numpy(
            rng.uniform(-1, 1, k).astype(np.float32))}

    class VisitSearch(Search):
        """the stub has no visit counts of its own: derive a spread-out table from the policy"""

        def results(self):
            vis = np.where(self.pi > 0.5, (self.pi * 40).astype(np.int32), -1).astype(np.int32)
            return {"visits": vis}

    sp = SelfPlay(actor, num_games=G, board_size=9, policy_distri_cutoff=6, never_resign_ratio=0.5, move_cutoff=14,
                  record_games=True, board=b, search=VisitSearch(b, "ai"), num_rollouts=32, num_rollouts_per_batch=4,
                  c_puct=1.5, virtual_loss=1, persistent_tree=1, root_epsilon=0.25, root_alpha=0.03, seed=5)
    sp.set_request(12, -1, 0.05, never_resign_prob=0.5)
    while len(sp.records) < 6:
        sp.step()
    return sp.records


@needs_ref9
def test_records_survive_reference_parser(selfplay_records):
    for rec in selfplay_records:
        back = oracles.ref_record_roundtrip(json.dumps(rec))
        assert back is not None, "the reference's Record::createFromJson refused the record"
        same(rec, json.loads(back))
        res = rec["result"]
        assert res["num_move"] == 13 and len(res["policies"]) == 6 and len(res["values"]) == 13  # cutoff 6, 13 moves
        assert rec["request"]["vers"]["mcts_opt"]["num_rollouts_per_thread"] == 32
        assert rec["request"]["vers"]["mcts_opt"]["alg_opt"]["c_puct"] == 1.5
    text = record.dumps(selfplay_records)
    assert oracles.ref_record_batch_count(text) == len(selfplay_records)
    broken = json.loads(text)
    del broken[0]["request"]["vers"]["mcts_opt"]["virtual_loss"]  # what an incomplete mcts_opt costs
    assert oracles.ref_record_batch_count(json.dumps(broken)) == len(selfplay_records) - 1


@needs_ref9
@needs_ref19
def test_policy_quantisation_matches_reference():
    rng = np.random.default_rng(2)
    for trial in range(200):
        n = 9 if trial % 2 else 19
        k = int(rng.integers(1, 40))
        acts = np.sort(rng.choice(n * n + 1, k, replace=False))
        vis = rng.integers(0, 800, k)
        if vis.sum() == 0:
            vis[0] = 1
        row = np.full(n * n + 1, -1, np.int32)
        row[acts] = vis
        want = oracles.ref_quantise_policy(acts, vis.astype(np.float32), n)
        got = np.array(record.quantise_policy(row, n), np.uint8)
        np.testing.assert_array_equal(got, want)


def check_against_reference(rb, recs_json, picks, n):
    out = rb.sample(picks)
    for b, (i, m, code) in enumerate(picks):
        want = oracles.ref_offline_sample(recs_json[i], m, code, rb.K, n)
        assert isinstance(want, dict), want
        np.testing.assert_array_equal(out["s"][b], want["s"])
        np.testing.assert_array_equal(out["offline_a"][b], want["offline_a"])
        np.testing.assert_array_equal(out["mcts_scores"][b], want["mcts_scores"])  # same float32 arithmetic
        assert out["winner"][b] == want["winner"] and out["move_idx"][b] == want["move_idx"] == m
        assert out["num_move"][b] == want["num_move"] and out["aug_code"][b] == want["aug_code"] == code
        assert out["selfplay_ver"][b] == want["selfplay_ver"]
        assert out["predicted_value"][b] == pytest.approx(want["predicted_value"], abs=1e-7)
    return out


@needs_ref9
@pytest.mark.parametrize("K", [1, 3])
def test_replay_samples_match_reference_extractors(selfplay_records, oracle_lib, K):
    n, B = 9, 8
    rb = replay.ReplayBatch(B, board_size=n, num_future_actions=K, seed=1, board=Boards(B, oracle_lib))
    assert rb.add_records(record.dumps(selfplay_records)) == len(selfplay_records)
    recs_json = [json.dumps(r) for r in selfplay_records]
    L = 13
    # all eight symmetries; first / stored-policy / beyond-the-cutoff / last admissible move
    for moves in ([0, 1, 5, 6, 7, L - K, 3, 2], [L - K, 4, 0, 9, 6, 5, 1, 8]):
        picks = [(b % len(recs_json), moves[b], b) for b in range(B)]
        out = check_against_reference(rb, recs_json, picks, n)
        assert out["selfplay_ver"].tolist() == [12] * B and set(out["winner"]) <= {1.0, -1.0}
        # a stored policy is a distribution, a missing one is one-hot on the move played
        np.testing.assert_allclose(out["mcts_scores"].sum(1), 1.0, rtol=1e-5)
        for b, (_, m, _) in enumerate(picks):
            if m >= 6:
                assert out["mcts_scores"][b].max() == 1.0 and out["mcts_scores"][b].argmax() == out["offline_a"][b, 0]
    # random draws stay inside switchRandomMove's range and also agree
    picks = rb.draw()
    assert all(0 <= m <= L - K and 0 <= c < 8 for _, m, c in picks)
    check_against_reference(rb, recs_json, picks, n)
    with pytest.raises(ValueError):
        rb.sample([(0, L - K + 1, 0)] * B)
    assert oracles.ref_offline_sample(recs_json[0], L - K + 1, 0, K, n) == -2


@needs_ref19
def test_replay_19x19_with_captures(oracle_lib):
    """a longer 19x19 record (random legal game, synthetic policies for the first 30 moves)"""
    n, B, K = 19, 4, 2
    rng = np.random.default_rng(4)
    o = oracles.Oracle(n, oracle_lib)
    rec = record.GameRecorder(n, 0, 30, mcts_opt=dict(num_rollouts=800))
    for t in range(150):
        legal = np.flatnonzero(o.legal())
        a = int(rng.choice(legal)) if len(legal) and t % 37 != 36 else n * n
        row = np.full(n * n + 1, -1, np.int32)
        row[rng.choice(n * n + 1, 25, replace=False)] = rng.integers(1, 300, 25)
        rec.on_move(int(o.info()[0]), a, row, float(rng.uniform(-1, 1)))
        assert o.forward(a)
    assert o.info()[2] + o.info()[3] > 0  # stones were captured along the way
    r = rec.finish(-1.0, True, model_ver=3)
    text = json.dumps(r)
    same(r, json.loads(oracles.ref_record_roundtrip(text, 19)))
    rb = replay.ReplayBatch(B, board_size=n, num_future_actions=K, board=Boards(B, oracle_lib, n))
    rb.add_records([r])
    check_against_reference(rb, [text], [(0, 148, 5), (0, 0, 3), (0, 29, 6), (0, 97, 7)], n)


REF_UTILS = "/root/reference/src_py/elf/utils_elf.py"


def test_train_label_through_compat_surface(selfplay_records, oracle_lib):
    """mode == "train": the reference's GCWrapper pumps `train` batches out of the replay engine"""
    import importlib.util
    import os

    from elf_b200 import compat

    if not os.path.exists(REF_UTILS):
        pytest.skip("reference tree not present")
    spec = importlib.util.spec_from_file_location("ref_utils_elf_train", REF_UTILS)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    B, K = 6, 2
    rb = replay.ReplayBatch(B, board_size=9, num_future_actions=K, seed=9, board=Boards(B, oracle_lib))
    rb.add_records(selfplay_records)
    twin = replay.ReplayBatch(B, board_size=9, num_future_actions=K, seed=9, board=Boards(B, oracle_lib))
    twin.add_records(selfplay_records)
    GC = compat.GameContext(compat.TrainEngine(rb), batchsize=B)
    desc = {"train": dict(input=["s", "offline_a", "winner", "mcts_scores", "move_idx", "selfplay_ver"], reply=None)}
    gcw = ref.GCWrapper(GC, B, desc, num_recv=2, gpu=None, use_numpy=False, params=GC.getParams())
    seen = []
    gcw.reg_callback("train", lambda batch: seen.append({k: batch[k].clone() for k in desc["train"]["input"]}))
    gcw.start()
    for _ in range(3):
        gcw.run()
    gcw.stop()
    assert len(seen) == 3
    for got in seen:
        want = twin.sample()
        assert got["s"].shape == (B, 18, 9, 9) and got["offline_a"].shape == (B, K) and got["offline_a"].dtype == torch.int64
        assert got["move_idx"].dtype == torch.int32 and got["mcts_scores"].shape == (B, 82)
        for k in got:
            np.testing.assert_array_equal(got[k].numpy(), want[k])
