"""ReplayBatch: training samples from self-play records, featurised on the GPU boards.

Reference: ``GoGameTrain::act`` (``src_cpp/elfgames/go/train/game_train.cc:22-53``: sample a record,
``fromRecord``, ``switchRandomMove``, ``generateD4Code``, send the ``train`` label),
``GoStateExtOffline`` (``common/go_state_ext.h:259-335``) and the ``train`` extractors of
``GoFeature`` (``common/game_feature.h:75-139``): ``s`` (extractAGZ under the sample's D4 code),
``offline_a`` (the next ``num_future_actions`` moves in NN orientation), ``winner``, ``mcts_scores``
(the u8 policy of that move read through ``action2Coord`` and renormalised, or one-hot on the move
when the record holds no policy for it), ``move_idx``, ``num_move``, ``predicted_value``,
``aug_code``, ``selfplay_ver``.

The reference replays each record on one CPU thread; here the ``B`` samples of a batch are replayed
by one ``elfb200_replay`` launch on a ``GoBatch`` of ``B`` games (every game forwards its own move
list, positions in registers) and the planes of all of them come from one ``elfb200_features``
launch.  Everything else is host arithmetic on the record.
"""
import json

import numpy as np

from . import sgf as _sgf
from .record import action_to_coord


def d4_transform(x, y, n, code):
    """BoardFeature::Transform (base/board_feature.h:98-114): board (x, y) -> NN orientation"""
    rot, flip = code % 4, (code >> 2) == 1
    if rot == 1:
        x, y = y, n - x - 1
    elif rot == 2:
        x, y = n - x - 1, n - y - 1
    elif rot == 3:
        x, y = n - y - 1, x
    if flip:
        x, y = y, x
    return x, y


def d4_inverse(x, y, n, code):
    """BoardFeature::InvTransform (base/board_feature.h:116-130)"""
    rot, flip = code % 4, (code >> 2) == 1
    if flip:
        x, y = y, x
    if rot == 1:
        x, y = n - y - 1, x
    elif rot == 2:
        x, y = n - x - 1, n - y - 1
    elif rot == 3:
        x, y = y, n - x - 1
    return x, y


def board_to_nn_action(a, n, code):
    """BoardFeature::coord2Action (board_feature.h:132-137) in action space"""
    if a == n * n:
        return a
    x, y = d4_transform(a // n, a % n, n, code)
    return x * n + y


def nn_to_board_action(a, n, code):
    """BoardFeature::action2Coord (board_feature.h:139-146) in action space"""
    if a == n * n:
        return a
    x, y = d4_inverse(a // n, a % n, n, code)
    return x * n + y


class ReplayBatch:
    def __init__(self, num_states, board_size=19, device=0, num_future_actions=1, seed=0, board=None,
                 use_df_feature=False):
        # GameOptions::use_df_feature (common/game_feature.h:22-33): "s" carries the 25 DarkForest planes
        # (BoardFeature::extract) instead of the 18 AGZ planes
        self.use_df_feature = bool(use_df_feature)
        self.num_planes = 25 if self.use_df_feature else 18
        if board is None:
            from .board import GoBatch

            board = GoBatch(num_states, board_size=board_size, device=device)
        self.board = board
        self.B = board.num_games
        self.N = board.board_size
        self.K = int(num_future_actions)
        self.rng = np.random.default_rng(seed)
        self.records = []  # parsed: dict(moves, winner, policies, values, ver)

    # -- GoStateExtOffline::fromRecord -------------------------------------------------------------
    def add_records(self, records):
        """records: list of record dicts in the reference's JSON layout, or that JSON as a string"""
        if isinstance(records, (str, bytes)):
            records = json.loads(records)
        n = self.N
        for r in records:
            res = r["result"]
            moves = _sgf.sgfstr2actions(res["content"], n)
            if any(a < 0 or a > n * n for a in moves):
                # a malformed vertex parses to M_INVALID in the reference and GoState::forward(M_INVALID)
                # throws there (go_state.cc:74-77); reject the record up front instead of replaying it
                raise ValueError("record contains an invalid move: " + res["content"][:60])
            self.records.append({
                "moves": moves,
                "winner": 1.0 if res["reward"] > 0 else -1.0,
                "policies": res.get("policies", []),
                "values": res.get("values", []),
                "ver": int(r["request"]["vers"]["black_ver"]),
            })
        return len(self.records)

    def usable(self, rec):  # switchRandomMove's guard (go_state_ext.h:285-293)
        return len(rec["moves"]) > self.K - 1

    def draw(self):
        """the random choices of GoGameTrain::act for B samples: (record index, move_to, d4 code)"""
        ok = [i for i, r in enumerate(self.records) if self.usable(r)]
        if not ok:
            raise RuntimeError("no record with at least num_future_actions moves")
        picks = []
        for _ in range(self.B):
            i = ok[int(self.rng.integers(len(ok)))]
            m = len(self.records[i]["moves"])
            picks.append((i, int(self.rng.integers(m - self.K + 1)), int(self.rng.integers(8))))
        return picks

    # -- one `train` batch ---------------------------------------------------------------------------
    def sample(self, picks=None, s_out=None):
        """one ``train`` batch as a dict of arrays.  ``s_out``: optional float32 torch tensor
        ``[B, 18, N, N]`` (contiguous) on the boards' device: the feature kernel then writes the planes
        straight into it (``elfb200_features_dev``) and ``out["s"]`` is that tensor -- the 26 KB per
        position never cross PCIe; all other fields are small host arrays."""
        n, B, K, A = self.N, self.B, self.K, self.N * self.N + 1
        picks = self.draw() if picks is None else list(picks)
        assert len(picks) == B
        recs = [self.records[i] for i, _, _ in picks]
        move_to = np.array([m for _, m, _ in picks], np.int64)
        d4 = np.array([c for _, _, c in picks], np.int32)
        for r, m in zip(recs, move_to):
            if not (0 <= m <= len(r["moves"]) - K):
                raise ValueError("move index outside switchRandomMove's range")
        # switchBeforeMove for all samples at once
        if hasattr(self.board, "replay"):  # one launch: every game forwards its own move list (k_replay)
            self.board.replay([r["moves"][:m] for r, m in zip(recs, move_to)])
        else:  # boards without the replay entry point: one step per ply, finished samples idle
            self.board.reset(None)
            for t in range(int(move_to.max()) if B else 0):
                acts = np.array([r["moves"][t] if t < m else -1 for r, m in zip(recs, move_to)], np.int32)
                self.board.forward(acts)  # the reference ignores forward()'s verdict here as well
        # the reference's extractors index the labels with _state.getPly() - 1 (game_feature.h:75-139),
        # i.e. the number of moves the board ACCEPTED, not the sampled move_to: a refused move in a record
        # (forward()'s verdict is ignored by switchBeforeMove) shifts the labels with the position
        move_idx = self.board.info()[:, 0].astype(np.int64) - 1
        if s_out is not None:
            import torch

            assert tuple(s_out.shape) == (B, self.num_planes, n, n) and s_out.dtype == torch.float32 and s_out.is_contiguous()
            d4_dev = torch.as_tensor(d4, device=s_out.device)
            if s_out.is_cuda:
                torch.cuda.current_stream(s_out.device).synchronize()  # d4_dev is ready, s_out is free
            if self.use_df_feature:
                _l = self.board._lib
                from .lib import check

                check(_l, _l.elfb200_features_df_dev(self.board._ctx, d4_dev.data_ptr(), s_out.data_ptr()))
            else:
                self.board.features_dev(s_out.data_ptr(), d4_dev.data_ptr())
            self.board.synchronize()
        out = {
            "s": s_out if s_out is not None else (self.board.features_df(d4) if self.use_df_feature else self.board.features(d4)),
            "offline_a": np.zeros((B, K), np.int64),
            "winner": np.array([r["winner"] for r in recs], np.float32),
            "mcts_scores": np.zeros((B, A), np.float32),
            "move_idx": move_idx.astype(np.int32),
            "num_move": np.array([len(r["moves"]) for r in recs], np.int32),
            "predicted_value": np.array([r["values"][m] if m < len(r["values"]) else 0.0
                                         for r, m in zip(recs, move_idx)], np.float32),
            "aug_code": d4.copy(),
            "selfplay_ver": np.array([r["ver"] for r in recs], np.int64),
        }
        for b, (r, m, code) in enumerate(zip(recs, move_idx, d4)):
            m = int(m)
            for k in range(K):  # extractOfflineAction
                out["offline_a"][b, k] = board_to_nn_action(r["moves"][m + k], n, int(code))
            sc = out["mcts_scores"][b]
            if m < len(r["policies"]):  # extractMCTSPi
                prob = r["policies"][m]
                tot = np.float32(0)
                for a in range(A):
                    sc[a] = prob[action_to_coord(nn_to_board_action(a, n, int(code)), n)]
                    tot = np.float32(tot + sc[a])
                with np.errstate(divide="ignore", invalid="ignore"):
                    sc /= tot
            else:
                sc[board_to_nn_action(r["moves"][m], n, int(code))] = 1.0
        return out
