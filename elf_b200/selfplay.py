"""SelfPlay: the per-move driver of the reference's self-play game loop, for G games at once.

Mirrors ``GoGameSelfPlay::act`` (``src_cpp/elfgames/go/common/game_selfplay.cc:272-430``):
one MCTS per move, ``mcts_make_diverse_move`` (sample from the visit distribution while
``ply <= policy_distri_cutoff``, ``:80-95``), ``MCTSGoAI::getValue`` as predicted value,
``ResignCheck`` (``common/game_utils.h:15-54``: resign when the side to move's value is below
``-1 + resign_thres`` and ``ply >= 50``; a ``never_resign_ratio`` fraction of games never resigns),
``GoState::forward``, game end on two passes / ply cap / superko / ``move_cutoff`` with the final
value from ``GoState::evaluate(komi)`` (``go_state_ext.h:79-105``), then restart.

Randomness (move sampling, never-resign draw) comes from a numpy Generator, not from the
reference's per-game ``std::mt19937`` streams: distributions are the same, streams are not.
"""
import numpy as np

from .board import GoBatch
from .mcts import MctsBatch


class SelfPlay:
    def __init__(self, actor, num_games=4096, board_size=19, device=0, policy_distri_cutoff=20,
                 resign_thres=0.05, never_resign_ratio=0.1, move_cutoff=-1, komi=7.5, seed=0,
                 record_games=False, actor_white=None, **mcts_opts):
        self.gb = GoBatch(num_games, board_size=board_size, device=device)
        mcts_opts.setdefault("komi", komi)
        self.mcts = MctsBatch(self.gb, **mcts_opts)
        self.actor = actor
        # evaluation matches (GoGameSelfPlay::_ai2, game_selfplay.cc:366-367): a second AI with its own
        # tree plays white; both trees follow every move (MCTSAI_T::advanceMoves)
        self.actor_white = actor_white
        self.mcts2 = MctsBatch(self.gb, **mcts_opts) if actor_white is not None else None
        self.G = num_games
        self.N = board_size
        self.komi = komi
        self.policy_distri_cutoff = policy_distri_cutoff
        self.resign_thres = resign_thres
        self.never_resign_ratio = never_resign_ratio
        self.move_cutoff = move_cutoff
        self.rng = np.random.default_rng(seed)
        self._seed = int(seed)
        self._move_counter = 0
        self.never_resign = self.rng.random(num_games) < never_resign_ratio
        self.moves_played = 0
        self.games_finished = 0
        self.results = []  # (final_value, plies, reason) of finished games
        self.records = []  # reference-format game records (elf_b200.record) when record_games is on
        self.recorders = None
        if record_games:
            from .record import GameRecorder

            self.recorders = [GameRecorder(board_size, g, policy_distri_cutoff) for g in range(num_games)]

    def close(self):
        self.mcts.close()
        if self.mcts2 is not None:
            self.mcts2.close()
        self.gb.close()

    @staticmethod
    def merge_results(black_to_move, res_b, res_w):
        """root statistics of the AI that is to move in each game"""
        out = {}
        for k in res_b:
            m = black_to_move if res_b[k].ndim == 1 else black_to_move[:, None]
            out[k] = np.where(m, res_b[k], res_w[k])
        return out

    def step(self):
        """one move of every game; returns the number of moves played"""
        info = self.gb.info()
        if self.mcts2 is None:
            self.mcts.search(self.actor)
        else:
            black = info[:, 1] == 1
            self.mcts.search(self.actor, active=black.astype(np.uint8))
            self.mcts2.search(self.actor_white, active=(~black).astype(np.uint8))
        return self.finish_move(info)

    def finish_move(self, info, res=None):
        """everything GoGameSelfPlay::act does after the search returned (game_selfplay.cc:372-429):
        move choice and resign check on the device (``elfb200_mcts_choose``), ``GoState::forward``,
        tree advance, game end / restart.  ``info`` are the games' info words from before the search;
        ``res`` (root tables) is only needed, and fetched, when games are being recorded."""
        self._move_counter += 1
        seed = (self._seed << 20) ^ self._move_counter
        nr = self.never_resign.astype(np.uint8)
        acts, vals = self.mcts.choose(self.policy_distri_cutoff, self.resign_thres, nr, seed)
        if self.mcts2 is not None:
            black = info[:, 1] == 1
            a2, v2 = self.mcts2.choose(self.policy_distri_cutoff, self.resign_thres, nr, seed)
            acts = np.where(black, acts, a2)
            vals = np.where(black, vals, v2)
        resign = acts == -1
        assert (acts != -2).all(), "a game was not searched"
        if self.recorders is not None:
            if res is None:
                res = self.mcts.results()
                if self.mcts2 is not None:
                    res = self.merge_results(info[:, 1] == 1, res, self.mcts2.results())
            for g in range(self.G):
                self.recorders[g].on_move(int(info[g, 0]), int(acts[g]), res["visits"][g], float(vals[g]))
        ok = self.gb.forward(acts)
        assert ok[~resign].all(), "MCTS proposed an illegal move"
        self.mcts.advance(acts)
        if self.mcts2 is not None:
            self.mcts2.advance(acts)
        self.moves_played += int((~resign).sum())
        info2 = self.gb.info()
        done = resign | (info2[:, 9] == 1)
        if self.move_cutoff > 0:
            done |= info2[:, 0] >= self.move_cutoff
        if done.any():
            final = self.gb.evaluate(self.komi)
            for g in np.flatnonzero(done):
                if resign[g]:
                    fv, why = (1.0 if info[g, 1] == 2 else -1.0), "resign"
                else:
                    fv = float(final[g])
                    why = "two_pass" if info2[g, 10] else ("superko" if info2[g, 11] else "max_step")
                self.results.append((fv, int(info2[g, 0]), why))
                if self.recorders is not None:
                    self.records.append(self.recorders[g].finish(fv, bool(self.never_resign[g]),
                                                                 resign_thres=self.resign_thres,
                                                                 never_resign_prob=self.never_resign_ratio))
            m = done.astype(np.uint8)
            self.gb.reset(m)
            self.mcts.reset(m)
            if self.mcts2 is not None:
                self.mcts2.reset(m)
            self.never_resign[done] = self.rng.random(int(done.sum())) < self.never_resign_ratio
            self.games_finished += int(done.sum())
        return int((~resign).sum())
