"""RefStream: the random streams of the reference's self-play game threads, for G games.

Host side of ``include/elfb200_refstream.h``.  The reference draws every random decision of a
self-play game from two ``std::mt19937`` generators per game thread -- ``GoGameBase::_rng``
(``common/game_base.h:32-38``: the actor's seed, the sampled move, the never-resign draw) and
``MCTSActor::rng_`` (``go/mcts/mcts.h:49,170``: root Dirichlet noise, one D4 code per evaluated
leaf) -- and walks its root edges in the iteration order of a libstdc++ ``unordered_map``.  With
``GameOptions::seed`` set (every game thread is then seeded alike, ``game_base.h:32-38``) its games
are reproducible; this class makes ours the same games, move for move (``SelfPlay(rng="reference")``,
single search thread).
"""
import ctypes

import numpy as np

from . import lib as _l


def _u8(mask):
    return None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)


def _p(a):
    return None if a is None else a.ctypes.data


class RefStream:
    def __init__(self, num_games, board_size, seed):
        self._lib = _l.load_library()
        self.G, self.N, self.P1 = int(num_games), int(board_size), int(board_size) ** 2 + 1
        seeds = np.ascontiguousarray(np.broadcast_to(np.asarray(seed, np.uint64), (self.G,)))
        self._h = _l.vp()
        _l.check(self._lib, self._lib.elfb200_refstream_create(self.G, self.N, seeds.ctypes.data, ctypes.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.elfb200_refstream_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init_actor(self, which=0, mask=None):
        """GoGameSelfPlay::init_ai: ``params.seed = _rng()`` seeds actor ``which`` (0 = _ai, 1 = _ai2)"""
        m = _u8(mask)
        _l.check(self._lib, self._lib.elfb200_refstream_init_actor(self._h, int(which), _p(m)))

    def game_u32(self, mask=None):
        m, out = _u8(mask), np.zeros(self.G, np.uint32)
        _l.check(self._lib, self._lib.elfb200_refstream_game_u32(self._h, _p(m), out.ctypes.data))
        return out

    def game_uniform(self, mask=None, lo=0.0, hi=1.0):
        m, out = _u8(mask), np.zeros(self.G, np.float64)
        _l.check(self._lib, self._lib.elfb200_refstream_game_uniform(self._h, _p(m), float(lo), float(hi), out.ctypes.data))
        return out

    def actor_d4(self, which, count, mask=None):
        """uint8 [G, count]: the D4 codes the next evaluated leaves would get (nothing consumed)"""
        m, out = _u8(mask), np.zeros((self.G, int(count)), np.uint8)
        _l.check(self._lib, self._lib.elfb200_refstream_actor_d4(self._h, int(which), _p(m), int(count), out.ctypes.data))
        return out

    def actor_discard(self, which, counts, mask=None):
        m, c = _u8(mask), np.ascontiguousarray(counts, dtype=np.int32)
        _l.check(self._lib, self._lib.elfb200_refstream_actor_discard(self._h, int(which), _p(m), c.ctypes.data))

    def root_noise(self, which, n_edges, actions, priors, epsilon, alpha, mask=None):
        """NodeT::enhanceExploration; ``priors`` float32 [G, P1] (storage order) is updated in place"""
        assert priors.dtype == np.float32 and priors.flags.c_contiguous and priors.shape == (self.G, self.P1)
        m = _u8(mask)
        n = np.ascontiguousarray(n_edges, dtype=np.int32)
        a = np.ascontiguousarray(actions, dtype=np.int16)
        _l.check(self._lib, self._lib.elfb200_refstream_root_noise(
            self._h, int(which), _p(m), n.ctypes.data, a.ctypes.data, priors.ctypes.data, float(epsilon), float(alpha)))
        return priors

    def choose(self, n_edges, actions, visits, sample, mask=None):
        """(best_edge, chosen_edge) int32 [G], storage-order edge indices (-1: no edges): the first
        most-visited edge in the reference's container order, and the sampled one where ``sample``"""
        m = _u8(mask)
        n = np.ascontiguousarray(n_edges, dtype=np.int32)
        a = np.ascontiguousarray(actions, dtype=np.int16)
        v = np.ascontiguousarray(visits, dtype=np.int32)
        s = _u8(sample)
        best, cho = np.empty(self.G, np.int32), np.empty(self.G, np.int32)
        _l.check(self._lib, self._lib.elfb200_refstream_choose(
            self._h, _p(m), n.ctypes.data, a.ctypes.data, v.ctypes.data, _p(s), best.ctypes.data, cho.ctypes.data))
        return best, cho

    @staticmethod
    def edge_order(board_size, actions):
        """iteration order of the reference's edge container after inserting ``actions`` in order"""
        lib = _l.load_library()
        a = np.ascontiguousarray(actions, dtype=np.int16)
        out = np.empty(len(a), np.int32)
        _l.check(lib, lib.elfb200_refstream_edge_order(int(board_size), len(a), a.ctypes.data, out.ctypes.data))
        return out


class RefStreamSearch:
    """Mixin for the search host classes (``MctsBatch``; the tests' emulator twin): the device ends of
    the reference streams.  Needs ``self._lib``, ``self._m``, ``self.gb``, ``self.options``,
    ``self.waves_per_move``.

    ``attach_ref_stream(rs, which, root_epsilon, root_alpha)``: from then on ``begin_move`` draws the
    root noise from actor ``which``'s generator in the reference's container order
    (``NodeT::enhanceExploration`` as ``TreeSearchT::run`` applies it, tree_search.h:413-417: only a
    root that already has edges), hands the leaves of the move their D4 codes from the same
    generator (``rotation_flip``) and ``ref_choose`` picks moves the way ``MCTSResultT::addActions`` /
    ``sampleAction`` do.  The search must have been created with ``root_epsilon = 0`` (the built-in
    counter-based noise off)."""

    _ref = None
    _ref_noise = True

    def set_root_noise_enabled(self, enabled):
        """TreeSearchT::runPolicyOnly (tree_search.h:387-408) does not call enhanceExploration: the
        per-move driver switches the reference-stream root noise off around the ``begin_move`` of a
        policy-only phase.  No effect without an attached stream."""
        self._ref_noise = bool(enabled)

    def root_edges(self):
        """root edges in storage order: n_edges int32 [G]; actions int16, visits int32, wsum float32,
        priors float32, all [G, N*N+1]"""
        G, P1 = self.gb.num_games, self.gb.num_actions
        out = {"n_edges": np.zeros(G, np.int32), "actions": np.zeros((G, P1), np.int16),
               "visits": np.zeros((G, P1), np.int32), "wsum": np.zeros((G, P1), np.float32),
               "priors": np.zeros((G, P1), np.float32)}
        _l.check(self._lib, self._lib.elfb200_mcts_root_edges(
            self._m, out["n_edges"].ctypes.data, out["actions"].ctypes.data, out["visits"].ctypes.data,
            out["wsum"].ctypes.data, out["priors"].ctypes.data))
        return out

    def set_root_priors(self, priors, mask=None):
        p = np.ascontiguousarray(priors, dtype=np.float32)
        assert p.shape == (self.gb.num_games, self.gb.num_actions)
        m = _u8(mask)
        _l.check(self._lib, self._lib.elfb200_mcts_set_root_priors(self._m, _p(m), p.ctypes.data))

    def set_d4_stream(self, codes):
        if codes is None:
            _l.check(self._lib, self._lib.elfb200_mcts_set_d4_stream(self._m, None, 0))
            return
        c = np.ascontiguousarray(codes, dtype=np.uint8)
        assert c.ndim == 2 and c.shape[0] == self.gb.num_games
        _l.check(self._lib, self._lib.elfb200_mcts_set_d4_stream(self._m, c.ctypes.data, c.shape[1]))

    def d4_used(self):
        u = np.zeros(self.gb.num_games, np.int32)
        _l.check(self._lib, self._lib.elfb200_mcts_d4_used(self._m, u.ctypes.data))
        return u

    def attach_ref_stream(self, rs, which=0, root_epsilon=0.0, root_alpha=0.0):
        if self.options.root_epsilon != 0:
            raise ValueError("create the search with root_epsilon = 0: the noise comes from the reference stream")
        self._ref = (rs, int(which), float(root_epsilon), float(root_alpha))
        self._ref_pending = None

    def _ref_settle(self):
        """consume the D4 draws the last move used (the device counted them)"""
        if self._ref is None or self._ref_pending is None:
            return
        rs, which = self._ref[0], self._ref[1]
        rs.actor_discard(which, self.d4_used(), self._ref_pending)
        self._ref_pending = None

    def _ref_begin(self, active):
        """after the device's begin_move: root noise, then this move's D4 codes"""
        if self._ref is None:
            return
        rs, which, eps, alpha = self._ref
        act = np.ones(self.gb.num_games, np.uint8) if active is None else np.ascontiguousarray(active, dtype=np.uint8)
        if eps > 0 and self._ref_noise:
            e = self.root_edges()
            if (e["n_edges"][act.astype(bool)] > 0).any():
                rs.root_noise(which, e["n_edges"], e["actions"], e["priors"], eps, alpha, mask=act)
                self.set_root_priors(e["priors"], mask=act)
        if self.options.rotation_flip:
            B = int(self.options.num_rollouts_per_batch)
            self.set_d4_stream(rs.actor_d4(which, self.waves_per_move * B, mask=act))
            self._ref_pending = act

    def ref_choose(self, sample, mask=None, root_value=None):
        """the reference's move choice for the searched games: dict(action, best_action int32 [G] (-2 where
        the game has no root edges / is masked out), value float32 [G] = MCTSGoAI::getValue)"""
        self._ref_settle()
        rs = self._ref[0]
        e = self.root_edges()
        best, cho = rs.choose(e["n_edges"], e["actions"], e["visits"], sample, mask=mask)
        G = self.gb.num_games
        ok = best >= 0
        rows = np.arange(G)
        b = np.where(ok, best, 0)
        c = np.where(ok, cho, 0)
        tot = np.where(np.arange(e["visits"].shape[1])[None, :] < e["n_edges"][:, None], e["visits"], 0).sum(1)
        if root_value is None:
            root_value = self.results()["root_value"]
        with np.errstate(divide="ignore", invalid="ignore"):
            q = e["wsum"][rows, b] / e["visits"][rows, b].astype(np.float32)  # EdgeInfo::getQSA
        value = np.where(tot == 0, root_value, q).astype(np.float32)
        return {"action": np.where(ok, e["actions"][rows, c], -2).astype(np.int32),
                "best_action": np.where(ok, e["actions"][rows, b], -2).astype(np.int32),
                "value": np.where(ok, value, 0).astype(np.float32), "total_visits": tot.astype(np.int32)}
