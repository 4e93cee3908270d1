"""MctsBatch: the batched GPU tree search, mirroring the reference's MCTSGoAI for G games at once.

Reference: ``src_cpp/elf/ai/tree_search/mcts.h:29-170`` (MCTSAI_T::act), ``tree_search.h``
(TreeSearchT::run / batch_rollouts), ``src_cpp/elfgames/go/mcts/mcts.h:351-381`` (MCTSGoAI) and
the Python NN callback ``src_py/rlpytorch/trainer/trainer.py:73-115`` (Evaluator.actor), whose
contract is kept: ``actor(batch) -> {"pi": [n, N*N+1], "V": [n]}`` with ``batch["s"]`` a CUDA
float tensor ``[n, 18, N, N]``.
"""
import ctypes

import numpy as np

from . import lib as _l


class MctsBatch:
    def __init__(self, go_batch, **opts):
        import torch

        self.gb = go_batch
        self._lib = _l.load_library()
        o = _l.MctsOptions()
        _l.check(self._lib, self._lib.elfb200_mcts_default_options(ctypes.byref(o)))
        for k, v in opts.items():
            if not hasattr(o, k):
                raise TypeError(f"unknown MCTS option {k}")
            setattr(o, k, v)
        self.options = o
        self._m = _l.vp()
        _l.check(self._lib, self._lib.elfb200_mcts_create(go_batch._ctx, ctypes.byref(o), ctypes.byref(self._m)))
        import weakref

        go_batch._children.append(weakref.ref(self))
        n = go_batch.board_size
        self.waves_per_move = self._lib.elfb200_mcts_waves_per_move(self._m)
        self.max_leaves = self._lib.elfb200_mcts_max_leaves(self._m)
        self._torch = torch
        self.device = torch.device("cuda", go_batch.device)
        # the leaf feature batch handed to the network: lives on the device, written in place
        self.feat = torch.zeros((self.max_leaves, 18, n, n), dtype=torch.float32, device=self.device)
        self._stream = torch.cuda.ExternalStream(go_batch.stream, device=self.device)

    def close(self):
        if getattr(self, "_m", None):
            # the search handle points into the board context: if that is already gone (interpreter
            # shutdown can finalise objects in any order) the handle must not be touched any more
            if getattr(self.gb, "_ctx", None):
                self._lib.elfb200_mcts_destroy(self._m)
            self._m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        _l.check(self._lib, self._lib.elfb200_mcts_reset(self._m, m.ctypes.data if m is not None else None))

    def begin_move(self, active=None):
        a = None if active is None else np.ascontiguousarray(active, dtype=np.uint8)
        _l.check(self._lib, self._lib.elfb200_mcts_begin_move(self._m, a.ctypes.data if a is not None else None))

    def select(self):
        """one wave of descents; returns the CUDA feature tensor view [n, 18, N, N] of the leaves
        that need the network (n may be 0)"""
        n = ctypes.c_int32()
        _l.check(self._lib, self._lib.elfb200_mcts_select(self._m, self.feat.data_ptr(), ctypes.byref(n)))
        self._n = n.value
        return self.feat[: n.value]

    def leaf_info(self):
        n = self._n
        h = np.empty(n, np.uint64)
        g = np.empty(n, np.int32)
        p = np.empty(n, np.int32)
        self.leaf_d4 = np.empty(n, np.int32)
        _l.check(self._lib, self._lib.elfb200_mcts_leaf_info(self._m, h.ctypes.data, g.ctypes.data, p.ctypes.data,
                                                             self.leaf_d4.ctypes.data))
        return h, g, p

    def expand_backup(self, pi, v):
        """pi: CUDA float32 [n, N*N+1], v: CUDA float32 [n] (contiguous), produced on any stream
        the caller has synchronised with the context stream."""
        if self._n > 0:
            assert pi.is_cuda and v.is_cuda and pi.dtype == self._torch.float32 and v.dtype == self._torch.float32
            assert pi.is_contiguous() and v.is_contiguous() and pi.shape[0] >= self._n
            _l.check(self._lib, self._lib.elfb200_mcts_expand_backup(self._m, pi.data_ptr(), v.data_ptr()))
        else:
            _l.check(self._lib, self._lib.elfb200_mcts_expand_backup(self._m, None, None))

    def results(self):
        G, P1 = self.gb.num_games, self.gb.num_actions
        best = np.empty(G, np.int32)
        visits = np.empty((G, P1), np.int32)
        rootv = np.empty(G, np.float32)
        bestq = np.empty(G, np.float32)
        tot = np.empty(G, np.int32)
        _l.check(self._lib, self._lib.elfb200_mcts_results(
            self._m, best.ctypes.data, visits.ctypes.data, rootv.ctypes.data, bestq.ctypes.data, tot.ctypes.data))
        return {"best_action": best, "visits": visits, "root_value": rootv, "best_q": bestq, "total_visits": tot}

    def choose(self, policy_distri_cutoff, resign_thres, never_resign=None, seed=0):
        """device-side move choice (sample ~ visits while ply <= cutoff, else most visited; -1 =
        resign, -2 = not searched); returns (actions int32[G], values float32[G])"""
        G = self.gb.num_games
        a = np.empty(G, np.int32)
        v = np.empty(G, np.float32)
        nr = None if never_resign is None else np.ascontiguousarray(never_resign, dtype=np.uint8)
        _l.check(self._lib, self._lib.elfb200_mcts_choose(
            self._m, int(policy_distri_cutoff), float(resign_thres), nr.ctypes.data if nr is not None else None,
            int(seed) & 0xFFFFFFFFFFFFFFFF, a.ctypes.data, v.ctypes.data))
        return a, v

    def advance(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.int32)
        _l.check(self._lib, self._lib.elfb200_mcts_advance(self._m, a.ctypes.data))

    def root_priors(self):
        """float32 [G, N*N+1]: prior of every root edge by action, -1 where there is no edge"""
        o = np.empty((self.gb.num_games, self.gb.num_actions), np.float32)
        _l.check(self._lib, self._lib.elfb200_mcts_root_priors(self._m, o.ctypes.data))
        return o

    def errors(self):
        e = np.zeros(4, np.int32)
        _l.check(self._lib, self._lib.elfb200_mcts_errors(self._m, e.ctypes.data))
        return e

    def eval_count(self):
        return self._lib.elfb200_mcts_eval_count(self._m)

    def timings(self, reset=False):
        """(ms[4] = select, features, expand, backup kernel time; number of waves)"""
        ms = np.zeros(4, np.float64)
        w = ctypes.c_int64()
        _l.check(self._lib, self._lib.elfb200_mcts_timings(self._m, ms.ctypes.data, ctypes.byref(w), int(reset)))
        return ms, w.value

    def stats(self):
        """uint64[4]: descent steps, edge records read, nodes created, stored edges of visited nodes"""
        s = np.zeros(4, np.uint64)
        _l.check(self._lib, self._lib.elfb200_mcts_stats(self._m, s.ctypes.data))
        return s

    # -- MCTSAI_T::act for all games -----------------------------------------------------------
    def act(self, actor, active=None):
        """Run one full search (all waves) with ``actor(batch) -> {"pi", "V"}`` as the network
        callback and return the root statistics.  The callback sees ``batch["s"]`` on the GPU."""
        self.search(actor, active)
        return self.results()

    def search(self, actor, active=None, waves=None):
        """the search of ``act`` without fetching the root tables (use ``choose`` / ``results``).
        ``waves``: run only that many waves (1 on a fresh tree = evaluate and expand the root only,
        TreeSearchT::runPolicyOnly, tree_search.h:387-408); default: the whole move"""
        torch = self._torch
        self.begin_move(active)
        pad = int(getattr(actor, "batchsize", 0) or 0)
        for _ in range(self.waves_per_move if waves is None else int(waves)):
            s = self.select()
            if s.shape[0] > 0:
                if pad > 1:
                    # static NN shapes: round the batch up to a multiple of the actor's batch size
                    # (rows past n are stale features; their replies are never read)
                    n_pad = min(-(-s.shape[0] // pad) * pad, self.max_leaves)
                    s = self.feat[:n_pad]
                self.gb.synchronize()  # features written on the context stream
                with torch.no_grad():
                    reply = actor({"s": s})
                pi = reply["pi"].to(torch.float32).contiguous()
                v = reply["V"].to(torch.float32).reshape(-1).contiguous()
                torch.cuda.current_stream(self.device).synchronize()
                self.expand_backup(pi, v)
            else:
                self.expand_backup(None, None)
