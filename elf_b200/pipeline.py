"""WavePipeline: several search batches (parts of the game batch) share one network.

The reference hides the network's latency behind other game threads: while one thread blocks in
``waitEvaluation`` (tree_search_node.h:169-174) the others descend (``elf::Batcher``,
``src_py/elf/utils_elf.py:378-405`` is the round trip).  Here the same overlap comes from streams:
each part of the game batch owns a board context with its own CUDA stream; the network runs on one
more stream.  While the network evaluates the leaves of part A, the descents / leaf features /
expansion / backup of part B run beside it, and the single host wait of a wave (the leaf count of B)
happens while the GPU is busy with A's network batch -- so the network stream never drains.

Search semantics are untouched: every part runs exactly the waves ``MctsBatch.search`` runs.
"""
import torch


class WavePipeline:
    def __init__(self, searches, actor, nn_stream=None):
        assert len(searches) >= 1
        self.parts = list(searches)
        self.actor = actor
        self.device = self.parts[0].device
        self.nn_stream = nn_stream if nn_stream is not None else torch.cuda.Stream(self.device)
        self._pending = [None] * len(self.parts)  # network replies not yet expanded, per part
        self._armed = [False] * len(self.parts)   # a wave of this part is in flight
        pad = int(getattr(actor, "batchsize", 0) or 0)
        for p in self.parts:
            p._pad = pad

    @property
    def waves_per_move(self):
        return self.parts[0].waves_per_move

    def begin_move(self, actives=None):
        for i, p in enumerate(self.parts):
            p.begin_move(None if actives is None else actives[i])

    def waves(self, k):
        """k more waves of every part, interleaved; returns without waiting for the last network
        batches (call drain() before reading results)"""
        for _ in range(int(k)):
            for i, p in enumerate(self.parts):
                if self._armed[i]:
                    p.wave_finish(self._pending[i])  # ordered behind its network batch by an event
                s = p.wave_select()                  # host waits for THIS part only
                self._pending[i] = p.wave_eval(self.actor, s, self.nn_stream)
                self._armed[i] = True

    def drain(self):
        for i, p in enumerate(self.parts):
            if self._armed[i]:
                p.wave_finish(self._pending[i])
                self._pending[i] = None
                self._armed[i] = False

    def search(self, actives=None, waves=None):
        """MctsBatch.search for all parts: one move's waves (or ``waves`` of them)"""
        self.begin_move(actives)
        self.waves(self.waves_per_move if waves is None else waves)
        self.drain()
