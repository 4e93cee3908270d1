"""Record emission (reference wire format, common/record.h + go_state_ext.h:128-195): pure host
logic, checked against hand-derived values of the reference formulas."""
import json

import numpy as np

from elf_b200 import record


def test_sgf_string_and_coords():
    n = 19
    # action = x*N + y ; coord2str = 'a'+x, 'a'+y ; pass -> "" ; colours alternate by index
    assert record.moves_to_sgf([0, 19 * 3 + 15, n * n, 360], n) == "(;B[aa];W[dp];B[];W[ss])"
    assert record.action_to_coord(0, n) == 22            # (0,0) -> (0+1)*21 + (0+1)
    assert record.action_to_coord(n * n, n) == 0         # M_PASS
    assert record.action_to_coord(3 * n + 15, n) == 16 * 21 + 4
    assert record.action_to_coord(0, 9) == 12            # coord_test.cc: str2coord("aa") == 12


def test_policy_quantisation_follows_addMCTSPolicy():
    n = 9
    v = np.full(n * n + 1, -1, np.int32)
    v[[0, 5, 40, n * n]] = [10, 3, 0, 7]
    q = record.quantise_policy(v, n)
    assert len(q) == 121 and sum(1 for c in q if c) == 3
    p = np.array([10, 3, 0, 7], np.float32) / np.float32(20)
    exp = (p / p.max() * np.float32(255)).astype(np.float32).astype(int)
    assert q[record.action_to_coord(0, n)] == 255 == exp[0]
    assert q[record.action_to_coord(5, n)] == exp[1] == 76
    assert q[record.action_to_coord(40, n)] == 0
    assert q[0] == exp[3] == 178                          # pass at coordinate 0


def test_game_recorder_emits_reference_layout():
    r = record.GameRecorder(9, thread_id=3, policy_distri_cutoff=2)
    v = np.full(82, -1, np.int32)
    v[[1, 2]] = [4, 1]
    r.on_move(1, 1, v, 0.1)
    r.on_move(2, 2, v, -0.2)
    r.on_move(3, 81, v, 0.3)      # beyond the cutoff: value recorded, policy not
    r.on_move(4, -1, v, -0.9)     # resignation: no move appended
    rec = r.finish(-1.0, never_resign=False, model_ver=5)
    assert set(rec) == {"request", "result", "timestamp", "thread_id", "seq", "pri", "offline"}
    res = rec["result"]
    assert res["content"] == "(;B[ab];W[ac];B[])" and res["num_move"] == 3
    assert len(res["policies"]) == 2 and len(res["values"]) == 4 and res["reward"] == -1.0
    assert rec["thread_id"] == 3 and rec["seq"] == 0 and r.seq == 1
    json.loads(record.dumps([rec]))
