"""Self-play game records in the reference's wire format.

Mirrors ``GoStateExt::addMCTSPolicy`` / ``dumpRecord`` (``src_cpp/elfgames/go/common/
go_state_ext.h:128-195``) and the JSON layout of ``Record`` / ``MsgResult`` / ``MsgRequest``
(``common/record.h:20-260``): the move list as the SGF-like string of ``coords2sgfstr``
(``sgf/sgf.h:46-95``), one u8-quantised MCTS policy per recorded move indexed by the reference's
expanded coordinate (``(y+1)*(N+2) + (x+1)``, pass = 0), the predicted values, the final reward.
Host-side bookkeeping only: nothing here touches the GPU.
"""
import json
import time

import numpy as np


def action_to_coord(a, n):
    """action x*N+y (pass N*N) -> reference Coord (expanded (N+2)^2 board, M_PASS = 0)"""
    if a == n * n:
        return 0
    x, y = a // n, a % n
    return (y + 1) * (n + 2) + (x + 1)


def coord2str(a, n):  # sgf.h:46-55: 'a'+x, 'a'+y; pass -> ""
    if a == n * n:
        return ""
    return chr(97 + a // n) + chr(97 + a % n)


def moves_to_sgf(actions, n):  # coords2sgfstr, sgf.h:87-95 (colour alternates with the index)
    return "(" + "".join(";" + ("B" if i % 2 == 0 else "W") + "[" + coord2str(a, n) + "]"
                         for i, a in enumerate(actions)) + ")"


def quantise_policy(visits_row, n):
    """addMCTSPolicy: normalised visit policy (MCTSPolicy::normalize, t = 1) scaled so that the most
    visited move maps to 255, truncated to u8, stored at the move's expanded coordinate."""
    out = [0] * ((n + 2) * (n + 2))
    idx = np.flatnonzero(visits_row >= 0)
    v = visits_row[idx].astype(np.float32)
    s = np.float32(np.cumsum(v, dtype=np.float32)[-1]) if len(v) else np.float32(0)  # sequential float sum, edge order
    if s <= 0:
        return out
    p = v / s
    mx = np.float32(p.max())
    q = (p / mx * np.float32(255)).astype(np.float32)
    for a, c in zip(idx, q):
        out[action_to_coord(int(a), n)] = int(c)
    return out


def ts_options_json(opts=None):
    """``TSOptions::setJsonFields`` (``src_cpp/elf/ai/tree_search/tree_search_options.h:77-215``): the
    ``mcts_opt`` object inside ``request.vers``.  The reference's parser REQUIRES every field
    (``JSON_LOAD`` throws on a missing one and ``Record::createBatchFromJson`` then drops the record),
    so all of them are written.  ``opts``: keyword options of ``elfb200_mcts_options`` / MctsBatch."""
    o = dict(opts or {})
    return {
        "max_num_moves": 0, "num_threads": 1, "num_rollouts_per_thread": int(o.get("num_rollouts", 100)),
        "num_rollouts_per_batch": int(o.get("num_rollouts_per_batch", 8)), "verbose": False, "verbose_time": False,
        "seed": int(o.get("seed", 0)) & 0x7FFFFFFF, "persistent_tree": bool(o.get("persistent_tree", 0)),
        "pick_method": "most_visited", "log_prefix": "", "root_epsilon": float(o.get("root_epsilon", 0.0)),
        "root_alpha": float(o.get("root_alpha", 0.0)), "virtual_loss": int(o.get("virtual_loss", 0)),
        "alg_opt": {"use_prior": bool(o.get("use_prior", 1)), "c_puct": float(o.get("c_puct", 5.0)),
                    "unexplored_q_zero": bool(o.get("unexplored_q_zero", 0)),
                    "root_unexplored_q_zero": bool(o.get("root_unexplored_q_zero", 0))},
    }


def request_json(black_ver, white_ver=-1, mcts_opt=None, black_resign_thres=0.0, white_resign_thres=0.0,
                 never_resign_prob=0.0, player_swap=False, async_=False, num_game_thread_used=-1, client_type=1):
    """``MsgRequest::setJsonFields`` (``common/record.h:113-135``): what the training server sends to a
    client (``ModelPair`` vers + ``ClientCtrl``), as a dict"""
    return {
        "vers": {"black_ver": int(black_ver), "white_ver": int(white_ver), "mcts_opt": ts_options_json(mcts_opt)
                 if not (isinstance(mcts_opt, dict) and "alg_opt" in mcts_opt) else mcts_opt},
        "client_ctrl": {"client_type": int(client_type), "num_game_thread_used": int(num_game_thread_used),
                        "black_resign_thres": float(black_resign_thres), "white_resign_thres": float(white_resign_thres),
                        "never_resign_prob": float(never_resign_prob), "player_swap": bool(player_swap),
                        "async": bool(async_)},
    }


def parse_request(msg):
    """a MsgRequest (dict or JSON text, ``MsgRequest::createFromJson``) -> keyword arguments of
    ``SelfPlay.set_request`` plus ``mcts_opt`` (the TSOptions object, kept for the records).
    ``player_swap`` is optional for self-play requests and ``async`` is optional altogether, as in
    ``ClientCtrl::createFromJson`` (record.h:52-69)."""
    if isinstance(msg, (str, bytes)):
        msg = json.loads(msg)
    v, c = msg["vers"], msg["client_ctrl"]
    for k in ("black_ver", "white_ver", "mcts_opt"):
        if k not in v:
            raise KeyError(f"{k}cannot not be found!")  # the reference's message, typo included
    selfplay = v["black_ver"] >= 0 and v["white_ver"] == -1
    if not selfplay and "player_swap" not in c:
        raise KeyError("player_swapcannot not be found!")
    return {
        "black_ver": int(v["black_ver"]), "white_ver": int(v["white_ver"]),
        "black_resign_thres": float(c["black_resign_thres"]), "white_resign_thres": float(c["white_resign_thres"]),
        "never_resign_prob": float(c["never_resign_prob"]), "player_swap": bool(c.get("player_swap", False)),
        "async_": bool(c.get("async", False)), "num_game_thread_used": int(c["num_game_thread_used"]),
    }, v["mcts_opt"]


class GameRecorder:
    """per-slot accumulation of one game's record fields"""

    def __init__(self, n, thread_id, policy_distri_cutoff, policy_distri_training_for_all=False, mcts_opt=None):
        self.mcts_opt = ts_options_json(mcts_opt)
        self.n = n
        self.thread_id = thread_id
        self.cutoff = policy_distri_cutoff
        self.for_all = policy_distri_training_for_all
        self.seq = 0
        self.restart()

    def restart(self):
        self.moves = []
        self.policies = []
        self.values = []
        self.models = set()  # GoStateExt::using_models_

    def add_models(self, *versions):
        """GoStateExt::addCurrentModel (go_state_ext.h:66-71): at every (re)start and at every
        async model update"""
        self.models.update(int(v) for v in versions if v >= 0)

    def on_move(self, ply_before, action, visits_row, predicted_value):
        if visits_row is not None and (self.for_all or ply_before <= self.cutoff):  # mcts_make_diverse_move, game_selfplay.cc:88-93
            self.policies.append(quantise_policy(visits_row, self.n))
        self.values.append(float(predicted_value))  # addPredictedValue
        if action >= 0:
            self.moves.append(int(action))

    def finish(self, final_value, never_resign, model_ver=-1, resign_thres=0.0, never_resign_prob=0.0,
               white_ver=-1, player_swap=False, async_=False, num_game_thread_used=-1):
        """GoStateExt::dumpRecord (go_state_ext.h:128-147): the request the game was played under
        (MsgRequest: versions + client control), the models used (``using_models_``: every version
        >= 0 the game saw, ascending) and the result"""
        rec = {
            "request": {
                "vers": {"black_ver": int(model_ver), "white_ver": int(white_ver), "mcts_opt": self.mcts_opt},
                "client_ctrl": {"client_type": 1, "num_game_thread_used": int(num_game_thread_used),
                                "black_resign_thres": resign_thres, "white_resign_thres": resign_thres,
                                "never_resign_prob": never_resign_prob, "player_swap": bool(player_swap),
                                "async": bool(async_)},
            },
            "result": {
                "num_move": len(self.moves), "reward": float(final_value),
                "black_never_resign": bool(never_resign), "white_never_resign": bool(never_resign),
                "using_models": sorted(self.models | {int(v) for v in (model_ver, white_ver) if v >= 0}),
                "content": moves_to_sgf(self.moves, self.n),
                "policies": self.policies, "values": self.values,
            },
            "timestamp": int(time.time()), "thread_id": self.thread_id, "seq": self.seq, "pri": 0.0, "offline": False,
        }
        self.seq += 1
        self.restart()
        return rec


def dumps(records):
    return json.dumps(records)
