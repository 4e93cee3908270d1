"""SGF reading for the online / preload path.

Reference: ``src_cpp/elfgames/go/sgf/sgf.h:22-46,73-125,127-290`` and ``sgf.cc:27-238`` (class Sgf:
header of the root node + the main line of moves), used by ``GoGameSelfPlay::restart`` to preload a
game (``common/game_selfplay.cc:202-219``) and by ``act`` to follow it (``:387-400``).

Same observable behaviour as the reference reader on linear game records (the only kind its
callers feed it):

* the ROOT node is the header (``SZ KM HA RE PW PB WR BR C``); moves inside the root node are not
  part of the move list (sgf.cc:163-179);
* a move value is two letters ``xy`` with ``x = first - 'a'``, ``y = second - 'a'`` and the action
  is ``x*N + y`` (str2coord sgf.h:22-46, no 'i' skipping, no flip); a value shorter than two
  characters is a pass; letters off the board (e.g. ``tt`` on 19x19) are INVALID (-1), not a pass;
* ``num_moves`` stops counting at the second consecutive pass (sgf.cc:39-49);
* ``RE[B+3.5]`` -> winner black, margin 3.5; ``RE[W+R]`` -> winner white, reason ``R``.

Deviation, on purpose: on a file WITH variations this reader follows the first variation at every
branch (the SGF main line); the reference's recursive loader strings every variation it meets one
after the other (sgf.cc:206-238), which yields move lists that cannot be replayed.
"""
import re
from dataclasses import dataclass, field

S_BLACK, S_WHITE, S_OFF_BOARD = 1, 2, 3
INVALID = -1

_FLOAT_PREFIX = re.compile(r"\s*[-+]?(\d+\.?\d*([eE][-+]?\d+)?|\.\d+([eE][-+]?\d+)?)")
_INT_PREFIX = re.compile(r"\s*[-+]?\d+")


def _stof(s):
    """std::stof semantics: parse the longest numeric prefix, raise if there is none"""
    m = _FLOAT_PREFIX.match(s)
    if not m:
        raise ValueError(f"stof: no conversion for {s!r}")
    return float(m.group(0))


def _stoi(s):
    m = _INT_PREFIX.match(s)
    if not m:
        raise ValueError(f"stoi: no conversion for {s!r}")
    return int(m.group(0))


def _trim(s):
    """the reference's trim (sgf.cc:15-24) keeps up to r+1 characters from the first non-blank one,
    r = index of the last non-blank: trailing blanks survive when there were leading ones"""
    l = 0
    while l < len(s) and s[l] in " \n":
        l += 1
    r = len(s) - 1
    while r >= 0 and s[r] in " \n":
        r -= 1
    return s[l : l + r + 1]


def str2action(v, n):
    """str2coord (sgf.h:22-46) in action space: pass = n*n, off board = INVALID"""
    if len(v) < 2:
        return n * n
    t = [c for c in v if c not in " \n"]
    if len(t) < 2:
        return INVALID
    x, y = ord(t[0]) - 97, ord(t[1]) - 97
    if not (0 <= x < n and 0 <= y < n):
        return INVALID
    return x * n + y


def action2str(a, n):
    """coord2str (sgf.h:48-57)"""
    if a == n * n:
        return ""
    return chr(97 + a // n) + chr(97 + a % n)


@dataclass
class SgfHeader:  # sgf.h:141-170
    rule: int = 0
    size: int = 19
    komi: float = 7.5
    handi: int = 0
    white_name: str = ""
    black_name: str = ""
    white_rank: str = ""
    black_rank: str = ""
    comment: str = ""
    winner: int = S_OFF_BOARD
    win_margin: float = 0.0
    win_reason: str = ""


@dataclass
class SgfMove:
    player: int
    action: int
    comment: str = ""
    props: dict = field(default_factory=dict)


def _tokenise(text):
    """yield ('(',), (')',), (';',) and ('prop', key, value); backslash escapes the next char
    inside and outside values (get_key_values, sgf.cc:68-117)"""
    i, n = 0, len(text)
    key = []
    while i < n:
        c = text[i]
        if c == "\\":
            i += 2
            continue
        if c == "[":
            j = i + 1
            val = []
            while j < n and text[j] != "]":
                if text[j] == "\\":
                    j += 2
                    continue
                val.append(text[j])
                j += 1
            if j >= n:
                return  # unterminated value: the reference drops it as well
            yield ("prop", _trim("".join(key)), _trim("".join(val)))
            key = []
            i = j + 1
            continue
        if c in "();":
            key = []
            yield (c,)
        else:
            key.append(c)
        i += 1


class Sgf:
    """main line of one game record"""

    def __init__(self, board_size=None):
        self.header = SgfHeader()
        self.moves = []
        self.num_moves = 0
        self._n = board_size

    # -- Sgf::load ---------------------------------------------------------------------------
    @classmethod
    def load(cls, filename, board_size=None):
        with open(filename, "r", errors="replace") as f:
            return cls.loads(f.read(), board_size)

    @classmethod
    def loads(cls, text, board_size=None):
        """parse; raises ValueError where the reference's load() returns false"""
        self = cls(board_size)
        nodes = self._main_line_nodes(text)
        if nodes is None:
            raise ValueError("SGF: no header node")
        for k, v in nodes[0]:
            self._header_prop(k, v)
        n = self._n if self._n is not None else self.header.size
        self._n = n
        for props in nodes[1:]:
            mv = SgfMove(S_OFF_BOARD, INVALID)
            for k, v in props:
                if len(k) == 1:
                    if k == "B":
                        mv.player, mv.action = S_BLACK, str2action(v, n)
                    elif k == "W":
                        mv.player, mv.action = S_WHITE, str2action(v, n)
                    elif k == "C":
                        mv.comment = v
                else:
                    mv.props.setdefault(k, v)
            self.moves.append(mv)
        if not self.moves:
            raise ValueError("SGF: no moves")
        # sgf.cc:39-49
        self.num_moves = 0
        last = INVALID
        i = 0
        pas = n * n
        while i < len(self.moves):
            self.num_moves += 1
            i += 1
            nxt = self.moves[i].action if i < len(self.moves) else INVALID
            if nxt == pas and last == pas:
                break
            last = nxt
        return self

    @staticmethod
    def _main_line_nodes(text):
        nodes = []
        cur = None
        depth = 0
        skip_depth = None  # depth at which a non-first variation started
        closed_at = {}  # depth -> a variation already closed at this depth below the current node
        for tok in _tokenise(text):
            t = tok[0]
            if t == "(":
                depth += 1
                if skip_depth is None and closed_at.get(depth):
                    skip_depth = depth
                continue
            if t == ")":
                if skip_depth is not None and depth == skip_depth:
                    skip_depth = None
                else:
                    closed_at[depth] = True
                for d in [d for d in closed_at if d > depth]:
                    del closed_at[d]
                depth -= 1
                cur = None
                continue
            if skip_depth is not None:
                continue
            if t == ";":
                cur = []
                nodes.append(cur)
            elif t == "prop" and cur is not None:
                cur.append((tok[1], tok[2]))
        return nodes if nodes else None

    def _header_prop(self, k, v):  # save_sgf_header, sgf.cc:123-161
        h = self.header
        if k == "RE":
            if v:
                h.winner = S_BLACK if v[0] in "Bb" else S_WHITE
                if len(v) >= 3:
                    try:
                        h.win_margin = _stof(v[2:])
                    except ValueError:
                        h.win_reason = v[2:]
        elif k == "SZ":
            h.size = _stoi(v)
        elif k == "PW":
            h.white_name = v
        elif k == "PB":
            h.black_name = v
        elif k == "WR":
            h.white_rank = v
        elif k == "BR":
            h.black_rank = v
        elif k == "C":
            h.comment = v
        elif k == "KM":
            h.komi = _stof(v)
        elif k == "HA":
            h.handi = _stoi(v)

    # -- iteration ----------------------------------------------------------------------------
    def actions(self):
        return [m.action for m in self.moves]

    def players(self):
        return [m.player for m in self.moves]

    def __iter__(self):
        return iter(self.moves)

    def __len__(self):
        return len(self.moves)

    def print_main_variation(self):  # Sgf::printMainVariation, sgf.cc:258-274
        out = []
        n = self._n
        for i, m in enumerate(self.moves):
            who = {S_BLACK: "B", S_WHITE: "W"}.get(m.player, "-")
            s = f"[{i}]: {who} {action2str(m.action, n) if m.action >= 0 else '??'}"
            if m.comment:
                s += " Comment: " + m.comment
            out.append(s)
        return "\n".join(out) + "\n"


def sgfstr2actions(sgf, n):
    """sgfstr2coords (sgf.h:97-125): the compact '(;B[aa];W[bb])' strings of self-play records"""
    moves = []
    if not sgf or sgf[0] != "(":
        return moves
    i = 1
    while i < len(sgf) and sgf[i] == ";":
        while i < len(sgf) and sgf[i] != "[":
            i += 1
        if i == len(sgf):
            break
        i += 1
        j = i
        while j < len(sgf) and sgf[j] != "]":
            j += 1
        if j == len(sgf):
            break
        moves.append(str2action(sgf[i:j], n))
        i = j + 1
    return moves
