"""Known answers of the reference's own gtests, re-run against every implementation.

Ported from /root/reference/src_cpp/elfgames/go/base/test/{go_test.cc,board_feature_test.cc,
coord_test.cc} (9x9, MiniGo-derived).  Positions are set up exactly as the reference's
``loadBoard`` does (test_utils.h:49-69): stone by stone through ``forward`` with a pass whenever
the colour to place is not the side to move.

Backends: ``oracle`` (C restatement), ``ref`` (compiled reference, when oracle/_ref exists) -- CPU;
``gpu`` (CUDA path through the C ABI) -- marked gpu.
"""
import numpy as np
import pytest

from tests import oracles

N = 9
PASS = N * N


def A(x, y):
    return x * N + y


def S(s):  # str2coord, sgf.h:22-46: first letter x, second y
    return A(ord(s[0]) - 97, ord(s[1]) - 97)


class OracleGame:
    def __init__(self):
        self.o = oracles.Oracle(N)

    def forward(self, a):
        return self.o.forward(a)

    def ply(self):
        return int(self.o.info()[0])

    def info(self):
        return self.o.info()

    def stones(self):
        return self.o.stones()

    def evaluate(self, komi):
        return self.o.evaluate(komi)

    def features(self):
        return self.o.features(0)

    def legal(self):
        return self.o.legal()

    def group(self, a):
        return oracles.oracle_group(self.o, a)

    def num_groups(self):
        return oracles.oracle_num_groups(self.o)


class RefGame(OracleGame):
    def __init__(self):
        self.o = oracles.Ref(N)

    def group(self, a):
        return oracles.ref_group(self.o, a)

    def num_groups(self):
        return oracles.ref_num_groups(self.o)


class GpuGame:
    def __init__(self):
        import elf_b200

        self.gb = elf_b200.GoBatch(1, board_size=N)

    def forward(self, a):
        return bool(self.gb.forward(np.array([a], np.int32))[0])

    def ply(self):
        return int(self.gb.info()[0, 0])

    def info(self):
        return self.gb.info()[0]

    def stones(self):
        return self.gb.stones()[0]

    def evaluate(self, komi):
        return float(self.gb.evaluate(komi)[0])

    def features(self):
        return self.gb.features()[0]

    def legal(self):
        return self.gb.legal_mask()[0, :-1]

    def group(self, a):
        return None

    def num_groups(self):
        return None


BACKENDS = [
    pytest.param(OracleGame, id="oracle"),
    pytest.param(RefGame, id="ref", marks=pytest.mark.skipif(not oracles.have_ref(N), reason="oracle/_ref not built")),
    pytest.param(GpuGame, id="gpu", marks=pytest.mark.gpu),
]


def turn(g):  # getTurn, test_utils.h:28-35
    return 2 if g.ply() % 2 == 0 else 1


def give_turn(g, s):  # giveTurn, test_utils.h:37-47
    if turn(g) != s:
        g.forward(PASS)


def load_board(g, rows):  # loadBoard, test_utils.h:49-69
    s = "".join(rows)
    assert len(s) == N * N
    for i, ch in enumerate(s):
        if ch == ".":
            continue
        x, y = i % N, i // N
        want = 1 if ch == "X" else 2
        if turn(g) != want:
            g.forward(PASS)
        assert g.forward(A(x, y)), f"loadBoard: stone {ch} at ({x},{y}) refused"


def board_rows(g):
    st = g.stones()
    rows = []
    for y in range(N):
        rows.append("".join(".XO"[st[A(x, y)]] for x in range(N)))
    return rows


EMPTY_ROW = "." * N


@pytest.mark.parametrize("Game", BACKENDS)
def test_merge_multiple_groups(Game):  # go_test.cc:177-212
    g = Game()
    load_board(g, [".X.......", "X.X......", ".X......."] + [EMPTY_ROW] * 6)
    give_turn(g, 1)
    assert g.forward(S("bb"))
    if g.group(A(1, 1)) is not None:
        assert g.num_groups() == 2
        assert g.group(A(1, 1)) == (6, 5)  # liberties 6 (go_test.cc:211), 5 stones


@pytest.mark.parametrize("Game", BACKENDS)
def test_capture_multiple_groups(Game):  # go_test.cc:214-254
    g = Game()
    load_board(g, [".OX......", "OXX......", "XX......."] + [EMPTY_ROW] * 6)
    give_turn(g, 1)
    assert g.forward(A(0, 0))
    assert g.info()[2] == 2  # _b_cap
    assert board_rows(g)[:3] == ["X.X......", ".XX......", "XX......."]
    if g.group(A(0, 0)) is not None:
        assert g.num_groups() == 3
        assert g.group(A(0, 0)) == (2, 1)
        assert g.group(A(2, 0)) == (7, 5)


@pytest.mark.parametrize("Game", BACKENDS)
def test_capture_stone_and_many(Game):  # go_test.cc:256-343
    g = Game()
    load_board(g, [".X.......", "XO.......", ".X......."] + [EMPTY_ROW] * 6)
    give_turn(g, 1)
    assert g.forward(A(2, 1))
    assert g.info()[2] == 1 and g.stones()[A(1, 1)] == 0
    g = Game()
    load_board(g, [".XX......", "XOO......", ".XX......"] + [EMPTY_ROW] * 6)
    give_turn(g, 1)
    assert g.forward(A(3, 1))
    assert g.info()[2] == 2
    assert g.stones()[A(1, 1)] == 0 and g.stones()[A(2, 1)] == 0
    if g.group(A(0, 1)) is not None:
        assert g.num_groups() == 5
        assert g.group(A(0, 1)) == (3, 1)
        assert g.group(A(3, 1)) == (4, 1)
        assert g.group(A(1, 0)) == (4, 2)
        assert g.group(A(1, 2)) == (6, 2)


@pytest.mark.parametrize("Game", BACKENDS)
def test_same_group_neighbouring_twice(Game):  # go_test.cc:345-404
    g = Game()
    load_board(g, ["XX.......", "X........"] + [EMPTY_ROW] * 7)
    give_turn(g, 1)
    assert g.forward(A(1, 1))
    if g.group(A(0, 0)) is not None:
        assert g.num_groups() == 2 and g.group(A(0, 0)) == (4, 4)
    g = Game()
    load_board(g, ["XX.......", "X........"] + [EMPTY_ROW] * 7)
    give_turn(g, 2)
    assert g.forward(A(1, 1))
    if g.group(A(0, 0)) is not None:
        assert g.num_groups() == 3 and g.group(A(0, 0)) == (2, 3) and g.group(A(1, 1)) == (2, 1)


@pytest.mark.parametrize("Game", BACKENDS)
def test_position_pass_and_moves(Game):  # go_test.cc:406-437
    g = Game()
    load_board(g, [".X.....OO", "X........"] + [EMPTY_ROW] * 7)
    before = board_rows(g)
    g2 = Game()
    load_board(g2, [".X.....OO", "X........"] + [EMPTY_ROW] * 7)
    g2.forward(PASS)
    assert board_rows(g2) == before
    give_turn(g, 1)
    g.forward(S("ca"))
    g.forward(S("ib"))
    assert board_rows(g) == [".XX....OO", "X.......O"] + [EMPTY_ROW] * 7


@pytest.mark.parametrize("Game", BACKENDS)
def test_is_move_suicidal(Game):  # go_test.cc:439-468
    g = Game()
    load_board(g, ["...O.O...", "....O....", "XO.....O.", "OXO...OXO", "O.XO.OX.O", "OXO...OOX",
                   "XO.......", "......XXO", ".....XOO."])
    for s in ["ea", "he"]:
        give_turn(g, 1)
        assert not g.forward(S(s))
    for s in ["be", "ii", "aa"]:
        give_turn(g, 1)
        assert g.forward(S(s))


@pytest.mark.parametrize("Game", BACKENDS)
def test_legal_moves(Game):  # go_test.cc:470-533
    rows = [".O.O.XOX.", "O..OOOOOX", "......O.O", "OO.....OX", "XO.....X.", ".O.......", "OX.....OO",
            "XX...OOOX", ".....O.X."]
    for flip in (False, True):
        if flip:
            rows = [r.translate(str.maketrans("XO", "OX")) for r in rows]
        me = 2 if flip else 1
        g = Game()
        load_board(g, rows)
        give_turn(g, me)
        legal = g.legal()
        for s in ["aa", "ea", "ia"]:
            assert legal[S(s)] == 0
        for s in ["af", "gi", "ii", "hc"]:
            assert legal[S(s)] == 1
        for s in ["aa", "ea", "ia"]:
            give_turn(g, me)
            assert not g.forward(S(s))
        # every move the legal mask reports is accepted (fresh copies, as the gtest does)
        for a in np.flatnonzero(legal)[:12]:
            h = Game()
            load_board(h, rows)
            give_turn(h, me)
            assert h.forward(int(a))


@pytest.mark.parametrize("Game", BACKENDS)
def test_move_with_captures(Game):  # go_test.cc:535-563
    g = Game()
    load_board(g, [EMPTY_ROW] * 5 + ["XXXX.....", "XOOX.....", "O.OX.....", "OOXX....."])
    give_turn(g, 1)
    assert g.forward(S("bh"))
    assert board_rows(g) == [EMPTY_ROW] * 5 + ["XXXX.....", "X..X.....", ".X.X.....", "..XX....."]


@pytest.mark.parametrize("Game", BACKENDS)
def test_ko_move(Game):  # go_test.cc:565-597
    g = Game()
    load_board(g, [".OX......", "OX......."] + [EMPTY_ROW] * 7)
    give_turn(g, 1)
    assert g.forward(S("aa"))
    assert board_rows(g) == ["X.X......", "OX......."] + [EMPTY_ROW] * 7
    assert not g.forward(S("ba"))  # ko
    assert g.info()[6] == S("ba")
    assert g.forward(S("ii"))
    assert g.forward(S("ih"))
    assert g.forward(S("ba"))  # retake allowed after two other moves


@pytest.mark.parametrize("Game", BACKENDS)
def test_game_over_two_passes(Game):  # go_test.cc:599-607
    g = Game()
    assert g.info()[9] == 0
    g.forward(PASS)
    g.forward(PASS)
    assert g.info()[9] == 1 and g.info()[10] == 1
    assert not g.forward(A(3, 3))


@pytest.mark.parametrize("Game", BACKENDS)
def test_scoring(Game):  # go_test.cc:609-631, komi 6.5 -> 1.5 / 2.5
    rows = [".XX......", "OOXX.....", "OOOX...X.", "OXX......", "OOXXXXXX.", "OOOXOXOXX", ".O.OOXOOX",
            ".O.O.OOXX", "......OOO"]
    g = Game()
    load_board(g, rows)
    assert g.evaluate(6.5) == 1.5
    rows[0] = "XXX......"
    g = Game()
    load_board(g, rows)
    assert g.evaluate(6.5) == 2.5


@pytest.mark.parametrize("Game", BACKENDS)
def test_replay_position(Game):  # go_test.cc:633-665
    s = ("B[fd];W[cf];B[eg];W[dd];B[dc];W[cc];B[de];W[cd];B[ed];W[he];B[ce];W[be];B[df];W[bf];B[hd];W[ge];"
         "B[gd];W[gg];B[db];W[cb];B[cg];W[bg];B[gh];W[fh];B[hh];W[fg];B[eh];W[ei];B[di];W[fi];B[hg];W[dh];"
         "B[ch];W[ci];B[bh];W[ff];B[fe];W[hf];B[id];W[bi];B[ah];W[ef];B[dg];W[ee];B[di];W[ig];B[ai];W[ih];"
         "B[fb];W[hi];B[ag];W[ab];B[bd];W[bc];B[ae];W[ad];B[af];W[bd];B[ca];W[ba];B[da];W[ie]")
    g = Game()
    for tok in s.split(";"):
        g.forward(S(tok[2:4]))
    assert board_rows(g) == [".OXX.....", "O.OX.X...", ".OOX.....", "OOOOXXXXX", "XOXXOXOOO", "XOOXOO.O.",
                             "XOXXXOOXO", "XXX.XOXXO", "X..XOO.O."]


@pytest.mark.parametrize("Game", BACKENDS)
def test_agz_feature(Game):  # board_feature_test.cc:24-101
    g = Game()
    for a in [A(0, 0), A(0, 1), A(0, 2), A(0, 3), A(1, 1)]:
        assert g.forward(a)
    f = g.features().reshape(18, N * N)

    def plane(idx):
        v = np.zeros(N * N, np.float32)
        v[list(idx)] = 1.0
        return v

    np.testing.assert_array_equal(f[0], plane([3]))
    np.testing.assert_array_equal(f[1], plane([0, 2, 10]))
    np.testing.assert_array_equal(f[2], plane([1, 3]))
    np.testing.assert_array_equal(f[3], plane([0, 2]))
    np.testing.assert_array_equal(f[4], plane([1]))
    np.testing.assert_array_equal(f[5], plane([0, 2]))
    for i in range(10, 16):
        np.testing.assert_array_equal(f[i], plane([]))


def test_coord_encoding():  # coord_test.cc:24-36: str2coord("aa") == 12 on the 11-wide expanded board
    assert (0 + 1) * (N + 2) + (0 + 1) == 12
    assert S("aa") == 0 and S("ba") == N  # action = x*N + y
