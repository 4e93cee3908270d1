"""The feature kernels decompose the inverse D4 transform (BoardFeature::InvTransform,
board_feature.h:115-130 -- d4_inverse in elf_b200/csrc/common.cuh) into three bits per code:
an output row is a board row or a board COLUMN (transposed), taken at index tx or N-1-tx (rev_idx),
read forwards or backwards (rev_bits).  The three bit masks in features_cta must agree with the
transform itself for every code, cell and board size."""
import re
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def d4_inverse(n, d4, tx, ty):  # common.cuh: d4_inverse
    a, b = (ty, tx) if d4 & 4 else (tx, ty)
    r = d4 & 3
    if r == 1:
        return n - b - 1, a
    if r == 2:
        return n - a - 1, n - b - 1
    if r == 3:
        return b, n - a - 1
    return a, b


def test_d4_decomposition_masks_match_the_transform():
    src = open(os.path.join(ROOT, "elf_b200", "csrc", "common.cuh")).read()
    m = re.search(r"transposed = \((0x[0-9A-Fa-f]+)u >> d4\) & 1u, rev_idx = \((0x[0-9A-Fa-f]+)u >> d4\) & 1u, "
                  r"rev_bits = \((0x[0-9A-Fa-f]+)u >> d4\) & 1u", src)
    assert m, "the masks moved: update this test"
    T, RI, RB = (int(x, 16) for x in m.groups())
    rng = np.random.default_rng(0)
    for n in (9, 19):
        B = rng.integers(0, 2, (n, n))  # B[y][x]
        for d4 in range(8):
            tr, ri, rb = (T >> d4) & 1, (RI >> d4) & 1, (RB >> d4) & 1
            for tx in range(n):
                src_i = n - 1 - tx if ri else tx
                word = B[:, src_i] if tr else B[src_i, :]
                if rb:
                    word = word[::-1]
                for ty in range(n):
                    x, y = d4_inverse(n, d4, tx, ty)
                    assert word[ty] == B[y, x], (n, d4, tx, ty)
