"""CPU model of the playout kernel's ALGORITHM (not of the CUDA code): row bitboards, link-masked
flood fill, incremental safe/atari group masks, captures from the atari mask, dead-end legality --
exactly the scheme of board.cuh's legal_rows_cached / play_move_cached, written with Python ints.
It is run against the oracle on whole playouts (hash, captures and full legal mask every ply through
the playout checksum), so the scheme itself stays verifiable on a box without a GPU."""
import numpy as np
import pytest

from tests import oracles

GOLD = 0x9E3779B97F4A7C15
M64 = (1 << 64) - 1


def splitmix(x):
    x = (x + GOLD) & M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M64
    return x ^ (x >> 31)


def rotl(v, r):
    return ((v << r) | (v >> (64 - r))) & M64


class Model:
    def __init__(self, n, zob):
        self.n, self.zob = n, zob
        self.rm = (1 << n) - 1
        self.b = [0] * n
        self.w = [0] * n
        self.safe = [0] * n
        self.atari = [0] * n
        self.hash = 0
        self.ply, self.next = 1, 1
        self.last1 = self.last2 = -1  # -2 pass
        self.ko_pt, self.ko_color, self.ko_active = -1, 0, False
        self.cap = [0, 0, 0]
        self.superko = False
        self.record = []

    # ---- row-vector helpers (one list entry = one lane's register) ----
    def nbr4(self, v):
        n, rm = self.n, self.rm
        return [((v[y] << 1) | (v[y] >> 1) | (v[y - 1] if y else 0) | (v[y + 1] if y < n - 1 else 0)) & rm for y in range(n)]

    def links(self, own, opp):
        n = self.n
        up = lambda v, y: v[y - 1] if y else 0
        dn = lambda v, y: v[y + 1] if y < n - 1 else 0
        return ([(own[y] & (own[y] << 1)) | (opp[y] & (opp[y] << 1)) for y in range(n)],
                [(own[y] & (own[y] >> 1)) | (opp[y] & (opp[y] >> 1)) for y in range(n)],
                [(own[y] & up(own, y)) | (opp[y] & up(opp, y)) for y in range(n)],
                [(own[y] & dn(own, y)) | (opp[y] & dn(opp, y)) for y in range(n)])

    def grow_link(self, g, k):
        n = self.n
        l, r, u, d = k
        return [g[y] | ((g[y] << 1) & l[y]) | ((g[y] >> 1) & r[y]) | ((g[y - 1] if y else 0) & u[y]) |
                ((g[y + 1] if y < n - 1 else 0) & d[y]) for y in range(n)]

    def fill_link(self, g, k):
        while True:
            g2 = self.grow_link(g, k)
            if g2 == g:
                return g
            g = g2

    def flood(self, g, through):
        while True:
            nb = self.nbr4(g)
            g2 = [g[y] | (nb[y] & through[y]) for y in range(self.n)]
            if g2 == g:
                return g
            g = g2

    # ---- the scheme under test ----
    def legal_rows(self):
        n = self.n
        own, opp = (self.b, self.w) if self.next == 1 else (self.w, self.b)
        e = [~(own[y] | opp[y]) & self.rm for y in range(n)]
        en = self.nbr4(e)
        legal = [e[y] & en[y] for y in range(n)]
        hard = [e[y] & ~en[y] for y in range(n)]
        if any(hard):
            a = self.nbr4([self.safe[y] & own[y] for y in range(n)])
            c = self.nbr4([self.atari[y] & opp[y] for y in range(n)])
            legal = [legal[y] | (hard[y] & (a[y] | c[y])) for y in range(n)]
        if self.ko_active and self.ko_color == self.next:
            ky, kx = divmod(self.ko_pt, n)
            legal[ky] &= ~(1 << kx)
        return legal

    def true_eyes(self):
        n, rm = self.n, self.rm
        own, opp = (self.b, self.w) if self.next == 1 else (self.w, self.b)
        e = [~(own[y] | opp[y]) & rm for y in range(n)]
        notown = [~own[y] & rm for y in range(n)]
        nn = self.nbr4(notown)
        out = []
        for y in range(n):
            ou = opp[y - 1] if y else 0
            od = opp[y + 1] if y < n - 1 else 0
            d = [(ou << 1) & rm, ou >> 1, (od << 1) & rm, od >> 1]
            ge1 = d[0] | d[1] | d[2] | d[3]
            ge2 = (d[0] & (d[1] | d[2] | d[3])) | (d[1] & (d[2] | d[3])) | (d[2] & d[3])
            edge = rm if y in (0, n - 1) else (1 | (1 << (n - 1)))
            fake = (edge & ge1) | (~edge & ge2)
            out.append(e[y] & ~nn[y] & ~fake & rm)
        return out

    def zrow(self, y, bits, color):
        h = 0
        E = self.n + 2
        while bits:
            x = (bits & -bits).bit_length() - 1
            bits &= bits - 1
            h ^= self.zob[(y + 1) * E + x + 1]
        return h if color == 1 else (((h >> 32) | (h << 32)) & M64)

    def play(self, p):
        n = self.n
        player, oppc = self.next, 3 - self.next
        if p >= 0:
            own, opp = (self.b, self.w) if player == 1 else (self.w, self.b)
            y, x = divmod(p, n)
            mybit = [0] * n
            mybit[y] = 1 << x
            nb = self.nbr4(mybit)
            single = not any(nb[r] & own[r] for r in range(n))
            own[y] |= 1 << x
            dead = [0] * n
            dseed = [nb[r] & opp[r] & self.atari[r] for r in range(n)]
            ncap = 0
            if any(dseed):
                dead = self.flood(dseed, opp)
                ncap = sum(bin(v).count("1") for v in dead)
                for r in range(n):
                    opp[r] &= ~dead[r]
                    self.safe[r] &= ~dead[r]
                    self.atari[r] &= ~dead[r]
                    self.hash ^= self.zrow(r, dead[r], oppc)
            self.hash ^= self.zrow(y, 1 << x, player)
            self.cap[player] += ncap
            stones = [own[r] | opp[r] for r in range(n)]
            e2 = [~stones[r] & self.rm for r in range(n)]
            libs = sum(bin(nb[r] & e2[r]).count("1") for r in range(n))
            if ncap == 1 and single and libs == 1:
                r = next(r for r in range(n) if dead[r])
                self.ko_pt, self.ko_color, self.ko_active = r * n + dead[r].bit_length() - 1, oppc, True
            else:
                self.ko_active = False
            dnb = self.nbr4(dead)
            seeds = [(mybit[r] | nb[r] | dnb[r]) & stones[r] for r in range(n)]
            if single:
                tgt = self.atari if libs == 1 else self.safe
                tgt[y] |= 1 << x
                seeds[y] &= ~(1 << x)
            k = self.links(own, opp)
            while any(seeds):
                r = next(r for r in range(n) if seeds[r])
                g = [0] * n
                g[r] = seeds[r] & -seeds[r]
                g = self.fill_link(g, k)
                gn = self.nbr4(g)
                nl = sum(bin(gn[q] & e2[q]).count("1") for q in range(n))
                for q in range(n):
                    if nl == 1:
                        self.atari[q] |= g[q]
                        self.safe[q] &= ~g[q]
                    else:
                        self.safe[q] |= g[q]
                        self.atari[q] &= ~g[q]
                    seeds[q] &= ~g[q]
        self.next = oppc
        self.last2, self.last1 = self.last1, (p if p >= 0 else -2)
        self.ply += 1

    def terminated(self):
        return (self.last1 == -2 and self.last2 == -2) or self.ply >= 2 * self.n * self.n or self.superko

    def playout(self, seed, gid):
        n = self.n
        chk = 0
        t = 0
        while not self.terminated():
            legal = self.legal_rows()
            eyes = self.true_eyes()
            cand = [legal[y] & ~eyes[y] for y in range(n)]
            rx = 0
            for y in range(n):
                v = ((((y + 1) << 32) | legal[y]) * 0xBF58476D1CE4E5B9) & M64
                rx ^= v ^ (v >> 29)
            caps = (self.cap[1] & 0xFFFF) | ((self.cap[2] & 0xFFFF) << 16) | (self.next << 32)
            chk = splitmix(chk ^ self.hash ^ rotl(rx, 23) ^ ((caps * GOLD) & M64))
            cnt = sum(bin(v).count("1") for v in cand)
            p = -2
            if cnt:
                r = splitmix(seed ^ ((gid * GOLD) & M64) ^ self.ply)
                k = ((r >> 32) * cnt) >> 32
                for x in range(n):          # ascending action order a = x*N + y
                    for y in range(n):
                        if (cand[y] >> x) & 1:
                            if k == 0:
                                p = y * n + x
                            k -= 1
                        if p >= 0:
                            break
                    if p >= 0:
                        break
            pre = self.hash
            self.play(p)
            if p >= 0:
                if self.hash in self.record:
                    self.superko = True
                self.record.append(pre)
            t += 1
        chk = splitmix(splitmix(chk ^ self.hash) ^ self.ply)
        return t, chk


def _zobrist():
    import os
    import re

    inc = os.path.join(oracles.ROOT, "include", "elfb200_zobrist.inc")
    return [int(v, 16) for v in re.findall(r"0x[0-9a-f]+", open(inc).read())]


@pytest.mark.parametrize("n,games", [(9, 12), (19, 2)])
def test_incremental_group_status_scheme_matches_oracle(n, games, oracle_lib):
    zob = _zobrist()
    for gid in range(games):
        t, chk = Model(n, zob).playout(7, gid)
        et, echk, _ = oracles.oracle_playout(n, 7, gid, lib=oracle_lib)
        assert (t, chk) == (et, echk), f"game {gid}"
