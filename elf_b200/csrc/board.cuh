// board.cuh -- warp-cooperative Go board primitives for sm_100a.
//
// Representation ("row-per-lane bitboards"): one game is owned by N consecutive
// lanes of a warp (N = board size; 19x19 -> 1 game/warp on lanes 0..18, 9x9 ->
// 3 games/warp on lanes 0..26).  Lane `row` holds row y=row of the position as
// two N-bit words: `own`/`opp` or `b`/`w`, bit x = intersection (x, y).  With
// this layout
//   * left/right neighbours are register shifts, up/down neighbours one
//     __shfl_up/__shfl_down each, so a 4-neighbour dilation is ~8 instructions;
//   * group flood fill ("which stones are connected to ...") is dilation to a
//     fixpoint with a warp vote as the termination test;
//   * per-game reductions (capture count, Zobrist delta) are REDUX over the
//     game's lane segment.
// In HBM a position is N uint64 words (black row | white row << 32) so that a
// warp loads/stores it with one coalesced 8-byte access per lane.
//
// Semantics follow the reference rules engine bit-for-bit on the observable
// state (hash, legality, captures, ko, superko, score); the reference functions
// each routine reproduces are cited inline as
// /root/reference/src_cpp/elfgames/go/base/<file>:<line>.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "elfb200_playout_policy.h"

namespace elfb200 {

constexpr unsigned FULL = 0xFFFFFFFFu;

enum : int { S_EMPTY = 0, S_BLACK = 1, S_WHITE = 2 };
// encodings of "last move" in BoardMeta (points are p = y*N + x)
enum : int { MV_PASS = -2, MV_INVALID = -1 };

// flags in BoardMeta::flags
enum : uint8_t { F_KO_ACTIVE = 1, F_SUPERKO = 2 };

// 16-byte per-game metadata (SoA array, one uint4 load per game).
struct __align__(16) BoardMeta {
  uint16_t ply;      // reference Board::_ply, starts at 1 (board.cc:106)
  uint8_t next;      // side to move (board.cc:99)
  uint8_t flags;     // F_KO_ACTIVE <=> _ko_age == 0 (board.h:144), F_SUPERKO
  int16_t last1;     // Board::_last_move  (point, MV_PASS, MV_INVALID)
  int16_t last2;     // Board::_last_move2
  int16_t ko_pt;     // Board::_simple_ko as point, -1 if never set
  uint8_t ko_color;  // Board::_simple_ko_color
  uint8_t pad;
  uint16_t b_cap;    // Board::_b_cap
  uint16_t w_cap;    // Board::_w_cap
};
static_assert(sizeof(BoardMeta) == 16, "BoardMeta must be 16 bytes");

template <int N>
struct Geo {
  static constexpr int GPW = 32 / N;  // games per warp
  static constexpr int LANES = GPW * N;
  static constexpr uint32_t ROWMASK = (1u << N) - 1u;
  static constexpr int E = N + 2;      // expanded stride of the reference (board.h:45)
  static constexpr int P = N * N;
  static constexpr int MAX_PLY = 2 * N * N;  // BOARD_MAX_MOVE, go_common.h:15
  static constexpr int ZOB = E * E;
};

struct Lane {
  int lane;          // 0..31
  int sub;           // game slot inside the warp
  int row;           // board row y owned by this lane (0 for idle lanes)
  int base;          // first lane of this game's segment
  uint32_t rm;       // ROWMASK for active lanes, 0 for idle lanes
  uint32_t segmask;  // lane mask of this game's segment
  bool active;
};

template <int N>
__device__ __forceinline__ Lane make_lane() {
  Lane L;
  L.lane = threadIdx.x & 31;
  L.active = L.lane < Geo<N>::LANES;
  L.sub = L.active ? L.lane / N : 0;
  L.row = L.active ? L.lane - L.sub * N : 0;
  L.base = L.sub * N;
  L.rm = L.active ? Geo<N>::ROWMASK : 0u;
  L.segmask = L.active ? (Geo<N>::ROWMASK << L.base) : 0u;
  return L;
}

template <int N>
__device__ __forceinline__ Lane make_lane_single() {  // one game per warp, also for 9x9
  Lane L;
  L.lane = threadIdx.x & 31;
  L.active = L.lane < N;
  L.sub = 0;
  L.row = L.active ? L.lane : 0;
  L.base = 0;
  L.rm = L.active ? Geo<N>::ROWMASK : 0u;
  L.segmask = Geo<N>::ROWMASK;
  return L;
}


// ---- neighbour shifts -------------------------------------------------------
template <int N>
__device__ __forceinline__ uint32_t up_of(uint32_t v, const Lane& L) {  // value of row y-1
  uint32_t u = __shfl_up_sync(FULL, v, 1);
  return L.row == 0 ? 0u : u;
}
template <int N>
__device__ __forceinline__ uint32_t dn_of(uint32_t v, const Lane& L) {  // value of row y+1
  uint32_t d = __shfl_down_sync(FULL, v, 1);
  return (L.row == N - 1 || !L.active) ? 0u : d;
}
// union of the 4 neighbours of every set point (may carry bit N: AND with rm / a board mask)
template <int N>
__device__ __forceinline__ uint32_t nbr4(uint32_t v, const Lane& L) {
  return (v << 1) | (v >> 1) | up_of<N>(v, L) | dn_of<N>(v, L);
}

// Dilation for flood fills: result is always ANDed with a board mask and ORed with `v`, so for
// one-game-per-warp boards the segment-boundary selects can be dropped: lane 0's "up" is its
// own word (already in v), the last row's "down" comes from an idle lane that holds 0, and idle
// lanes are masked by `through` (0 there).
template <int N>
__device__ __forceinline__ uint32_t grow4(uint32_t v, const Lane& L) {
  if (Geo<N>::GPW == 1)
    return (v << 1) | (v >> 1) | __shfl_up_sync(FULL, v, 1) | __shfl_down_sync(FULL, v, 1);
  return nbr4<N>(v, L);
}

// ---- per-game reductions ----------------------------------------------------
template <int N>
__device__ __forceinline__ int game_sum(int v, const Lane& L) {
  if (Geo<N>::GPW == 1) return __reduce_add_sync(FULL, L.active ? v : 0);
  int r = 0;
  if (L.active) r = __reduce_add_sync(L.segmask, v);
  return r;
}
template <int N>
__device__ __forceinline__ uint64_t game_xor64(uint64_t v, const Lane& L) {
  uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
  if (Geo<N>::GPW == 1) {
    if (!L.active) lo = hi = 0;
    lo = __reduce_xor_sync(FULL, lo);
    hi = __reduce_xor_sync(FULL, hi);
  } else if (L.active) {
    lo = __reduce_xor_sync(L.segmask, lo);
    hi = __reduce_xor_sync(L.segmask, hi);
  }
  return ((uint64_t)hi << 32) | lo;
}
template <int N>
__device__ __forceinline__ bool game_any(bool pred, const Lane& L) {
  if (Geo<N>::GPW == 1) return __any_sync(FULL, pred && L.active);
  return (__ballot_sync(FULL, pred) & L.segmask) != 0u;
}

// ---- flood fill -------------------------------------------------------------
// Grow `g` through `through` to a fixpoint (4-connectivity).  Two dilations per
// vote.  All lanes of the warp iterate together; extra iterations are idempotent.
template <int N>
__device__ __forceinline__ uint32_t flood(uint32_t g, uint32_t through, const Lane& L) {
  while (true) {
    uint32_t n1 = g | (grow4<N>(g, L) & through);
    uint32_t n2 = n1 | (grow4<N>(n1, L) & through);
    bool ch = n2 != g;
    g = n2;
    if (!__any_sync(FULL, ch)) break;
  }
  return g;
}

// K independent fills advanced in lock step (ILP across fills, one vote/iteration).
template <int N, int K>
__device__ __forceinline__ void floodK(uint32_t (&g)[K], const uint32_t (&through)[K],
                                       const Lane& L) {
  while (true) {
    bool ch = false;
#pragma unroll
    for (int k = 0; k < K; ++k) {  // two dilations per vote
      const uint32_t n1 = g[k] | (grow4<N>(g[k], L) & through[k]);
      const uint32_t n2 = n1 | (grow4<N>(n1, L) & through[k]);
      ch |= (n2 != g[k]);
      g[k] = n2;
    }
    if (!__any_sync(FULL, ch)) break;
  }
}

// ---- Zobrist ----------------------------------------------------------------
// transform_hash (board.cc:24-36): black -> h, white -> rotate by 32 bits.
__device__ __forceinline__ uint64_t zob_color(uint64_t h, int color) {
  return color == S_WHITE ? ((h >> 32) | (h << 32)) : h;
}

// XOR of the raw table entries of all points set in row word `bits` of row `y`.
template <int N>
__device__ __forceinline__ uint64_t zob_row(const uint64_t* __restrict__ zob, int y, uint32_t bits) {
  uint64_t h = 0;
  const uint64_t* r = zob + (y + 1) * Geo<N>::E + 1;
  while (bits) {
    int x = __ffs(bits) - 1;
    bits &= bits - 1;
    h ^= r[x];
  }
  return h;
}

// ---- legality ---------------------------------------------------------------
// Same-colour link masks: bit x of l? is set iff the stone at (x, y) has a stone of the SAME colour
// as its left / right / upper / lower neighbour.  With them one dilation step grows groups of both
// colours at once without leaking across colours:
//   g' = g | (g << 1 & lL) | (g >> 1 & lR) | (up(g) & lU) | (down(g) & lD)
// and because the masks are zero at board / game-segment borders the raw shuffles need no selects.
struct Links {
  uint32_t l, r, u, d;
};

template <int N>
__device__ __forceinline__ Links make_links(uint32_t own, uint32_t opp, const Lane& L) {
  Links k;
  k.l = (own & (own << 1)) | (opp & (opp << 1));
  k.r = (own & (own >> 1)) | (opp & (opp >> 1));
  k.u = (own & up_of<N>(own, L)) | (opp & up_of<N>(opp, L));
  k.d = (own & dn_of<N>(own, L)) | (opp & dn_of<N>(opp, L));
  return k;
}

__device__ __forceinline__ uint32_t grow_link(uint32_t g, const Links& k) {
  return g | ((g << 1) & k.l) | ((g >> 1) & k.r) | (__shfl_up_sync(FULL, g, 1) & k.u) |
         (__shfl_down_sync(FULL, g, 1) & k.d);
}

// Own true eyes: isEye && !isFakeEye (board.cc:1850-1910).
template <int N>
__device__ __forceinline__ uint32_t true_eye_rows(uint32_t own, uint32_t opp, const Lane& L) {
  const uint32_t e = ~(own | opp) & L.rm;
  const uint32_t notown = ~own & L.rm;                       // on-board, not ours
  const uint32_t eyeish = e & ~nbr4<N>(notown, L);           // every on-board neighbour is ours
  const uint32_t o_u = up_of<N>(opp, L), o_d = dn_of<N>(opp, L);
  const uint32_t d1 = o_u << 1, d2 = o_u >> 1, d3 = o_d << 1, d4 = o_d >> 1;  // 4 diagonals
  const uint32_t ge1 = d1 | d2 | d3 | d4;
  const uint32_t ge2 = (d1 & (d2 | d3 | d4)) | (d2 & (d3 | d4)) | (d3 & d4);
  const uint32_t edge =
      (L.row == 0 || L.row == N - 1) ? Geo<N>::ROWMASK : (1u | (1u << (N - 1)));
  const uint32_t fake = (edge & ge1) | (~edge & ge2);
  return eyeish & ~fake & L.rm;
}

// simple_tt_scoring (go_state.h:32-93): black area minus white area, where a side's area is
// its stones plus the empties reachable from them through empties.
template <int N>
__device__ __forceinline__ int tt_score(uint32_t b, uint32_t w, const Lane& L) {
  const uint32_t e = ~(b | w) & L.rm;
  uint32_t g[2] = {b, w};
  const uint32_t thr[2] = {e, e};
  floodK<N, 2>(g, thr, L);
  return game_sum<N>(__popc(g[0] & ~g[1]) - __popc(g[1] & ~g[0]), L);
}

// ---- applying a move ----------------------------------------------------------
// `p` = y*N+x, MV_PASS, or MV_NONE (leave the game untouched).
enum : int { MV_NONE = -3 };

// ---- legality and moves on the incremental group status --------------------------------------------
// Every stored position (board batch, search node, the playout kernel's registers) carries two masks over ALL stones: `safe` (stone belongs to a
// group with >= 2 liberties) and `atari` (exactly 1).  A group's liberty count only changes when a
// stone lands on one of its liberties, when it merges, or when stones next to it are captured, so
// after a move only the groups touching the new stone or the captured stones are recounted
// (typically 2-4 small fills) instead of classifying every group near a dead-end point (~7 fills
// per ply in random play, 95 % of them groups in atari).  The masks also give:
//   * captures: the enemy neighbour groups of the new stone that are in `atari` (their single
//     liberty is necessarily the point just played) -- no global "still alive" fill;
//   * legality of dead-end points straight from the masks.
// TryPlay semantics for legality (board.cc:788-827: empty, not the active simple-ko point for this
// player, board.cc:234-240, not suicide, board.cc:201-232) and Play (board.cc:1297-1401) for the move;
// the per-ply parity tests (hash, captures, full legal mask against the oracle and the compiled
// reference) pin both.
template <int N>
__device__ __forceinline__ uint32_t legal_rows_cached(uint32_t own, uint32_t opp, uint32_t safe,
                                                      uint32_t atari, const Lane& L, bool ko_applies,
                                                      int ko_pt) {
  const uint32_t e = ~(own | opp) & L.rm;
  const uint32_t en = nbr4<N>(e, L);
  uint32_t legal = e & en;
  const uint32_t hard = e & ~en;
  if (__any_sync(FULL, hard != 0u))
    legal |= hard & (nbr4<N>(safe & own, L) | nbr4<N>(atari & opp, L));
  if (ko_applies) {
    int ky = ko_pt / N, kx = ko_pt - ky * N;
    if (L.row == ky) legal &= ~(1u << kx);
  }
  return legal & L.rm;
}

template <int N>
__device__ __forceinline__ int play_move_cached(uint32_t& b, uint32_t& w, BoardMeta& meta, uint64_t& hash,
                                                int p, const uint64_t* __restrict__ zob, const Lane& L,
                                                uint32_t& safe, uint32_t& atari) {
  const int player = meta.next;
  const int oppc = S_BLACK + S_WHITE - player;
  const bool is_stone = p >= 0;
  uint32_t own = player == S_BLACK ? b : w;
  uint32_t opp = player == S_BLACK ? w : b;
  const int y = is_stone ? p / N : -9, x = is_stone ? p - y * N : 0;
  // the new stone and its (<= 4) neighbour points, straight from (x, y): no shuffles
  const int dy = L.row - y;
  const uint32_t xb = 1u << x;
  const uint32_t mybit = (dy == 0 && L.active) ? xb : 0u;
  const uint32_t nb = (dy == 0 ? ((xb << 1) | (xb >> 1)) : ((dy == 1 || dy == -1) ? xb : 0u)) & L.rm;
  const bool single = !game_any<N>((nb & own) != 0u, L);
  own |= mybit;
  uint64_t dh = 0;
  int ncap = 0;
  uint32_t dead = 0, dead_nb = 0;
  // captures (board.cc:1346-1369): enemy neighbour groups whose only liberty was this point
  const uint32_t dseed = nb & opp & atari;
  if (__any_sync(FULL, dseed != 0u)) {
    dead = flood<N>(dseed, opp, L);
    ncap = game_sum<N>(__popc(dead), L);
    opp &= ~dead;
    safe &= ~dead;
    atari &= ~dead;
    dh = zob_color(game_xor64<N>(zob_row<N>(zob, L.row, dead), L), oppc);
    dead_nb = nbr4<N>(dead, L);
  }
  if (is_stone) {
    hash ^= dh ^ zob_color(zob[(y + 1) * Geo<N>::E + (x + 1)], player);
    if (player == S_BLACK) {
      b = own; w = opp; meta.b_cap += ncap;
    } else {
      w = own; b = opp; meta.w_cap += ncap;
    }
  }
  const uint32_t stones = own | opp;
  const uint32_t e2 = ~stones & L.rm;
  const int libs = game_sum<N>(__popc(nb & e2), L);
  if (__any_sync(FULL, dead != 0u)) {
    uint32_t bal = __ballot_sync(FULL, dead != 0u) & L.segmask;
    int src = __ffs(bal) - 1;
    int dx = __shfl_sync(FULL, __ffs(dead) - 1, src & 31);
    if (is_stone && ncap == 1 && single && libs == 1) {  // simple ko, board.cc:1384-1393
      meta.ko_pt = (int16_t)((src - L.base) * N + dx);
      meta.ko_color = (uint8_t)oppc;
      meta.flags |= F_KO_ACTIVE;
    } else if (is_stone) {
      meta.flags &= ~F_KO_ACTIVE;
    }
  } else if (is_stone) {
    meta.flags &= ~F_KO_ACTIVE;  // no capture, no new ko: _ko_age++ (board.cc:1391)
  }
  // recount the groups whose liberties may have changed
  uint32_t seeds = (mybit | nb | dead_nb) & stones;
  if (single) {
    // the new stone is a group of its own: its liberties are the empty neighbour points
    if (libs == 1) atari |= mybit; else safe |= mybit;
    seeds &= ~mybit;
  }
  if (__any_sync(FULL, seeds != 0u)) {
    const Links k = make_links<N>(own, opp, L);
    const uint32_t linked = k.l | k.r | k.u | k.d;  // stones with a same-colour neighbour
    while (true) {
      const uint32_t bal = __ballot_sync(FULL, seeds != 0u) & L.segmask;
      const int src = __ffs(bal) - 1;
      uint32_t grp = (L.lane == src) ? (seeds & (0u - seeds)) : 0u;
      if (__any_sync(FULL, (grp & linked) != 0u)) {  // single stones need no fill
        while (true) {
          const uint32_t g1 = grow_link(grp, k);
          const uint32_t g2 = grow_link(g1, k);
          const bool ch = g2 != grp;
          grp = g2;
          if (!__any_sync(FULL, ch)) break;
        }
      }
      const int nl = game_sum<N>(__popc(nbr4<N>(grp, L) & e2), L);
      if (nl == 1) {
        atari |= grp;
        safe &= ~grp;
      } else {
        safe |= grp;
        atari &= ~grp;
      }
      seeds &= ~grp;
      if (!__any_sync(FULL, seeds != 0u)) break;
    }
  }
  if (p != MV_NONE) {
    meta.next = (uint8_t)oppc;
    meta.last2 = meta.last1;
    meta.last1 = (int16_t)p;
    meta.ply++;
  }
  return ncap;
}

// GoState::terminated (go_state.h:145-147) from the cached superko flag.
template <int N>
__device__ __forceinline__ bool is_terminated(const BoardMeta& m) {
  return (m.last1 == MV_PASS && m.last2 == MV_PASS) || m.ply >= Geo<N>::MAX_PLY ||
         (m.flags & F_SUPERKO);
}

// GoState::_check_superko (go_state.cc:96-111): does `hash` equal any recorded pre-move
// position hash?  (64-bit hash equality stands in for the reference's hash + 2-bit-board
// comparison; see DESIGN.md.)  `hist` points at this game's record, `n` entries.
template <int N>
__device__ __forceinline__ bool superko_scan(const uint64_t* __restrict__ hist, int n, uint64_t hash,
                                             const Lane& L) {
  bool found = false;
  if (L.active)
    for (int i = L.row; i < n; i += N) found |= (hist[i] == hash);
  return game_any<N>(found, L);
}
// Same test for the one-game-per-warp kernels (k_select): all 32 lanes scan, four independent
// loads in flight per lane, so a 250-entry record costs two round trips instead of fourteen.
// Every lane scans, but only the game's own lanes are guaranteed to hold the position hash (with three
// 9x9 games per warp in the board kernels the per-game reductions run on a game's lanes only, and a move that
// captures leaves the other lanes with a hash that lacks the capture's delta): the scan takes lane 0's copy.
__device__ __forceinline__ bool superko_scan_warp(const uint64_t* __restrict__ hist, int n, uint64_t hash) {
  const int lane = threadIdx.x & 31;
  hash = ((uint64_t)__shfl_sync(FULL, (uint32_t)(hash >> 32), 0) << 32) | __shfl_sync(FULL, (uint32_t)hash, 0);
  bool found = false;
  int i = lane;
  for (; i + 96 < n; i += 128) {
    const uint64_t h0 = hist[i], h1 = hist[i + 32], h2 = hist[i + 64], h3 = hist[i + 96];
    found |= (h0 == hash) | (h1 == hash) | (h2 == hash) | (h3 == hash);
  }
  for (; i < n; i += 32) found |= (hist[i] == hash);
  return __any_sync(FULL, found);
}

// k-th (0-based) set point of `cand` in ascending ACTION order a = x*N + y (x outer, y inner);
// returns the point p = y*N + x.  Requires k < number of candidates of this game.
// Binary search over the column x with per-game REDUX counts, then one ballot for the column.
template <int N>
__device__ __forceinline__ int select_kth_action_order(uint32_t cand, int k, const Lane& L) {
  int lo = 0, hi = N - 1;  // smallest x with count(columns <= x) > k
#pragma unroll
  for (int it = 0; it < 5; ++it) {  // ceil(log2(19)) = 5 halvings (also enough for 9)
    const int mid = (lo + hi) >> 1;
    const int c = game_sum<N>(__popc(cand & ((2u << mid) - 1u)), L);
    if (c > k)
      hi = mid;
    else
      lo = mid + 1;
  }
  const int x = lo < N ? lo : N - 1;
  const int before = game_sum<N>(__popc(cand & ((1u << x) - 1u)), L);
  const uint32_t colmask = (__ballot_sync(FULL, (cand >> x) & 1u) >> L.base) & Geo<N>::ROWMASK;
  const int y = (int)__fns(colmask, 0, k - before + 1);
  return y * N + x;
}

}  // namespace elfb200
