/* oracle/go_oracle.c -- TEST INFRASTRUCTURE ONLY (see go_oracle.h).
 *
 * A from-scratch CPU restatement of the reference's Go rules path.  The
 * reference keeps linked-list groups with incremental liberty counters
 * (board.cc:526-782, 1297-1401); this restatement recomputes groups by flood
 * fill on a plain colour array, which yields the same observable behaviour
 * (hash, legality, captures, ko, superko, score, features) and is checked
 * against the compiled reference in tests/test_oracle_vs_ref.py.
 *
 * Internal point index p = y*N + x; action a = x*N + y; expanded coord
 * (y+1)*(N+2) + (x+1) indexes the Zobrist table (board.h:183-184, hash_num.h).
 */
#include "go_oracle.h"

#include <stdlib.h>
#include <string.h>

#include "elfb200_playout_policy.h"

#define MAXN 19
#define MAXP (MAXN * MAXN)
#define HIST 8 /* MAX_NUM_AGZ_HISTORY, board_feature.h:39 */

static const uint64_t kZobrist[441] = {
#include "elfb200_zobrist.inc"
};

enum { EMPTY = 0, BLACK = 1, WHITE = 2 };
enum { MV_PASS = -2, MV_INVALID = -1 }; /* internal encodings of last moves */

typedef struct {
  uint64_t hash;
  uint8_t bits[(MAXP + 3) / 4]; /* 2-bit packed position, mirrors Board::_bits use in go_state.cc:96-111 */
} PosRecord;

struct GoOracle {
  int N;
  uint8_t color[MAXP];
  uint64_t hash;
  int ply;  /* starts at 1, board.cc:106 */
  int next; /* BLACK first, board.cc:99 */
  int b_cap, w_cap;
  int last[4]; /* p, MV_PASS or MV_INVALID; board.cc:100-103 */
  int ko_pt;   /* point or -1; board.h:144-146 */
  int ko_color;
  int ko_age;
  /* AGZ history: ring of the last <=8 positions AFTER each accepted move (go_state.cc:90-92) */
  uint8_t hist[HIST][MAXP];
  int hist_n;   /* number valid (<=8) */
  int hist_pos; /* next write slot */
  /* superko table: pre-move positions of every non-pass move (go_state.cc:113-121) */
  PosRecord* sk;
  int sk_n, sk_cap;
  /* GoState::_moves (go_state.h:217): every accepted action, for moves_since (go_state.h:158-168) */
  int16_t moves[2 * MAXP];
  int n_moves;
};

/* ---- helpers ---------------------------------------------------------- */
static inline int nbrs(int N, int p, int out[4]) {
  int x = p % N, y = p / N, n = 0;
  if (x > 0) out[n++] = p - 1;      /* L */
  if (y > 0) out[n++] = p - N;      /* T */
  if (x < N - 1) out[n++] = p + 1;  /* R */
  if (y < N - 1) out[n++] = p + N;  /* B */
  return n;
}

/* transform_hash, board.cc:24-36 */
static inline uint64_t zob(int N, int p, int color) {
  int E = N + 2;
  uint64_t h = kZobrist[(p / N + 1) * E + (p % N + 1)];
  if (color == BLACK) return h;
  if (color == WHITE) return (h >> 32) | (h << 32);
  return 0;
}

static void pack_bits(const GoOracle* s, uint8_t* bits) {
  int P = s->N * s->N;
  memset(bits, 0, (MAXP + 3) / 4);
  for (int p = 0; p < P; ++p) bits[p >> 2] |= (uint8_t)(s->color[p] << ((p & 3) * 2));
}

/* flood-fill the group containing p; returns stone count, *libs = #distinct liberties,
 * stones listed in out[] */
static int group_of(const GoOracle* s, int p, int* out, int* libs) {
  int N = s->N;
  uint8_t seen[MAXP];
  memset(seen, 0, (size_t)(N * N));
  int c = s->color[p], n = 0, head = 0, nl = 0;
  out[n++] = p;
  seen[p] = 1;
  while (head < n) {
    int q = out[head++], nb[4];
    int k = nbrs(N, q, nb);
    for (int i = 0; i < k; ++i) {
      int r = nb[i];
      if (seen[r]) continue;
      if (s->color[r] == c) {
        seen[r] = 1;
        out[n++] = r;
      } else if (s->color[r] == EMPTY) {
        seen[r] = 1;
        nl++;
      }
    }
  }
  if (libs) *libs = nl;
  return n;
}

/* ---- lifecycle -------------------------------------------------------- */
void go_reset(GoOracle* s) { /* GoState::reset go_state.cc:134-141 + clearBoard board.cc:79-107 */
  int N = s->N;
  PosRecord* sk = s->sk;
  int cap = s->sk_cap;
  memset(s, 0, sizeof(*s));
  s->N = N;
  s->sk = sk;
  s->sk_cap = cap;
  s->ply = 1;
  s->next = BLACK;
  for (int i = 0; i < 4; ++i) s->last[i] = MV_INVALID;
  s->ko_pt = -1;
}

GoOracle* go_new(int board_size) {
  if (board_size < 2 || board_size > MAXN) return NULL;
  GoOracle* s = (GoOracle*)calloc(1, sizeof(GoOracle));
  s->N = board_size;
  go_reset(s);
  return s;
}

void go_free(GoOracle* s) {
  if (!s) return;
  free(s->sk);
  free(s);
}

GoOracle* go_clone(const GoOracle* src) { /* GoState copy-ctor go_state.h:117-124 */
  GoOracle* s = (GoOracle*)malloc(sizeof(GoOracle));
  memcpy(s, src, sizeof(GoOracle));
  s->sk_cap = src->sk_n > 0 ? src->sk_n + 16 : 0;
  s->sk = NULL;
  if (s->sk_cap) {
    s->sk = (PosRecord*)malloc(sizeof(PosRecord) * (size_t)s->sk_cap);
    memcpy(s->sk, src->sk, sizeof(PosRecord) * (size_t)src->sk_n);
  }
  return s;
}

int go_board_size(const GoOracle* s) { return s->N; }
uint64_t go_hash(const GoOracle* s) { return s->hash; }
int go_ply(const GoOracle* s) { return s->ply; }
int go_next_player(const GoOracle* s) { return s->next; }

static inline int p2a(int N, int p) { return (p % N) * N + p / N; }
static inline int a2p(int N, int a) { return (a % N) * N + a / N; }

static int mv2action(int N, int m) {
  if (m == MV_PASS) return N * N;
  if (m == MV_INVALID) return -1;
  return p2a(N, m);
}
int go_last_move(const GoOracle* s) { return mv2action(s->N, s->last[0]); }
int go_num_moves(const GoOracle* s) { return s->n_moves; }
int go_move_at(const GoOracle* s, int i) { return (i >= 0 && i < s->n_moves) ? s->moves[i] : -1; }

/* ---- rules ------------------------------------------------------------ */
/* GoState::_check_superko go_state.cc:96-111 */
static int superko(const GoOracle* s) {
  if (s->last[0] == MV_PASS) return 0;
  uint8_t bits[(MAXP + 3) / 4];
  int packed = 0;
  for (int i = 0; i < s->sk_n; ++i) {
    if (s->sk[i].hash != s->hash) continue;
    if (!packed) {
      pack_bits(s, bits);
      packed = 1;
    }
    if (memcmp(bits, s->sk[i].bits, sizeof(bits)) == 0) return 1;
  }
  return 0;
}

static int two_pass(const GoOracle* s) { /* go_state.h:141-143 */
  return s->last[0] == MV_PASS && s->last[1] == MV_PASS;
}

int go_terminated(const GoOracle* s) { /* go_state.h:145-147, go_common.h:15 */
  return two_pass(s) || s->ply >= 2 * s->N * s->N || superko(s);
}

/* TryPlay, board.cc:788-827: empty, not simple-ko, not suicide */
static int legal_point(const GoOracle* s, int p, int player) {
  if (s->color[p] != EMPTY) return 0;
  /* isSimpleKoViolation board.cc:234-240 */
  if (s->ko_pt == p && s->ko_age == 0 && s->ko_color == player) return 0;
  int nb[4], k = nbrs(s->N, p, nb);
  for (int i = 0; i < k; ++i)
    if (s->color[nb[i]] == EMPTY) return 1; /* isSuicideMove: liberty > 0, board.cc:203 */
  int grp[MAXP];
  for (int i = 0; i < k; ++i) {
    int libs;
    group_of(s, nb[i], grp, &libs);
    if (s->color[nb[i]] == player) {
      if (libs > 1) return 1; /* board.cc:213-215 */
    } else {
      if (libs == 1) return 1; /* board.cc:216-218 */
    }
  }
  return 0;
}

int go_check_move(const GoOracle* s, int action) {
  int N = s->N;
  if (action == N * N) return 1; /* pass always accepted by TryPlay board.cc:794-800 */
  if (action < 0 || action > N * N) return 0;
  return legal_point(s, a2p(N, action), s->next);
}

static void set_stone(GoOracle* s, int p, int c) { /* set_color board.cc:38-51 */
  s->hash ^= zob(s->N, p, s->color[p]);
  s->color[p] = (uint8_t)c;
  s->hash ^= zob(s->N, p, c);
}

/* update_next_move board.cc:1225-1238 */
static void advance(GoOracle* s, int mv) {
  s->next = BLACK + WHITE - s->next;
  s->last[3] = s->last[2];
  s->last[2] = s->last[1];
  s->last[1] = s->last[0];
  s->last[0] = mv;
  s->ply++;
}

/* Play, board.cc:1297-1401 (legality already established) */
static void play_point(GoOracle* s, int p) {
  int N = s->N, player = s->next, opp = BLACK + WHITE - player;
  set_stone(s, p, player);
  int nb[4], k = nbrs(N, p, nb);
  int grp[MAXP], total_capture = 0, capture_pt = -1;
  for (int i = 0; i < k; ++i) {
    int q = nb[i];
    if (s->color[q] != opp) continue;
    int libs, n = group_of(s, q, grp, &libs);
    if (libs == 0) { /* board.cc:1346-1369 */
      for (int j = 0; j < n; ++j) set_stone(s, grp[j], EMPTY);
      total_capture += n;
      capture_pt = q;
      if (player == BLACK)
        s->b_cap += n;
      else
        s->w_cap += n;
    }
  }
  int libs, n = group_of(s, p, grp, &libs);
  if (libs == 1 && n == 1 && total_capture == 1) { /* board.cc:1385-1393 */
    s->ko_pt = capture_pt;
    s->ko_color = opp;
    s->ko_age = 0;
  } else {
    s->ko_age++;
  }
  advance(s, p);
}

int go_forward(GoOracle* s, int action) {
  int N = s->N;
  if (go_terminated(s)) return 0; /* go_state.cc:78-79 */
  if (action == N * N) {          /* pass: board.cc:1306-1309; no superko record go_state.cc:114 */
    advance(s, MV_PASS);
  } else {
    if (action < 0 || action > N * N) return 0;
    int p = a2p(N, action);
    if (!legal_point(s, p, s->next)) return 0;
    /* _add_board_hash: record the PRE-move position, go_state.cc:85,113-121 */
    if (s->sk_n == s->sk_cap) {
      s->sk_cap = s->sk_cap ? s->sk_cap * 2 : 64;
      s->sk = (PosRecord*)realloc(s->sk, sizeof(PosRecord) * (size_t)s->sk_cap);
    }
    s->sk[s->sk_n].hash = s->hash;
    pack_bits(s, s->sk[s->sk_n].bits);
    s->sk_n++;
    play_point(s, p);
  }
  s->moves[s->n_moves++] = (int16_t)action; /* _moves.push_back(c), go_state.cc:89 */
  /* _history.emplace_back(_board), trimmed to 8: go_state.cc:90-92 */
  memcpy(s->hist[s->hist_pos], s->color, (size_t)(N * N));
  s->hist_pos = (s->hist_pos + 1) % HIST;
  if (s->hist_n < HIST) s->hist_n++;
  return 1;
}

void go_info(const GoOracle* s, int32_t* out) {
  int N = s->N;
  out[0] = s->ply;
  out[1] = s->next;
  out[2] = s->b_cap;
  out[3] = s->w_cap;
  out[4] = mv2action(N, s->last[0]);
  out[5] = mv2action(N, s->last[1]);
  out[6] = (s->ko_age == 0 && s->ko_pt >= 0) ? p2a(N, s->ko_pt) : -1;
  out[7] = s->ko_color;
  out[8] = s->ko_age;
  out[9] = go_terminated(s);
  out[10] = two_pass(s);
  out[11] = superko(s);
}

void go_stones(const GoOracle* s, uint8_t* out) {
  int N = s->N;
  for (int a = 0; a < N * N; ++a) out[a] = s->color[a2p(N, a)];
}

void go_legal_mask(const GoOracle* s, uint8_t* out) { /* FindAllValidMoves board.cc:949-968 */
  int N = s->N;
  for (int a = 0; a < N * N; ++a) out[a] = (uint8_t)legal_point(s, a2p(N, a), s->next);
}

/* isTrueEye = isEye && !isFakeEye, board.cc:1850-1910 */
static int true_eye(const GoOracle* s, int p, int player) {
  int N = s->N;
  if (s->color[p] != EMPTY) return 0;
  int nb[4], k = nbrs(N, p, nb);
  for (int i = 0; i < k; ++i)
    if (s->color[nb[i]] != player) return 0;
  int x = p % N, y = p / N, opp = BLACK + WHITE - player, n_opp = 0, n_off = 0;
  for (int dy = -1; dy <= 1; dy += 2)
    for (int dx = -1; dx <= 1; dx += 2) {
      int xx = x + dx, yy = y + dy;
      if (xx < 0 || xx >= N || yy < 0 || yy >= N)
        n_off++;
      else if (s->color[yy * N + xx] == opp)
        n_opp++;
    }
  int fake = (n_off > 0 && n_opp >= 1) || (n_off == 0 && n_opp >= 2);
  return !fake;
}

void go_true_eye_mask(const GoOracle* s, int player, uint8_t* out) {
  int N = s->N;
  for (int a = 0; a < N * N; ++a) out[a] = (uint8_t)true_eye(s, a2p(N, a), player);
}

/* simple_flood_fill go_state.h:32-73: player's stones plus empties reachable through empties */
static void reach(const GoOracle* s, int player, uint8_t* f) {
  int N = s->N, P = N * N, q[MAXP], n = 0, head = 0;
  memset(f, 0, (size_t)P);
  for (int p = 0; p < P; ++p)
    if (s->color[p] == player) {
      f[p] = 1;
      q[n++] = p;
    }
  while (head < n) {
    int p = q[head++], nb[4], k = nbrs(N, p, nb);
    for (int i = 0; i < k; ++i)
      if (s->color[nb[i]] == EMPTY && !f[nb[i]]) {
        f[nb[i]] = 1;
        q[n++] = nb[i];
      }
  }
}

int go_tt_score(const GoOracle* s) { /* simple_tt_scoring go_state.h:75-93 */
  uint8_t b[MAXP], w[MAXP];
  reach(s, BLACK, b);
  reach(s, WHITE, w);
  int P = s->N * s->N, bv = 0, wv = 0;
  for (int p = 0; p < P; ++p) {
    if (b[p] && !w[p])
      bv++;
    else if (w[p] && !b[p])
      wv++;
  }
  return bv - wv;
}

float go_evaluate(const GoOracle* s, float komi) { /* GoState::evaluate go_state.h:194-203 */
  if (superko(s)) return s->next == BLACK ? 1.0f : -1.0f;
  return (float)go_tt_score(s) - komi;
}

/* BoardFeature::Transform, board_feature.h:97-113 */
static void d4_fwd(int N, int d4, int x, int y, int* ox, int* oy) {
  int rot = d4 & 3, flip = (d4 >> 2) == 1, a, b;
  switch (rot) {
    case 1: a = y; b = N - x - 1; break;
    case 2: a = N - x - 1; b = N - y - 1; break;
    case 3: a = N - y - 1; b = x; break;
    default: a = x; b = y; break;
  }
  if (flip) { int t = a; a = b; b = t; }
  *ox = a;
  *oy = b;
}

/* BoardFeature::InvTransform, board_feature.h:115-130 */
static void d4_inv(int N, int d4, int x, int y, int* ox, int* oy) {
  int rot = d4 & 3, flip = (d4 >> 2) == 1, a = x, b = y, c, d;
  if (flip) { int t = a; a = b; b = t; }
  switch (rot) {
    case 1: c = N - b - 1; d = a; break;
    case 2: c = N - a - 1; d = N - b - 1; break;
    case 3: c = b; d = N - a - 1; break;
    default: c = a; d = b; break;
  }
  *ox = c;
  *oy = d;
}

int go_d4_action2action(int N, int d4, int nn_action) { /* action2Coord board_feature.h:139-144 */
  if (nn_action == -1 || nn_action == N * N) return N * N;
  int x, y;
  d4_inv(N, d4, nn_action / N, nn_action % N, &x, &y);
  return x * N + y;
}

void go_features_agz(const GoOracle* s, int d4, float* out) { /* extractAGZ board_feature.cc:247-290 */
  int N = s->N, P = N * N;
  memset(out, 0, sizeof(float) * 18 * (size_t)P);
  int me = s->next, opp = BLACK + WHITE - me;
  for (int t = 0; t < s->hist_n; ++t) {
    const uint8_t* pos = s->hist[(s->hist_pos - 1 - t + 2 * HIST) % HIST];
    for (int p = 0; p < P; ++p) {
      if (pos[p] == EMPTY) continue;
      int tx, ty;
      d4_fwd(N, d4, p % N, p / N, &tx, &ty);
      int plane = 2 * t + (pos[p] == me ? 0 : 1);
      (void)opp;
      out[plane * P + tx * N + ty] = 1.0f;
    }
  }
  float* ind = out + (me == BLACK ? 16 : 17) * P;
  for (int p = 0; p < P; ++p) ind[p] = 1.0f;
}

/* group liberties / stones at `action` (-1 if empty): what Board::_groups[id] holds in the
 * reference (board.h:71-76), recomputed by flood fill here. */
int go_group_liberties(const GoOracle* s, int action) {
  int p = a2p(s->N, action), grp[MAXP], libs;
  if (s->color[p] == EMPTY) return -1;
  group_of(s, p, grp, &libs);
  return libs;
}

int go_group_stones(const GoOracle* s, int action) {
  int p = a2p(s->N, action), grp[MAXP], libs;
  if (s->color[p] == EMPTY) return -1;
  return group_of(s, p, grp, &libs);
}

/* number of groups + 1, as Board::_num_groups counts (board.h:111-113) */
int go_num_groups(const GoOracle* s) {
  int N = s->N, P = N * N, n = 1, grp[MAXP];
  uint8_t seen[MAXP];
  memset(seen, 0, sizeof(seen));
  for (int p = 0; p < P; ++p) {
    if (s->color[p] == EMPTY || seen[p]) continue;
    int k = group_of(s, p, grp, NULL);
    for (int i = 0; i < k; ++i) seen[grp[i]] = 1;
    n++;
  }
  return n;
}

/* ---- the deterministic playout workload (include/elfb200_playout_policy.h) ---- */
int go_playout(int N, uint64_t seed, uint64_t game_id, int max_plies, int32_t* moves,
               uint64_t* hashes, int32_t* caps, uint64_t* out_chk, int32_t* out_score) {
  GoOracle* s = go_new(N);
  uint64_t chk = 0;
  int t = 0, P = N * N;
  uint8_t legal[MAXP];
  while (!go_terminated(s) && t < max_plies) {
    uint32_t rows[32];
    memset(rows, 0, sizeof(rows));
    int n = 0;
    for (int a = 0; a < P; ++a) {
      int p = a2p(N, a);
      legal[a] = (uint8_t)legal_point(s, p, s->next);
      if (legal[a]) {
        rows[a % N] |= 1u << (a / N);
        if (true_eye(s, p, s->next))
          legal[a] = 2; /* legal but excluded from candidates */
        else
          n++;
      }
    }
    chk = pp_fold_position(chk, s->hash, (uint32_t)s->b_cap, (uint32_t)s->w_cap, (uint32_t)s->next,
                           rows, N);
    int action = P;
    if (n > 0) {
      uint32_t k = pp_pick(seed, game_id, (uint32_t)s->ply, (uint32_t)n);
      for (int a = 0; a < P; ++a)
        if (legal[a] == 1) {
          if (k == 0) {
            action = a;
            break;
          }
          --k;
        }
    }
    if (!go_forward(s, action)) break;
    if (moves) moves[t] = action;
    if (hashes) hashes[t] = s->hash;
    if (caps) {
      caps[2 * t] = s->b_cap;
      caps[2 * t + 1] = s->w_cap;
    }
    ++t;
  }
  chk = pp_fold_final(chk, s->hash, (uint32_t)s->ply);
  if (out_chk) *out_chk = chk;
  if (out_score) *out_score = go_tt_score(s);
  go_free(s);
  return t;
}

int64_t go_playout_many(int N, uint64_t seed, uint64_t first_id, int n_games, int max_plies,
                        uint64_t* chks, int32_t* plies, int32_t* scores) {
  int64_t total = 0;
  for (int g = 0; g < n_games; ++g) {
    uint64_t chk;
    int32_t sc;
    int t = go_playout(N, seed, first_id + (uint64_t)g, max_plies, NULL, NULL, NULL, &chk, &sc);
    if (chks) chks[g] = chk;
    if (plies) plies[g] = t;
    if (scores) scores[g] = sc;
    total += t;
  }
  return total;
}

/* steady-state playouts of one slot (elfb200_playout_stream): games first_id+slot, +G, +2G, ...
 * until exactly `budget` plies have been played; fold of the games' checksums in order. */
int go_playout_stream(int N, uint64_t seed, uint64_t first_id, int slot, int G, int budget,
                      uint64_t* out_acc, int32_t* out_games) {
  uint64_t acc = 0, gid = first_id + (uint64_t)slot;
  int remaining = budget, games = 0;
  while (remaining > 0) {
    uint64_t chk;
    int32_t sc;
    int t = go_playout(N, seed, gid, remaining < 2 * N * N ? remaining : 2 * N * N, NULL, NULL, NULL, &chk, &sc);
    acc = pp_splitmix64(acc ^ chk);
    remaining -= t;
    gid += (uint64_t)G;
    games++;
    if (t == 0) break;
  }
  if (out_acc) *out_acc = acc;
  if (out_games) *out_games = games;
  return budget - remaining;
}
