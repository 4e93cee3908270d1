// oracle/ref_shim.cc -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A thin extern "C" wrapper around the UNMODIFIED reference sources under
// /root/reference (compiled in place by oracle/Makefile into oracle/_ref/).
// It exposes the reference's own GoState / Board / BoardFeature behaviour
// (src_cpp/elfgames/go/base/{board.cc,go_state.cc,board_feature.cc}) through
// plain pointers so that tests can pin the C restatement (oracle/go_oracle.c)
// and the CUDA path against the real thing.  No reference code is copied here:
// every function below only *calls* the reference API.
//
// Board size is fixed at compile time by the reference (-DBOARD9x9 -> 9x9).

#include <cstdint>
#include <cstring>
#include <vector>

#include "elfgames/go/base/board.h"
#include "elfgames/go/base/board_feature.h"
#include "elfgames/go/base/go_state.h"
#include "elfgames/go/sgf/sgf.h"

#include "elfb200_playout_policy.h"

namespace {

// GoState keeps the superko test protected (go_state.h:225); a subclass can
// read it without touching the reference.
struct RefState : public GoState {
  RefState() : GoState() {}
  RefState(const RefState& s) : GoState(s) {}
  bool superko() const {
    return _check_superko();
  }
};

inline Coord action_to_coord(int a) {
  const int N = BOARD_SIZE;
  if (a == N * N)
    return M_PASS;
  return OFFSETXY(a / N, a % N);  // action = x*N + y  (board.h:189)
}

} // namespace

extern "C" {

int ref_board_size() {
  return BOARD_SIZE;
}

void* ref_new() {
  return new RefState();
}

void ref_free(void* p) {
  delete static_cast<RefState*>(p);
}

void* ref_clone(void* p) {
  return new RefState(*static_cast<RefState*>(p));
}

void ref_reset(void* p) {
  static_cast<RefState*>(p)->reset();
}

// GoState::forward (go_state.cc:74) with an action index (x*N+y, N*N = pass).
int ref_forward(void* p, int action) {
  return static_cast<RefState*>(p)->forward(action_to_coord(action)) ? 1 : 0;
}

uint64_t ref_hash(void* p) {
  return static_cast<RefState*>(p)->getHashCode();
}

// out[0..11]: ply, next_player, b_cap, w_cap, last_move(action or -1),
// last_move2(action or -1), ko_action(-1 if none active), ko_color, ko_age,
// terminated, two_pass, superko
void ref_info(void* p, int32_t* out) {
  const RefState& s = *static_cast<RefState*>(p);
  const Board& b = s.board();
  const int N = BOARD_SIZE;
  auto c2a = [&](Coord c) -> int32_t {
    if (c == M_PASS)
      return N * N;
    if (c == M_INVALID || c == M_RESIGN)
      return -1;
    return EXPORT_OFFSET(c);
  };
  out[0] = b._ply;
  out[1] = b._next_player;
  out[2] = b._b_cap;
  out[3] = b._w_cap;
  out[4] = c2a(b._last_move);
  out[5] = c2a(b._last_move2);
  out[6] = (b._ko_age == 0 && b._simple_ko != 0) ? EXPORT_OFFSET(b._simple_ko) : -1;
  out[7] = b._simple_ko_color;
  out[8] = b._ko_age;
  out[9] = s.terminated() ? 1 : 0;
  out[10] = s.isTwoPass() ? 1 : 0;
  out[11] = s.superko() ? 1 : 0;
}

// colours by action index: 0 empty, 1 black, 2 white
void ref_stones(void* p, uint8_t* out) {
  const Board& b = static_cast<RefState*>(p)->board();
  const int N = BOARD_SIZE;
  for (int a = 0; a < N * N; ++a)
    out[a] = b._infos[action_to_coord(a)].color;
}

// GoState::checkMove (go_state.cc:123) for every board action; out[N*N].
void ref_legal_mask(void* p, uint8_t* out) {
  const RefState& s = *static_cast<RefState*>(p);
  const int N = BOARD_SIZE;
  for (int a = 0; a < N * N; ++a)
    out[a] = s.checkMove(action_to_coord(a)) ? 1 : 0;
}

// FindAllValidMoves (board.cc:949) -> list of actions, returns count.
int ref_find_all_valid_moves(void* p, int32_t* out) {
  const RefState& s = *static_cast<RefState*>(p);
  AllMoves m;
  FindAllValidMoves(&s.board(), s.nextPlayer(), &m);
  for (int i = 0; i < m.num_moves; ++i)
    out[i] = EXPORT_OFFSET(m.moves[i]);
  return m.num_moves;
}

// isTrueEye (board.cc:1908) for `player` at every action; out[N*N].
void ref_true_eye_mask(void* p, int player, uint8_t* out) {
  const Board& b = static_cast<RefState*>(p)->board();
  const int N = BOARD_SIZE;
  for (int a = 0; a < N * N; ++a)
    out[a] = isTrueEye(&b, action_to_coord(a), (Stone)player) ? 1 : 0;
}

int ref_tt_score(void* p) {
  return simple_tt_scoring(static_cast<RefState*>(p)->board());
}

float ref_evaluate(void* p, float komi) {
  return static_cast<RefState*>(p)->evaluate(komi);
}

// BoardFeature::extractAGZ (board_feature.cc:247) with a D4 code (0..7).
void ref_features_agz(void* p, int d4, float* out) {
  BoardFeature bf(*static_cast<RefState*>(p));
  bf.setD4Code(d4);
  bf.extractAGZ(out);
}

// BoardFeature::extract (board_feature.cc:209-237): the 25-plane DarkForest feature set.
void ref_features_df(void* p, int d4, float* out) {
  BoardFeature bf(*static_cast<RefState*>(p));
  bf.setD4Code(d4);
  bf.extract(out);
}

// BoardFeature::action2Coord / coord2Action under a D4 code, in action space.
int ref_d4_action2action(int d4, int nn_action) {
  RefState s;
  BoardFeature bf(s);
  bf.setD4Code(d4);
  Coord c = bf.action2Coord(nn_action);
  if (c == M_PASS)
    return BOARD_SIZE * BOARD_SIZE;
  return EXPORT_OFFSET(c);
}

// liberties / stones of the group at `action` from the reference's own group table
// (Board::_groups, board.h:71-76); -1 if the point is empty.  Used to port the
// reference gtests' liberty assertions (base/test/go_test.cc).
int ref_group_liberties(void* p, int action) {
  const Board& b = static_cast<RefState*>(p)->board();
  unsigned char id = b._infos[action_to_coord(action)].id;
  if (id == 0 || id == MAX_GROUP)
    return -1;
  return b._groups[id].liberties;
}

int ref_group_stones(void* p, int action) {
  const Board& b = static_cast<RefState*>(p)->board();
  unsigned char id = b._infos[action_to_coord(action)].id;
  if (id == 0 || id == MAX_GROUP)
    return -1;
  return b._groups[id].stones;
}

int ref_num_groups(void* p) {
  return static_cast<RefState*>(p)->board()._num_groups;
}


// One deterministic random-policy playout (SURVEY 8d "config 1"), driven
// through the reference GoState.  Per ply t (0-based) writes, if non-null:
//   moves[t]  action played, hashes[t] hash AFTER the move, caps[2t],caps[2t+1]
// and folds (hash, caps, legal mask of the position BEFORE the move, next
// player) into a running checksum with the shared playout_policy.h mixer.
// Returns number of plies played; *out_chk = checksum; *out_score = tt score.
int ref_playout(
    uint64_t seed,
    uint64_t game_id,
    int max_plies,
    int32_t* moves,
    uint64_t* hashes,
    int32_t* caps,
    uint64_t* out_chk,
    int32_t* out_score) {
  RefState s;
  const int N = BOARD_SIZE;
  uint64_t chk = 0;
  int t = 0;
  std::vector<uint8_t> legal(N * N), eye(N * N);
  while (!s.terminated() && t < max_plies) {
    ref_legal_mask(&s, legal.data());
    ref_true_eye_mask(&s, s.nextPlayer(), eye.data());
    uint32_t rows[32];
    std::memset(rows, 0, sizeof(rows));
    int n = 0;
    for (int a = 0; a < N * N; ++a) {
      if (legal[a]) {
        rows[a % N] |= 1u << (a / N);  // row y, bit x
        if (!eye[a])
          n++;
      }
    }
    chk = pp_fold_position(chk, s.getHashCode(), s.board()._b_cap,
                           s.board()._w_cap, s.nextPlayer(), rows, N);
    int action = N * N;
    if (n > 0) {
      uint32_t k = pp_pick(seed, game_id, (uint32_t)s.getPly(), (uint32_t)n);
      for (int a = 0; a < N * N; ++a) {
        if (legal[a] && !eye[a]) {
          if (k == 0) {
            action = a;
            break;
          }
          --k;
        }
      }
    }
    if (!s.forward(action_to_coord(action)))
      break;
    if (moves)
      moves[t] = action;
    if (hashes)
      hashes[t] = s.getHashCode();
    if (caps) {
      caps[2 * t] = s.board()._b_cap;
      caps[2 * t + 1] = s.board()._w_cap;
    }
    ++t;
  }
  chk = pp_fold_final(chk, s.getHashCode(), (uint32_t)s.getPly());
  if (out_chk)
    *out_chk = chk;
  if (out_score)
    *out_score = simple_tt_scoring(s.board());
  return t;
}

// GoState::showBoard (go_state.h:187-192): the board picture the online console prints.
int ref_show_board(void* p, char* out, int cap) {
  std::string s = static_cast<RefState*>(p)->showBoard();
  if ((int)s.size() + 1 > cap)
    return -1;
  std::memcpy(out, s.c_str(), s.size() + 1);
  return (int)s.size();
}

// Sgf::load(filename, game_string) + iteration over the main line (sgf.h:158-290, sgf.cc:27-58).
// Per entry: actions[i] = x*N+y, N*N for pass, -1 for an off-board / unparsable coordinate;
// players[i] = S_BLACK(1) / S_WHITE(2) / other.  header_i = {size, handi, winner, num_moves},
// header_f = {komi, win_margin}.  Returns the number of entries on the main line (capped), or -1
// if the reference refuses the text.
int ref_sgf_parse(const char* text, int32_t* actions, int32_t* players, int cap, int32_t* header_i,
                  float* header_f) {
  Sgf sgf;
  if (!sgf.load("<mem>", std::string(text)))
    return -1;
  const SgfHeader& h = sgf.getHeader();
  header_i[0] = h.size;
  header_i[1] = h.handi;
  header_i[2] = (int)h.winner;
  header_i[3] = sgf.numMoves();
  header_f[0] = h.komi;
  header_f[1] = h.win_margin;
  int n = 0;
  for (auto it = sgf.begin(); !it.done() && n < cap; ++it) {
    SgfMove m = it.getCurrMove();
    int a;
    if (m.move == M_PASS)
      a = BOARD_SIZE * BOARD_SIZE;
    else if (m.move == M_INVALID || !ON_BOARD(X(m.move), Y(m.move)))
      a = -1;
    else
      a = X(m.move) * BOARD_SIZE + Y(m.move);
    actions[n] = a;
    players[n] = (int)m.player;
    ++n;
  }
  return n;
}

// coord2str2 (sgf.h:73-86): the GTP-style vertex the console prints for the last move.
int ref_vertex_str(int action, char* out, int cap) {
  std::string s = coord2str2(action_to_coord(action));
  if ((int)s.size() + 1 > cap)
    return -1;
  std::memcpy(out, s.c_str(), s.size() + 1);
  return (int)s.size();
}

} // extern "C"
