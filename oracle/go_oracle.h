/* oracle/go_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the reference's Go board path
 * (/root/reference/src_cpp/elfgames/go/base/{board.cc,go_state.{h,cc},board_feature.{h,cc}}).
 * It is the parity checker for the CUDA path: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference leg may load it.  The product
 * (elf_b200/, include/) never links or calls it.
 *
 * Parity pinning: tests/test_oracle_vs_ref.py checks every function below against
 * the compiled UNMODIFIED reference (oracle/_ref, built by oracle/Makefile) and
 * tests/test_oracle_golden.py against tests/golden/ (reference gtest known answers
 * + fixtures generated from the compiled reference by scripts/gen_golden.py).
 *
 * Conventions: action a = x*N + y (reference EXPORT_OFFSET_XY, board.h:189),
 * pass = N*N; colours 0 empty / 1 black / 2 white (common.h:37-40).
 */
#ifndef GO_ORACLE_H_
#define GO_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct GoOracle GoOracle;

GoOracle* go_new(int board_size);
void go_free(GoOracle* s);
GoOracle* go_clone(const GoOracle* s);
void go_reset(GoOracle* s);
int go_board_size(const GoOracle* s);

int go_forward(GoOracle* s, int action);          /* GoState::forward, go_state.cc:74 */
int go_check_move(const GoOracle* s, int action); /* GoState::checkMove, go_state.cc:123 */
uint64_t go_hash(const GoOracle* s);
/* out[0..11] = ply, next_player, b_cap, w_cap, last_move, last_move2, ko_action,
 * ko_color, ko_age, terminated, two_pass, superko (same as ref_info) */
void go_info(const GoOracle* s, int32_t* out);
void go_stones(const GoOracle* s, uint8_t* out);
void go_legal_mask(const GoOracle* s, uint8_t* out);
void go_true_eye_mask(const GoOracle* s, int player, uint8_t* out);
int go_tt_score(const GoOracle* s);
float go_evaluate(const GoOracle* s, float komi);
void go_features_agz(const GoOracle* s, int d4, float* out);
int go_d4_action2action(int board_size, int d4, int nn_action);
int go_terminated(const GoOracle* s);
int go_ply(const GoOracle* s);
int go_next_player(const GoOracle* s);
int go_last_move(const GoOracle* s);
int go_num_moves(const GoOracle* s);
int go_move_at(const GoOracle* s, int i);

int go_group_liberties(const GoOracle* s, int action);
int go_group_stones(const GoOracle* s, int action);
int go_num_groups(const GoOracle* s);

int go_playout(int board_size, uint64_t seed, uint64_t game_id, int max_plies, int32_t* moves,
               uint64_t* hashes, int32_t* caps, uint64_t* out_chk, int32_t* out_score);

/* run `n_games` playouts with game ids first_id.., return total plies (timing helper) */
int64_t go_playout_many(int board_size, uint64_t seed, uint64_t first_id, int n_games,
                        int max_plies, uint64_t* chks, int32_t* plies, int32_t* scores);

int go_playout_stream(int board_size, uint64_t seed, uint64_t first_id, int slot, int num_slots, int budget,
                      uint64_t* out_acc, int32_t* out_games);

#ifdef __cplusplus
}
#endif
#endif
