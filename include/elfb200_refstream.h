/* elfb200_refstream.h -- the random streams of the reference's self-play game threads (host side).
 *
 * The reference draws every random decision of a self-play game from two std::mt19937 generators
 * per game thread and walks its root edges in the iteration order of a libstdc++
 * std::unordered_map<Coord, EdgeInfo>:
 *
 *   GoGameBase::_rng            (common/game_base.h:32-38,62; seeded with GameOptions::seed)
 *     -> MCTSActorParams::seed = _rng()               GoGameSelfPlay::init_ai   (game_selfplay.cc:47)
 *     -> policy.sampleAction(&_rng)                   mcts_make_diverse_move    (game_selfplay.cc:80-95)
 *     -> ResignCheck::check(value, &_rng)             never-resign draw         (game_utils.h:25-30)
 *   MCTSActor::rng_             (go/mcts/mcts.h:49,170; seeded with MCTSActorParams::seed)
 *     -> NodeT::enhanceExploration(eps, alpha, rng)   root Dirichlet noise      (tree_search_node.h:132-155)
 *     -> BoardFeature::RandomShuffle(s, &rng_)        D4 code per evaluated leaf (board_feature.h:74-78)
 *
 * This header is the host-side counterpart: G game threads' generators and the four consumers,
 * built on the same standard-library facilities (std::mt19937, std::uniform_real_distribution<>,
 * std::gamma_distribution<>, std::unordered_map<unsigned short, ...>), so that with
 * GameOptions::seed set a batch of games reproduces, move for move, what the reference's game
 * threads play (single search thread).  Everything here is host logic on
 * small per-move tables -- the reference does the same work on its game threads; the search itself
 * stays on the device (elfb200_mcts.h: elfb200_mcts_root_edges / _set_root_priors / _set_d4_stream
 * are the device ends of these calls).
 *
 * All tables are host memory, row-major with row stride P1 = board_size^2 + 1, one row per game,
 * edges in STORAGE order (= the order MCTSActor::pi2response inserts them, go/mcts/mcts.h:255-332:
 * descending prior).  Actions are action indices a = x*N + y, pass = N*N (board.h:189).
 * `mask` (uint8[G], may be NULL = all games) selects the games a call applies to.
 * Return value: ELFB200_OK or an error code; elfb200_last_error() has the text. */
#ifndef ELFB200_REFSTREAM_H
#define ELFB200_REFSTREAM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct elfb200_refstream elfb200_refstream;

/* G game generators, game g seeded with seeds[g] (GoGameBase: _rng.seed(_seed)). */
int elfb200_refstream_create(int num_games, int board_size, const uint64_t* seeds, elfb200_refstream** out);
void elfb200_refstream_destroy(elfb200_refstream* rs);

/* GoGameSelfPlay::init_ai: the actor generator `which` (0 = _ai, 1 = _ai2) is re-seeded with the
 * game generator's next output. */
int elfb200_refstream_init_actor(elfb200_refstream* rs, int which, const uint8_t* mask);

/* (*rng)() on the game generator (out may be NULL: draw and drop). */
int elfb200_refstream_game_u32(elfb200_refstream* rs, const uint8_t* mask, uint32_t* out);

/* std::uniform_real_distribution<>(lo, hi)(game generator): the never-resign draw uses (0, 1). */
int elfb200_refstream_game_uniform(elfb200_refstream* rs, const uint8_t* mask, double lo, double hi, double* out);

/* The next `count` D4 codes ((*rng)() % 8) the actor generator WOULD hand to evaluated leaves,
 * codes[G][count]; nothing is consumed.  elfb200_refstream_actor_discard consumes counts[g] draws
 * once the device reported how many leaves the move evaluated. */
int elfb200_refstream_actor_d4(elfb200_refstream* rs, int which, const uint8_t* mask, int count, uint8_t* codes);
int elfb200_refstream_actor_discard(elfb200_refstream* rs, int which, const uint8_t* mask, const int32_t* counts);

/* NodeT::enhanceExploration on the root edges of every selected game with n_edges[g] > 0:
 * priors[G][P1] (storage order) are updated in place. */
int elfb200_refstream_root_noise(elfb200_refstream* rs, int which, const uint8_t* mask, const int32_t* n_edges,
                                 const int16_t* actions, float* priors, float epsilon, float alpha);

/* MCTSResultT::addActions (most visited: FIRST maximum in container order,
 * tree_search_base.h:237-294) and, where sample[g] != 0, MCTSPolicy::normalize + sampleAction on the
 * game generator (tree_search_base.h:193-209, utils.h:158-181).  Outputs are edge indices in
 * storage order (-1 for games without edges); chosen_edge = best_edge where sample[g] == 0. */
int elfb200_refstream_choose(elfb200_refstream* rs, const uint8_t* mask, const int32_t* n_edges,
                             const int16_t* actions, const int32_t* visits, const uint8_t* sample,
                             int32_t* best_edge, int32_t* chosen_edge);

/* Iteration order of the reference's edge container after inserting `n` actions in the given order:
 * order[i] = storage index of the i-th edge visited.  (Exposed for the parity tests.) */
int elfb200_refstream_edge_order(int board_size, int n, const int16_t* actions, int32_t* order);

#ifdef __cplusplus
}
#endif
#endif
