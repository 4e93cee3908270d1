/* elfb200_playout_policy.h -- the deterministic random-playout workload spec.
 *
 * Shared verbatim by the CUDA playout kernel, the CPU oracle and the reference
 * shim so that all three walk the SAME move stream (SURVEY.md 8d, configs
 * 1/2/5): at every ply the candidates are the legal moves of the side to move
 * (reference GoState::checkMove, go_state.cc:123) that are not that side's own
 * true eyes (reference isTrueEye, board.cc:1908), enumerated in ascending
 * action index a = x*N + y (reference EXPORT_OFFSET_XY, board.h:189).  With n
 * candidates the k-th is played, k = pp_pick(...); with none the side passes.
 *
 * The position checksum folds, per ply, the board hash, capture counters, side
 * to move and the full legal-move mask (as N row words: row y, bit x), so one
 * 64-bit value per game pins bit-exact agreement of every intermediate
 * position of a playout.
 *
 * Plain C99 / C++ / CUDA.  No dependencies.
 */
#ifndef ELFB200_PLAYOUT_POLICY_H_
#define ELFB200_PLAYOUT_POLICY_H_

#include <stdint.h>

#if defined(__CUDACC__)
#define PP_HD __host__ __device__ __forceinline__
#else
#define PP_HD static inline
#endif

PP_HD uint64_t pp_splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}

/* index of the candidate to play among n (n >= 1); ply is the reference's
 * 1-based Board::_ply of the position being moved from. */
PP_HD uint32_t pp_pick(uint64_t seed, uint64_t game_id, uint32_t ply, uint32_t n) {
  uint64_t r = pp_splitmix64(seed ^ (game_id * 0x9E3779B97F4A7C15ULL) ^ (uint64_t)ply);
  return (uint32_t)(((r >> 32) * (uint64_t)n) >> 32);
}

/* contribution of one legal-mask row (order independent across rows): one multiply and one
 * xor-shift of (row index, row bits) -- cheap on the GPU's integer pipe, and not GF(2)-linear, so
 * the XOR over rows still separates different masks */
PP_HD uint64_t pp_row_term(uint32_t y, uint32_t row_bits) {
  uint64_t v = ((((uint64_t)(y + 1)) << 32) | (uint64_t)row_bits) * 0xBF58476D1CE4E5B9ULL;
  return v ^ (v >> 29);
}

PP_HD uint64_t pp_rotl64(uint64_t v, int r) { return (v << r) | (v >> (64 - r)); }

/* fold one position into the running checksum: a single splitmix64 round over the XOR of
 * bijectively transformed fields (hash, legal-mask digest, capture counters + side to move) */
PP_HD uint64_t pp_fold3(uint64_t chk, uint64_t hash, uint64_t rows_xor, uint32_t b_cap,
                        uint32_t w_cap, uint32_t next_player) {
  const uint64_t caps = (uint64_t)(b_cap & 0xFFFF) | ((uint64_t)(w_cap & 0xFFFF) << 16) |
                        ((uint64_t)next_player << 32);
  return pp_splitmix64(chk ^ hash ^ pp_rotl64(rows_xor, 23) ^ (caps * 0x9E3779B97F4A7C15ULL));
}

PP_HD uint64_t pp_fold_position(uint64_t chk, uint64_t hash, uint32_t b_cap, uint32_t w_cap,
                                uint32_t next_player, const uint32_t* legal_rows, int n_rows) {
  uint64_t m = 0;
  for (int y = 0; y < n_rows; ++y) m ^= pp_row_term((uint32_t)y, legal_rows[y]);
  return pp_fold3(chk, hash, m, b_cap, w_cap, next_player);
}

PP_HD uint64_t pp_fold_final(uint64_t chk, uint64_t hash, uint32_t ply) {
  return pp_splitmix64(pp_splitmix64(chk ^ hash) ^ (uint64_t)ply);
}

#endif /* ELFB200_PLAYOUT_POLICY_H_ */
