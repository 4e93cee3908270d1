// refstream.cu -- host side of include/elfb200_refstream.h: the random streams of the reference's
// self-play game threads.  No device code in this file (it is a .cu only so that it is built with
// the rest of the library).
//
// What is restated here, with the reference lines it follows:
//   * container order of the root edges: the reference keeps them in
//     std::unordered_map<Coord, EdgeInfo> (tree_search_node.h:310), filled by NodeT::setEvaluation
//     (:176-203) in the order of NodeResponse::pi.  The same map type with the same keys, filled in
//     the same order, is built here, so the iteration order is the library's own;
//   * NodeT::enhanceExploration (tree_search_node.h:132-155);
//   * MCTSResultT::addActions (tree_search_base.h:237-294), MCTSPolicy::normalize / sampleAction
//     (:193-209) and elf_utils::sample_multinomial (elf/utils/utils.h:158-181);
//   * where the game threads draw from which generator (game_base.h:32-38, game_selfplay.cc:47,80-95,
//     game_utils.h:25-30, board_feature.h:74-78).
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <new>
#include <random>
#include <unordered_map>
#include <vector>

#include "elfb200.h"
#include "elfb200_refstream.h"

int elfb200_fail(int code, const char* fmt, ...);  // elfb200.cu

typedef unsigned short Coord;  // base/common.h:34

struct elfb200_refstream {
  int G = 0, N = 0, P1 = 0;
  std::vector<std::mt19937> game;      // GoGameBase::_rng
  std::vector<std::mt19937> actor[2];  // MCTSActor::rng_ of _ai / _ai2
};

namespace {

// action index a = x*N + y (board.h:189) -> expanded coordinate (board.h:183-184); pass = M_PASS = 0
inline Coord action_to_coord(int a, int N) {
  if (a >= N * N) return 0;
  const int x = a / N, y = a % N;
  return (Coord)((y + 1) * (N + 2) + (x + 1));
}

// the root's stateActions_ after setEvaluation: key = move, value = storage index of the edge
inline void fill_container(std::unordered_map<Coord, int>& m, const int16_t* actions, int n, int N) {
  for (int i = 0; i < n; ++i) m.insert(std::make_pair(action_to_coord(actions[i], N), i));
}

inline bool selected(const uint8_t* mask, int g) { return !mask || mask[g]; }

}  // namespace

extern "C" {

int elfb200_refstream_create(int num_games, int board_size, const uint64_t* seeds, elfb200_refstream** out) {
  if (!out || !seeds) return elfb200_fail(ELFB200_ERR_ARG, "NULL argument");
  *out = nullptr;
  if (num_games <= 0 || (board_size != 9 && board_size != 19))
    return elfb200_fail(ELFB200_ERR_ARG, "num_games must be > 0 and board_size 9 or 19");
  elfb200_refstream* rs = new (std::nothrow) elfb200_refstream();
  if (!rs) return elfb200_fail(ELFB200_ERR_STATE, "out of memory");
  rs->G = num_games;
  rs->N = board_size;
  rs->P1 = board_size * board_size + 1;
  rs->game.resize(num_games);
  for (int g = 0; g < num_games; ++g) rs->game[g].seed(seeds[g]);  // _rng.seed(_seed), game_base.h:38
  for (auto& a : rs->actor) a.resize(num_games);
  *out = rs;
  return ELFB200_OK;
}

void elfb200_refstream_destroy(elfb200_refstream* rs) { delete rs; }

int elfb200_refstream_init_actor(elfb200_refstream* rs, int which, const uint8_t* mask) {
  if (!rs || which < 0 || which > 1) return elfb200_fail(ELFB200_ERR_ARG, "bad refstream / actor index");
  for (int g = 0; g < rs->G; ++g)
    if (selected(mask, g)) {
      const uint64_t seed = rs->game[g]();  // params.seed = _rng()   (game_selfplay.cc:47)
      rs->actor[which][g] = std::mt19937(seed);  // rng_(params.seed)     (go/mcts/mcts.h:49)
    }
  return ELFB200_OK;
}

int elfb200_refstream_game_u32(elfb200_refstream* rs, const uint8_t* mask, uint32_t* out) {
  if (!rs) return elfb200_fail(ELFB200_ERR_ARG, "refstream is NULL");
  for (int g = 0; g < rs->G; ++g)
    if (selected(mask, g)) {
      const uint32_t v = (uint32_t)rs->game[g]();
      if (out) out[g] = v;
    }
  return ELFB200_OK;
}

int elfb200_refstream_game_uniform(elfb200_refstream* rs, const uint8_t* mask, double lo, double hi, double* out) {
  if (!rs || !out) return elfb200_fail(ELFB200_ERR_ARG, "NULL argument");
  for (int g = 0; g < rs->G; ++g)
    if (selected(mask, g)) {
      std::uniform_real_distribution<> dis(lo, hi);
      out[g] = dis(rs->game[g]);
    }
  return ELFB200_OK;
}

int elfb200_refstream_actor_d4(elfb200_refstream* rs, int which, const uint8_t* mask, int count, uint8_t* codes) {
  if (!rs || !codes || which < 0 || which > 1 || count < 0) return elfb200_fail(ELFB200_ERR_ARG, "bad argument");
  for (int g = 0; g < rs->G; ++g) {
    uint8_t* row = codes + (size_t)g * count;
    if (!selected(mask, g)) {
      memset(row, 0, count);
      continue;
    }
    std::mt19937 peek = rs->actor[which][g];                         // a copy: nothing is consumed
    for (int i = 0; i < count; ++i) row[i] = (uint8_t)(peek() % 8);  // bf.setD4Code((*rng)() % 8)
  }
  return ELFB200_OK;
}

int elfb200_refstream_actor_discard(elfb200_refstream* rs, int which, const uint8_t* mask, const int32_t* counts) {
  if (!rs || !counts || which < 0 || which > 1) return elfb200_fail(ELFB200_ERR_ARG, "bad argument");
  for (int g = 0; g < rs->G; ++g)
    if (selected(mask, g) && counts[g] > 0) rs->actor[which][g].discard((unsigned long long)counts[g]);
  return ELFB200_OK;
}

int elfb200_refstream_root_noise(elfb200_refstream* rs, int which, const uint8_t* mask, const int32_t* n_edges,
                                 const int16_t* actions, float* priors, float epsilon, float alpha) {
  if (!rs || !n_edges || !actions || !priors || which < 0 || which > 1)
    return elfb200_fail(ELFB200_ERR_ARG, "bad argument");
  if (epsilon == 0.0) return ELFB200_OK;  // tree_search_node.h:135-137
  std::vector<float> etas;
  for (int g = 0; g < rs->G; ++g) {
    const int n = n_edges[g];
    if (!selected(mask, g) || n <= 0) continue;  // a root without edges: the loops below are empty
    if (n > rs->P1) return elfb200_fail(ELFB200_ERR_ARG, "game %d: %d edges", g, n);
    const int16_t* act = actions + (size_t)g * rs->P1;
    float* pr = priors + (size_t)g * rs->P1;
    std::unordered_map<Coord, int> edges;
    fill_container(edges, act, n, rs->N);
    std::gamma_distribution<> dis(alpha);
    etas.assign(n, 0.f);
    float Z = 1e-10;
    for (int i = 0; i < n; ++i) {
      etas[i] = dis(rs->actor[which][g]);
      Z += etas[i];
    }
    int i = 0;
    for (auto& p : edges) {
      // (1 - epsilon) * P + epsilon * eta / Z in float.  The one product feeding the sum is
      // contracted into a fused multiply-add by gcc at the reference's -O3 -march=native on any
      // FMA host (and in oracle/_ref); fmaf states that choice instead of leaving it to flags.
      const float t = epsilon * etas[i] / Z;
      pr[p.second] = fmaf(1 - epsilon, pr[p.second], t);
      i++;
    }
  }
  return ELFB200_OK;
}

int elfb200_refstream_choose(elfb200_refstream* rs, const uint8_t* mask, const int32_t* n_edges,
                             const int16_t* actions, const int32_t* visits, const uint8_t* sample,
                             int32_t* best_edge, int32_t* chosen_edge) {
  if (!rs || !n_edges || !actions || !visits || !best_edge || !chosen_edge)
    return elfb200_fail(ELFB200_ERR_ARG, "NULL argument");
  std::vector<std::pair<int, float>> policy;
  for (int g = 0; g < rs->G; ++g) {
    best_edge[g] = chosen_edge[g] = -1;
    const int n = n_edges[g];
    if (!selected(mask, g) || n <= 0) continue;
    if (n > rs->P1) return elfb200_fail(ELFB200_ERR_ARG, "game %d: %d edges", g, n);
    const int16_t* act = actions + (size_t)g * rs->P1;
    const int32_t* vis = visits + (size_t)g * rs->P1;
    std::unordered_map<Coord, int> edges;
    fill_container(edges, act, n, rs->N);
    // addActions, MOST_VISITED: score = num_visits, first strict maximum in container order
    float max_score = std::numeric_limits<float>::lowest();
    int best = -1;
    policy.clear();
    for (const auto& e : edges) {
      const float score = vis[e.second];
      policy.push_back(std::make_pair(e.second, score));
      if (score > max_score) {
        max_score = score;
        best = e.second;
      }
    }
    best_edge[g] = chosen_edge[g] = best;
    if (sample && sample[g]) {
      // MCTSPolicy::normalize(t = 1)
      const float t = 1;
      float exp_sum = 0;
      for (auto& entry : policy) {
        float e = std::pow(entry.second, 1.0 / t);
        entry.second = e;
        exp_sum += e;
      }
      for (auto& entry : policy) entry.second /= exp_sum;
      // sample_multinomial
      float Z = 0.0;
      for (const auto& vv : policy) Z += vv.second;
      std::uniform_real_distribution<> dis(0, Z);
      const float rd = dis(rs->game[g]);
      size_t pick = policy.size() - 1;
      float accu = 0;
      for (size_t i = 0; i < policy.size(); ++i) {
        accu = policy[i].second + accu;
        if (rd < accu) {
          pick = i;
          break;
        }
      }
      chosen_edge[g] = policy[pick].first;
    }
  }
  return ELFB200_OK;
}

int elfb200_refstream_edge_order(int board_size, int n, const int16_t* actions, int32_t* order) {
  if (!actions || !order || n < 0 || (board_size != 9 && board_size != 19))
    return elfb200_fail(ELFB200_ERR_ARG, "bad argument");
  std::unordered_map<Coord, int> edges;
  fill_container(edges, actions, n, board_size);
  int i = 0;
  for (const auto& e : edges) order[i++] = e.second;
  return ELFB200_OK;
}

}  // extern "C"
