// oracle/ref_mcts_shim.cc -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// extern "C" wrapper around the UNMODIFIED reference search: elf::ai::tree_search::MCTSAI_T /
// TreeSearchT (src_cpp/elf/ai/tree_search/{mcts.h,tree_search.h,tree_search_node.h}) driven by
// the reference's own Go actor logic (src_cpp/elfgames/go/mcts/mcts.h: pre_evaluate,
// remove_pass_if_dangerous, pi2response), with the network replaced by either the
// deterministic oracle/fakenet.h or a caller-supplied callback.  The reference's MCTSActor keeps
// those routines private and talks to the net through the ELF batching client; the shim only
// re-routes that one call (act_batch) -- every other line executed is the reference's.
//
// The access-control override below applies to this translation unit only.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <queue>
#include <random>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include <nlohmann/json.hpp>
#include <pybind11/pybind11.h>
#include <spdlog/spdlog.h>

#define private public
#define protected public
#include "elfgames/go/mcts/mcts.h"
#undef private
#undef protected
#include "elfgames/go/common/game_utils.h"

#include "fakenet.h"


// ---- link-time stubs ---------------------------------------------------------------------------
// go/mcts/mcts.h drags in the ELF batching client (elf/base/context.h -> elf/comm -> TBB
// concurrent_hash_map); its virtual methods are instantiated but never called here because the
// shim bypasses the client.  libtbb is not built for the oracle, so the handful of out-of-line TBB
// entry points those instantiations reference are defined as traps: reaching one is a bug.
#include <cstdlib>
namespace tbb {
bool spin_rw_mutex_v3::internal_upgrade() { std::abort(); }
void spin_rw_mutex_v3::internal_acquire_reader() { std::abort(); }
bool spin_rw_mutex_v3::internal_try_acquire_reader() { std::abort(); }
bool spin_rw_mutex_v3::internal_try_acquire_writer() { std::abort(); }
namespace internal {
void* NFS_Allocate(size_t, size_t, void*) { std::abort(); }
void throw_exception_v4(exception_id) { std::abort(); }
void* allocate_via_handler_v3(size_t) { std::abort(); }
void deallocate_via_handler_v3(void*) { std::abort(); }
} // namespace internal
} // namespace tbb

namespace {

using elf::ai::tree_search::TSOptions;

// float pi[n][P+1] (NN action order), float v[n]
typedef void (*eval_cb_t)(int n, const float* feats, const uint64_t* hashes, float* pi, float* v);

class ShimActor {
 public:
  using Action = Coord;
  using State = GoState;
  using NodeResponse = elf::ai::tree_search::NodeResponseT<Coord>;

  ShimActor(const MCTSActorParams& params, eval_cb_t cb, std::atomic<long>* n_evals)
      : inner_(nullptr, params), cb_(cb), n_evals_(n_evals) {}

  // Same control flow as MCTSActor::evaluate (go/mcts/mcts.h:73-121) with ai_->act_batch
  // replaced by the fake net / callback.
  void evaluate(const std::vector<const GoState*>& states, std::vector<NodeResponse>* p_resps) {
    if (states.empty())
      return;
    auto& resps = *p_resps;
    resps.resize(states.size());
    std::vector<BoardFeature> sel_bfs;
    std::vector<size_t> sel_indices;
    for (size_t i = 0; i < states.size(); i++) {
      auto res = inner_.pre_evaluate(*states[i], &resps[i]);
      if (res == MCTSActor::EVAL_NEED_NN) {
        sel_bfs.push_back(inner_.get_extractor(*states[i]));
        sel_indices.push_back(i);
      }
    }
    if (sel_bfs.empty())
      return;
    const int P1 = BOARD_NUM_ACTION;
    std::vector<GoReply> replies;
    for (size_t i = 0; i < sel_bfs.size(); ++i)
      replies.emplace_back(sel_bfs[i]);
    const size_t n = sel_bfs.size();
    if (cb_) {
      const size_t F = MAX_NUM_AGZ_FEATURE * BOARD_SIZE * BOARD_SIZE;
      std::vector<float> feats(n * F), pi(n * P1), v(n);
      std::vector<uint64_t> hashes(n);
      for (size_t i = 0; i < n; ++i) {
        sel_bfs[i].extractAGZ(&feats[i * F]);
        hashes[i] = sel_bfs[i].state().getHashCode();
      }
      cb_((int)n, feats.data(), hashes.data(), pi.data(), v.data());
      for (size_t i = 0; i < n; ++i) {
        for (int a = 0; a < P1; ++a)
          replies[i].pi[a] = pi[i * P1 + a];
        replies[i].value = v[i];
      }
    } else {
      for (size_t i = 0; i < n; ++i) {
        const uint64_t h = sel_bfs[i].state().getHashCode();
        for (int a = 0; a < P1; ++a)
          replies[i].pi[a] = fakenet_pi(h, a);
        replies[i].value = fakenet_value(h);
      }
    }
    if (n_evals_)
      n_evals_->fetch_add((long)n);
    for (size_t i = 0; i < n; i++)
      inner_.post_nn_result(replies[i], &resps[sel_indices[i]]);
  }

  void evaluate(const GoState& s, NodeResponse* resp) {
    std::vector<const GoState*> v{&s};
    std::vector<NodeResponse> r;
    evaluate(v, &r);
    *resp = r[0];
  }

  bool forward(GoState& s, Coord a) {
    return inner_.forward(s, a);
  }
  float reward(const GoState& s, float value) const {
    return inner_.reward(s, value);
  }
  std::mt19937* rng() {
    return inner_.rng();
  }
  std::string info() const {
    return inner_.info();
  }
  void setID(int) {}

 private:
  MCTSActor inner_;
  eval_cb_t cb_;
  std::atomic<long>* n_evals_;
};

struct RefMcts {
  std::unique_ptr<elf::ai::tree_search::MCTSAI_T<ShimActor>> ai;
  std::atomic<long> n_evals{0};
};

inline Coord a2c(int a) {
  const int N = BOARD_SIZE;
  if (a == N * N)
    return M_PASS;
  return OFFSETXY(a / N, a % N);
}
inline int c2a(Coord c) {
  if (c == M_PASS)
    return BOARD_SIZE * BOARD_SIZE;
  return EXPORT_OFFSET(c);
}

} // namespace

namespace elf {
namespace ai {
namespace tree_search {
template <>
struct ActorTrait<ShimActor> {
  static std::string to_string(const ShimActor& a) {
    return a.info();
  }
};
} // namespace tree_search
} // namespace ai
} // namespace elf

extern "C" {

// iopts: [0] num_rollouts_per_thread, [1] num_rollouts_per_batch, [2] virtual_loss,
//        [3] persistent_tree, [4] use_prior, [5] unexplored_q_zero, [6] root_unexplored_q_zero,
//        [7] ply_pass_enabled, [8] remove_pass_if_dangerous, [9] seed, [10] num_threads
// fopts: [0] c_puct, [1] komi, [2] root_epsilon, [3] root_alpha
static void* ref_mcts_build(const int32_t* iopts, const float* fopts, eval_cb_t cb, int rotation_flip,
                            uint64_t seed);

void* ref_mcts_new(const int32_t* iopts, const float* fopts, eval_cb_t cb) {
  return ref_mcts_build(iopts, fopts, cb, 0, (uint64_t)iopts[9]);
}

// Same, with the two things GoGameSelfPlay::init_ai leaves to the game thread's generator and the
// defaults: rotation_flip as MCTSActorParams has it (true, go/mcts/mcts.h:26) and
// params.seed = _rng() (game_selfplay.cc:47), a full 32-bit value.
void* ref_mcts_new_ex(const int32_t* iopts, const float* fopts, eval_cb_t cb, int rotation_flip, uint32_t seed) {
  return ref_mcts_build(iopts, fopts, cb, rotation_flip, (uint64_t)seed);
}

static void* ref_mcts_build(const int32_t* iopts, const float* fopts, eval_cb_t cb, int rotation_flip,
                            uint64_t seed) {
  TSOptions opt;
  opt.num_threads = iopts[10] > 0 ? iopts[10] : 1;
  opt.num_rollouts_per_thread = iopts[0];
  opt.num_rollouts_per_batch = iopts[1];
  opt.virtual_loss = iopts[2];
  opt.persistent_tree = iopts[3] != 0;
  opt.alg_opt.use_prior = iopts[4] != 0;
  opt.alg_opt.unexplored_q_zero = iopts[5] != 0;
  opt.alg_opt.root_unexplored_q_zero = iopts[6] != 0;
  opt.alg_opt.c_puct = fopts[0];
  opt.root_epsilon = fopts[2];
  opt.root_alpha = fopts[3];
  opt.pick_method = "most_visited";
  opt.seed = iopts[9];

  MCTSActorParams params;
  params.actor_name = "shim";
  params.ply_pass_enabled = iopts[7];
  params.remove_pass_if_dangerous = iopts[8] != 0;
  params.seed = seed;
  params.komi = fopts[1];
  params.rotation_flip = rotation_flip != 0;  // D4 is drawn from the per-actor mt19937
  params.required_version = -1;

  RefMcts* m = new RefMcts();
  std::atomic<long>* ctr = &m->n_evals;
  m->ai.reset(new elf::ai::tree_search::MCTSAI_T<ShimActor>(
      opt, [params, cb, ctr](int) { return new ShimActor(params, cb, ctr); }));
  return m;
}

void ref_mcts_free(void* p) {
  delete static_cast<RefMcts*>(p);
}

// Outputs of the last search by ACTION index (size N*N+1): visits (-1 where the root has no such
// edge), edge reward sums W, priors; best_q = MCTSGoAI::getValue (go/mcts/mcts.h:358-365).
static int report_last_result(RefMcts* m, Coord c, int32_t* visits, float* wsum, float* prior,
                              float* root_value, float* best_q, int32_t* total_visits) {
  const auto& res = m->ai->getLastResult();
  const int P1 = BOARD_NUM_ACTION;
  for (int a = 0; a < P1; ++a) {
    if (visits) visits[a] = -1;
    if (wsum) wsum[a] = 0;
    if (prior) prior[a] = 0;
  }
  for (const auto& ae : res.action_edge_pairs) {
    int a = c2a(ae.first);
    if (visits) visits[a] = ae.second.num_visits;
    if (wsum) wsum[a] = ae.second.reward;
    if (prior) prior[a] = ae.second.prior_probability;
  }
  if (root_value) *root_value = res.root_value;
  if (best_q) *best_q = res.total_visits == 0 ? res.root_value : res.best_edge_info.getQSA();
  if (total_visits) *total_visits = res.total_visits;
  return c2a(c);
}

// MCTSAI_T::act (elf/ai/tree_search/mcts.h:59-81) on reference state `state` (a RefState from
// ref_shim.cc, which IS-A GoState).  Returns the chosen action.
int ref_mcts_act(void* p, void* state, int32_t* visits, float* wsum, float* prior,
                 float* root_value, float* best_q, int32_t* total_visits) {
  RefMcts* m = static_cast<RefMcts*>(p);
  const GoState& s = *static_cast<GoState*>(state);
  Coord c = M_INVALID;
  m->ai->act(s, &c);
  return report_last_result(m, c, visits, wsum, prior, root_value, best_q, total_visits);
}

// MCTSAI_T::actPolicyOnly (elf/ai/tree_search/mcts.h:83-89 -> TreeSearchT::runPolicyOnly,
// tree_search.h:387-408): the move of a colour with black/white_use_policy_network_only
// (game_selfplay.cc:359-371).  Same outputs as ref_mcts_act.
int ref_mcts_act_policy_only(void* p, void* state, int32_t* visits, float* wsum, float* prior,
                             float* root_value, float* best_q, int32_t* total_visits) {
  RefMcts* m = static_cast<RefMcts*>(p);
  const GoState& s = *static_cast<GoState*>(state);
  Coord c = M_INVALID;
  m->ai->actPolicyOnly(s, &c);
  return report_last_result(m, c, visits, wsum, prior, root_value, best_q, total_visits);
}

// The root edges of the last search in the order MCTSResultT::addActions walked the root's
// stateActions_ (tree_search_base.h:237-294): actions[i] = action index of the i-th edge.
int ref_mcts_last_order(void* p, int32_t* actions) {
  RefMcts* m = static_cast<RefMcts*>(p);
  const auto& res = m->ai->getLastResult();
  int i = 0;
  for (const auto& ae : res.action_edge_pairs) actions[i++] = c2a(ae.first);
  return i;
}

// A game thread's generator (GoGameBase::_rng, game_base.h:38,62).
void* ref_rng_new(uint64_t seed) {
  std::mt19937* r = new std::mt19937();
  r->seed(seed);
  return r;
}
void ref_rng_free(void* r) { delete static_cast<std::mt19937*>(r); }
uint32_t ref_rng_next(void* r) { return (uint32_t)(*static_cast<std::mt19937*>(r))(); }

// The sampling lines of GoGameSelfPlay::mcts_make_diverse_move (game_selfplay.cc:80-88):
// MCTSGoAI::getMCTSPolicy (go/mcts/mcts.h:367-372) then MCTSPolicy::sampleAction on the game
// thread's generator.
int ref_mcts_sample(void* p, void* rng) {
  RefMcts* m = static_cast<RefMcts*>(p);
  auto policy = m->ai->getLastResult().mcts_policy;
  policy.normalize();
  return c2a(policy.sampleAction(static_cast<std::mt19937*>(rng)));
}

// GoStateExt::shouldResign on a persistent ResignCheck (go_state_ext.h:207-214, game_utils.h:15-54):
// the never-resign draw happens at the first call after a reset, from the game thread's generator.
void* ref_resign_new(float thres, float never_resign_ratio) { return new ResignCheck(thres, never_resign_ratio); }
void ref_resign_free(void* rc) { delete static_cast<ResignCheck*>(rc); }
void ref_resign_reset(void* rc) { static_cast<ResignCheck*>(rc)->reset(); }
int ref_resign_check(void* rc, float value, int next_player, void* rng) {
  ResignCheck* c = static_cast<ResignCheck*>(rc);
  std::mt19937* r = static_cast<std::mt19937*>(rng);
  return (next_player == S_BLACK ? c->check(value, r) : c->check(-value, r)) ? 1 : 0;
}

long ref_mcts_num_evals(void* p) {
  return static_cast<RefMcts*>(p)->n_evals.load();
}

void ref_mcts_end_game(void* p, void* state) {
  static_cast<RefMcts*>(p)->ai->endGame(*static_cast<GoState*>(state));
}

} // extern "C"
