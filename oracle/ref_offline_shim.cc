// oracle/ref_offline_shim.cc -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// extern "C" wrapper around the UNMODIFIED reference record / offline-training code:
//   Record::createFromJson / setJsonFields        src_cpp/elfgames/go/common/record.h:20-330
//   GoStateExtOffline::fromRecord/switchBeforeMove  common/go_state_ext.h:259-335
//   GoFeature::extract{StateExtAGZ,OfflineAction,MCTSPi,Winner,MoveIdx,NumMove,PredictedValue,
//              StateSelfplayVersion,AugCode}        common/game_feature.h:75-139
// so that tests can (a) feed the records elf_b200.record writes to the reference's own parser and
// (b) pin the replay featuriser (elf_b200/replay.py) on the reference's training extractors.
// Every statement executed below the extern "C" line is a call into the reference.
//
// GoStateExtOffline keeps its BoardFeature private and only exposes a random D4 draw; the
// access-control override (this translation unit only) lets the shim set the D4 code it is asked for.
#include <cstdint>
#include <cstring>
#include <string>

#include <nlohmann/json.hpp>
#include <pybind11/pybind11.h>
#include <spdlog/spdlog.h>

#define private public
#define protected public
#include "elfgames/go/common/game_feature.h"
#undef private
#undef protected

extern "C" {

// Parse one record with the reference and serialise it again.  Returns the length written, -1 if
// the reference's parser throws (which is how createBatchFromJson silently drops a record), -2 if
// `cap` is too small.
int ref_record_roundtrip(const char* text, char* out, int cap) {
  try {
    Record r = Record::createFromJson(json::parse(std::string(text)));
    json j;
    r.setJsonFields(j);
    std::string s = j.dump();
    if ((int)s.size() + 1 > cap)
      return -2;
    std::memcpy(out, s.c_str(), s.size() + 1);
    return (int)s.size();
  } catch (...) {
    return -1;
  }
}

// MsgRequest::createFromJson + setJsonFields (record.h:113-135): the server's request message.
int ref_request_roundtrip(const char* text, char* out, int cap) {
  try {
    MsgRequest r = MsgRequest::createFromJson(json::parse(std::string(text)));
    std::string s = r.setJsonFields();
    if ((int)s.size() + 1 > cap)
      return -2;
    std::memcpy(out, s.c_str(), s.size() + 1);
    return (int)s.size();
  } catch (...) {
    return -1;
  }
}

// MCTSPolicy::normalize (tree_search_base.h:193-203) + GoStateExt::addMCTSPolicy
// (go_state_ext.h:168-190) on `n` (action, visit count) pairs given in edge order; out = the u8
// policy indexed by the reference Coord, BOUND_COORD entries.  Returns BOUND_COORD.
int ref_quantise_policy(int n, const int32_t* actions, const float* visits, uint8_t* out) {
  elf::ai::tree_search::MCTSPolicy<Coord> pol;
  for (int i = 0; i < n; ++i) {
    const int a = actions[i];
    const Coord c = a == BOARD_SIZE * BOARD_SIZE ? M_PASS : OFFSETXY(a / BOARD_SIZE, a % BOARD_SIZE);
    pol.addAction(c, visits[i]);
  }
  pol.normalize();
  GameOptions opt;
  GoStateExt st(0, opt);
  st.addMCTSPolicy(pol);
  std::memcpy(out, st._mcts_policies.back().prob, BOUND_COORD);
  return (int)BOUND_COORD;
}

// Number of records Record::createBatchFromJson keeps out of a JSON array (record.h:283-296).
int ref_record_batch_count(const char* text) {
  try {
    return (int)Record::createBatchFromJson(std::string(text)).size();
  } catch (...) {
    return -1;
  }
}

// One training sample the way GoGameTrain::act builds it (train/game_train.cc:22-45) with the
// random choices (move index, D4 code) supplied by the caller.
//   s[18*N*N], offline_a[num_future], mcts_scores[N*N+1], scalars = {winner, predicted_value},
//   ints = {move_idx, num_move, aug_code}, *selfplay_ver
// Returns 0; -1 record refused by the parser; -2 move_to out of range for num_future.
int ref_offline_sample(const char* record_json, int move_to, int d4, int num_future, float* s,
                       int64_t* offline_a, float* mcts_scores, float* scalars, int32_t* ints,
                       int64_t* selfplay_ver) {
  Record r;
  try {
    r = Record::createFromJson(json::parse(std::string(record_json)));
  } catch (...) {
    return -1;
  }
  GameOptions opt;
  opt.num_future_actions = num_future;
  GoStateExtOffline st(0, opt);
  st.fromRecord(r);
  const int n_moves = st.getNumMoves();
  if (n_moves <= num_future - 1 || move_to < 0 || move_to > n_moves - num_future)
    return -2;  // switchRandomMove's range (go_state_ext.h:285-298)
  st.switchBeforeMove((size_t)move_to);
  st._bf.setD4Code(d4);
  GoFeature::extractStateExtAGZ(st, s);
  GoFeature::extractOfflineAction(st, offline_a);
  GoFeature::extractMCTSPi(st, mcts_scores);
  GoFeature::extractWinner(st, &scalars[0]);
  scalars[1] = 0.f;
  if (move_to < (int)r.result.values.size())
    GoFeature::extractPredictedValue(st, &scalars[1]);
  GoFeature::extractMoveIdx(st, &ints[0]);
  GoFeature::extractNumMove(st, &ints[1]);
  GoFeature::extractAugCode(st, &ints[2]);
  GoFeature::extractStateSelfplayVersion(st, selfplay_ver);
  return 0;
}

// The resignation rule as the reference evaluates it: GoStateExt::shouldResign (go_state_ext.h:207-214)
// on the reference's own ResignCheck (game_utils.h:15-54), then the ply test of GoGameSelfPlay::act
// (game_selfplay.cc:387-391).  `never_resign` fixes ResignCheck's single random draw (ratio 1 -> the game
// never resigns, ratio 0 -> it may).  `value` is the black-perspective predicted value.
int ref_should_resign(float resign_thres, int never_resign, float value, int next_player, int ply) {
  ResignCheck rc(resign_thres, never_resign ? 1.0f : 0.0f);
  std::mt19937 rng(0);
  const bool r = next_player == S_BLACK ? rc.check(value, &rng) : rc.check(-value, &rng);
  return (r && ply >= 50) ? 1 : 0;
}

} // extern "C"
