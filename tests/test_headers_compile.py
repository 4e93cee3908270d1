"""The drop-in boundary is a C ABI: the headers must compile as plain C and as C++, and a
reference-side caller written against them (the C++ sketch of INTEGRATION.md §3) must compile and
link against libelfb200.so.  Nothing is executed here (no GPU)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")
LIB = os.path.join(ROOT, "elf_b200", "libelfb200.so")

pytestmark = pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")


def run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]


@pytest.mark.parametrize("std", ["c99", "c11"])
def test_headers_are_plain_c(tmp_path, std):
    src = tmp_path / "use.c"
    src.write_text('#include "elfb200.h"\n#include "elfb200_mcts.h"\n#include "elfb200_playout_policy.h"\n'
                   '#include "elfb200_refstream.h"\n'
                   "int main(void) { elfb200_mcts_options o; elfb200_refstream* r = 0; (void)o; (void)r;\n"
                   "  return (int)pp_pick(1, 2, 3, 4) + ELFB200_OK; }\n")
    run(["gcc", f"-std={std}", "-Wall", "-Wextra", "-Werror", "-pedantic", "-fsyntax-only", "-I", INC, str(src)])


@pytest.mark.skipif(not os.path.exists(LIB), reason="libelfb200.so not built")
def test_cpp_caller_of_integration_md_links(tmp_path):
    src = tmp_path / "driver.cc"
    src.write_text(r'''
#include <cstdint>
#include <vector>
#include "elfb200.h"
#include "elfb200_mcts.h"
#include "elfb200_refstream.h"
// INTEGRATION.md 3: one move of every game, the network round trip left to the caller
static void run_network(const float*, int, float*, float*) {}
int drive(int num_games, float* feat_dev, float* pi_dev, float* v_dev) {
  elfb200_ctx* games = nullptr;
  elfb200_mcts* search = nullptr;
  if (elfb200_create(19, num_games, 0, &games)) return 1;
  elfb200_mcts_options opt;
  elfb200_mcts_default_options(&opt);
  opt.num_rollouts = 800; opt.num_rollouts_per_batch = 8; opt.c_puct = 1.5f; opt.virtual_loss = 1;
  opt.persistent_tree = 1; opt.komi = 7.5f;
  if (elfb200_mcts_create(games, &opt, &search)) return 2;
  elfb200_mcts_begin_move(search, nullptr);
  for (int w = 0; w < elfb200_mcts_waves_per_move(search); ++w) {
    int32_t n = 0;
    elfb200_mcts_select(search, feat_dev, &n);
    run_network(feat_dev, n, pi_dev, v_dev);
    elfb200_mcts_expand_backup(search, pi_dev, v_dev);
  }
  std::vector<int32_t> actions(num_games);
  std::vector<float> values(num_games);
  elfb200_mcts_choose(search, 20, 0.05f, nullptr, 1, actions.data(), values.data());
  std::vector<uint8_t> ok(num_games);
  elfb200_step(games, actions.data(), ok.data());
  elfb200_mcts_advance(search, actions.data());
  std::vector<int16_t> moves(num_games * 4, 0);
  std::vector<int32_t> count(num_games, 4);
  elfb200_replay(games, moves.data(), 4, count.data());
  elfb200_mcts_destroy(search);
  elfb200_destroy(games);
  return 0;
}
// INTEGRATION.md 3: the same move on the reference game threads' random streams
int drive_on_reference_streams(elfb200_mcts* search, int num_games, int rollouts, const uint64_t* seeds) {
  const size_t P1 = 19 * 19 + 1;
  elfb200_refstream* rs = nullptr;
  if (elfb200_refstream_create(num_games, 19, seeds, &rs)) return 1;
  elfb200_refstream_init_actor(rs, 0, nullptr);
  std::vector<int32_t> n_edges(num_games), visits(num_games * P1), used(num_games), best(num_games), chosen(num_games);
  std::vector<int16_t> actions(num_games * P1);
  std::vector<float> wsum(num_games * P1), priors(num_games * P1);
  std::vector<uint8_t> codes((size_t)num_games * rollouts), sample(num_games, 1);
  std::vector<double> u(num_games);
  elfb200_mcts_begin_move(search, nullptr);
  elfb200_mcts_root_edges(search, n_edges.data(), actions.data(), nullptr, nullptr, priors.data());
  elfb200_refstream_root_noise(rs, 0, nullptr, n_edges.data(), actions.data(), priors.data(), 0.25f, 0.03f);
  elfb200_mcts_set_root_priors(search, nullptr, priors.data());
  elfb200_refstream_actor_d4(rs, 0, nullptr, rollouts, codes.data());
  elfb200_mcts_set_d4_stream(search, codes.data(), rollouts);
  /* ... the waves ... */
  elfb200_mcts_d4_used(search, used.data());
  elfb200_refstream_actor_discard(rs, 0, nullptr, used.data());
  elfb200_mcts_root_edges(search, n_edges.data(), actions.data(), visits.data(), wsum.data(), nullptr);
  elfb200_refstream_choose(rs, nullptr, n_edges.data(), actions.data(), visits.data(), sample.data(), best.data(), chosen.data());
  elfb200_refstream_game_uniform(rs, nullptr, 0.0, 1.0, u.data());
  std::vector<uint32_t> r32(num_games);
  elfb200_refstream_game_u32(rs, nullptr, r32.data());
  elfb200_refstream_destroy(rs);
  return 0;
}
int main() { return 0; }
''')
    out = tmp_path / "driver"
    run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", INC, str(src), "-o", str(out), LIB,
         "-Wl,-rpath," + os.path.dirname(LIB), "-Wl,--allow-shlib-undefined"])
