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
                   "int main(void) { elfb200_mcts_options o; (void)o; return (int)pp_pick(1, 2, 3, 4) + ELFB200_OK; }\n")
    run(["gcc", f"-std={std}", "-Wall", "-Wextra", "-Werror", "-pedantic", "-fsyntax-only", "-I", INC, str(src)])


@pytest.mark.skipif(not os.path.exists(LIB), reason="libelfb200.so not built")
def test_cpp_caller_of_integration_md_links(tmp_path):
    src = tmp_path / "driver.cc"
    src.write_text(r'''
#include <cstdint>
#include <vector>
#include "elfb200.h"
#include "elfb200_mcts.h"
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
int main() { return 0; }
''')
    out = tmp_path / "driver"
    run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", INC, str(src), "-o", str(out), LIB,
         "-Wl,-rpath," + os.path.dirname(LIB), "-Wl,--allow-shlib-undefined"])
