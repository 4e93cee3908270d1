// tests/cxx/stdsort_check.cc -- TEST INFRASTRUCTURE ONLY.  Pins oracle/stdsort_emul.h (the restated
// libstdc++ std::sort) against the real std::sort of this toolchain, with the comparator and element
// type of MCTSActor::pi2response (go/mcts/mcts.h:289-295): pairs (move, probability), descending
// probability.  Usage: stdsort_check <seed> <rounds>; prints "ok <cases> heap_sorts <k>" or the first
// difference and exits 1.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <utility>
#include <vector>

#include "../../oracle/stdsort_emul.h"
// the product's copy of the restatement (device code: one lane of k_expand<N, true> runs it), compiled
// for the host here and held to the same checks
#define __device__
#include "../../elf_b200/csrc/stdsort.cuh"
#undef __device__

typedef std::pair<unsigned short, float> data_type;

static long g_heap = 0;

static bool check(const std::vector<float>& p) {
  const int n = (int)p.size();
  std::vector<data_type> ref(n);
  for (int i = 0; i < n; ++i) ref[i] = std::make_pair((unsigned short)i, p[i]);
  std::sort(ref.begin(), ref.end(), [](const data_type& d1, const data_type& d2) { return d1.second > d2.second; });
  std::vector<uint32_t> key(n);
  for (int i = 0; i < n; ++i) memcpy(&key[i], &p[i], 4);
  std::vector<uint16_t> v(n);
  for (int i = 0; i < n; ++i) v[i] = (uint16_t)i;
  SseCtx c = {key.data(), 0};
  sse_sort(&c, v.data(), n);
  g_heap += c.heap_sorts;
  for (int i = 0; i < n; ++i)
    if (v[i] != ref[i].first) {
      printf("DIFF n=%d at %d: emul %d (p=%g) std::sort %d (p=%g)\n", n, i, v[i], p[v[i]], ref[i].first, ref[i].second);
      return false;
    }
  if (n <= 511) {  // the device copy's explicit stack is sized for n <= 511 (it sorts 82 or 362 elements)
    std::vector<uint16_t> d(n);
    for (int i = 0; i < n; ++i) d[i] = (uint16_t)i;
    elfb200::StdSortCtx dc{key.data()};
    elfb200::ss_sort(&dc, d.data(), n);
    for (int i = 0; i < n; ++i)
      if (d[i] != ref[i].first) {
        printf("DIFF (stdsort.cuh) n=%d at %d: %d vs std::sort %d\n", n, i, d[i], ref[i].first);
        return false;
      }
  }
  return true;
}

// McIlroy's adversary ("A killer adversary for quicksort", 1999) run against std::sort itself: the
// values it freezes make THIS library's pivot choices as bad as possible, so the depth limit is hit
static std::vector<int> adv_val;
static int adv_gas, adv_nsolid, adv_candidate;
static bool adv_less(int x, int y) {
  if (adv_val[x] == adv_gas && adv_val[y] == adv_gas) {
    if (x == adv_candidate)
      adv_val[x] = adv_nsolid++;
    else
      adv_val[y] = adv_nsolid++;
  }
  if (adv_val[x] == adv_gas)
    adv_candidate = x;
  else if (adv_val[y] == adv_gas)
    adv_candidate = y;
  return adv_val[x] < adv_val[y];
}
static std::vector<float> killer(int n, bool descending_cmp) {
  adv_val.assign(n, n - 1);
  adv_gas = n - 1;
  adv_nsolid = adv_candidate = 0;
  std::vector<int> ptr(n);
  for (int i = 0; i < n; ++i) ptr[i] = i;
  if (descending_cmp)
    std::sort(ptr.begin(), ptr.end(), [](int a, int b) { return adv_less(b, a); });
  else
    std::sort(ptr.begin(), ptr.end(), [](int a, int b) { return adv_less(a, b); });
  std::vector<float> p(n);
  for (int i = 0; i < n; ++i) p[i] = (float)adv_val[i] / (float)n;
  return p;
}

int main(int argc, char** argv) {
  const unsigned seed = argc > 1 ? (unsigned)atoi(argv[1]) : 1u;
  const int rounds = argc > 2 ? atoi(argv[2]) : 2000;
  std::mt19937 rng(seed);
  long cases = 0;
  for (int r = 0; r < rounds; ++r) {
    const int n = (int)(rng() % 401);
    const int kind = (int)(rng() % 6);
    std::vector<float> p(n);
    for (int i = 0; i < n; ++i) {
      switch (kind) {
        case 0: p[i] = (float)(rng() % 1000003) / 1000003.0f; break;         // (almost) distinct
        case 1: p[i] = (float)(rng() % 8) / 8.0f; break;                      // 8 values: ties everywhere
        case 2: p[i] = (float)(rng() % 64) / 64.0f; break;                    // fp16-like quantisation
        case 3: p[i] = (rng() % 10) ? 0.0f : (float)(rng() % 100) / 100.0f; break;  // mostly zeros
        case 4: p[i] = (float)i / 512.0f + ((rng() % 4) ? 0.0f : 0.25f); break;     // nearly sorted with steps
        default: p[i] = (float)((n - i) / 3) / 200.0f; break;                 // descending runs of equal values
      }
    }
    if (!check(p)) return 1;
    cases++;
  }
  // 362 = the size pi2response sorts; a few more
  const int sizes[] = {17, 33, 82, 100, 362, 363, 400, 1000, 4000};
  for (int n : sizes)
    for (int d = 0; d < 2; ++d) {
      if (!check(killer(n, d != 0))) return 1;
      cases++;
    }
  printf("ok %ld heap_sorts %ld\n", cases, g_heap);
  return 0;
}
