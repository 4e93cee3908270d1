/* oracle/fakenet.h -- TEST INFRASTRUCTURE ONLY.
 *
 * A deterministic stand-in for the policy/value net used by the MCTS parity tests: the raw
 * policy and value are exact float32 functions of the 64-bit board hash, so the reference
 * search (oracle/_ref), the C restatement and the CUDA search (fed from numpy in the test)
 * all see bit-identical network outputs with pairwise-distinct priors.
 *   u(a)  = ((splitmix64(hash ^ (a+1)*GOLD) >> 40) + 1) * 2^-24      in (0, 1]
 *   pi[a] = u^8  (three float32 squarings; sharp enough that 800 rollouts go several plies deep)
 *   V     = (splitmix64(hash ^ 0x5EED5EED) >> 40) * 2^-23 - 1         in [-1, 1)
 */
#ifndef ORACLE_FAKENET_H_
#define ORACLE_FAKENET_H_

#include <stdint.h>

#include "elfb200_playout_policy.h" /* pp_splitmix64 */

static inline float fakenet_pi(uint64_t hash, int action) {
  uint64_t r = pp_splitmix64(hash ^ ((uint64_t)(action + 1) * 0x9E3779B97F4A7C15ULL));
  float u = (float)((r >> 40) + 1) * (1.0f / 16777216.0f);
  float t = u * u;
  t = t * t;
  t = t * t;
  return t;
}

static inline float fakenet_value(uint64_t hash) {
  uint64_t r = pp_splitmix64(hash ^ 0x5EED5EEDULL);
  return (float)(r >> 40) * (1.0f / 8388608.0f) - 1.0f;
}

#endif
