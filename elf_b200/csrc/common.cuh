// common.cuh -- pieces shared by the translation units of libelfb200.so (board path: elfb200.cu,
// search path: mcts.cu): Zobrist table, launch geometry, the HBM state descriptor and the context.
#pragma once

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>

#include "board.cuh"
#include "elfb200.h"

namespace elfb200 {

static __device__ const uint64_t g_zobrist[441] = {
#include "elfb200_zobrist.inc"
};

constexpr int BLOCK = 128;  // 4 warps per CTA
constexpr int WARPS = BLOCK / 32;

template <int N>
__device__ __forceinline__ void load_zobrist(uint64_t* s_zob) {
  for (int i = threadIdx.x; i < Geo<N>::ZOB; i += blockDim.x) s_zob[i] = g_zobrist[i];
  __syncthreads();
}

template <int N>
__device__ __forceinline__ int warp_game(const Lane& L, int G, bool& valid) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int g = warp * Geo<N>::GPW + L.sub;
  valid = L.active && g < G;
  return g;
}

__device__ __forceinline__ BoardMeta initial_meta() {
  BoardMeta m;
  m.ply = 1;
  m.next = S_BLACK;
  m.flags = 0;
  m.last1 = MV_INVALID;
  m.last2 = MV_INVALID;
  m.ko_pt = -1;
  m.ko_color = 0;
  m.pad = 0;
  m.b_cap = 0;
  m.w_cap = 0;
  return m;
}

__device__ __forceinline__ BoardMeta load_meta(const BoardMeta* p) {
  uint4 v = *reinterpret_cast<const uint4*>(p);
  BoardMeta m;
  memcpy(&m, &v, 16);
  return m;
}
__device__ __forceinline__ void store_meta(BoardMeta* p, const BoardMeta& m) {
  uint4 v;
  memcpy(&v, &m, 16);
  *reinterpret_cast<uint4*>(p) = v;
}

struct DevState {
  uint64_t* cur;
  uint64_t* ring;
  uint32_t* legal;
  uint64_t* hash;
  BoardMeta* meta;
  uint64_t* sk;
  int32_t* sk_n;
  int G;
};


// ---- AGZ feature planes -------------------------------------------------------------------------
// BoardFeature::extractAGZ (board_feature.cc:247-290) for one position whose <=8 history positions
// (newest first) are staged in shared memory as rows[t][y] = black_row | white_row << 32.
// One thread per OUTPUT cell: the thread resolves its cell through the inverse D4 once
// (InvTransform, board_feature.h:115-130), then emits the 16 stone planes and the 2 side-to-move
// planes; for a fixed plane consecutive threads write consecutive floats (coalesced 128 B/warp).
__device__ __forceinline__ void d4_inverse(int N, int d4, int tx, int ty, int& x, int& y) {
  int a = tx, b = ty;
  if (d4 & 4) { int t = a; a = b; b = t; }
  switch (d4 & 3) {
    case 1: x = N - b - 1; y = a; break;
    case 2: x = N - a - 1; y = N - b - 1; break;
    case 3: x = b; y = N - a - 1; break;
    default: x = a; y = b; break;
  }
}

template <int N>
__device__ __forceinline__ void write_agz_planes(const uint64_t (*rows)[N], int hn, int next, int d4,
                                                 float* __restrict__ out) {
  constexpr int P = Geo<N>::P;
  for (int cell = threadIdx.x; cell < P; cell += blockDim.x) {
    const int tx = cell / N, ty = cell - tx * N;
    int x, y;
    d4_inverse(N, d4, tx, ty, x, y);
    const bool black_first = next == S_BLACK;  // even planes = side to move (board_feature.cc:268-281)
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      float mine = 0.f, theirs = 0.f;
      if (t < hn) {
        const uint64_t r = rows[t][y];
        const float bl = (float)(((uint32_t)r >> x) & 1u), wh = (float)(((uint32_t)(r >> 32) >> x) & 1u);
        mine = black_first ? bl : wh;
        theirs = black_first ? wh : bl;
      }
      out[(2 * t) * P + cell] = mine;
      out[(2 * t + 1) * P + cell] = theirs;
    }
    out[16 * P + cell] = black_first ? 1.0f : 0.0f;
    out[17 * P + cell] = black_first ? 0.0f : 1.0f;
  }
}

}  // namespace elfb200

// ---- host side ---------------------------------------------------------------------------------
int elfb200_fail(int code, const char* fmt, ...);  // records the message for elfb200_last_error()

#define CK(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess)                                                                        \
      return elfb200_fail(ELFB200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                          __FILE__, __LINE__);                                                    \
  } while (0)

#define DISPATCH_N(ctx, expr19, expr9) \
  do {                                 \
    if ((ctx)->N == 19) {              \
      expr19;                          \
    } else {                           \
      expr9;                           \
    }                                  \
  } while (0)

struct elfb200_ctx {
  int N = 0, G = 0, device = 0;
  cudaStream_t stream = nullptr;
  elfb200::DevState st{};
  // scratch
  int32_t* d_actions = nullptr;
  uint8_t* d_ok = nullptr;
  uint8_t* d_bytes = nullptr;   // G * (P+1) export buffer
  int32_t* d_words = nullptr;   // G * 12 export buffer
  int32_t* d_d4 = nullptr;
  float* d_feat = nullptr;      // lazily allocated G*18*P floats
  // playout outputs
  uint64_t* d_po_sk = nullptr;
  uint64_t* d_po_chk = nullptr;
  uint64_t* d_po_hash = nullptr;
  int32_t* d_po_plies = nullptr;
  int32_t* d_po_score = nullptr;
  // pinned staging
  void* h_pin = nullptr;
  size_t h_pin_bytes = 0;
  int64_t launches = 0;
  // move lists of elfb200_replay (grown on demand)
  int16_t* d_replay = nullptr;
  size_t d_replay_bytes = 0;
};
