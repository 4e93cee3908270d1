// tests/simt_emu/cuda_runtime.h -- TEST INFRASTRUCTURE ONLY.
//
// A host stand-in for <cuda_runtime.h> that lets the kernel SOURCES of elf_b200/csrc (board.cuh,
// elfb200.cu, mcts.cu) be compiled with g++ and executed under a small SIMT emulator, so that the
// kernels' LOGIC (lane predication, warp collectives with partial masks, shared-memory staging,
// node-pool bookkeeping) can be checked against the oracle on machines without a GPU.
//
// It is never part of the product: tests/simt_emu/build.py compiles it into
// tests/simt_emu/_build/libelfb200_emu.so, which only tests/test_emu_kernels.py loads (by explicit
// path).  elf_b200.lib loads elf_b200/libelfb200.so and nothing else; without that CUDA library the
// package fails loudly.  Nothing measured or shipped runs here, and it says nothing about timing,
// memory ordering between warps or races -- only about what each thread computes.
//
// Model: one kernel launch = blocks run one after another; the threads of a block are ucontext
// fibers on one OS thread, scheduled round-robin; a fiber yields only inside a warp/block
// collective.  A collective over `mask` completes when every not-yet-exited lane of the mask has
// arrived at a collective with that mask (exited lanes are not waited for, as on hardware); lanes
// that arrive with different operations under the same mask abort the run ("divergent
// collective"), and a scheduling pass without progress aborts as a deadlock.
#pragma once

#include <ucontext.h>

#include <cfloat>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <string>
#include <utility>
#include <vector>

// ---- qualifiers ---------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static
#ifndef __restrict__
#define __restrict__ __restrict
#endif
#define __align__(n) alignas(n)

// ---- vector types ---------------------------------------------------------------------------------
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint2 { unsigned x, y; };
struct uint3 { unsigned x, y, z; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(8) int2 { int x, y; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) ulonglong2 { unsigned long long x, y; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return ulonglong2{x, y}; }

namespace simt {

enum Op { OP_SYNC, OP_SHFL_IDX, OP_SHFL_UP, OP_SHFL_DOWN, OP_SHFL_XOR, OP_BALLOT, OP_ANY, OP_ALL,
          OP_RADD, OP_RMIN, OP_RMAX, OP_RXOR, OP_RAND, OP_ROR, OP_RADD_U, OP_RMIN_U, OP_RMAX_U, OP_MATCH };

struct Slot {
  uint32_t mask = 0, arrived = 0;
  uint64_t gen = 0;
  int op = -1;
  uint64_t in[32];
  int aux[32];
  uint64_t out[2][32];
};

struct Warp {
  uint32_t exited = 0;  // lanes that returned (or never existed)
  std::deque<Slot> slots;  // references stay valid while lanes sleep inside a collective
  Slot& slot(uint32_t mask) {
    for (auto& s : slots)
      if (s.mask == mask) return s;
    slots.emplace_back();
    slots.back().mask = mask;
    return slots.back();
  }
};

struct Fiber {
  ucontext_t ctx;
  std::vector<char> stack;
  unsigned tid = 0;
  bool done = false;
};

struct State {
  dim3 grid, block, bidx;
  std::vector<Fiber> fibers;
  std::vector<Warp> warps;
  Fiber* cur = nullptr;
  ucontext_t sched;
  std::function<void()> body;
  uint64_t progress = 0;
  // block barrier
  unsigned bar_arrived = 0;
  uint64_t bar_gen = 0;
  unsigned live = 0;
  const char* kernel = "?";
};

inline State& S() {
  static State s;
  return s;
}

[[noreturn]] inline void die(const char* what) {
  State& s = S();
  std::fprintf(stderr, "[simt_emu] %s (kernel %s, block %u, thread %u)\n", what, s.kernel, s.bidx.x,
               s.cur ? s.cur->tid : 0u);
  std::abort();
}

inline void yield() {
  State& s = S();
  swapcontext(&s.cur->ctx, &s.sched);
}

// per-kernel counts of completed warp collectives by operation (one count per warp-wide collective)
struct Counters {
  uint64_t n[32] = {0};
};
inline std::vector<std::pair<std::string, Counters>>& counters() {
  static std::vector<std::pair<std::string, Counters>> c;
  return c;
}
inline Counters& counters_for(const char* kernel) {
  for (auto& kv : counters())
    if (kv.first == kernel) return kv.second;
  counters().emplace_back(kernel, Counters());
  return counters().back().second;
}

inline void complete(Slot& sl, uint32_t need) {
  // all needed lanes are here: compute every lane's result
  uint64_t* out = sl.out[sl.gen & 1];
  const int op = sl.op;
  uint32_t ballot = 0;
  int64_t sadd = 0, smin = INT64_MAX, smax = INT64_MIN;
  uint64_t uadd = 0, umin = UINT64_MAX, umax = 0, rx = 0, ra = ~0ull, ro = 0;
  for (int l = 0; l < 32; ++l) {
    if (!((need >> l) & 1)) continue;
    const uint64_t v = sl.in[l];
    if (v & 1) ballot |= 1u << l;
    const int32_t sv = (int32_t)(uint32_t)v;
    sadd += sv;
    smin = sv < smin ? sv : smin;
    smax = sv > smax ? sv : smax;
    const uint32_t uv = (uint32_t)v;
    uadd += uv;
    umin = uv < umin ? uv : umin;
    umax = uv > umax ? uv : umax;
    rx ^= uv;
    ra &= uv;
    ro |= uv;
  }
  for (int l = 0; l < 32; ++l) {
    if (!((need >> l) & 1)) continue;
    uint64_t r = 0;
    int src;
    switch (op) {
      case OP_SYNC: break;
      case OP_SHFL_IDX:
        src = sl.aux[l] & 31;
        r = ((need >> src) & 1) ? sl.in[src] : 0;
        break;
      case OP_SHFL_UP:
        src = l - sl.aux[l];
        r = src >= 0 ? (((need >> src) & 1) ? sl.in[src] : 0) : sl.in[l];
        break;
      case OP_SHFL_DOWN:
        src = l + sl.aux[l];
        r = src < 32 ? (((need >> src) & 1) ? sl.in[src] : 0) : sl.in[l];
        break;
      case OP_SHFL_XOR:
        src = l ^ sl.aux[l];
        r = (src < 32 && ((need >> src) & 1)) ? sl.in[src] : (src < 32 ? 0 : sl.in[l]);
        break;
      case OP_BALLOT: r = ballot; break;
      case OP_ANY: r = ballot != 0; break;
      case OP_ALL: r = ballot == need; break;
      case OP_RADD: r = (uint32_t)(int32_t)sadd; break;
      case OP_RMIN: r = (uint32_t)(int32_t)smin; break;
      case OP_RMAX: r = (uint32_t)(int32_t)smax; break;
      case OP_RADD_U: r = (uint32_t)uadd; break;
      case OP_RMIN_U: r = (uint32_t)umin; break;
      case OP_RMAX_U: r = (uint32_t)umax; break;
      case OP_RXOR: r = (uint32_t)rx; break;
      case OP_RAND: r = (uint32_t)ra; break;
      case OP_ROR: r = (uint32_t)ro; break;
      case OP_MATCH: {
        uint32_t m = 0;
        for (int k = 0; k < 32; ++k)
          if (((need >> k) & 1) && sl.in[k] == sl.in[l]) m |= 1u << k;
        r = m;
        break;
      }
      default: die("unknown collective");
    }
    out[l] = r;
  }
  counters_for(S().kernel).n[op]++;
  sl.arrived = 0;
  sl.op = -1;
  sl.gen++;
  S().progress++;
}

inline uint64_t collective(uint32_t mask, int op, uint64_t val, int aux) {
  State& s = S();
  const unsigned tid = s.cur->tid;
  const int lane = tid & 31;
  Warp& w = s.warps[tid >> 5];
  if (!((mask >> lane) & 1)) die("a lane called a collective whose mask does not include it");
  Slot& sl = w.slot(mask);
  if (sl.op != -1 && sl.op != op) die("divergent collective: lanes of one mask arrived with different operations");
  sl.op = op;
  sl.in[lane] = val;
  sl.aux[lane] = aux;
  sl.arrived |= 1u << lane;
  const uint64_t my_gen = sl.gen;
  for (;;) {
    if (sl.gen != my_gen) break;
    const uint32_t need = mask & ~w.exited;
    if ((sl.arrived & need) == need) {
      complete(sl, need);
      break;
    }
    yield();
  }
  return sl.out[my_gen & 1][lane];
}

inline void block_barrier() {
  State& s = S();
  const uint64_t g = s.bar_gen;
  s.bar_arrived++;
  for (;;) {
    if (s.bar_gen != g) return;
    if (s.bar_arrived >= s.live) {
      s.bar_arrived = 0;
      s.bar_gen++;
      s.progress++;
      return;
    }
    yield();
  }
}

inline void on_exit_lane() {
  State& s = S();
  const unsigned tid = s.cur->tid;
  Warp& w = s.warps[tid >> 5];
  w.exited |= 1u << (tid & 31);
  s.live--;
  s.progress++;
  for (auto& sl : w.slots) {  // lanes waiting on this one no longer have to
    const uint32_t need = sl.mask & ~w.exited;
    if (sl.arrived && need && (sl.arrived & need) == need) complete(sl, need);
  }
  if (s.bar_arrived && s.bar_arrived >= s.live && s.live) {
    s.bar_arrived = 0;
    s.bar_gen++;
  }
}

inline void fiber_main() {
  State& s = S();
  s.body();
  s.cur->done = true;
  on_exit_lane();
  swapcontext(&s.cur->ctx, &s.sched);
}

inline int& order_mode_ref() {  // 0 ascending, 1 reverse, 2 random; initial value from SIMT_EMU_ORDER
  static int mode = [] {
    const char* e = std::getenv("SIMT_EMU_ORDER");
    return !e ? 0 : (e[0] == 'r' && e[1] == 'e') ? 1 : (e[0] == 'r' && e[1] == 'a') ? 2 : 0;
  }();
  return mode;
}

template <class F>
inline void launch(const char* name, dim3 grid, dim3 block, F&& f) {
  State& s = S();
  s.kernel = name;
  s.grid = grid;
  s.block = block;
  s.body = std::function<void()>(f);
  const unsigned nt = block.x * block.y * block.z;
  if (s.fibers.size() < nt) s.fibers.resize(nt);
  for (unsigned b = 0; b < grid.x; ++b) {
    s.bidx = dim3(b, 0, 0);
    s.warps.assign((nt + 31) / 32, Warp());
    if (nt & 31) s.warps.back().exited = ~0u << (nt & 31);
    s.live = nt;
    s.bar_arrived = 0;
    for (unsigned t = 0; t < nt; ++t) {
      Fiber& fb = s.fibers[t];
      if (fb.stack.empty()) fb.stack.resize(256 * 1024);
      fb.tid = t;
      fb.done = false;
      getcontext(&fb.ctx);
      fb.ctx.uc_stack.ss_sp = fb.stack.data();
      fb.ctx.uc_stack.ss_size = fb.stack.size();
      fb.ctx.uc_link = nullptr;
      makecontext(&fb.ctx, (void (*)())fiber_main, 0);
    }
    unsigned remaining = nt;
    // Scheduling order of the lanes between two collectives.  Hardware gives no ordering between
    // the lanes of a warp except at __syncwarp / *_sync; code that reads what another lane wrote
    // without such a barrier only works under one particular order.  SIMT_EMU_ORDER=reverse /
    // random runs the lanes in a different order (a poor man's racecheck for intra-warp hazards).
    const int order_mode = order_mode_ref();
    static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
    std::vector<unsigned> order(nt);
    for (unsigned t = 0; t < nt; ++t) order[t] = order_mode == 1 ? nt - 1 - t : t;
    while (remaining) {
      const uint64_t before = s.progress;
      if (order_mode == 2) {
        for (unsigned t = nt; t > 1; --t) {
          rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull;
          std::swap(order[t - 1], order[(rng_state >> 33) % t]);
        }
      }
      for (unsigned oi = 0; oi < nt; ++oi) {
        const unsigned t = order[oi];
        Fiber& fb = s.fibers[t];
        if (fb.done) continue;
        s.cur = &fb;
        swapcontext(&s.sched, &fb.ctx);
        if (fb.done) remaining--;
      }
      if (remaining && s.progress == before) {
        s.cur = nullptr;
        die("deadlock: a full scheduling pass made no progress (lanes wait at a collective nobody else reaches)");
      }
    }
  }
  s.cur = nullptr;
}

struct Idx {
  unsigned x, y, z;
};
inline Idx thread_idx() {
  State& s = S();
  return Idx{s.cur->tid % s.block.x, 0, 0};
}
inline Idx block_idx() { return Idx{S().bidx.x, 0, 0}; }
inline Idx block_dim() { return Idx{S().block.x, S().block.y, S().block.z}; }
inline Idx grid_dim() { return Idx{S().grid.x, S().grid.y, S().grid.z}; }

template <class T>
inline uint64_t to_bits(T v) {
  uint64_t b = 0;
  static_assert(sizeof(T) <= 8, "collective operand too wide");
  std::memcpy(&b, &v, sizeof(T));
  return b;
}
template <class T>
inline T from_bits(uint64_t b) {
  T v;
  std::memcpy(&v, &b, sizeof(T));
  return v;
}

}  // namespace simt

// collective counters: text dump "kernel op count" per line into buf; reset afterwards if asked
extern "C" __attribute__((used, visibility("default"))) inline int simt_emu_counters(char* buf, int cap, int reset) {
  static const char* names[] = {"syncwarp", "shfl_idx", "shfl_up", "shfl_down", "shfl_xor", "ballot", "any", "all",
                                "redux_add", "redux_min", "redux_max", "redux_xor", "redux_and", "redux_or",
                                "redux_add_u", "redux_min_u", "redux_max_u", "match_any"};
  int len = 0;
  for (auto& kv : simt::counters())
    for (int op = 0; op < 18; ++op)
      if (kv.second.n[op]) {
        const int k = std::snprintf(buf + len, cap > len ? cap - len : 0, "%s %s %llu\n", kv.first.c_str(), names[op],
                                    (unsigned long long)kv.second.n[op]);
        if (k < 0 || len + k >= cap) return -1;
        len += k;
      }
  if (reset) simt::counters().clear();
  return len;
}

extern "C" __attribute__((used, visibility("default"))) inline void simt_emu_set_order(int mode) {
  simt::order_mode_ref() = mode;
}

#define threadIdx (simt::thread_idx())
#define blockIdx (simt::block_idx())
#define blockDim (simt::block_dim())
#define gridDim (simt::grid_dim())
#define warpSize 32

// ---- warp / block primitives ------------------------------------------------------------------------
template <class T> inline T __shfl_sync(unsigned m, T v, int src, int = 32) { return simt::from_bits<T>(simt::collective(m, simt::OP_SHFL_IDX, simt::to_bits(v), src)); }
template <class T> inline T __shfl_up_sync(unsigned m, T v, unsigned d, int = 32) { return simt::from_bits<T>(simt::collective(m, simt::OP_SHFL_UP, simt::to_bits(v), (int)d)); }
template <class T> inline T __shfl_down_sync(unsigned m, T v, unsigned d, int = 32) { return simt::from_bits<T>(simt::collective(m, simt::OP_SHFL_DOWN, simt::to_bits(v), (int)d)); }
template <class T> inline T __shfl_xor_sync(unsigned m, T v, int x, int = 32) { return simt::from_bits<T>(simt::collective(m, simt::OP_SHFL_XOR, simt::to_bits(v), x)); }
inline unsigned __ballot_sync(unsigned m, int p) { return (unsigned)simt::collective(m, simt::OP_BALLOT, p ? 1 : 0, 0); }
inline int __any_sync(unsigned m, int p) { return (int)simt::collective(m, simt::OP_ANY, p ? 1 : 0, 0); }
inline int __all_sync(unsigned m, int p) { return (int)simt::collective(m, simt::OP_ALL, p ? 1 : 0, 0); }
inline void __syncwarp(unsigned m = 0xffffffffu) { simt::collective(m, simt::OP_SYNC, 0, 0); }
inline void __syncthreads() { simt::block_barrier(); }
inline int __reduce_add_sync(unsigned m, int v) { return (int)(uint32_t)simt::collective(m, simt::OP_RADD, (uint32_t)v, 0); }
inline int __reduce_min_sync(unsigned m, int v) { return (int)(uint32_t)simt::collective(m, simt::OP_RMIN, (uint32_t)v, 0); }
inline int __reduce_max_sync(unsigned m, int v) { return (int)(uint32_t)simt::collective(m, simt::OP_RMAX, (uint32_t)v, 0); }
inline unsigned __reduce_add_sync(unsigned m, unsigned v) { return (unsigned)simt::collective(m, simt::OP_RADD_U, v, 0); }
inline unsigned __reduce_min_sync(unsigned m, unsigned v) { return (unsigned)simt::collective(m, simt::OP_RMIN_U, v, 0); }
inline unsigned __reduce_max_sync(unsigned m, unsigned v) { return (unsigned)simt::collective(m, simt::OP_RMAX_U, v, 0); }
inline unsigned __reduce_xor_sync(unsigned m, unsigned v) { return (unsigned)simt::collective(m, simt::OP_RXOR, v, 0); }
inline unsigned __reduce_and_sync(unsigned m, unsigned v) { return (unsigned)simt::collective(m, simt::OP_RAND, v, 0); }
inline unsigned __reduce_or_sync(unsigned m, unsigned v) { return (unsigned)simt::collective(m, simt::OP_ROR, v, 0); }
template <class T> inline unsigned __match_any_sync(unsigned m, T v) { return (unsigned)simt::collective(m, simt::OP_MATCH, simt::to_bits(v), 0); }
inline unsigned __activemask() { return 0xffffffffu; }

// ---- scalar intrinsics ------------------------------------------------------------------------------
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
inline unsigned __brev(unsigned v) {
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i);
  return r;
}
// __fns(mask, base, offset): position of the |offset|-th set bit of mask at or above (offset > 0) /
// at or below (offset < 0) bit `base`; 0xffffffff if there is none; offset 0: base if set
inline unsigned __fns(unsigned mask, unsigned base, int offset) {
  if (offset == 0) return ((mask >> base) & 1u) ? base : 0xffffffffu;
  if (offset > 0) {
    for (unsigned i = base; i < 32; ++i)
      if (((mask >> i) & 1u) && --offset == 0) return i;
  } else {
    for (int i = (int)base; i >= 0; --i)
      if (((mask >> i) & 1u) && ++offset == 0) return (unsigned)i;
  }
  return 0xffffffffu;
}
inline int __float_as_int(float f) { return simt::from_bits<int>(simt::to_bits(f)); }
inline float __int_as_float(int i) { return simt::from_bits<float>(simt::to_bits(i)); }
inline unsigned __float_as_uint(float f) { return simt::from_bits<unsigned>(simt::to_bits(f)); }
inline float __uint_as_float(unsigned i) { return simt::from_bits<float>(simt::to_bits(i)); }
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline float __fdividef(float a, float b) { return a / b; }
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
inline float cospif(float x) { return std::cos(3.14159265358979323846f * x); }
inline float sinpif(float x) { return std::sin(3.14159265358979323846f * x); }
inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
template <class T> inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T> inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <class T> inline T atomicCAS(T* p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }
inline void __threadfence() {}
inline void __threadfence_block() {}
using std::max;
using std::min;

// ---- runtime API (host memory stands in for device memory) -------------------------------------------
typedef int cudaError_t;
#define cudaSuccess 0
typedef struct simt_stream* cudaStream_t;
typedef struct simt_event {
  std::chrono::steady_clock::time_point t;
}* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
#define cudaStreamNonBlocking 1
inline const char* cudaGetErrorString(cudaError_t) { return "simt_emu: no error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) { *p = (T*)std::calloc(n ? n : 1, 1); return *p ? cudaSuccess : 2; }
template <class T> inline cudaError_t cudaMallocHost(T** p, size_t n) { *p = (T*)std::calloc(n ? n : 1, 1); return *p ? cudaSuccess : 2; }
inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
enum { cudaHostAllocMapped = 2 };
template <class T> inline cudaError_t cudaHostAlloc(T** p, size_t n, unsigned) { *p = (T*)std::calloc(n ? n : 1, 1); return *p ? cudaSuccess : 2; }
template <class T> inline cudaError_t cudaHostGetDevicePointer(T** d, void* h, unsigned) { *d = (T*)h; return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = (cudaStream_t)std::malloc(8); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { return cudaStreamCreate(s); }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { std::free(s); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new simt_event(); return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline void __threadfence_system() {}
template <class T> inline T __ldcg(const T* p) { return *p; }
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift) {
  return (unsigned)(((((unsigned long long)hi) << 32) | lo) >> (shift & 31));
}
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
  return cudaSuccess;
}
