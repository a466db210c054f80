// cuda_emu.h -- TEST INFRASTRUCTURE ONLY: a single-threaded CTA emulator that lets the kernel sources under
// recbole-cdr_b200/csrc compile with g++ (-DXDR_EMU) and run on the CPU, so that kernel logic written while no GPU is
// reachable (index math, MMA fragment ownership, barrier placement, tile bookkeeping) can be checked against the oracle
// by `pytest -m "not gpu"`.  Nothing in the product imports this; libxdr.so never contains it.
//
// Model: one CTA at a time; each CUDA thread is a ucontext fiber; a fiber runs until it reaches a CTA barrier
// (__syncthreads) or a warp-collective (mma.sync, __shfl_*_sync) and then yields.  The scheduler resumes fibers either
// round-robin or in a seeded random order (emu::set_schedule_seed) -- under the random order a missing __syncthreads shows
// up as a result that changes with the seed.  Atomics are plain read-modify-writes (one OS thread).  What this does NOT
// model: memory-ordering bugs, bank conflicts, register pressure, timing, TMA/mbarrier/tcgen05 -- hardware runs remain
// the parity gate for those (tests marked `gpu`).
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(x) __attribute__((aligned(x)))
#define __shared__ static

struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) float2 { float x, y; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline float2 make_float2(float x, float y) { return float2{x, y}; }
struct uint3_emu { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

// the sliver of the CUDA runtime API that the host side of the kernel files touches
typedef void* cudaStream_t;
typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <typename F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }

namespace emu {

constexpr size_t kStackBytes = 256 * 1024;
constexpr size_t kGuardBytes = 4096;  // guard zone behind the dynamic shared memory of a CTA

struct Fiber {
  ucontext_t ctx;
  std::vector<char> stack;
  bool done = false;
  unsigned tid = 0;
};

struct State {
  ucontext_t sched;
  std::vector<Fiber> fibers;
  int current = -1;
  unsigned nthreads = 0;
  // CTA barrier
  unsigned cta_arrived = 0, cta_gen = 0;
  // per-warp collectives
  std::vector<unsigned> warp_arrived, warp_gen;
  std::vector<uint32_t> warp_buf;  // [n_warps][32][8] exchange words
  std::vector<char> dyn_smem;
  std::function<void()> body;
  uint64_t rng = 0;  // 0 = round-robin
};

inline State& st() {
  static State s;
  return s;
}

inline void set_schedule_seed(uint64_t seed) { st().rng = seed; }

inline void yield() {
  State& s = st();
  swapcontext(&s.fibers[s.current].ctx, &s.sched);
}

inline void cta_barrier() {
  State& s = st();
  const unsigned gen = s.cta_gen;
  if (++s.cta_arrived == s.nthreads) {
    s.cta_arrived = 0;
    ++s.cta_gen;
    return;
  }
  while (s.cta_gen == gen) yield();
}

inline void warp_barrier() {
  State& s = st();
  const unsigned w = s.fibers[s.current].tid >> 5;
  const unsigned lanes = std::min(32u, s.nthreads - w * 32);
  const unsigned gen = s.warp_gen[w];
  if (++s.warp_arrived[w] == lanes) {
    s.warp_arrived[w] = 0;
    ++s.warp_gen[w];
    return;
  }
  while (s.warp_gen[w] == gen) yield();
}

inline uint32_t* warp_slot(unsigned lane) {
  State& s = st();
  const unsigned w = s.fibers[s.current].tid >> 5;
  return &s.warp_buf[(size_t)(w * 32 + lane) * 8];
}

inline void* dyn_smem() { return st().dyn_smem.data(); }

}  // namespace emu

// CUDA built-in index variables: refreshed by the scheduler before a fiber resumes
inline uint3_emu threadIdx, blockIdx, blockDim, gridDim;

namespace emu {

inline void fiber_entry() {
  State& s = st();
  s.body();
  s.fibers[s.current].done = true;
  swapcontext(&s.fibers[s.current].ctx, &s.sched);
}

// Runs `body` once per thread of every CTA of the grid; CTAs execute one after another (x fastest), 1-D blocks only.
inline void launch(dim3 grid, dim3 block3, size_t smem_bytes, std::function<void()> body) {
  State& s = st();
  if (block3.y != 1 || block3.z != 1) { std::fprintf(stderr, "cuda_emu: only 1-D blocks are emulated\n"); std::abort(); }
  const unsigned block = block3.x;
  s.body = std::move(body);
  gridDim = {grid.x, grid.y, grid.z};
  blockDim = {block, 1, 1};
  const unsigned n_ctas = grid.x * grid.y * grid.z;
  for (unsigned b = 0; b < n_ctas; ++b) {
    blockIdx = {b % grid.x, (b / grid.x) % grid.y, b / (grid.x * grid.y)};
    s.nthreads = block;
    s.cta_arrived = 0;
    const unsigned nwarps = (block + 31) / 32;
    s.warp_arrived.assign(nwarps, 0);
    s.warp_gen.assign(nwarps, 0);
    s.warp_buf.assign((size_t)nwarps * 32 * 8, 0);
    // poison: reads of unwritten shared memory become visible; the tail is a guard zone checked after the CTA has run
    s.dyn_smem.assign(smem_bytes + kGuardBytes, (char)0xCD);
    s.fibers.clear();
    s.fibers.resize(block);
    for (unsigned t = 0; t < block; ++t) {
      Fiber& f = s.fibers[t];
      f.tid = t;
      f.stack.resize(kStackBytes);
      getcontext(&f.ctx);
      f.ctx.uc_stack.ss_sp = f.stack.data();
      f.ctx.uc_stack.ss_size = f.stack.size();
      f.ctx.uc_link = &s.sched;
      makecontext(&f.ctx, (void (*)())fiber_entry, 0);
    }
    unsigned remaining = block;
    std::vector<unsigned> order(block);
    for (unsigned t = 0; t < block; ++t) order[t] = t;
    uint64_t rng = s.rng ? (s.rng * 0x9E3779B97F4A7C15ull + b + 1) : 0;
    while (remaining) {
      if (rng) {  // seeded Fisher-Yates reshuffle of the resume order each round
        for (unsigned i = block - 1; i > 0; --i) {
          rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
          std::swap(order[i], order[rng % (i + 1)]);
        }
      }
      for (unsigned i = 0; i < block; ++i) {
        const unsigned t = order[i];
        Fiber& f = s.fibers[t];
        if (f.done) continue;
        s.current = (int)t;
        threadIdx = {t, 0, 0};
        swapcontext(&s.sched, &f.ctx);
        if (f.done) --remaining;
      }
    }
    s.current = -1;
    for (size_t g = smem_bytes; g < smem_bytes + kGuardBytes; ++g)
      if (s.dyn_smem[g] != (char)0xCD) {
        std::fprintf(stderr, "cuda_emu: CTA %u wrote %zu bytes past its %zu bytes of dynamic shared memory\n", b,
                     g - smem_bytes + 1, smem_bytes);
        std::abort();
      }
  }
}

// round-to-nearest (ties away) conversion to TF32, as cvt.rna.tf32.f32
inline uint32_t to_tf32(float x) {
  uint32_t u;
  std::memcpy(&u, &x, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return u;  // inf / nan unchanged
  u += 0x1000u;
  return u & 0xffffe000u;
}
inline float as_float(uint32_t u) {
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}

// mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 over the 32 fibers of the calling warp
inline void mma_m16n8k8_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  const unsigned lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  uint32_t* mine = warp_slot(lane);
  for (int i = 0; i < 4; ++i) mine[i] = a[i];
  mine[4] = b[0];
  mine[5] = b[1];
  warp_barrier();
  auto A = [&](unsigned row, unsigned k) {  // a0=(g,t) a1=(g+8,t) a2=(g,t+4) a3=(g+8,t+4)
    const unsigned src = (row & 7) * 4 + (k & 3), reg = (row >= 8 ? 1 : 0) + (k >= 4 ? 2 : 0);
    return as_float(warp_slot(src)[reg] & 0xffffe000u);
  };
  auto B = [&](unsigned k, unsigned n) {  // b0=(k=t,n=g) b1=(k=t+4,n=g)
    const unsigned src = n * 4 + (k & 3), reg = 4 + (k >= 4 ? 1 : 0);
    return as_float(warp_slot(src)[reg] & 0xffffe000u);
  };
  float r[4];
  const unsigned rows[4] = {g, g, g + 8, g + 8}, cols[4] = {2 * t, 2 * t + 1, 2 * t, 2 * t + 1};
  for (int i = 0; i < 4; ++i) {
    double s = c[i];
    for (unsigned k = 0; k < 8; ++k) s += (double)A(rows[i], k) * (double)B(k, cols[i]);
    r[i] = (float)s;
  }
  warp_barrier();
  for (int i = 0; i < 4; ++i) c[i] = r[i];
}

template <typename T>
inline T shfl_exchange(T v, unsigned src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  const unsigned lane = threadIdx.x & 31;
  std::memcpy(warp_slot(lane) + 6, &v, sizeof(T));
  warp_barrier();
  T out;
  std::memcpy(&out, warp_slot(src_lane & 31) + 6, sizeof(T));
  warp_barrier();
  return out;
}

}  // namespace emu

inline void __syncthreads() { emu::cta_barrier(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier(); }
inline void __threadfence() {}
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask) { return emu::shfl_exchange(v, (threadIdx.x & 31) ^ (unsigned)lane_mask); }
template <typename T>
inline T __shfl_sync(unsigned, T v, int src) { return emu::shfl_exchange(v, (unsigned)src); }
template <typename T>
inline T __shfl_down_sync(unsigned, T v, unsigned d) {
  const unsigned lane = threadIdx.x & 31;
  T o = emu::shfl_exchange(v, lane + d < 32 ? lane + d : lane);
  return o;
}
template <typename T>
inline T __ldg(const T* p) { return *p; }
template <typename T>
inline T __ldcg(const T* p) { return *p; }
template <typename T>
inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <typename T>
inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <typename T>
inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T>
inline T atomicCAS(T* p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }
template <typename T>
inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <typename T>
inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) {
  return (unsigned long long)(((unsigned __int128)a * b) >> 64);
}
inline float __uint_as_float(uint32_t u) { return emu::as_float(u); }
inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline int __float_as_int(float f) { int u; std::memcpy(&u, &f, 4); return u; }
inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
inline float __fdividef(float a, float b) { return a / b; }
inline float __expf(float x) { return std::exp(x); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
inline unsigned __ballot_sync(unsigned, int pred) {
  const unsigned lane = threadIdx.x & 31;
  emu::warp_slot(lane)[6] = pred ? 1u : 0u;
  emu::warp_barrier();
  unsigned bits = 0;
  for (unsigned l = 0; l < 32; ++l)
    if (emu::warp_slot(l)[6]) bits |= 1u << l;
  emu::warp_barrier();
  return bits;
}
using std::max;
using std::min;
