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
#include <deque>
#include <functional>
#include <unordered_map>
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
struct alignas(16) uint4 { unsigned x, y, z, w; };
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
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }

namespace emu {

constexpr size_t kStackBytes = 192 * 1024;
constexpr size_t kGuardBytes = 4096;  // guard zone behind the dynamic shared memory of a CTA

struct MBar {  // emulated mbarrier: phase completes when every expected arrival has come and no transaction bytes are pending
  uint32_t expected = 0, pending = 0, phase = 0;
  int64_t tx = 0;
  void check() {
    if (pending == 0 && tx == 0) { phase ^= 1u; pending = expected; }
  }
};

struct LaunchRec {  // one kernel launch: geometry + the code every thread runs
  dim3 grid;
  unsigned block = 0;
  size_t smem_bytes = 0;
  std::function<void()> body;
};

struct Cta {
  unsigned nthreads = 0, index = 0;
  int launch = 0;                                // index into State::launches
  unsigned arrived = 0, gen = 0;               // CTA barrier
  std::vector<unsigned> warp_arrived, warp_gen;  // per-warp collectives
  std::vector<uint32_t> warp_buf;                // [n_warps][32][8] exchange words
  std::vector<char> dyn_smem;
  size_t smem_bytes = 0;
  std::unordered_map<const void*, MBar> mbars;
  std::vector<uint32_t> tmem;   // tensor memory: 128 lanes x 512 columns of 32 bits, allocated on first tcgen05 use
  unsigned tmem_next = 0;       // bump allocator (columns)
  // the tensor core works ASYNCHRONOUSLY: tcgen05.mma / tcgen05.commit are queued at issue and executed in order, one item at a
  // time, while the scheduler runs other fibers -- operands are read and the accumulator is written when the item executes
  std::deque<std::function<void(Cta&)>> tc_queue;
  unsigned tc_delay = 0;
  std::deque<std::function<void(Cta&)>> dma_queue;   // bulk copies in flight (cp.async.bulk), completed in issue order
  unsigned dma_delay = 0;
  // what the tensor core (async proxy) sees of shared memory: the image as of the last fence.proxy.async executed by a thread
  // of the CTA.  Generic-proxy stores that no such fence followed are NOT visible to tcgen05.mma operand reads.
  std::vector<char> smem_async;
};

struct Fiber {
  ucontext_t ctx;
  char* stack = nullptr;
  bool done = false;
  unsigned tid = 0, cta = 0;
  const char* site = "";        // debug: what the fiber last waited for (printed by the deadlock report)
  const void* site_arg = nullptr;
  std::vector<std::pair<uint32_t*, uint32_t>> tmem_pending;   // tcgen05.ld results not yet released by tcgen05.wait::ld
};

struct State {
  ucontext_t sched;
  std::vector<Fiber> fibers;
  std::vector<Cta> ctas;
  int current = -1;
  std::vector<LaunchRec> launches;  // the launch(es) whose CTAs are resident now
  bool grouping = false;            // group_begin() .. group_run(): concurrent launches are queued, then run TOGETHER
  std::vector<LaunchRec> queued;
  uint64_t rng = 0;  // 0 = round-robin
  uint64_t clock = 0;
};

inline State& st() {
  static State s;
  return s;
}

// Work counters of everything launched since the last reset (per warp-level instruction for the MMAs, bytes for the row
// traffic helpers of xdr_common.cuh): a pre-measurement estimate of instruction mix and algorithmic traffic.
struct Counters {
  uint64_t mma_tf32 = 0, mma_bf16 = 0, umma_tf32 = 0, umma_bf16 = 0;   // warp-level mma.sync instructions; tcgen05.mma instructions
  uint64_t row_load_bytes = 0, row_red_bytes = 0;       // ld_row4 / ldg_row4 ; red_add4
  uint64_t cta_barriers = 0;
};
inline Counters& counters() {
  static Counters c;
  return c;
}

inline void set_schedule_seed(uint64_t seed) { st().rng = seed; }

inline Fiber& self() { return st().fibers[st().current]; }
inline Cta& cta() { return st().ctas[self().cta]; }

inline void yield() {
  State& s = st();
  swapcontext(&s.fibers[s.current].ctx, &s.sched);
}

inline void cta_barrier() {
  Cta& c = cta();
  const unsigned gen = c.gen;
  if (++c.arrived == c.nthreads) {
    ++counters().cta_barriers;
    c.arrived = 0;
    ++c.gen;
    return;
  }
  while (cta().gen == gen) yield();
}

inline void warp_barrier() {
  Cta& c = cta();
  const unsigned w = self().tid >> 5;
  const unsigned lanes = std::min(32u, c.nthreads - w * 32);
  const unsigned gen = c.warp_gen[w];
  if (++c.warp_arrived[w] == lanes) {
    c.warp_arrived[w] = 0;
    ++c.warp_gen[w];
    return;
  }
  while (cta().warp_gen[w] == gen) yield();
}

inline uint32_t* warp_slot(unsigned lane) {
  Cta& c = cta();
  const unsigned w = self().tid >> 5;
  return &c.warp_buf[(size_t)(w * 32 + lane) * 8];
}

inline void* dyn_smem() { return cta().dyn_smem.data(); }

// ---- mbarrier / bulk-copy model (enough for the TMA id-tile ring and the mbarrier hand-offs of steps_persistent.cu) ----
inline void mbar_init(const void* bar, uint32_t count) {
  MBar& m = cta().mbars[bar];
  m.expected = m.pending = count;
  m.phase = 0;
  m.tx = 0;
}
inline void mbar_arrive(const void* bar) { MBar& m = cta().mbars[bar]; --m.pending; m.check(); }
inline void mbar_expect_tx(const void* bar, uint32_t bytes) { MBar& m = cta().mbars[bar]; m.tx += bytes; --m.pending; m.check(); }
inline void mbar_complete_tx(const void* bar, uint32_t bytes) { MBar& m = cta().mbars[bar]; m.tx -= bytes; m.check(); }
inline bool mbar_test(const void* bar, uint32_t parity) { return cta().mbars[bar].phase != (parity & 1u); }
inline void mbar_wait(const void* bar, uint32_t parity) {
  self().site = parity ? "mbar_wait(parity 1)" : "mbar_wait(parity 0)";
  self().site_arg = bar;
  while (!mbar_test(bar, parity)) yield();
  self().site = "";
}
inline void note_site(const char* what, const void* arg = nullptr) { self().site = what; self().site_arg = arg; }
// cp.async.bulk (global -> shared, mbarrier completion): the copy engine is asynchronous too -- the bytes land, and the
// barrier's transaction count drops, some scheduler slices after the issue (same queue discipline as the tensor core)
inline void bulk_g2s(void* dst, const void* src, uint32_t bytes, const void* bar) {
  cta().dma_queue.push_back([=](Cta& c) {
    std::memcpy(dst, src, bytes);
    MBar& m = c.mbars[bar];
    m.tx -= bytes;
    m.check();
  });
}

// ---- tcgen05 / TMEM model -------------------------------------------------------------------------------------------------
// What is modelled is THIS REPOSITORY'S READING of the interface (scripts/ubench_tcgen05.cu is the hardware experiment that
// confirms or corrects it): SWIZZLE_NONE canonical operand layouts addressed through 64-bit shared-memory descriptors
// (start >> 4 in bits [0,14), leading byte offset >> 4 in [16,30), stride byte offset >> 4 in [32,46)), the instruction
// descriptor's major / N / M fields, kind::tf32 (operands truncated to 10 mantissa bits, fp32 accumulation), D in tensor
// memory with lane = row and column = n, tcgen05.ld 32x32b (warp w of a warpgroup reads lanes 32*(w%4)..+31).  MMAs
// complete instantly, so tcgen05.commit is a plain mbarrier arrive.
inline float as_float(uint32_t u);
constexpr unsigned kTmemLanes = 128, kTmemCols = 512;
inline unsigned threadIdx_lane_for_tmem() { return self().tid & 31u; }
inline uint32_t smem_addr(const void* p) { return (uint32_t)(static_cast<const char*>(p) - cta().dyn_smem.data()); }
inline char* smem_ptr(uint32_t addr) { return cta().dyn_smem.data() + addr; }
inline uint32_t tmem_alloc(unsigned ncols) {
  Cta& c = cta();
  if (c.tmem.empty()) c.tmem.assign((size_t)kTmemLanes * kTmemCols, 0xCDCDCDCDu);
  if (c.tmem_next + ncols > kTmemCols) { std::fprintf(stderr, "cuda_emu: tensor memory exhausted\n"); std::abort(); }
  const uint32_t base = c.tmem_next;   // lane 0, column `base`
  c.tmem_next += ncols;
  return base;
}
inline void tmem_dealloc(uint32_t, unsigned ncols) { cta().tmem_next -= ncols; }
inline uint32_t& tmem_at(Cta& k, uint32_t taddr, unsigned lane, unsigned col) {
  const unsigned l = (taddr >> 16) + lane, c = (taddr & 0xffffu) + col;
  if (l >= kTmemLanes || c >= kTmemCols || k.tmem.empty()) { std::fprintf(stderr, "cuda_emu: tensor memory access out of range\n"); std::abort(); }
  return k.tmem[(size_t)l * kTmemCols + c];
}
inline uint32_t& tmem_at(uint32_t taddr, unsigned lane, unsigned col) { return tmem_at(cta(), taddr, lane, col); }
inline const char* smem_at(Cta& k, uint32_t addr, unsigned bytes) {
  if ((size_t)addr + bytes > k.smem_bytes) { std::fprintf(stderr, "cuda_emu: tcgen05 operand read outside shared memory\n"); std::abort(); }
  if (k.smem_async.empty()) {
    std::fprintf(stderr, "cuda_emu: tcgen05.mma reads shared memory, but no thread of the CTA executed fence.proxy.async\n");
    std::abort();
  }
  return k.smem_async.data() + addr;
}
// fence.proxy.async.shared::cta: this thread's earlier generic-proxy writes become visible to the async proxy.  Modelled per
// CTA: the tensor core's view of shared memory is refreshed (writes by threads that fence later are included early -- the
// model errs on the permissive side; writes followed by NO fence before the MMAs execute are caught).
inline void proxy_fence() {
  Cta& k = cta();
  k.smem_async.assign(k.dyn_smem.begin(), k.dyn_smem.begin() + (long)k.smem_bytes);
}
inline float umma_operand(Cta& c, uint64_t desc, bool mn_major, unsigned r, unsigned k) {
  const uint32_t start = (uint32_t)(desc & 0x3fffu) << 4, lbo = (uint32_t)((desc >> 16) & 0x3fffu) << 4,
                 sbo = (uint32_t)((desc >> 32) & 0x3fffu) << 4;
  const uint32_t off = mn_major ? (k / 8) * lbo + (r / 4) * sbo + (k % 8) * 16 + (r % 4) * 4
                                : (k / 4) * lbo + (r / 8) * sbo + (r % 8) * 16 + (k % 4) * 4;
  uint32_t u;
  std::memcpy(&u, smem_at(c, start + off, 4), 4);
  return as_float(u & 0xffffe000u);
}
// 16-bit operands (8 elements per 16-byte chunk): the same canonical layouts with T = 8
inline float umma_operand_bf16(Cta& c, uint64_t desc, bool mn_major, unsigned r, unsigned k) {
  const uint32_t start = (uint32_t)(desc & 0x3fffu) << 4, lbo = (uint32_t)((desc >> 16) & 0x3fffu) << 4,
                 sbo = (uint32_t)((desc >> 32) & 0x3fffu) << 4;
  const uint32_t off = mn_major ? (k / 8) * lbo + (r / 8) * sbo + (k % 8) * 16 + (r % 8) * 2
                                : (k / 8) * lbo + (r / 8) * sbo + (r % 8) * 16 + (k % 8) * 2;
  uint16_t h;
  std::memcpy(&h, smem_at(c, start + off, 2), 2);
  return as_float((uint32_t)h << 16);
}
// tcgen05.mma.cta_group::1.kind::f16 with BF16 operands: D[M x N] (+)= A[M x 16] * B[N x 16]^T, fp32 accumulation
inline void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
  ++counters().umma_bf16;
  const bool a_mn = (idesc >> 15) & 1u, b_mn = (idesc >> 16) & 1u;
  const unsigned N = ((idesc >> 17) & 0x3fu) << 3, M = ((idesc >> 24) & 0x1fu) << 4;
  if (M != 128 || ((idesc >> 4) & 3u) != 1u || ((idesc >> 7) & 7u) != 1u || ((idesc >> 10) & 7u) != 1u) {
    std::fprintf(stderr, "cuda_emu: only M = 128, F32 accumulate, BF16 operands are modelled for kind::f16 (idesc %08x)\n", idesc);
    std::abort();
  }
  cta().tc_queue.push_back([=](Cta& c) {
    for (unsigned m = 0; m < M; ++m)
      for (unsigned n = 0; n < N; ++n) {
        double s = 0;
        for (unsigned k = 0; k < 16; ++k)
          s += (double)umma_operand_bf16(c, desc_a, a_mn, m, k) * (double)umma_operand_bf16(c, desc_b, b_mn, n, k);
        uint32_t& d = tmem_at(c, tmem_d, m, n);
        const float r = (accumulate ? as_float(d) : 0.f) + (float)s;
        std::memcpy(&d, &r, 4);
      }
  });
}
// tcgen05.mma.cta_group::1.kind::tf32: D[M x N] (+)= A[M x 8] * B[N x 8]^T
inline void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
  ++counters().umma_tf32;
  const bool a_mn = (idesc >> 15) & 1u, b_mn = (idesc >> 16) & 1u;
  const unsigned N = ((idesc >> 17) & 0x3fu) << 3, M = ((idesc >> 24) & 0x1fu) << 4;
  if (M != 128 || ((idesc >> 4) & 3u) != 1u || ((idesc >> 7) & 7u) != 2u || ((idesc >> 10) & 7u) != 2u) {
    std::fprintf(stderr, "cuda_emu: only M = 128, F32 accumulate, TF32 operands are modelled (idesc %08x)\n", idesc);
    std::abort();
  }
  cta().tc_queue.push_back([=](Cta& c) {
    for (unsigned m = 0; m < M; ++m)
      for (unsigned n = 0; n < N; ++n) {
        double s = 0;
        for (unsigned k = 0; k < 8; ++k)
          s += (double)umma_operand(c, desc_a, a_mn, m, k) * (double)umma_operand(c, desc_b, b_mn, n, k);
        uint32_t& d = tmem_at(c, tmem_d, m, n);
        const float r = (accumulate ? as_float(d) : 0.f) + (float)s;
        std::memcpy(&d, &r, 4);
      }
  });
}
// tcgen05.commit: the mbarrier arrive happens when every MMA queued before it has executed
inline void tc_commit(const void* bar) {
  cta().tc_queue.push_back([=](Cta& c) {
    MBar& m = c.mbars[bar];
    --m.pending;
    m.check();
  });
}
// one step of every CTA's tensor core: called by the scheduler between fiber slices
inline void tc_tick(std::vector<Cta>& ctas, uint64_t& rng) {
  for (Cta& c : ctas) {
    if (!c.dma_queue.empty()) {
      if (c.dma_delay > 0) {
        --c.dma_delay;
      } else {
        auto op = std::move(c.dma_queue.front());
        c.dma_queue.pop_front();
        op(c);
        if (rng) {
          rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
          c.dma_delay = (unsigned)(rng % 7);
        }
      }
    }
    if (c.tc_queue.empty()) continue;
    if (c.tc_delay > 0) { --c.tc_delay; continue; }
    auto op = std::move(c.tc_queue.front());
    c.tc_queue.pop_front();
    op(c);
    if (rng) {
      rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
      c.tc_delay = (unsigned)(rng % 5);    // seeded schedules: completion times vary
    }
  }
}
// tcgen05.ld.sync.aligned.32x32b.xN: thread `lane` of warp w gets N consecutive columns of TMEM lane 32*(w%4) + lane
// The destination registers are only defined after tcgen05.wait::ld: until then they hold poison.
inline void tmem_ld(uint32_t taddr, uint32_t* out, unsigned n) {
  const unsigned lane = threadIdx_lane_for_tmem();
  Fiber& f = self();
  for (unsigned j = 0; j < n; ++j) {
    f.tmem_pending.emplace_back(out + j, tmem_at(taddr, lane, j));
    out[j] = 0x7fc0dead;   // a NaN
  }
}
inline void tmem_ld_wait() {
  Fiber& f = self();
  for (auto& pv : f.tmem_pending) *pv.first = pv.second;
  f.tmem_pending.clear();
}

}  // namespace emu

// CUDA built-in index variables: refreshed by the scheduler before a fiber resumes
inline uint3_emu threadIdx, blockIdx, blockDim, gridDim;

namespace emu {

inline void fiber_entry() {
  State& s = st();
  s.launches[s.ctas[s.fibers[s.current].cta].launch].body();
  s.fibers[s.current].done = true;
  swapcontext(&s.fibers[s.current].ctx, &s.sched);
}

// Runs the fibers of the given (launch, CTA index) pairs to completion under one scheduler.
inline void run_ctas(const std::vector<std::pair<int, unsigned>>& which) {
  State& s = st();
  const unsigned count = (unsigned)which.size();
  s.ctas.assign(count, Cta());
  unsigned n = 0;
  for (unsigned c = 0; c < count; ++c) {
    const LaunchRec& L = s.launches[which[c].first];
    Cta& k = s.ctas[c];
    const unsigned nwarps = (L.block + 31) / 32;
    k.nthreads = L.block;
    k.index = which[c].second;
    k.launch = which[c].first;
    k.warp_arrived.assign(nwarps, 0);
    k.warp_gen.assign(nwarps, 0);
    k.warp_buf.assign((size_t)nwarps * 32 * 8, 0);
    k.smem_bytes = L.smem_bytes;
    // poison: reads of unwritten shared memory become visible; the tail is a guard zone checked after the CTA has run
    k.dyn_smem.assign(L.smem_bytes + kGuardBytes, (char)0xCD);
    n += L.block;
  }
  for (Fiber& f : s.fibers) std::free(f.stack);
  s.fibers.clear();
  s.fibers.resize(n);
  unsigned i = 0;
  for (unsigned c = 0; c < count; ++c)
    for (unsigned t = 0; t < s.ctas[c].nthreads; ++t, ++i) {
      Fiber& f = s.fibers[i];
      f.cta = c;
      f.tid = t;
      f.stack = static_cast<char*>(std::malloc(kStackBytes));  // untouched pages stay uncommitted
      getcontext(&f.ctx);
      f.ctx.uc_stack.ss_sp = f.stack;
      f.ctx.uc_stack.ss_size = kStackBytes;
      f.ctx.uc_link = &s.sched;
      makecontext(&f.ctx, (void (*)())fiber_entry, 0);
    }
  unsigned remaining = n;
  std::vector<unsigned> order(n);
  for (unsigned q = 0; q < n; ++q) order[q] = q;
  uint64_t rng = s.rng ? (s.rng * 0x9E3779B97F4A7C15ull + which[0].second + 1) : 0;
  uint64_t rounds = 0;
  while (remaining) {
    static const uint64_t kRoundScale = std::getenv("XDR_EMU_PATIENCE") ? std::strtoull(std::getenv("XDR_EMU_PATIENCE"), nullptr, 10) : 1;
    if (++rounds > kRoundScale * (200000000ull / (n ? n : 1) + 100000ull)) {
      std::fprintf(stderr, "cuda_emu: no fiber finished for too long -- deadlock in the emulated kernel?\n");
      if (std::getenv("XDR_EMU_REPORT"))   // lane 0 of every unfinished warp: what it last waited for
        for (const Fiber& f : s.fibers)
          if (!f.done)
            std::fprintf(stderr, "  cta %u warp %u: %s %+ld\n", f.cta, f.tid >> 5, f.site,
                         f.site_arg ? (long)((const char*)f.site_arg - (const char*)s.ctas[f.cta].dyn_smem.data()) : 0l);
      std::abort();
    }
    if (rng) {  // seeded Fisher-Yates reshuffle of the resume order each round
      for (unsigned q = n - 1; q > 0; --q) {
        rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
        std::swap(order[q], order[rng % (q + 1)]);
      }
    }
    for (unsigned q = 0; q < n; ++q) {
      Fiber& f = s.fibers[order[q]];
      if (f.done) continue;
      s.current = (int)order[q];
      const Cta& k = s.ctas[f.cta];
      const LaunchRec& L = s.launches[k.launch];
      const unsigned b = k.index;
      threadIdx = {f.tid, 0, 0};
      blockIdx = {b % L.grid.x, (b / L.grid.x) % L.grid.y, b / (L.grid.x * L.grid.y)};
      blockDim = {L.block, 1, 1};
      gridDim = {L.grid.x, L.grid.y, L.grid.z};
      swapcontext(&s.sched, &f.ctx);
      if (f.done) { --remaining; rounds = 0; }
      tc_tick(s.ctas, rng);
    }
  }
  s.current = -1;
  for (Cta& k : s.ctas) {    // copies / MMAs nobody waited for still complete on the hardware
    while (!k.dma_queue.empty()) {
      auto op = std::move(k.dma_queue.front());
      k.dma_queue.pop_front();
      op(k);
    }
    while (!k.tc_queue.empty()) {
      auto op = std::move(k.tc_queue.front());
      k.tc_queue.pop_front();
      op(k);
    }
  }
  for (const Cta& k : s.ctas)
    for (size_t g = k.smem_bytes; g < k.smem_bytes + kGuardBytes; ++g)
      if (k.dyn_smem[g] != (char)0xCD) {
        std::fprintf(stderr, "cuda_emu: CTA %u wrote %zu bytes past its %zu bytes of dynamic shared memory\n", k.index,
                     g - k.smem_bytes + 1, k.smem_bytes);
        std::abort();
      }
}

// Runs `body` once per thread of every CTA of the grid (1-D blocks only).  concurrent == false: the CTAs execute one after
// another (x fastest) -- `static` stand-ins for __shared__ variables are then private to the running CTA.  concurrent ==
// true: all CTAs are resident at once, as a persistent kernel with grid-wide hand-offs needs (such kernels must keep their
// CTA-local state in dynamic shared memory).  Between group_begin() and group_run(), concurrent launches are only QUEUED
// (the body must own its arguments) and then run all together: several "GPUs" whose persistent kernels talk to each other
// through peer memory.
inline void launch(dim3 grid, dim3 block3, size_t smem_bytes, std::function<void()> body, bool concurrent = false) {
  State& s = st();
  if (block3.y != 1 || block3.z != 1) { std::fprintf(stderr, "cuda_emu: only 1-D blocks are emulated\n"); std::abort(); }
  LaunchRec rec;
  rec.grid = grid;
  rec.block = block3.x;
  rec.smem_bytes = smem_bytes;
  rec.body = std::move(body);
  const unsigned n_ctas = grid.x * grid.y * grid.z;
  if (concurrent && s.grouping) {
    s.queued.push_back(std::move(rec));
    return;
  }
  s.launches.assign(1, std::move(rec));
  std::vector<std::pair<int, unsigned>> which;
  if (concurrent) {
    for (unsigned b = 0; b < n_ctas; ++b) which.emplace_back(0, b);
    run_ctas(which);
  } else {
    for (unsigned b = 0; b < n_ctas; ++b) {
      which.assign(1, std::make_pair(0, b));
      run_ctas(which);
    }
  }
}

inline void group_begin() {
  st().grouping = true;
  st().queued.clear();
}
inline void group_run() {
  State& s = st();
  s.grouping = false;
  s.launches = std::move(s.queued);
  s.queued.clear();
  std::vector<std::pair<int, unsigned>> which;
  for (int l = 0; l < (int)s.launches.size(); ++l) {
    const dim3& g = s.launches[l].grid;
    for (unsigned b = 0; b < g.x * g.y * g.z; ++b) which.emplace_back(l, b);
  }
  if (!which.empty()) run_ctas(which);
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
  if (lane == 0) ++counters().mma_tf32;
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

// mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 over the 32 fibers of the calling warp (two bf16 per register)
inline void mma_m16n8k16_bf16(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  const unsigned lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  if (lane == 0) ++counters().mma_bf16;
  uint32_t* mine = warp_slot(lane);
  for (int i = 0; i < 4; ++i) mine[i] = a[i];
  mine[4] = b[0];
  mine[5] = b[1];
  warp_barrier();
  auto half = [](uint32_t reg, unsigned hi) { return as_float((hi ? (reg >> 16) : (reg & 0xffffu)) << 16); };
  auto A = [&](unsigned row, unsigned k) {  // a0=(g,2t..) a1=(g+8,2t..) a2=(g,2t+8..) a3=(g+8,2t+8..)
    const unsigned kk = k & 7, src = (row & 7) * 4 + (kk >> 1), reg = (row >= 8 ? 1 : 0) + (k >= 8 ? 2 : 0);
    return half(warp_slot(src)[reg], kk & 1);
  };
  auto B = [&](unsigned k, unsigned n) {  // b0=(k=2t..2t+1,n=g) b1=(k=2t+8..,n=g)
    const unsigned kk = k & 7, src = n * 4 + (kk >> 1), reg = 4 + (k >= 8 ? 1 : 0);
    return half(warp_slot(src)[reg], kk & 1);
  };
  float r[4];
  const unsigned rows[4] = {g, g, g + 8, g + 8}, cols[4] = {2 * t, 2 * t + 1, 2 * t, 2 * t + 1};
  for (int i = 0; i < 4; ++i) {
    double s = c[i];
    for (unsigned k = 0; k < 16; ++k) s += (double)A(rows[i], k) * (double)B(k, cols[i]);
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
inline void __nanosleep(unsigned) { emu::yield(); }
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
inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0u; }
using std::max;
using std::min;
