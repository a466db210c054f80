// steps_persistent.cu -- the trainer-step hot loop as ONE persistent, software-pipelined launch over K batches.
//
// Replaces K iterations of recbole Trainer._train_epoch's inner loop [recbole-1.0.1] around
// EMCDR.calculate_source_loss/target_loss (reference emcdr.py:110-154; CMF per-domain term cmf.py:75-98):
//     for batch in batches: loss = calculate_loss(batch); loss.backward()
// i.e. per batch: gather -> dot score -> loss (+EmbLoss) -> row gradients -> scatter-add, with the per-batch
// losses written to out8[k].  Results per batch are identical to xdr_bpr_fwd + xdr_bpr_bwd on that batch
// (gradient accumulation over the K batches when dst is a gradient table; asynchronous SGD with bounded
// staleness (a few steps) when dst is the weight table and scale = -lr).
//
// Why: one B=8192 step moves 12.8 MB = 1.95 us at the HBM roofline, while the dependent chain
// ids -> rows -> grid-wide EmbLoss norm -> atomics is ~4 us of latency and a launch ~2 us.  So one launch
// covers K steps and the chain is overlapped across steps (profiles/r1_*: v1 of this kernel staged rows in
// shared memory and spent 60% of its time in CTA barriers and stage waits; this is v2).
//
//   grid      one CTA per SM (persistent); CTA c owns interactions [c*S, (c+1)*S) of EVERY step
//   warp 0    producer: TMA bulk copies (cp.async.bulk -> UBLKCP) of the CTA's id tiles (and labels) of step t into
//             an 8-deep shared-memory ring, several steps ahead of everyone else
//   warp 1    publisher: per step, sums the CTA's task partials (fixed order) and publishes them to global memory as
//             8-byte {value, step-tag} words (no fences: data and flag travel in one atomic store)
//   warps 2,3 gatherers (even / odd steps): the step's root CTA (rotating, s mod grid) polls all CTAs' words (one
//             batch of loads per poll round), reduces in fp64 in a fixed order and republishes two result words;
//             every other CTA polls just those; the norm factors reach the scatter side through shared memory
//   STAGED kernel (the default; profiles/r1_steps_trace.md shows why gather and scatter must be decoupled):
//     loaders   a task = 32/LPR interactions of one step; LPR lanes own one interaction.  Rows are gathered with
//               LDG.128 (L1-bypassing) into registers, two tasks in flight per warp; score / loss term / squared
//               norms by shuffles; then rows + scores are stashed in a 4-deep shared-memory stage ring.  Loaders
//               never wait for norms, so the gather stream (and DRAM) stays busy during the norm exchange.
//     scatterers  when a step's norm factors arrive: rows from the stage ring, row gradients, REDG.E.ADD.F32x4.
//   REGISTER kernel (fallback when 3 stages of a CTA slice do not fit shared memory, e.g. dim 128 at B=8192):
//     workers keep the rows in registers from gather to scatter (two tasks per warp) and wait for the norms.
//   All hand-offs are mbarriers (ids landed / partial written / norms ready / slots free); no CTA-wide barrier and
//   no grid-wide barrier anywhere.
//
// HBM roofline: 1560 algorithmic bytes per BPR interaction at dim 64 (ids + 3 rows gathered + 3 rows scattered).
#include <string.h>
#include <mutex>
#include <vector>
#include "xdr_common.cuh"

namespace xdr {

constexpr int kServiceWarps = 4;   // producer, publisher, two gatherers
constexpr int kWorkerWarps = 20;   // register kernel: workers
#ifndef XDR_LOADER_WARPS
#define XDR_LOADER_WARPS 14
#endif
#ifndef XDR_SCATTER_WARPS
#define XDR_SCATTER_WARPS 6
#endif
constexpr int kLoaderWarpsFull = XDR_LOADER_WARPS;  // staged kernel: loaders when the CTA owns the whole SM ...
constexpr int kLoaderWarpsLite = 10;                // ... and when item rows are pre-staged by the peer-gather kernel, whose
                                                    // CTAs must fit next to this one (register file: 576 x 80 + 256 x 40)
constexpr int kScatterWarps = XDR_SCATTER_WARPS;    // scatterers
constexpr int kRegThreads = (kServiceWarps + kWorkerWarps) * 32;
__host__ __device__ constexpr int staged_threads(int loaders) { return (kServiceWarps + loaders + kScatterWarps) * 32; }
// register cap of the staged kernel: 768 threads -> 80 registers; so that in the lite configuration a 256-thread peer-gather
// CTA still fits in the SM's register file next to this CTA
constexpr int kStagedBound = staged_threads(kLoaderWarpsFull) > 768 ? staged_threads(kLoaderWarpsFull) : 768;
constexpr int kMaxCtaPerLane = 5;  // gatherer lanes poll <= 5 CTAs each: grid <= 160
constexpr int kMaxStages = 4;      // staged kernel: stage ring depth (3 or 4)
constexpr int kRing = 8;            // id-tile / partial / norm ring depth (steps)
constexpr int kHotMax = 64;        // hot rows (both tables together) whose gradients a CTA accumulates in shared memory
constexpr int kHotTable = 128;     // open-addressing table over them (power of two, load factor <= 1/2)
constexpr int kHotBytes = 16 * 1024;  // accumulators: n_hot * row_f * 4 bytes must fit

struct StepsArgs {
  Shards user_tab, item_tab;  // gather sources
  Shards user_dst, item_dst;  // scatter-add destinations
  int log2g;
  const float* stage_a;       // optional dense pre-gathered first-item rows  [n_steps][batch][dim] (row = position in batch)
  const float* stage_b;       // optional dense pre-gathered second-item rows (pairwise)
  int64_t n_users, n_items;   // GLOBAL row counts
  int nv;                 // float4 per row
  const int64_t* user;    // [n_steps] arrays of `batch` ids, consecutive steps `step_stride` elements apart
  const int64_t* item_a;
  const int64_t* item_b;  // pairwise only
  const float* label;     // pointwise only, same stride
  int64_t step_stride;
  int64_t batch;
  int n_steps;
  int loss_kind;
  float gamma, reg_weight;
  float* out8;  // [n_steps, 8]
  const float* grad_loss;
  float scale;
  unsigned long long* words;  // [n_steps][gridDim.x][3] partials then [n_steps][2] results, {fp32, step tag}
  unsigned int tag_base;      // step s carries tag tag_base + s + 1: tags grow from launch to launch on one workspace, so a
                              // word left by an earlier launch can never match and the workspace needs no per-launch zeroing
  int slice;                  // S: interactions per CTA per step (multiple of 4)
  // Lazily zeroed gradient tables (optional, single GPU, staged kernel): 2 bits per destination row, 16 rows per word --
  // bit 0 "claimed" (some CTA's filler warp owns the zero-fill of the row), bit 1 "filled" (the zeros are in L2; scatter-adds
  // may follow).  A row whose bits are clear counts as zero whatever it holds: its first touch stores a full row of zeros
  // (full-line writes allocate in L2 without a DRAM read) and the REDs then hit L2 -- no read-modify-write of gradient lines.
  unsigned int* touch_u;
  unsigned int* touch_i;
  // Hot rows (optional, staged kernel, plain scatter-add mode): rows the data names so often (popular items of a Zipf-like
  // catalogue) that their REDs serialise in one L2 slice.  Every CTA adds its contributions to such a row in SHARED memory
  // over the whole launch and flushes them once at the end (one RED per CTA and row instead of one per interaction).
  const int64_t* hot_u;
  const int64_t* hot_i;
  int n_hot_u, n_hot_i;
  size_t hot_off;   // byte offset of the hot region in dynamic shared memory (0 = none)
  int32_t* oob;
  unsigned long long* trace;  // optional [n_steps][grid][8] globaltimer stamps (debug; NULL in production)
};

// ---- PTX wrappers -------------------------------------------------------------------------------------------------
#ifdef XDR_EMU
// CPU CTA emulator (tests/emu): the same eight primitives on the emulator's mbarrier / bulk-copy model; polls yield.
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { emu::mbar_init(bar, count); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { emu::mbar_expect_tx(bar, bytes); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { emu::mbar_arrive(bar); }
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) { return emu::mbar_test(bar, parity); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { emu::mbar_wait(bar, parity); }
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  emu::bulk_g2s(dst_smem, src_gmem, bytes, bar);
}
__device__ __forceinline__ unsigned long long gtime() { return ++emu::st().clock; }
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  emu::note_site("poll exchange word");
  emu::yield();  // every poll gives the other fibers a turn
  return *p;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) { *p = v; }
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  emu::note_site("poll touch word");
  emu::yield();
  return *p;
}
__device__ __forceinline__ void fence_acq_rel_gpu() {}
#else
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t a = smem_u32(bar);
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  }
}
// TMA 1-D bulk copy global -> shared, completion (bytes) on an mbarrier.  16-byte aligned src/dst/size.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// Strong (L2-coherent) load of a touch-map word.  relaxed, not acquire: an acquire load is followed by an L1 invalidation
// (CCTL.IVALL) per use; what the scatter warps need is only that their REDs -- issued after, and control-dependent on, the
// observed "filled" bit -- reach L2 after the zero row did, and the filler set the bit only after the row was complete in L2.
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
#endif  // XDR_EMU
// coherent (L2) 128-bit row load: the table may be the scatter destination of this very launch (fused SGD)
__device__ __forceinline__ float4 ldcg_row4(const float* row, int col4) {
  return __ldcg(reinterpret_cast<const float4*>(row) + col4);
}

// ---- shared-memory layout -------------------------------------------------------------------------------------------
struct SmemLayout {
  int slice, rows_per, tasks, row_f, stages;  // S, 2|3, tasks per step per CTA (upper bound), floats per row, stage count
  __host__ __device__ SmemLayout(int s, int r, int t, int rf = 0, int ns = 0)
      : slice(s), rows_per(r), tasks(t), row_f(rf), stages(ns) {}
  // [0, 320): 4 x kRing + kMaxStages mbarriers.  [320, 384): norms [kRing][2].  then partials, id ring, stage ring.
  __host__ __device__ size_t bars_off() const { return 0; }
  __host__ __device__ size_t norms_off() const { return 320; }
  __host__ __device__ size_t part_off() const { return 384; }
  __host__ __device__ size_t ids_off() const {
    return (part_off() + (size_t)kRing * tasks * sizeof(float4) + 127) & ~(size_t)127;
  }
  __host__ __device__ size_t ids_slot_bytes() const {  // ids [rows_per][S] int64 + labels [S] fp32
    return (((size_t)rows_per * slice * 8 + (size_t)slice * 4) + 127) & ~(size_t)127;
  }
  __host__ __device__ size_t stage_off() const { return ids_off() + (size_t)kRing * ids_slot_bytes(); }
  __host__ __device__ size_t stage_slot_bytes() const {  // rows [rows_per][S][row_f] fp32 + scores [2][S] fp32
    return (((size_t)rows_per * slice * row_f * 4 + (size_t)2 * slice * 4) + 127) & ~(size_t)127;
  }
  __host__ __device__ size_t bytes() const { return stage_off() + (size_t)stages * stage_slot_bytes(); }
};

struct Bars {
  uint64_t *idsf, *ifree, *adone, *normf, *sfree;
  __device__ explicit Bars(unsigned char* smem) {
    idsf = reinterpret_cast<uint64_t*>(smem);  // id tile landed        (tx barrier)
    ifree = idsf + kRing;                      // id slot free          (count = consumers of the ids per step)
    adone = ifree + kRing;                     // task partials written (count = tasks)
    normf = adone + kRing;                     // norms ready           (count = 1)
    sfree = normf + kRing;                     // stage slot free       (count = scatter warps)   [kMaxStages]
  }
};

// One task's registers: ids, the rows (VEC float4 columns per lane) and the scores.
template <int VEC, bool PAIRWISE>
struct TaskRegs {
  float4 u[VEC], a[VEC], b[PAIRWISE ? VEC : 1];
  int iu, ia, ib;  // row ids (tables have < 2^31 rows); -1 = padding lane or out-of-range id (row reads as zeros)
  float sa, sb, label;
  int s, q;
  // lazily zeroed destination tables: lanes 0..R-1 of an interaction's group hold the claim of its user / item+ / item- row
  unsigned int old;  // the row's touch-map word as the claim found it (see lazy_claim_of for the word and the bit)
};

// the touch-map word and "claimed" bit of the row lane `sub` of a group looks after (nullptr: none -- lane >= R or a bad id)
template <int VEC, bool PAIRWISE>
__device__ __forceinline__ unsigned int* lazy_claim_of(const TaskRegs<VEC, PAIRWISE>& r, const StepsArgs& a, int sub,
                                                       unsigned int& cbit) {
  const int id = sub == 0 ? r.iu : (sub == 1 ? r.ia : ((PAIRWISE && sub == 2) ? r.ib : -1));
  cbit = 1u << (((unsigned int)id & 15u) * 2u);
  return id < 0 ? nullptr : (sub == 0 ? a.touch_u : a.touch_i) + (id >> 4);
}

// ---- service warps (shared by both kernels) ---------------------------------------------------------------------------
template <bool PAIRWISE>
__device__ __forceinline__ void service_producer(const StepsArgs& a, const SmemLayout& L, const Bars& B,
                                                 unsigned char* ids_ring, int64_t first, int cnt, int lane) {
  constexpr int R = PAIRWISE ? 3 : 2;
  if (lane != 0) return;
  const uint32_t idb = (uint32_t)cnt * 8u, lbb = (uint32_t)cnt * 4u;  // cnt % 4 == 0 -> multiples of 16 bytes
  const bool has_label = !PAIRWISE && a.label != nullptr;
  for (int t = 0; t < a.n_steps; ++t) {
    const int slot = t % kRing;
    const uint32_t par = (uint32_t)((t / kRing) & 1);
    int64_t* ids = reinterpret_cast<int64_t*>(ids_ring + (size_t)slot * L.ids_slot_bytes());
    float* lab = reinterpret_cast<float*>(ids + (size_t)R * L.slice);
    mbar_wait(&B.ifree[slot], par ^ 1u);  // passes immediately the first time round the ring
    const int64_t off = (int64_t)t * a.step_stride + first;
    mbar_expect_tx(&B.idsf[slot], idb * R + (has_label ? lbb : 0u));
    bulk_g2s(ids, a.user + off, idb, &B.idsf[slot]);
    bulk_g2s(ids + L.slice, a.item_a + off, idb, &B.idsf[slot]);
    if (PAIRWISE) bulk_g2s(ids + 2 * L.slice, a.item_b + off, idb, &B.idsf[slot]);
    if (has_label) bulk_g2s(lab, a.label + off, lbb, &B.idsf[slot]);
  }
}

// per step: sum this CTA's task partials in a fixed order and publish them as three 8-byte {value, tag} words
__device__ __forceinline__ void service_publisher(const StepsArgs& a, const SmemLayout& L, const Bars& B,
                                                  const float4* part, int tasks, int lane) {
  const unsigned int n_cta = gridDim.x;
  for (int s = 0; s < a.n_steps; ++s) {
    const int slot = s % kRing;
    mbar_wait(&B.adone[slot], (uint32_t)((s / kRing) & 1));
    if (a.trace && lane == 0) a.trace[((size_t)s * n_cta + blockIdx.x) * 8 + 0] = gtime();
    float p0 = 0.f, p1 = 0.f, p2 = 0.f;
    for (int q = lane; q < tasks; q += 32) {
      const float4 v = part[slot * L.tasks + q];
      p0 += v.x;
      p1 += v.y;
      p2 += v.z;
    }
    p0 = warp_sum(p0);
    p1 = warp_sum(p1);
    p2 = warp_sum(p2);
    if (lane < 3) {
      const float v = lane == 0 ? p0 : (lane == 1 ? p1 : p2);
      st_relaxed_u64(a.words + ((size_t)s * n_cta + blockIdx.x) * 3 + lane,
                     ((unsigned long long)(a.tag_base + (unsigned int)(s + 1)) << 32) | (unsigned long long)__float_as_uint(v));
    }
    __syncwarp();
  }
}

// (Round 2, call 18: an "all-read" form -- every CTA polls all CTAs' words itself, no republished result -- was measured and
// is SLOWER: 148 CTAs polling 444 words each through a memory system the gather keeps saturated stretched the exchange from
// 3.6 to 4.4 us and the K = 20 step from 3.52 to 3.58 us.  The rotating-root form stays.)
// steps s = which, which+2, ...: root CTA (s mod grid) gathers + reduces + republishes; the others poll the result
__device__ __forceinline__ void service_gatherer(const StepsArgs& a, const Bars& B, float2* norms, int which, int lane) {
  const unsigned int n_cta = gridDim.x;
  const float inv_b = 1.0f / (float)a.batch;
  const float g = (a.grad_loss ? __ldg(a.grad_loss) : 1.0f) * a.scale;
  unsigned long long* finals = a.words + (size_t)a.n_steps * n_cta * 3;  // [n_steps][2] {factor, tag}
  for (int s = which; s < a.n_steps; s += 2) {
    const int slot = s % kRing;
    const unsigned int tag = a.tag_base + (unsigned int)(s + 1);
    float cu = 0.f, ci = 0.f;
    if ((unsigned int)s % n_cta == blockIdx.x) {
      const unsigned long long* base = a.words + (size_t)s * n_cta * 3;
      unsigned long long wv[kMaxCtaPerLane][3];
      bool all_ok;
      do {
#pragma unroll
        for (int i = 0; i < kMaxCtaPerLane; ++i) {
          const unsigned int c = lane + 32u * i;
          if (c < n_cta) {
#pragma unroll
            for (int k = 0; k < 3; ++k) wv[i][k] = ld_relaxed_u64(base + (size_t)c * 3 + k);
          }
        }
        all_ok = true;
#pragma unroll
        for (int i = 0; i < kMaxCtaPerLane; ++i) {
          const unsigned int c = lane + 32u * i;
          if (c < n_cta) {
#pragma unroll
            for (int k = 0; k < 3; ++k) all_ok = all_ok && ((unsigned int)(wv[i][k] >> 32) == tag);
          }
        }
        all_ok = __all_sync(0xffffffffu, all_ok);
      } while (!all_ok);
      double t0 = 0.0, t1 = 0.0, t2 = 0.0;
#pragma unroll
      for (int i = 0; i < kMaxCtaPerLane; ++i) {
        const unsigned int c = lane + 32u * i;
        if (c < n_cta) {
          t0 += (double)__uint_as_float((unsigned int)wv[i][0]);
          t1 += (double)__uint_as_float((unsigned int)wv[i][1]);
          t2 += (double)__uint_as_float((unsigned int)wv[i][2]);
        }
      }
      t0 = warp_sum(t0);
      t1 = warp_sum(t1);
      t2 = warp_sum(t2);
      const float nu = (float)sqrt(t1), ni = (float)sqrt(t2);
      // d(reg_weight * (||U||_F + ||I||_F)/B)/dU_r = reg_weight/(B*||U||_F) * U_r   (torch: 0 when the norm is 0)
      cu = (a.reg_weight != 0.f && nu > 0.f) ? g * a.reg_weight * inv_b / nu : 0.f;
      ci = (a.reg_weight != 0.f && ni > 0.f) ? g * a.reg_weight * inv_b / ni : 0.f;
      if (lane < 2)
        st_relaxed_u64(finals + (size_t)s * 2 + lane,
                       ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(lane == 0 ? cu : ci));
      if (lane == 0) {
        const float data = (float)(t0 / (double)a.batch);
        const float reg = (float)(((double)nu + (double)ni) / (double)a.batch);
        float* o = a.out8 + (size_t)s * 8;
        o[0] = data + a.reg_weight * reg;
        o[1] = data;
        o[2] = nu;
        o[3] = ni;
        o[4] = reg;
        o[5] = 0.f;
        o[6] = 0.f;
        o[7] = 0.f;
      }
    } else {
      unsigned long long f0 = 0, f1 = 0;
      if (lane == 0) {
        const unsigned long long* f = finals + (size_t)s * 2;
        for (;;) {
          f0 = ld_relaxed_u64(f);
          f1 = ld_relaxed_u64(f + 1);
          if ((unsigned int)(f0 >> 32) == tag && (unsigned int)(f1 >> 32) == tag) break;
          __nanosleep(64);
        }
      }
      cu = __uint_as_float((unsigned int)__shfl_sync(0xffffffffu, f0, 0));
      ci = __uint_as_float((unsigned int)__shfl_sync(0xffffffffu, f1, 0));
    }
    if (a.trace && lane == 0) a.trace[((size_t)s * n_cta + blockIdx.x) * 8 + 1] = gtime();
    if (lane == 0) {
      norms[slot] = make_float2(cu, ci);
      mbar_arrive(&B.normf[slot]);  // release: the norms are visible to everyone who observes this phase
    }
    __syncwarp();
  }
}

// ---- task primitives (shared) ------------------------------------------------------------------------------------------
// request the rows of CTA-local task `lt` (ids from the shared-memory ring)
template <int LPR, int VEC, bool PAIRWISE, bool LAZY = false>
__device__ __forceinline__ void task_issue(TaskRegs<VEC, PAIRWISE>& r, int lt, const StepsArgs& a, const SmemLayout& L,
                                           const Bars& B, const unsigned char* ids_ring, int tasks, int cnt, int lane) {
  constexpr int IPW = 32 / LPR;
  constexpr int R = PAIRWISE ? 3 : 2;
  const int sub = lane % LPR, grp = lane / LPR;
  const int row_f = a.nv * 4;
  r.s = lt / tasks;
  r.q = lt - r.s * tasks;
  const int slot = r.s % kRing;
  mbar_wait(&B.idsf[slot], (uint32_t)((r.s / kRing) & 1));
  if (a.trace && lane == 0 && r.q == 0) a.trace[((size_t)r.s * gridDim.x + blockIdx.x) * 8 + 2] = gtime();
  const int64_t* ids = reinterpret_cast<const int64_t*>(ids_ring + (size_t)slot * L.ids_slot_bytes());
  const float* lab = reinterpret_cast<const float*>(ids + (size_t)R * L.slice);
  const int j = r.q * IPW + grp;
  const bool live = j < cnt;
  const int jj = live ? j : 0;
  const int64_t iu = ids[jj], ia = ids[L.slice + jj], ib = PAIRWISE ? ids[2 * L.slice + jj] : 0;
  r.label = (!PAIRWISE && a.label != nullptr) ? lab[jj] : 0.f;
  const bool oku = live && (uint64_t)iu < (uint64_t)a.n_users;
  const bool oka = live && (uint64_t)ia < (uint64_t)a.n_items;
  const bool okb = PAIRWISE && live && (uint64_t)ib < (uint64_t)a.n_items;
  if (live && a.oob && sub == 0 && (!oku || !oka || (PAIRWISE && !okb))) *a.oob = 1;
  r.iu = oku ? (int)iu : -1;
  r.ia = oka ? (int)ia : -1;
  r.ib = okb ? (int)ib : -1;
  r.sb = live ? 1.f : 0.f;  // until task_score overwrites it with the negative score: "this lane owns an interaction"
  const float* pu = shard_row(a.user_tab, a.log2g, oku ? iu : 0, row_f);
  const float* pa = shard_row(a.item_tab, a.log2g, oka ? ia : 0, row_f);
  const float* pb = shard_row(a.item_tab, a.log2g, okb ? ib : 0, row_f);
  if (a.stage_a != nullptr) {
    // item rows were pulled over NVLink into a dense local block by the peer-gather kernel one chunk ahead: sequential,
    // local reads here; the ids are still needed for the scatter side
    const int64_t pos = ((int64_t)r.s * a.batch + (int64_t)blockIdx.x * a.slice + jj) * row_f;
    pa = a.stage_a + pos;
    if (PAIRWISE) pb = a.stage_b + pos;
  }
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    const int c = sub + v * LPR;
    const bool on = c < a.nv;
    r.u[v] = (oku && on) ? ldcg_row4(pu, c) : z4;
    r.a[v] = (oka && on) ? ldcg_row4(pa, c) : z4;
    if (PAIRWISE) r.b[v] = (okb && on) ? ldcg_row4(pb, c) : z4;
  }
  if (LAZY) {
    // claim the task's destination rows (bit 0 of the row's pair in the touch map); the answers travel with the row loads
    // and are consumed by task_resolve
    unsigned int cbit;
    unsigned int* wp = lazy_claim_of(r, a, sub, cbit);
    r.old = wp ? atomicOr(wp, cbit) : 0u;
  }
}

// Lazily zeroed destination tables, loader side ("claim and fill at gather time").  A row's claim is requested together
// with its gather (task_issue), by the loader that gathers it:
//   won        -> nobody has touched the row since the map was cleared: the loader stores a row of ZEROS (whole 128-byte
//                 lines: L2 allocates them without reading DRAM) and, after its fence, sets the row's "filled" bit;
//   lost, filled  -> nothing to do: the zeros (or earlier sums) are in L2;
//   lost, not yet filled -> the winner -- another loader, a few microseconds ahead -- is between its claim and its
//                 publication: wait for the bit here.
// When a window is resolved every destination row of its tasks is ready for plain REDs, so the scatter warps are exactly
// those of the plain kernel.  A loader resolves the claims of a WINDOW of two tasks at once (one claim round trip and one
// fence per window).  Deadlock freedom: between a claim and the publication of its "filled" bit a loader waits for memory
// only, never for a barrier (the window waits for both id tiles first, issues, resolves, and only then turns to the stage
// ring), and it publishes before it polls.
// Measured (round 2, calls 18-23, profiles/r2_lazy_tables.md): 4.5 us/step at K = 200 against 3.0 for plain scatter-add --
// under load every dependent memory round trip (claim answers, fence) costs 2.5-4 us on a warp that then gathers nothing.
// Variants that moved the fence to a publisher warp, kept the rolling two-task pipeline with guarded barrier waits, or let
// the scatter warps wait for the bits were all slower or equal; this is the simplest of the family.  Opt-in.
template <int LPR, int VEC, bool PAIRWISE>
__device__ __forceinline__ bool lazy_fill(const TaskRegs<VEC, PAIRWISE>& r, const StepsArgs& a, int lane) {
  const int sub = lane % LPR, grp = lane / LPR;
  const int64_t row_f = (int64_t)a.nv * 4;
  unsigned int cbit;
  const bool won = lazy_claim_of(r, a, sub, cbit) != nullptr && (r.old & cbit) == 0u;
  const unsigned int won_mask = __ballot_sync(0xffffffffu, won);
  if (won_mask != 0u) {
    const unsigned int gw = won_mask >> (grp * LPR);   // bits 0..2: the group's user / item+ / item- row was won
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float* du = (gw & 1u) ? shard_row(a.user_dst, 0, r.iu, row_f) : nullptr;
    float* da = (gw & 2u) ? shard_row(a.item_dst, 0, r.ia, row_f) : nullptr;
    float* db = (PAIRWISE && (gw & 4u)) ? shard_row(a.item_dst, 0, r.ib, row_f) : nullptr;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int c = sub + v * LPR;
      if (c >= a.nv) continue;
      if (du) st4(du, c, z4);
      if (da) st4(da, c, z4);
      if (db) st4(db, c, z4);
    }
  }
  return won_mask != 0u;
}
template <int VEC, bool PAIRWISE>
__device__ __forceinline__ void lazy_publish(const TaskRegs<VEC, PAIRWISE>& r, const StepsArgs& a, int sub) {
  unsigned int cbit;
  unsigned int* wp = lazy_claim_of(r, a, sub, cbit);
  if (wp != nullptr && (r.old & cbit) == 0u) atomicOr(wp, cbit << 1);
}
template <int VEC, bool PAIRWISE>
__device__ __forceinline__ void lazy_wait_filled(const TaskRegs<VEC, PAIRWISE>& r, const StepsArgs& a, int sub) {
  unsigned int cbit;
  unsigned int* wp = lazy_claim_of(r, a, sub, cbit);
  const bool pending = wp != nullptr && (r.old & cbit) != 0u && (r.old & (cbit << 1)) == 0u;
  if (__any_sync(0xffffffffu, pending)) {
    for (;;) {
      const bool ok = !pending || (ld_acquire_u32(wp) & (cbit << 1)) != 0u;
      if (__all_sync(0xffffffffu, ok)) break;
    }
  }
}
template <int LPR, int VEC, bool PAIRWISE>
__device__ __forceinline__ void lazy_resolve(TaskRegs<VEC, PAIRWISE>& r0, TaskRegs<VEC, PAIRWISE>& r1, bool two,
                                             const StepsArgs& a, int lane) {
  bool any = lazy_fill<LPR, VEC, PAIRWISE>(r0, a, lane);
  if (two) any = lazy_fill<LPR, VEC, PAIRWISE>(r1, a, lane) || any;
  if (any) {
    fence_acq_rel_gpu();   // every lane's zeros are in L2 ...
    __syncwarp();          // ... and every lane is past its fence
    lazy_publish(r0, a, lane % LPR);
    if (two) lazy_publish(r1, a, lane % LPR);
  }
  lazy_wait_filled(r0, a, lane % LPR);
  if (two) lazy_wait_filled(r1, a, lane % LPR);
}

// score, loss term and squared norms of the task (consumes the row loads); returns the task partial in every lane
template <int LPR, int VEC, bool PAIRWISE>
__device__ __forceinline__ float4 task_score(TaskRegs<VEC, PAIRWISE>& r, const StepsArgs& a) {
  float da = 0.f, db = 0.f, uu = 0.f, aa = 0.f;
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    da += dot4(r.u[v], r.a[v]);
    uu += dot4(r.u[v], r.u[v]);
    aa += dot4(r.a[v], r.a[v]);
    if (PAIRWISE) db += dot4(r.u[v], r.b[v]);
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) {
    da += __shfl_xor_sync(0xffffffffu, da, o);
    uu += __shfl_xor_sync(0xffffffffu, uu, o);
    aa += __shfl_xor_sync(0xffffffffu, aa, o);
    if (PAIRWISE) db += __shfl_xor_sync(0xffffffffu, db, o);
  }
  const bool live = r.sb != 0.f;
  r.sa = da;
  r.sb = db;
  float term = 0.f;
  if (live) {
    if (PAIRWISE) {
      term = -logf(a.gamma + sigmoidf_(da - db));
    } else if (a.loss_kind == XDR_LOSS_MSE) {
      const float d = da - r.label;
      term = d * d;
    } else if (a.loss_kind == XDR_LOSS_BCE_SIGMOID) {
      const float p = sigmoidf_(da);
      term = -(r.label * fmaxf(logf(p), -100.f) + (1.f - r.label) * fmaxf(logf(1.f - p), -100.f));
    }
  } else {
    uu = 0.f;
    aa = 0.f;
  }
  // sum over the interactions of the task (fixed shuffle tree -> deterministic)
#pragma unroll
  for (int o = 16; o >= LPR; o >>= 1) {
    term += __shfl_xor_sync(0xffffffffu, term, o);
    uu += __shfl_xor_sync(0xffffffffu, uu, o);
    aa += __shfl_xor_sync(0xffffffffu, aa, o);
  }
  return make_float4(term, uu, aa, 0.f);
}

// g * dL_data/dscore_a for one interaction (BPR: dscore_b = -c)
template <bool PAIRWISE>
__device__ __forceinline__ float score_coeff(const StepsArgs& a, float g, float inv_b, float sa, float sb, float label) {
  if (PAIRWISE) {
    const float sg = sigmoidf_(sa - sb);
    return -g * inv_b * (sg * (1.f - sg)) / (a.gamma + sg);
  } else if (a.loss_kind == XDR_LOSS_MSE) {
    return g * inv_b * 2.f * (sa - label);
  } else if (a.loss_kind == XDR_LOSS_BCE_SIGMOID) {
    const float p = sigmoidf_(sa);
    const float pq = p * (1.f - p);
    return g * inv_b * (p - label) / fmaxf(pq, 1e-12f) * pq;
  }
  return 0.f;
}

// How a gradient row reaches its destination.  kRowAdd: RED (scatter-add).  kRowStore: plain 128-bit stores -- the row's FIRST
// touch since the touch map was cleared (lazily zeroed tables): whatever the destination holds counts as zero and is
// overwritten, whole lines at a time, so L2 allocates them without reading DRAM.  kRowSkip: nothing (bad id, or deferred).
enum RowMode : int { kRowSkip = 0, kRowAdd = 1, kRowStore = 2, kRowHot = 3 };   // kRowHot: add into a shared-memory accumulator row
__device__ __forceinline__ void emit_row4(int mode, float* row, int cidx, float4 v) {
  if (mode == kRowAdd) red_add4(row, cidx, v);
  else if (mode == kRowStore) st4(row, cidx, v);
  else if (mode == kRowHot) {
    float* q = row + 4 * cidx;
    atomicAdd(q, v.x);
    atomicAdd(q + 1, v.y);
    atomicAdd(q + 2, v.z);
    atomicAdd(q + 3, v.w);
  }
}

// hot rows: open-addressing table in shared memory, key = 2 * row id + table (0 user, 1 item); -1 = not a hot row
struct HotRows {
  long long* keys;       // [kHotTable] (-1 = empty)
  int* slots;            // [kHotTable] accumulator index of the key
  long long* slot_key;   // [kHotMax] key of every accumulator (for the flush)
  float* acc;            // [n][row_f]
  __device__ HotRows(unsigned char* base)
      : keys(reinterpret_cast<long long*>(base)), slots(reinterpret_cast<int*>(base + kHotTable * 8)),
        slot_key(reinterpret_cast<long long*>(base + kHotTable * 12)), acc(reinterpret_cast<float*>(base + kHotTable * 12 + kHotMax * 8)) {}
  __device__ __forceinline__ static unsigned int hash(long long key) {
    return ((unsigned int)key * 2654435761u) >> 25;   // 7 bits
  }
  __device__ __forceinline__ int find(int table, int id) const {
    const long long key = 2ll * id + table;
    unsigned int h = hash(key);
    for (;;) {
      const long long k = keys[h];
      if (k == key) return slots[h];
      if (k < 0) return -1;
      h = (h + 1) & (kHotTable - 1);
    }
  }
};
__host__ __device__ inline size_t hot_region_bytes(int n_hot, int row_f) {
  return (size_t)kHotTable * 12 + (size_t)kHotMax * 8 + (size_t)n_hot * row_f * 4;
}

template <bool PAIRWISE>
__device__ __forceinline__ void scatter_cols(const StepsArgs& a, int cidx, float c, float cu, float ci, int iu, int ia,
                                             int ib, float4 ru, float4 ra, float4 rb, int mu = kRowAdd, int ma = kRowAdd,
                                             int mb = kRowAdd, float* hu = nullptr, float* ha = nullptr, float* hb = nullptr) {
  const int64_t row_f = (int64_t)a.nv * 4;
  if (PAIRWISE) {
    if (iu >= 0) emit_row4(mu, hu ? hu : shard_row(a.user_dst, a.log2g, iu, row_f), cidx, axpy4(cu, ru, scale4(c, sub4(ra, rb))));
    if (ia >= 0) emit_row4(ma, ha ? ha : shard_row(a.item_dst, a.log2g, ia, row_f), cidx, axpy4(ci, ra, scale4(c, ru)));
    if (ib >= 0) emit_row4(mb, hb ? hb : shard_row(a.item_dst, a.log2g, ib, row_f), cidx, scale4(-c, ru));
  } else {
    if (iu >= 0) emit_row4(mu, hu ? hu : shard_row(a.user_dst, a.log2g, iu, row_f), cidx, axpy4(cu, ru, scale4(c, ra)));
    if (ia >= 0) emit_row4(ma, ha ? ha : shard_row(a.item_dst, a.log2g, ia, row_f), cidx, axpy4(ci, ra, scale4(c, ru)));
  }
}

__device__ __forceinline__ void init_bars(const Bars& B, int tasks, int ifree_count, int stages, int sfree_count) {
  for (int i = 0; i < kRing; ++i) {
    mbar_init(&B.idsf[i], 1);
    mbar_init(&B.ifree[i], (uint32_t)ifree_count);
    mbar_init(&B.adone[i], (uint32_t)tasks);
    mbar_init(&B.normf[i], 1);
  }
  for (int i = 0; i < stages; ++i) mbar_init(&B.sfree[i], (uint32_t)sfree_count);
#ifndef XDR_EMU
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
}

// =====================================================================================================================
// STAGED kernel: loaders (gather -> score -> stash) and scatterers (norms -> gradients -> RED) are different warps
// =====================================================================================================================
// EARLY (opt-in, reg_weight == 0 only; parity-green on a B200, tests/test_gpu_engines.py): without the EmbLoss term the row gradients do not depend on
// the batch-wide norms, so the scatterers do not wait for the step's norm exchange before they issue the REDs and free the
// stage slot; they still wait for it before they free the id slot, which keeps the id / partial / norm rings in step (the
// per-step loss is still the exchanged batch mean).
template <int LPR, int VEC, bool PAIRWISE, int kLoaderWarps, bool EARLY = false, bool LAZY = false, bool HOT = false>
__global__ void __launch_bounds__(kStagedBound, 1) train_steps_staged_kernel(StepsArgs a, int n_stages) {
  constexpr int IPW = 32 / LPR;
  constexpr int R = PAIRWISE ? 3 : 2;
  XDR_DYN_SMEM_ALIGNED(unsigned char, smem_raw, 128);
  const int64_t first = (int64_t)blockIdx.x * a.slice;
  const int cnt = (int)min((int64_t)a.slice, a.batch - first);  // > 0: the host launches ceil(batch/slice) CTAs
  const int tasks = (cnt + IPW - 1) / IPW;
  const int row_f = a.nv * 4;
  const SmemLayout L(a.slice, R, (a.slice + IPW - 1) / IPW, row_f, n_stages);
  const Bars B(smem_raw + L.bars_off());
  float2* norms = reinterpret_cast<float2*>(smem_raw + L.norms_off());
  float4* part = reinterpret_cast<float4*>(smem_raw + L.part_off());
  unsigned char* ids_ring = smem_raw + L.ids_off();
  unsigned char* stage_ring = smem_raw + L.stage_off();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // hot rows: table + zeroed accumulators in shared memory (built by everybody before the roles split)
  constexpr bool hot_on = HOT && !LAZY;   // (a template parameter: the plain instantiation carries none of this)
  const HotRows hot(smem_raw + a.hot_off);
  uint64_t* hot_done = reinterpret_cast<uint64_t*>(smem_raw + 288);   // [288, 296): every scatter warp is past its last step
  if constexpr (hot_on) {
    const int n_hot = a.n_hot_u + a.n_hot_i;
    for (int i = threadIdx.x; i < kHotTable; i += blockDim.x) hot.keys[i] = -1;
    for (int i = threadIdx.x; i < n_hot * row_f; i += blockDim.x) hot.acc[i] = 0.f;
    __syncthreads();
    if ((int)threadIdx.x < n_hot) {
      const int t = (int)threadIdx.x < a.n_hot_u ? 0 : 1;
      const int64_t id = t == 0 ? a.hot_u[threadIdx.x] : a.hot_i[threadIdx.x - a.n_hot_u];
      long long key = -2;   // ids outside the table (or duplicates in the list) get an accumulator nobody finds
      if ((uint64_t)id < (uint64_t)(t == 0 ? a.n_users : a.n_items)) key = 2ll * id + t;
      hot.slot_key[threadIdx.x] = key;
      if (key >= 0) {
        unsigned int h = HotRows::hash(key);
        for (;;) {
          const long long prev = (long long)atomicCAS(reinterpret_cast<unsigned long long*>(&hot.keys[h]), (unsigned long long)-1ll,
                                                      (unsigned long long)key);
          if (prev == -1ll) { hot.slots[h] = (int)threadIdx.x; break; }
          if (prev == key) { hot.slot_key[threadIdx.x] = -2; break; }   // listed twice: the first entry owns the row
          h = (h + 1) & (kHotTable - 1);
        }
      }
    }
  }
  if (threadIdx.x == 0) {
    init_bars(B, tasks, kScatterWarps, n_stages, kScatterWarps);
    if (hot_on) mbar_init(hot_done, kScatterWarps);
#ifndef XDR_EMU
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
  }
  __syncthreads();

  if (warp == 0) {
    service_producer<PAIRWISE>(a, L, B, ids_ring, first, cnt, lane);
  } else if (warp == 1) {
    service_publisher(a, L, B, part, tasks, lane);
  } else if (warp < kServiceWarps) {
    service_gatherer(a, B, norms, warp - 2, lane);
  } else if (warp < kServiceWarps + kLoaderWarps) {
    // ------------------------------------------------ loaders ---------------------------------------------------------
    // at most 3*tasks loaders are active so that a loader's consecutive tasks are <= 3 steps apart: with two tasks in
    // flight the steps it touches span <= 6 < kRing, so the id ring is never lapped
    const int n_loaders = min(kLoaderWarps, 3 * tasks);
    const int w = warp - kServiceWarps;
    if (w >= n_loaders) return;
    const int sub = lane % LPR, grp = lane / LPR;
    const int total = a.n_steps * tasks;
    using Regs = TaskRegs<VEC, PAIRWISE>;
    auto issue = [&](Regs& r, int lt) { task_issue<LPR, VEC, PAIRWISE, LAZY>(r, lt, a, L, B, ids_ring, tasks, cnt, lane); };
    auto score_and_stash = [&](Regs& r) {
      const float4 p = task_score<LPR, VEC, PAIRWISE>(r, a);
      const int st = r.s % n_stages;
      mbar_wait(&B.sfree[st], (uint32_t)(((r.s / n_stages) & 1) ^ 1));  // first time round the ring: passes at once
      float* rows = reinterpret_cast<float*>(stage_ring + (size_t)st * L.stage_slot_bytes());
      float* sc = rows + (size_t)R * L.slice * row_f;
      const int j = r.q * IPW + grp;
      if (j < cnt) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          const int c = sub + v * LPR;
          if (c >= a.nv) continue;
          st4(rows + (size_t)j * row_f, c, r.u[v]);
          st4(rows + ((size_t)L.slice + j) * row_f, c, r.a[v]);
          if (PAIRWISE) st4(rows + ((size_t)2 * L.slice + j) * row_f, c, r.b[v]);
        }
        if (sub == 0) {
          sc[j] = r.sa;
          sc[L.slice + j] = r.sb;
        }
      }
      __syncwarp();  // every lane's stage writes are ordered before lane 0's release-arrive
      if (lane == 0) {
        part[(r.s % kRing) * L.tasks + r.q] = p;
        mbar_arrive(&B.adone[r.s % kRing]);
        if (a.trace && r.q == 0) a.trace[((size_t)r.s * gridDim.x + blockIdx.x) * 8 + 3] = gtime();
      }
    };
    Regs r0, r1;
    if constexpr (LAZY) {
      // lazily zeroed tables: windows of two tasks -- wait for both id tiles, gather + claim both, resolve the claims (zero
      // fill of first-touch rows, fence, "filled" bits), then score and stash both (see lazy_resolve)
      for (int lt = w; lt < total; lt += 2 * n_loaders) {
        const int l1 = lt + n_loaders;
        const bool two = l1 < total;
        { const int s0 = lt / tasks; mbar_wait(&B.idsf[s0 % kRing], (uint32_t)((s0 / kRing) & 1)); }
        if (two) { const int s1 = l1 / tasks; mbar_wait(&B.idsf[s1 % kRing], (uint32_t)((s1 / kRing) & 1)); }
        issue(r0, lt);
        if (two) issue(r1, l1);
        lazy_resolve<LPR, VEC, PAIRWISE>(r0, r1, two, a, lane);
        score_and_stash(r0);
        if (two) score_and_stash(r1);
      }
      return;
    }
    int lt = w;
    if (lt < total) issue(r0, lt);
    while (lt < total) {
      int nx = lt + n_loaders;
      if (nx < total) issue(r1, nx);
      score_and_stash(r0);
      lt = nx;
      if (lt >= total) break;
      nx = lt + n_loaders;
      if (nx < total) issue(r0, nx);
      score_and_stash(r1);
      lt = nx;
    }
  } else {
    // ------------------------------------------------ scatterers ------------------------------------------------------
    // (Lazily zeroed destination tables need nothing here: the loaders have claimed the step's rows, zero-filled the first-touch
    // ones and waited for those somebody else was filling before they stashed the task, so every row is ready for plain REDs.)
    const int x = warp - kServiceWarps - kLoaderWarps;
    const int sub = lane % LPR, grp = lane / LPR;
    const float inv_b = 1.0f / (float)a.batch;
    const float g = (a.grad_loss ? __ldg(a.grad_loss) : 1.0f) * a.scale;
    for (int s = 0; s < a.n_steps; ++s) {
      const int slot = s % kRing, st = s % n_stages;
      const uint32_t par = (uint32_t)((s / kRing) & 1);
      mbar_wait(&B.idsf[slot], par);   // (long complete) acquire: the TMA-written ids are visible to this warp
      const int64_t* ids = reinterpret_cast<const int64_t*>(ids_ring + (size_t)slot * L.ids_slot_bytes());
      const float* lab = reinterpret_cast<const float*>(ids + (size_t)R * L.slice);
      const float* rows = reinterpret_cast<const float*>(stage_ring + (size_t)st * L.stage_slot_bytes());
      const float* sc = rows + (size_t)R * L.slice * row_f;
      // one task: the rows of 32 / LPR interactions -> gradient rows -> destination
      auto scatter_task = [&](int q, float2 nf) {
        int mu = kRowAdd, ma = kRowAdd, mb = kRowAdd;
        const int j = q * IPW + grp;
        if (j >= cnt) return;
        const int64_t iu64 = ids[j], ia64 = ids[L.slice + j], ib64 = PAIRWISE ? ids[2 * L.slice + j] : 0;
        const int iu = (uint64_t)iu64 < (uint64_t)a.n_users ? (int)iu64 : -1;
        const int ia = (uint64_t)ia64 < (uint64_t)a.n_items ? (int)ia64 : -1;
        const int ib = (PAIRWISE && (uint64_t)ib64 < (uint64_t)a.n_items) ? (int)ib64 : -1;
        const float label = (!PAIRWISE && a.label != nullptr) ? lab[j] : 0.f;
        const float c = score_coeff<PAIRWISE>(a, g, inv_b, sc[j], sc[L.slice + j], label);
        float *hu = nullptr, *ha = nullptr, *hb = nullptr;
        if constexpr (hot_on) {   // rows on the hot list go to this CTA's shared-memory accumulators
          int sl;
          if (iu >= 0 && (sl = hot.find(0, iu)) >= 0) { hu = hot.acc + (size_t)sl * row_f; mu = kRowHot; }
          if (ia >= 0 && (sl = hot.find(1, ia)) >= 0) { ha = hot.acc + (size_t)sl * row_f; ma = kRowHot; }
          if (PAIRWISE && ib >= 0 && (sl = hot.find(1, ib)) >= 0) { hb = hot.acc + (size_t)sl * row_f; mb = kRowHot; }
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          const int cidx = sub + v * LPR;
          if (cidx >= a.nv) continue;
          const float4 ru = ld_row4(rows + (size_t)j * row_f, cidx);
          const float4 ra = ld_row4(rows + ((size_t)L.slice + j) * row_f, cidx);
          const float4 rb = PAIRWISE ? ld_row4(rows + ((size_t)2 * L.slice + j) * row_f, cidx) : ru;
          scatter_cols<PAIRWISE>(a, cidx, c, nf.x, nf.y, iu, ia, ib, ru, ra, rb, mu, ma, mb, hu, ha, hb);
        }
      };
      if (!EARLY) mbar_wait(&B.normf[slot], par);  // => every CTA (this one included) has scored and stashed the step
      mbar_wait(&B.adone[slot], par);  // (long complete) acquire: the loaders' stage writes are visible to this warp
      if (a.trace && lane == 0 && x == 0) a.trace[((size_t)s * gridDim.x + blockIdx.x) * 8 + 5] = gtime();
      const float2 nf = EARLY ? make_float2(0.f, 0.f) : norms[slot];
      for (int q = x; q < tasks; q += kScatterWarps) scatter_task(q, nf);
      __syncwarp();
      if (a.trace && lane == 0 && x == 0) a.trace[((size_t)s * gridDim.x + blockIdx.x) * 8 + 6] = gtime();
      if (EARLY) {
        if (lane == 0) mbar_arrive(&B.sfree[st]);
        mbar_wait(&B.normf[slot], par);  // the step's exchange is over: its id, partial and norm slots may be reused
        if (lane == 0) mbar_arrive(&B.ifree[slot]);
      } else if (lane == 0) {
        mbar_arrive(&B.sfree[st]);    // the stage slot may be overwritten by the loaders
        mbar_arrive(&B.ifree[slot]);  // the id slot may be refilled by the producer
      }
    }
    if constexpr (hot_on) {   // one flush per CTA and launch: the accumulated hot rows go to their destination rows
      __syncwarp();
      if (lane == 0) mbar_arrive(hot_done);
      if (x == 0) {
        mbar_wait(hot_done, 0u);
        const int n_hot = a.n_hot_u + a.n_hot_i;
        for (int sl = 0; sl < n_hot; ++sl) {
          const long long key = hot.slot_key[sl];
          if (key < 0) continue;
          float* row = shard_row((key & 1) ? a.item_dst : a.user_dst, 0, (int64_t)(key >> 1), row_f);
          for (int c4 = lane; c4 < a.nv; c4 += 32) red_add4(row, c4, *reinterpret_cast<const float4*>(hot.acc + (size_t)sl * row_f + 4 * c4));
        }
      }
    }
  }
}

// =====================================================================================================================
// REGISTER kernel: workers keep the rows in registers from gather to scatter
// =====================================================================================================================
template <int LPR, int VEC, bool PAIRWISE>
__global__ void __launch_bounds__(kRegThreads, 1) train_steps_regs_kernel(StepsArgs a) {
  constexpr int IPW = 32 / LPR;
  constexpr int R = PAIRWISE ? 3 : 2;
  XDR_DYN_SMEM_ALIGNED(unsigned char, smem_raw, 128);
  const int64_t first = (int64_t)blockIdx.x * a.slice;
  const int cnt = (int)min((int64_t)a.slice, a.batch - first);
  const int tasks = (cnt + IPW - 1) / IPW;
  const SmemLayout L(a.slice, R, (a.slice + IPW - 1) / IPW);
  const Bars B(smem_raw + L.bars_off());
  float2* norms = reinterpret_cast<float2*>(smem_raw + L.norms_off());
  float4* part = reinterpret_cast<float4*>(smem_raw + L.part_off());
  unsigned char* ids_ring = smem_raw + L.ids_off();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) init_bars(B, tasks, tasks, 0, 1);
  __syncthreads();

  if (warp == 0) {
    service_producer<PAIRWISE>(a, L, B, ids_ring, first, cnt, lane);
  } else if (warp == 1) {
    service_publisher(a, L, B, part, tasks, lane);
  } else if (warp < kServiceWarps) {
    service_gatherer(a, B, norms, warp - 2, lane);
  } else {
    // Only min(kWorkerWarps, 3*tasks) workers are active so that a worker's consecutive tasks are at most 3 steps
    // apart: with two tasks held per warp the steps in flight span <= 7 < kRing, so no ring is ever lapped.
    // The host guarantees tasks <= 2*kWorkerWarps (each worker owns at most two tasks of any step, and it scores
    // both before it waits for that step's norms).
    const int n_workers = min(kWorkerWarps, 3 * tasks);
    const int w = warp - kServiceWarps;
    if (w >= n_workers) return;
    const int sub = lane % LPR;
    const float inv_b = 1.0f / (float)a.batch;
    const float g = (a.grad_loss ? __ldg(a.grad_loss) : 1.0f) * a.scale;
    const int total = a.n_steps * tasks;
    using Regs = TaskRegs<VEC, PAIRWISE>;
    auto issue = [&](Regs& r, int lt) { task_issue<LPR, VEC, PAIRWISE>(r, lt, a, L, B, ids_ring, tasks, cnt, lane); };
    auto phase_a = [&](Regs& r) {
      const float4 p = task_score<LPR, VEC, PAIRWISE>(r, a);
      if (lane == 0) {
        part[(r.s % kRing) * L.tasks + r.q] = p;
        mbar_arrive(&B.adone[r.s % kRing]);  // release: the partial is visible to the publisher
        if (a.trace && r.q == 0) a.trace[((size_t)r.s * gridDim.x + blockIdx.x) * 8 + 3] = gtime();
      }
    };
    auto phase_b = [&](Regs& r) {
      const int slot = r.s % kRing;
      if (a.trace && lane == 0 && r.q == 0) a.trace[((size_t)r.s * gridDim.x + blockIdx.x) * 8 + 4] = gtime();
      mbar_wait(&B.normf[slot], (uint32_t)((r.s / kRing) & 1));
      if (a.trace && lane == 0 && r.q == 0) a.trace[((size_t)r.s * gridDim.x + blockIdx.x) * 8 + 5] = gtime();
      const float2 nf = norms[slot];
      const float c = score_coeff<PAIRWISE>(a, g, inv_b, r.sa, r.sb, r.label);
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const int cidx = sub + v * LPR;
        if (cidx >= a.nv) continue;
        scatter_cols<PAIRWISE>(a, cidx, c, nf.x, nf.y, r.iu, r.ia, r.ib, r.u[v], r.a[v], PAIRWISE ? r.b[v] : r.u[v]);
      }
      __syncwarp();
      if (a.trace && lane == 0 && r.q == 0) a.trace[((size_t)r.s * gridDim.x + blockIdx.x) * 8 + 6] = gtime();
      if (lane == 0) mbar_arrive(&B.ifree[slot]);  // this task no longer needs the step's id tile / partial slot
    };
    // Two register sets per warp; scoring runs ahead of the norm wait:  B(r0) issue(r0') B(r1) issue(r1') A(r0') A(r1')
    // Hazard rule: before waiting for the norms of step s, every task of this warp with step <= s must be scored.
    Regs r0, r1;
    bool live0 = false, live1 = false;  // the set holds a task (issued or scored)
    bool iss0 = false, iss1 = false;    // ... whose rows are requested but not scored yet
    int next = w;
    if (next < total) { issue(r0, next); next += n_workers; live0 = iss0 = true; }
    if (next < total) { issue(r1, next); next += n_workers; live1 = iss1 = true; }
    while (live0 || live1) {
      if (iss0) { phase_a(r0); iss0 = false; }
      if (iss1) { phase_a(r1); iss1 = false; }
      if (live0) {
        phase_b(r0);
        live0 = false;
        if (next < total) { issue(r0, next); next += n_workers; live0 = iss0 = true; }
      }
      if (live1) {
        // never block on norms while holding requested-but-unscored rows: other CTAs (and the hazard rule) may be
        // waiting for exactly that partial
        if (iss0 && (r0.s <= r1.s || !mbar_test(&B.normf[r1.s % kRing], (uint32_t)((r1.s / kRing) & 1)))) {
          phase_a(r0);
          iss0 = false;
        }
        phase_b(r1);
        live1 = false;
        if (next < total) { issue(r1, next); next += n_workers; live1 = iss1 = true; }
      }
    }
  }
}

struct StepsPlan {
  int grid, slice, lpr, vec, stages;  // stages > 0: staged kernel; 0: register kernel
  size_t smem;
};

// Lanes per row / float4 columns per lane for a row of nv float4s (at most 2 columns per lane so that two tasks stay
// in registers), CTA slice, and which kernel fits.  Returns false when neither does (use the per-step kernels).
static bool plan_steps(int64_t batch, int nv, bool pairwise, StepsPlan* plan) {
  int lpr, vec;
  if (nv <= 8) { lpr = 8; vec = 1; }
  else if (nv <= 16) { lpr = 8; vec = 2; }
  else if (nv <= 32) { lpr = 16; vec = 2; }
  else if (nv <= 64) { lpr = 32; vec = 2; }
  else return false;
  const int sms = sm_count();
  int64_t slice = (batch + sms - 1) / sms;
  slice = (slice + 3) & ~(int64_t)3;  // multiple of 4: 16-byte granular TMA id tiles, whole warp tasks
  if (slice < 4) slice = 4;
  const int64_t grid = (batch + slice - 1) / slice;
  if (grid > sms || grid > 32 * kMaxCtaPerLane) return false;
  const int ipw = 32 / lpr;
  const int tasks = (int)((slice + ipw - 1) / ipw);
  plan->grid = (int)grid;
  plan->slice = (int)slice;
  plan->lpr = lpr;
  plan->vec = vec;
  const size_t smem_cap = 220 * 1024;
  for (int ns = kMaxStages; ns >= 3; --ns) {
    SmemLayout L((int)slice, pairwise ? 3 : 2, tasks, nv * 4, ns);
    if (L.bytes() <= smem_cap) {
      plan->stages = ns;
      plan->smem = L.bytes();
      return true;
    }
  }
  // register kernel: a worker scores at most two tasks before it waits for a step's norms, so every step's tasks
  // must fit 2 x workers, else the step could never complete
  SmemLayout L((int)slice, pairwise ? 3 : 2, tasks);
  if (tasks > 2 * kWorkerWarps || L.bytes() > smem_cap) return false;
  plan->stages = 0;
  plan->smem = L.bytes();
  return true;
}

// hot-row lists of the next launches (xdr_steps_set_hot_rows): device pointers to int64 row ids, users then items
static const int64_t* g_hot_u = nullptr;
static const int64_t* g_hot_i = nullptr;
static int g_n_hot_u = 0, g_n_hot_i = 0;

static int g_early_scatter = 0;  // opt-in (xdr_steps_set_early_scatter): reg_weight == 0 launches take the EARLY staged kernel

template <int LPR, int VEC, bool PW>
static int launch_steps(const StepsArgs& a, const StepsPlan& plan, cudaStream_t s) {
  const bool lazy = a.touch_u != nullptr;
  if (plan.stages > 0 && a.stage_a != nullptr) {
    auto kern = train_steps_staged_kernel<LPR, VEC, PW, kLoaderWarpsLite>;
    XDR_LAUNCH_COOP((kern), plan.grid, staged_threads(kLoaderWarpsLite), plan.smem, s, a, plan.stages);
  } else if (plan.stages > 0 && lazy) {
    auto kern = train_steps_staged_kernel<LPR, VEC, PW, kLoaderWarpsFull, false, true>;
    XDR_LAUNCH_COOP((kern), plan.grid, staged_threads(kLoaderWarpsFull), plan.smem, s, a, plan.stages);
  } else if (plan.stages > 0 && a.hot_off != 0) {
    auto kern = train_steps_staged_kernel<LPR, VEC, PW, kLoaderWarpsFull, false, false, true>;
    XDR_LAUNCH_COOP((kern), plan.grid, staged_threads(kLoaderWarpsFull), plan.smem, s, a, plan.stages);
  } else if (plan.stages > 0 && g_early_scatter && a.reg_weight == 0.f) {
    auto kern = train_steps_staged_kernel<LPR, VEC, PW, kLoaderWarpsFull, true>;
    XDR_LAUNCH_COOP((kern), plan.grid, staged_threads(kLoaderWarpsFull), plan.smem, s, a, plan.stages);
  } else if (plan.stages > 0) {
    auto kern = train_steps_staged_kernel<LPR, VEC, PW, kLoaderWarpsFull>;
    XDR_LAUNCH_COOP((kern), plan.grid, staged_threads(kLoaderWarpsFull), plan.smem, s, a, plan.stages);
  } else {
    auto kern = train_steps_regs_kernel<LPR, VEC, PW>;
    XDR_LAUNCH_COOP((kern), plan.grid, kRegThreads, plan.smem, s, a);
  }
  return XDR_OK;
}

template <bool PW>
static int dispatch_steps(const StepsArgs& a, const StepsPlan& plan, cudaStream_t s) {
  if (plan.lpr == 8 && plan.vec == 1) return launch_steps<8, 1, PW>(a, plan, s);
  if (plan.lpr == 8 && plan.vec == 2) return launch_steps<8, 2, PW>(a, plan, s);
  if (plan.lpr == 16) return launch_steps<16, 2, PW>(a, plan, s);
  return launch_steps<32, 2, PW>(a, plan, s);
}

struct WsRecord {
  int dev;
  void* ptr;
  size_t zeroed_bytes;     // prefix of the workspace this library has zeroed (and owns the tags of) since it first saw it
  unsigned int next_tag;   // tags handed out so far
};
static std::mutex g_ws_mu;
static std::vector<WsRecord> g_ws_seen;

static unsigned long long* g_trace = nullptr;  // debug only, see xdr_debug_set_steps_trace
static int g_force_regs = 0;                   // debug only: force the register kernel where both fit

}  // namespace xdr

using namespace xdr;

extern "C" {

// Debug hooks (not part of the drop-in surface).  Trace: device buffer of n_steps*grid*8 u64 that the next
// xdr_train_steps launches fill with globaltimer stamps: [0] CTA partial ready, [1] norms known, [2] task 0 rows
// requested, [3] task 0 scored (+stashed), [4] task 0 starts waiting for norms (register kernel), [5] norms arrived at
// the scatter side, [6] scatter issued.
XDR_API void xdr_debug_set_steps_trace(void* buf) { g_trace = reinterpret_cast<unsigned long long*>(buf); }
XDR_API void xdr_debug_force_register_kernel(int on) { g_force_regs = on; }
// Opt-in (validated on hardware; the default kernel stays the one every measurement was taken on): launches with reg_weight == 0 (CMF's yaml default lambda = gamma = 0, reg-free BPR)
// scatter without waiting for the step's norm exchange (train_steps_staged_kernel<..., EARLY = true>).
XDR_API void xdr_steps_set_early_scatter(int on) { g_early_scatter = on; }

// Hot rows of the following xdr_train_steps launches (plain scatter-add destinations, single GPU): device arrays of int64 row
// ids -- e.g. the most popular items of the catalogue, computed once per dataset.  Every CTA accumulates the gradients of
// these rows in shared memory over the whole launch and adds them to the destination once, at the end (same sums, other
// order).  At most 64 rows (fewer for rows wider than 64 floats); the arrays must stay valid while launches use them.
// n_hot_users = n_hot_items = 0 switches it off.
int xdr_steps_set_hot_rows(const int64_t* hot_users, int n_hot_users, const int64_t* hot_items, int n_hot_items) {
  XDR_REQUIRE(n_hot_users >= 0 && n_hot_items >= 0, "xdr_steps_set_hot_rows: negative count");
  XDR_REQUIRE((n_hot_users == 0 || hot_users) && (n_hot_items == 0 || hot_items), "xdr_steps_set_hot_rows: null id array");
  g_hot_u = hot_users; g_hot_i = hot_items; g_n_hot_u = n_hot_users; g_n_hot_i = n_hot_items;
  return XDR_OK;
}

// 2 bits per row, 16 rows per 32-bit word; every table's part is padded to a multiple of 4 words (16 bytes)
static inline size_t touch_words_of(int64_t n_rows) { return (size_t)(((n_rows + 15) / 16 + 3) & ~(int64_t)3); }

size_t xdr_touch_map_bytes(int64_t n_users, int64_t n_items) {
  if (n_users < 0 || n_items < 0) return 0;
  return (touch_words_of(n_users) + touch_words_of(n_items)) * sizeof(uint32_t);
}

size_t xdr_steps_workspace_bytes(int n_steps) {
  if (n_steps < 0) return 0;
  return (size_t)n_steps * ((size_t)sm_count() * 3 + 2) * sizeof(unsigned long long);
}

static int train_steps_core(const Shards& user_tab, const Shards& item_tab, const Shards& user_dst, const Shards& item_dst,
                            int log2g, int64_t n_users, int64_t n_items, int dim, const int64_t* user,
                            const int64_t* item_a, const int64_t* item_b, const float* label, int64_t step_stride,
                            int64_t batch, int n_steps, int pairwise, int loss_kind, float gamma, float reg_weight,
                            const float* grad_loss, float scale, float* out8, void* steps_ws, size_t steps_ws_bytes,
                            const float* stage_a, const float* stage_b, int32_t* oob, xdr_stream_t stream,
                            uint32_t* touch = nullptr, int touch_clear = 0) {
  const char* fn = "xdr_train_steps";
  XDR_REQUIRE(dim_ok(dim), "%s: dim=%d must be a multiple of 4 in (0, 256]", fn, dim);
  XDR_REQUIRE(batch > 0 && n_steps >= 0, "%s: batch=%lld n_steps=%d", fn, (long long)batch, n_steps);
  if (n_steps == 0) return XDR_OK;
  XDR_REQUIRE(user && item_a && out8 && steps_ws, "%s: null pointer", fn);
  for (int g = 0; g < (1 << log2g); ++g) {
    XDR_REQUIRE(user_tab.p[g] && item_tab.p[g] && user_dst.p[g] && item_dst.p[g], "%s: null table shard %d", fn, g);
    XDR_REQUIRE(aligned16(user_tab.p[g]) && aligned16(item_tab.p[g]) && aligned16(user_dst.p[g]) && aligned16(item_dst.p[g]),
                "%s: table shards must be 16-byte aligned", fn);
  }
  XDR_REQUIRE(!pairwise || item_b, "%s: pairwise needs the negative-item ids", fn);
  XDR_REQUIRE(pairwise || loss_kind == XDR_LOSS_NONE || label, "%s: label is required for this loss kind", fn);
  XDR_REQUIRE(pairwise || (loss_kind >= XDR_LOSS_MSE && loss_kind <= XDR_LOSS_NONE), "%s: bad loss_kind", fn);
  XDR_REQUIRE(step_stride >= batch, "%s: step_stride=%lld < batch", fn, (long long)step_stride);
  XDR_REQUIRE(steps_ws_bytes >= xdr_steps_workspace_bytes(n_steps), "%s: steps_ws too small (%zu < %zu)", fn,
              steps_ws_bytes, xdr_steps_workspace_bytes(n_steps));
  StepsPlan plan;
  // TMA bulk copies of the id tiles need 16-byte aligned sources and sizes
  const bool tma_ok = (batch % 4) == 0 && (step_stride % 4) == 0 && aligned16(user) && aligned16(item_a) &&
                      (!pairwise || aligned16(item_b)) && (label == nullptr || aligned16(label));
  if (!tma_ok || !plan_steps(batch, dim / 4, pairwise != 0, &plan)) {
    set_error("%s: batch=%lld dim=%d (batch and step_stride must be multiples of 4, id arrays 16-byte aligned, and a "
              "CTA slice must fit the stage ring or the register kernel); use the per-step entry points",
              fn, (long long)batch, dim);
    return XDR_ERR_UNSUPPORTED;
  }
  if (touch != nullptr) {
    XDR_REQUIRE(log2g == 0, "%s: lazily zeroed gradient tables (touch map) are a single-GPU feature", fn);
    XDR_REQUIRE(aligned16(touch), "%s: the touch map must be 16-byte aligned", fn);
    XDR_REQUIRE(user_dst.p[0] != user_tab.p[0] && item_dst.p[0] != item_tab.p[0],
                "%s: a touch map zero-fills destination rows on first touch; the destination cannot be the weight table", fn);
    if (plan.stages == 0) {
      set_error("%s: the touch map needs the staged kernel, which does not fit batch=%lld dim=%d; zero the gradient tables "
                "instead", fn, (long long)batch, dim);
      return XDR_ERR_UNSUPPORTED;
    }
  }
  if (g_force_regs && plan.stages > 0 && touch == nullptr) {
    const int ipw = 32 / plan.lpr;
    const int tasks = (plan.slice + ipw - 1) / ipw;
    if (tasks <= 2 * kWorkerWarps) {
      plan.stages = 0;
      plan.smem = SmemLayout(plan.slice, pairwise ? 3 : 2, tasks).bytes();
    }
  }
  StepsArgs a{};
  a.user_tab = user_tab; a.item_tab = item_tab; a.user_dst = user_dst; a.item_dst = item_dst; a.log2g = log2g;
  XDR_REQUIRE(stage_a == nullptr || (aligned16(stage_a) && (!pairwise || (stage_b && aligned16(stage_b)))),
              "%s: staged item rows must be 16-byte aligned (and both given for pairwise)", fn);
  a.stage_a = stage_a; a.stage_b = stage_b;
  a.n_users = n_users; a.n_items = n_items; a.nv = dim / 4;
  a.user = user; a.item_a = item_a; a.item_b = item_b; a.label = label; a.step_stride = step_stride; a.batch = batch;
  a.n_steps = n_steps; a.loss_kind = loss_kind; a.gamma = gamma; a.reg_weight = reg_weight; a.out8 = out8;
  a.grad_loss = grad_loss; a.scale = scale; a.slice = plan.slice; a.oob = oob;
  a.words = reinterpret_cast<unsigned long long*>(steps_ws);
  a.trace = g_trace;
  cudaStream_t s = (cudaStream_t)stream;
  if (touch == nullptr && log2g == 0 && plan.stages > 0 && stage_a == nullptr && g_n_hot_u + g_n_hot_i > 0) {
    // hot rows: as many of the listed rows as fit 16 KB of accumulators (users first), if the CTA still fits shared memory
    const int row_f = dim;
    int nu_h = g_n_hot_u, ni_h = g_n_hot_i;
    const int cap = (int)((size_t)kHotBytes / ((size_t)row_f * 4)) < kHotMax ? (int)((size_t)kHotBytes / ((size_t)row_f * 4)) : kHotMax;
    if (nu_h > cap) nu_h = cap;
    if (ni_h > cap - nu_h) ni_h = cap - nu_h;
    const size_t extra = hot_region_bytes(nu_h + ni_h, row_f);
    const size_t base = (plan.smem + 127) & ~(size_t)127;
    if (nu_h + ni_h > 0 && base + extra <= (size_t)220 * 1024) {
      a.hot_u = g_hot_u; a.hot_i = g_hot_i; a.n_hot_u = nu_h; a.n_hot_i = ni_h;
      a.hot_off = base;
      plan.smem = base + extra;
    }
  }
  if (touch != nullptr) {
    a.touch_u = touch;
    a.touch_i = touch + touch_words_of(n_users);
    if (touch_clear) XDR_CUDA_OK(cudaMemsetAsync(touch, 0, xdr_touch_map_bytes(n_users, n_items), s));
  }
  // Step tags grow from launch to launch on one workspace (tag_base + s + 1), so a word left behind by an earlier launch can
  // never match: the workspace is zeroed only when the library sees it for the first time (or the 32-bit tags would wrap),
  // not per launch.  The caller must not write to it between launches (xdr.h).
  {
    const size_t need = xdr_steps_workspace_bytes(n_steps);
    int dev = 0;
    XDR_CUDA_OK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_ws_mu);
    WsRecord* rec = nullptr;
    for (auto& r : g_ws_seen)
      if (r.dev == dev && r.ptr == steps_ws) rec = &r;
    if (!rec) {
      if (g_ws_seen.size() >= 64) g_ws_seen.clear();
      g_ws_seen.push_back(WsRecord{dev, steps_ws, 0, 0});
      rec = &g_ws_seen.back();
    }
    if (rec->zeroed_bytes < need || rec->next_tag > 0xf0000000u - (unsigned int)n_steps) {
      XDR_CUDA_OK(cudaMemsetAsync(steps_ws, 0, need, s));
      rec->zeroed_bytes = need;
      rec->next_tag = 0;
    }
    a.tag_base = rec->next_tag;
    rec->next_tag += (unsigned int)n_steps;
  }
  const int rc = pairwise ? dispatch_steps<true>(a, plan, s) : dispatch_steps<false>(a, plan, s);
  if (rc != XDR_OK) return rc;
  XDR_LAUNCH_OK();
  return XDR_OK;
}

int xdr_train_steps(const float* user_tab, const float* item_tab, int64_t n_users, int64_t n_items, int dim,
                    const int64_t* user, const int64_t* item_a, const int64_t* item_b, const float* label,
                    int64_t step_stride, int64_t batch, int n_steps, int pairwise, int loss_kind, float gamma,
                    float reg_weight, const float* grad_loss, float scale, float* user_dst, float* item_dst, float* out8,
                    void* steps_ws, size_t steps_ws_bytes, int32_t* oob, xdr_stream_t stream) {
  Shards ut{}, it{}, du{}, di{};
  ut.p[0] = const_cast<float*>(user_tab);
  it.p[0] = const_cast<float*>(item_tab);
  du.p[0] = user_dst;
  di.p[0] = item_dst;
  return train_steps_core(ut, it, du, di, 0, n_users, n_items, dim, user, item_a, item_b, label, step_stride, batch,
                          n_steps, pairwise, loss_kind, gamma, reg_weight, grad_loss, scale, out8, steps_ws,
                          steps_ws_bytes, nullptr, nullptr, oob, stream);
}

int xdr_train_steps_lazy(const float* user_tab, const float* item_tab, int64_t n_users, int64_t n_items, int dim,
                         const int64_t* user, const int64_t* item_a, const int64_t* item_b, const float* label,
                         int64_t step_stride, int64_t batch, int n_steps, int pairwise, int loss_kind, float gamma,
                         float reg_weight, const float* grad_loss, float scale, float* user_dst, float* item_dst, float* out8,
                         void* steps_ws, size_t steps_ws_bytes, uint32_t* touch_map, int clear_map, int32_t* oob,
                         xdr_stream_t stream) {
  if (touch_map == nullptr) {
    set_error("xdr_train_steps_lazy: null touch map (use xdr_train_steps for plain scatter-add destinations)");
    return XDR_ERR_INVALID;
  }
  Shards ut{}, it{}, du{}, di{};
  ut.p[0] = const_cast<float*>(user_tab);
  it.p[0] = const_cast<float*>(item_tab);
  du.p[0] = user_dst;
  di.p[0] = item_dst;
  return train_steps_core(ut, it, du, di, 0, n_users, n_items, dim, user, item_a, item_b, label, step_stride, batch,
                          n_steps, pairwise, loss_kind, gamma, reg_weight, grad_loss, scale, out8, steps_ws,
                          steps_ws_bytes, nullptr, nullptr, oob, stream, touch_map, clear_map);
}

int xdr_train_steps_sharded(const float* const* user_shards, const float* const* item_shards, float* const* user_dst_shards,
                            float* const* item_dst_shards, int n_shards, int64_t n_users, int64_t n_items, int dim,
                            const int64_t* user, const int64_t* item_a, const int64_t* item_b, const float* label,
                            int64_t step_stride, int64_t batch, int n_steps, int pairwise, int loss_kind, float gamma,
                            float reg_weight, const float* grad_loss, float scale, float* out8, void* steps_ws,
                            size_t steps_ws_bytes, const float* staged_item_a, const float* staged_item_b, int32_t* oob,
                            xdr_stream_t stream) {
  XDR_REQUIRE(n_shards >= 1 && n_shards <= kMaxShards && (n_shards & (n_shards - 1)) == 0,
              "xdr_train_steps_sharded: n_shards=%d must be a power of two <= %d", n_shards, kMaxShards);
  XDR_REQUIRE(user_shards && item_shards && user_dst_shards && item_dst_shards, "xdr_train_steps_sharded: null pointer");
  Shards ut{}, it{}, du{}, di{};
  int log2g = 0;
  while ((1 << log2g) < n_shards) ++log2g;
  for (int g = 0; g < n_shards; ++g) {
    ut.p[g] = const_cast<float*>(user_shards[g]);
    it.p[g] = const_cast<float*>(item_shards[g]);
    du.p[g] = user_dst_shards[g];
    di.p[g] = item_dst_shards[g];
  }
  return train_steps_core(ut, it, du, di, log2g, n_users, n_items, dim, user, item_a, item_b, label, step_stride, batch,
                          n_steps, pairwise, loss_kind, gamma, reg_weight, grad_loss, scale, out8, steps_ws,
                          steps_ws_bytes, staged_item_a, staged_item_b, oob, stream);
}

#ifndef XDR_EMU  // streams, events and host copies have no emulator counterpart
// ---- one chunk of K steps fed from PINNED HOST id blocks, enqueued with ONE call -------------------------------------------
int xdr_train_steps_host(const float* user_tab, const float* item_tab, int64_t n_users, int64_t n_items, int dim,
                         const int64_t* host_ids, int64_t* dev_ids, int64_t batch, int n_steps, float gamma, float reg_weight,
                         const float* grad_loss, float scale, float* user_dst, float* item_dst, float* out8, float* host_out8,
                         void* steps_ws, size_t steps_ws_bytes, int32_t* oob, xdr_stream_t copy_stream, void* buf_free_event,
                         void* ids_ready_event, void* launch_done_event, xdr_stream_t stream) {
  const char* fn = "xdr_train_steps_host";
  XDR_REQUIRE(user_tab && item_tab && host_ids && dev_ids && user_dst && item_dst && out8 && steps_ws && ids_ready_event,
              "%s: null pointer", fn);
  XDR_REQUIRE(dim_ok(dim), "%s: dim=%d must be a multiple of 4 in (0, 256]", fn, dim);
  XDR_REQUIRE(batch > 0 && batch % 4 == 0 && n_steps > 0, "%s: batch=%lld (a positive multiple of 4) n_steps=%d (> 0)", fn,
              (long long)batch, n_steps);
  XDR_REQUIRE(aligned16(dev_ids), "%s: the device id block must be 16-byte aligned", fn);
  cudaStream_t cs = (cudaStream_t)copy_stream, ms = (cudaStream_t)stream;
  if (buf_free_event) XDR_CUDA_OK(cudaStreamWaitEvent(cs, (cudaEvent_t)buf_free_event, 0));   // the buffer's previous launch is over
  XDR_CUDA_OK(cudaMemcpyAsync(dev_ids, host_ids, (size_t)n_steps * 3 * (size_t)batch * sizeof(int64_t), cudaMemcpyHostToDevice, cs));
  XDR_CUDA_OK(cudaEventRecord((cudaEvent_t)ids_ready_event, cs));
  XDR_CUDA_OK(cudaStreamWaitEvent(ms, (cudaEvent_t)ids_ready_event, 0));
  const int rc = xdr_train_steps(user_tab, item_tab, n_users, n_items, dim, dev_ids, dev_ids + batch, dev_ids + 2 * batch, nullptr,
                                 3 * batch, batch, n_steps, 1, XDR_LOSS_NONE, gamma, reg_weight, grad_loss, scale, user_dst, item_dst,
                                 out8, steps_ws, steps_ws_bytes, oob, stream);
  if (rc != XDR_OK) return rc;
  if (host_out8)
    XDR_CUDA_OK(cudaMemcpyAsync(host_out8, out8, (size_t)n_steps * 8 * sizeof(float), cudaMemcpyDeviceToHost, ms));
  if (launch_done_event) XDR_CUDA_OK(cudaEventRecord((cudaEvent_t)launch_done_event, ms));
  return XDR_OK;
}

// (CUDA IPC has no emulator counterpart either)
// ---- peer-memory plumbing for row-sharded tables (CUDA IPC; one process per GPU) --------------------------------------
int xdr_ipc_export(const void* dev_ptr, unsigned char* handle64_host, int64_t* offset_host) {
  XDR_REQUIRE(dev_ptr && handle64_host && offset_host, "xdr_ipc_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaPointerAttributes attr;
  XDR_CUDA_OK(cudaPointerGetAttributes(&attr, dev_ptr));
  XDR_REQUIRE(attr.type == cudaMemoryTypeDevice, "xdr_ipc_export: not a device pointer");
  // the IPC handle names the whole allocation: find its base with a page-granular probe of the driver's range query
  void* base = nullptr;
  size_t size = 0;
  {
    typedef int (*range_fn)(void**, size_t*, void*);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    XDR_CUDA_OK(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres));
    XDR_REQUIRE(fn != nullptr, "xdr_ipc_export: cuMemGetAddressRange unavailable");
    const int rc = reinterpret_cast<range_fn>(fn)(&base, &size, const_cast<void*>(dev_ptr));
    XDR_REQUIRE(rc == 0 && base != nullptr, "xdr_ipc_export: cuMemGetAddressRange failed (%d)", rc);
  }
  cudaIpcMemHandle_t h;
  XDR_CUDA_OK(cudaIpcGetMemHandle(&h, base));
  memcpy(handle64_host, &h, 64);
  *offset_host = (int64_t)(reinterpret_cast<const char*>(dev_ptr) - reinterpret_cast<const char*>(base));
  return XDR_OK;
}

int xdr_ipc_open(const unsigned char* handle64_host, void** base_out_host) {
  XDR_REQUIRE(handle64_host && base_out_host, "xdr_ipc_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64_host, 64);
  XDR_CUDA_OK(cudaIpcOpenMemHandle(base_out_host, h, cudaIpcMemLazyEnablePeerAccess));
  return XDR_OK;
}

int xdr_ipc_close(void* base) {
  XDR_REQUIRE(base, "xdr_ipc_close: null pointer");
  XDR_CUDA_OK(cudaIpcCloseMemHandle(base));
  return XDR_OK;
}
#endif  // !XDR_EMU

}  // extern "C"
