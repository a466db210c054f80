// steps_persistent.cu -- the trainer-step hot loop as ONE persistent, software-pipelined launch over K batches.
//
// Replaces K iterations of recbole Trainer._train_epoch's inner loop [recbole-1.0.1] around
// EMCDR.calculate_source_loss/target_loss (reference emcdr.py:110-154; CMF per-domain term cmf.py:75-98):
//     for batch in batches: loss = calculate_loss(batch); loss.backward()
// i.e. per batch: gather -> dot score -> loss (+EmbLoss) -> row gradients -> scatter-add, with the per-batch
// losses written to out8[k].  Results per batch are identical to xdr_bpr_fwd + xdr_bpr_bwd on that batch
// (gradient accumulation over the K batches when dst is a gradient table; asynchronous SGD with bounded
// staleness <= kStages steps when dst is the weight table and scale = -lr).
//
// Why: one B=8192 step moves 12.8 MB = 1.95 us at the HBM roofline while a dependent load chain
// (ids -> rows -> grid reduction for the EmbLoss norms -> atomics) is > 3 us of latency and a launch ~2 us.
// So the launch is amortised over K steps and the latencies are overlapped across steps:
//
//   grid    one CTA per SM (persistent); CTA c owns interactions [c*S, (c+1)*S) of EVERY step (S = ceil(B/grid))
//   warp 0  producer: for step t (running ahead by up to kStages-2 steps)
//             TMA bulk copy (cp.async.bulk, UBLKCP) of the CTA's id tiles -> smem, wait, then one bulk copy per
//             embedding row (dim*4 bytes, 256 B at dim 64) HBM -> smem stage, completion on an mbarrier
//   warps 1..16 consumers: iteration s does
//             phase A(s)   rows from smem (8 lanes per interaction, LDS.128), dots + squared norms via shuffles,
//                          per-CTA partial sums -> global, arrive on the step counter
//             phase B(s-1) (lagging one step so the grid-wide EmbLoss norm of step s-1 is already complete)
//                          fixed-order fp64 reduction of all CTAs' partials, row gradients from the rows still in
//                          smem, REDG.E.ADD.F32x4 scatter-add to the destination tables, release the stage
//   No grid-wide barrier anywhere; the only cross-CTA dependency is the per-step arrival counter, which is
//   waited on one full iteration after it was signalled.
//
// HBM roofline: same algorithmic bytes as the two-kernel path (1560 B per BPR interaction at dim 64) but the
// backward re-gather disappears (rows stay in smem), so DRAM traffic ~= algorithmic bytes.
#include "xdr_common.cuh"

namespace xdr {

constexpr int kConsumerWarps = 16;
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kStepThreads = kConsumerThreads + 32;  // + producer warp
constexpr int kStages = 4;

struct StepsArgs {
  const float* user_tab;
  const float* item_tab;
  int64_t n_users, n_items;
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
  float* user_dst;
  float* item_dst;
  float4* partials;      // [n_steps][gridDim.x]
  unsigned int* arrive;  // [n_steps], zeroed by the host wrapper before the launch
  int slice;             // S
  int32_t* oob;
};

// ---- mbarrier / bulk-copy PTX wrappers ---------------------------------------------------------------------------
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
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory"); }

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Shared-memory stage: ids, (labels), rows and scores of one step's slice.
template <bool PAIRWISE>
struct StageLayout {
  static constexpr int kRowsPer = PAIRWISE ? 3 : 2;
  int slice, row_f;
  __host__ __device__ StageLayout(int s, int nv) : slice(s), row_f(nv * 4) {}
  // all offsets in bytes, each region 16-byte aligned (slice is rounded up to a multiple of 4 by the host)
  __host__ __device__ size_t ids_off() const { return 0; }
  __host__ __device__ size_t label_off() const { return ids_off() + (size_t)kRowsPer * slice * 8; }
  __host__ __device__ size_t score_off() const { return label_off() + (size_t)slice * 4; }
  __host__ __device__ size_t rows_off() const { return score_off() + (size_t)2 * slice * 4; }
  __host__ __device__ size_t bytes() const {
    size_t b = rows_off() + (size_t)kRowsPer * slice * row_f * 4;
    return (b + 127) & ~(size_t)127;
  }
};

constexpr size_t kHeaderBytes = 256;  // mbarriers + cross-warp reduction scratch

template <int VEC, bool PAIRWISE>
__global__ void __launch_bounds__(kStepThreads, 1) train_steps_kernel(StepsArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);  // [kStages] rows landed
  uint64_t* idsf = full + kStages;                         // [kStages] id tiles landed
  uint64_t* empty = idsf + kStages;                        // [kStages] stage released by the consumers
  float* red = reinterpret_cast<float*>(smem_raw + 3 * kStages * 8);  // [4*kConsumerWarps] + 4 broadcast slots... see below
  // header budget: 3*4*8 = 96 B of barriers; reduction scratch lives right after the header
  const StageLayout<PAIRWISE> L(a.slice, a.nv);
  constexpr int R = StageLayout<PAIRWISE>::kRowsPer;
  unsigned char* stage0 = smem_raw + kHeaderBytes + 4 * sizeof(float) * (kConsumerWarps + 2);
  stage0 = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(stage0) + 127) & ~(uintptr_t)127);
  float* wred = reinterpret_cast<float*>(smem_raw + kHeaderBytes);  // [3][kConsumerWarps] partial sums
  float* bcast = wred + 3 * kConsumerWarps;                          // [4] cu, ci for phase B (+ spare)
  (void)red;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t first = (int64_t)blockIdx.x * a.slice;
  const int cnt = (int)min((int64_t)a.slice, a.batch - first);  // > 0: the host launches ceil(batch/slice) CTAs
  const int row_f = a.nv * 4;
  const uint32_t row_bytes = (uint32_t)row_f * 4u;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&idsf[i], 1);
      mbar_init(&empty[i], kConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == 0) {
    // =========================================== producer ===========================================
    for (int t = 0; t < a.n_steps; ++t) {
      const int slot = t % kStages;
      const uint32_t par = (uint32_t)((t / kStages) & 1);
      unsigned char* st = stage0 + (size_t)slot * L.bytes();
      int64_t* ids = reinterpret_cast<int64_t*>(st + L.ids_off());
      float* lab = reinterpret_cast<float*>(st + L.label_off());
      float* rows = reinterpret_cast<float*>(st + L.rows_off());
      mbar_wait(&empty[slot], par ^ 1u);  // passes immediately the first time round the ring
      const int64_t off = (int64_t)t * a.step_stride + first;
      // id tiles: cnt*8 bytes each; cnt is even except possibly in the last CTA -> round the copy up to 16 bytes
      // (the host guarantees the id arrays are padded/aligned so the rounded copy stays inside the allocation)
      const uint32_t idb = ((uint32_t)cnt * 8u + 15u) & ~15u;
      const uint32_t lbb = ((uint32_t)cnt * 4u + 15u) & ~15u;
      if (lane == 0) {
        const bool has_label = !PAIRWISE && a.label != nullptr;
        mbar_expect_tx(&idsf[slot], idb * R + (has_label ? lbb : 0u));
        bulk_g2s(ids, a.user + off, idb, &idsf[slot]);
        bulk_g2s(ids + L.slice, a.item_a + off, idb, &idsf[slot]);
        if (PAIRWISE) bulk_g2s(ids + 2 * L.slice, a.item_b + off, idb, &idsf[slot]);
        if (has_label) bulk_g2s(lab, a.label + off, lbb, &idsf[slot]);
      }
      __syncwarp();
      mbar_wait(&idsf[slot], par);
      if (lane == 0) mbar_expect_tx(&full[slot], (uint32_t)(cnt * R) * row_bytes);
      __syncwarp();
      for (int r = lane; r < cnt * R; r += 32) {
        const int which = r / cnt, j = r - which * cnt;
        int64_t id = ids[which * L.slice + j];
        const int64_t n_rows = which == 0 ? a.n_users : a.n_items;
        if ((uint64_t)id >= (uint64_t)n_rows) {
          if (a.oob) *a.oob = 1;
          id = 0;  // keep the copy in bounds; the consumers zero the row through the same validity test
        }
        const float* src = (which == 0 ? a.user_tab : a.item_tab) + id * (int64_t)row_f;
        bulk_g2s(rows + ((size_t)which * L.slice + j) * row_f, src, row_bytes, &full[slot]);
      }
    }
  } else {
    // =========================================== consumers ==========================================
    const int cw = warp - 1;
    const int sub = lane & (kLanesPerRow - 1), grp = lane >> 3;
    const int ctid = threadIdx.x - 32;
    const float inv_b = 1.0f / (float)a.batch;
    const float g = (a.grad_loss ? __ldg(a.grad_loss) : 1.0f) * a.scale;
    const unsigned int n_cta = gridDim.x;
    for (int s = 0; s <= a.n_steps; ++s) {
      if (s < a.n_steps) {
        // ---------------- phase A(s) ----------------
        const int slot = s % kStages;
        const uint32_t par = (uint32_t)((s / kStages) & 1);
        unsigned char* st = stage0 + (size_t)slot * L.bytes();
        const int64_t* ids = reinterpret_cast<const int64_t*>(st + L.ids_off());
        const float* lab = reinterpret_cast<const float*>(st + L.label_off());
        float* sc = reinterpret_cast<float*>(st + L.score_off());
        const float* rows = reinterpret_cast<const float*>(st + L.rows_off());
        mbar_wait(&idsf[slot], par);
        mbar_wait(&full[slot], par);
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
        for (int base = cw * kRowsPerWarp; base < cnt; base += kConsumerWarps * kRowsPerWarp) {
          const int j = base + grp;
          const bool live = j < cnt;
          const int jj = live ? j : 0;
          const bool oku = live && (uint64_t)ids[jj] < (uint64_t)a.n_users;
          const bool oka = live && (uint64_t)ids[L.slice + jj] < (uint64_t)a.n_items;
          const bool okb = PAIRWISE && live && (uint64_t)ids[2 * L.slice + jj] < (uint64_t)a.n_items;
          const float* pu = rows + (size_t)jj * row_f;
          const float* pa = rows + ((size_t)L.slice + jj) * row_f;
          const float* pb = rows + ((size_t)2 * L.slice + jj) * row_f;
          float da = 0.f, db = 0.f, uu = 0.f, aa = 0.f;
          const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            const int c = sub + v * kLanesPerRow;
            const bool on = c < a.nv;
            const float4 ru = (oku && on) ? ld_row4(pu, c) : z;
            const float4 ra = (oka && on) ? ld_row4(pa, c) : z;
            da += dot4(ru, ra);
            uu += dot4(ru, ru);
            aa += dot4(ra, ra);
            if (PAIRWISE) {
              const float4 rb = (okb && on) ? ld_row4(pb, c) : z;
              db += dot4(ru, rb);
            }
          }
          da = group8_sum(da);
          if (PAIRWISE) db = group8_sum(db);
          uu = group8_sum(uu);
          aa = group8_sum(aa);
          if (live && sub == 0) {
            sc[j] = da;
            float term;
            if (PAIRWISE) {
              sc[L.slice + j] = db;
              term = -logf(a.gamma + sigmoidf_(da - db));
            } else if (a.loss_kind == XDR_LOSS_MSE) {
              const float d = da - lab[j];
              term = d * d;
            } else if (a.loss_kind == XDR_LOSS_BCE_SIGMOID) {
              const float p = sigmoidf_(da), y = lab[j];
              term = -(y * fmaxf(logf(p), -100.f) + (1.f - y) * fmaxf(logf(1.f - p), -100.f));
            } else {
              term = 0.f;
            }
            acc0 += term;
            acc1 += uu;
            acc2 += aa;
          }
        }
        acc0 = warp_sum(acc0);
        acc1 = warp_sum(acc1);
        acc2 = warp_sum(acc2);
        if (lane == 0) {
          wred[cw] = acc0;
          wred[kConsumerWarps + cw] = acc1;
          wred[2 * kConsumerWarps + cw] = acc2;
        }
      }
      consumer_bar();  // (1) this CTA's phase-A partial sums are in smem
      if (cw == 0) {
        if (s < a.n_steps) {
          float p0 = lane < kConsumerWarps ? wred[lane] : 0.f;
          float p1 = lane < kConsumerWarps ? wred[kConsumerWarps + lane] : 0.f;
          float p2 = lane < kConsumerWarps ? wred[2 * kConsumerWarps + lane] : 0.f;
          p0 = warp_sum(p0);
          p1 = warp_sum(p1);
          p2 = warp_sum(p2);
          if (lane == 0) {
            a.partials[(size_t)s * n_cta + blockIdx.x] = make_float4(p0, p1, p2, 0.f);
            __threadfence();
            atomicAdd(&a.arrive[s], 1u);
          }
        }
        if (s >= 1) {
          // ---- wait (normally already satisfied) for every CTA's phase A of step s-1, reduce in a fixed order
          const int sp = s - 1;
          if (lane == 0) {
            while (ld_acquire_u32(&a.arrive[sp]) < n_cta) __nanosleep(32);
          }
          __syncwarp();
          double t0 = 0.0, t1 = 0.0, t2 = 0.0;
          for (unsigned int c = lane; c < n_cta; c += 32) {
            const float4 p = __ldcg(&a.partials[(size_t)sp * n_cta + c]);
            t0 += (double)p.x;
            t1 += (double)p.y;
            t2 += (double)p.z;
          }
          t0 = warp_sum(t0);
          t1 = warp_sum(t1);
          t2 = warp_sum(t2);
          if (lane == 0) {
            const float nu = (float)sqrt(t1), ni = (float)sqrt(t2);
            bcast[0] = (a.reg_weight != 0.f && nu > 0.f) ? g * a.reg_weight * inv_b / nu : 0.f;
            bcast[1] = (a.reg_weight != 0.f && ni > 0.f) ? g * a.reg_weight * inv_b / ni : 0.f;
            if (blockIdx.x == 0) {
              const float data = (float)(t0 / (double)a.batch);
              const float reg = (float)(((double)nu + (double)ni) / (double)a.batch);
              float* o = a.out8 + (size_t)sp * 8;
              o[0] = data + a.reg_weight * reg;
              o[1] = data;
              o[2] = nu;
              o[3] = ni;
              o[4] = reg;
              o[5] = 0.f;
              o[6] = 0.f;
              o[7] = 0.f;
            }
          }
        }
      }
      consumer_bar();  // (2) norms of step s-1 are in smem
      if (s >= 1) {
        // ---------------- phase B(s-1) ----------------
        const int sp = s - 1;
        const int slot = sp % kStages;
        unsigned char* st = stage0 + (size_t)slot * L.bytes();
        const int64_t* ids = reinterpret_cast<const int64_t*>(st + L.ids_off());
        const float* lab = reinterpret_cast<const float*>(st + L.label_off());
        const float* sc = reinterpret_cast<const float*>(st + L.score_off());
        const float* rows = reinterpret_cast<const float*>(st + L.rows_off());
        const float cu = bcast[0], ci = bcast[1];
        for (int base = cw * kRowsPerWarp; base < cnt; base += kConsumerWarps * kRowsPerWarp) {
          const int j = base + grp;
          if (j >= cnt) continue;
          const int64_t u = ids[j], ia = ids[L.slice + j];
          const int64_t ib = PAIRWISE ? ids[2 * L.slice + j] : 0;
          const bool oku = (uint64_t)u < (uint64_t)a.n_users;
          const bool oka = (uint64_t)ia < (uint64_t)a.n_items;
          const bool okb = PAIRWISE && (uint64_t)ib < (uint64_t)a.n_items;
          float c;
          if (PAIRWISE) {
            const float sg = sigmoidf_(sc[j] - sc[L.slice + j]);
            c = -g * inv_b * (sg * (1.f - sg)) / (a.gamma + sg);
          } else if (a.loss_kind == XDR_LOSS_MSE) {
            c = g * inv_b * 2.f * (sc[j] - lab[j]);
          } else if (a.loss_kind == XDR_LOSS_BCE_SIGMOID) {
            const float p = sigmoidf_(sc[j]), y = lab[j];
            const float pq = p * (1.f - p);
            c = g * inv_b * (p - y) / fmaxf(pq, 1e-12f) * pq;
          } else {
            c = 0.f;
          }
          const float* pu = rows + (size_t)j * row_f;
          const float* pa = rows + ((size_t)L.slice + j) * row_f;
          const float* pb = rows + ((size_t)2 * L.slice + j) * row_f;
          const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            const int cidx = sub + v * kLanesPerRow;
            if (cidx >= a.nv) continue;
            const float4 ru = oku ? ld_row4(pu, cidx) : z;
            const float4 ra = oka ? ld_row4(pa, cidx) : z;
            if (PAIRWISE) {
              const float4 rb = okb ? ld_row4(pb, cidx) : z;
              if (oku) red_add4(a.user_dst + u * (int64_t)row_f, cidx, axpy4(cu, ru, scale4(c, sub4(ra, rb))));
              if (oka) red_add4(a.item_dst + ia * (int64_t)row_f, cidx, axpy4(ci, ra, scale4(c, ru)));
              if (okb) red_add4(a.item_dst + ib * (int64_t)row_f, cidx, scale4(-c, ru));
            } else {
              if (oku) red_add4(a.user_dst + u * (int64_t)row_f, cidx, axpy4(cu, ru, scale4(c, ra)));
              if (oka) red_add4(a.item_dst + ia * (int64_t)row_f, cidx, axpy4(ci, ra, scale4(c, ru)));
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);  // stage may be refilled by the producer
      }
      (void)ctid;
    }
  }
}

template <bool PAIRWISE>
static size_t steps_smem_bytes(int slice, int nv) {
  StageLayout<PAIRWISE> L(slice, nv);
  return kHeaderBytes + 4 * sizeof(float) * (kConsumerWarps + 2) + 128 + (size_t)kStages * L.bytes();
}

struct StepsPlan {
  int grid, slice;
  size_t smem;
};

// Choose the slice so that kStages stages fit in shared memory; returns false when the configuration does not fit
// the persistent kernel (large batch or wide rows) -- the caller then uses the per-step kernels.
static bool plan_steps(int64_t batch, int nv, bool pairwise, StepsPlan* plan) {
  const int sms = sm_count();
  int64_t slice = (batch + sms - 1) / sms;
  slice = (slice + 3) & ~(int64_t)3;  // multiple of 4: id tiles 32-byte granular, one warp pass = 4 interactions
  if (slice < 4) slice = 4;
  const int64_t grid = (batch + slice - 1) / slice;
  const size_t smem = pairwise ? steps_smem_bytes<true>((int)slice, nv) : steps_smem_bytes<false>((int)slice, nv);
  if (smem > 220 * 1024 || grid > sms) return false;
  plan->grid = (int)grid;
  plan->slice = (int)slice;
  plan->smem = smem;
  return true;
}

template <int VEC, bool PW>
static int launch_steps(const StepsArgs& a, const StepsPlan& plan, cudaStream_t s) {
  auto kern = train_steps_kernel<VEC, PW>;
  XDR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
  kern<<<plan.grid, kStepThreads, plan.smem, s>>>(a);
  return XDR_OK;
}

}  // namespace xdr

using namespace xdr;

extern "C" {

size_t xdr_steps_workspace_bytes(int n_steps) {
  if (n_steps < 0) return 0;
  // [n_steps] arrival counters (padded to 256 B) + [n_steps][<= 2048 CTAs -> sm count] float4 partials
  const size_t counters = (((size_t)n_steps * sizeof(unsigned int)) + 255) & ~(size_t)255;
  return counters + (size_t)n_steps * (size_t)sm_count() * sizeof(float4);
}

int xdr_train_steps(const float* user_tab, const float* item_tab, int64_t n_users, int64_t n_items, int dim,
                    const int64_t* user, const int64_t* item_a, const int64_t* item_b, const float* label,
                    int64_t step_stride, int64_t batch, int n_steps, int pairwise, int loss_kind, float gamma,
                    float reg_weight, const float* grad_loss, float scale, float* user_dst, float* item_dst, float* out8,
                    void* steps_ws, size_t steps_ws_bytes, int32_t* oob, xdr_stream_t stream) {
  XDR_REQUIRE(dim_ok(dim), "xdr_train_steps: dim=%d must be a multiple of 4 in (0, 256]", dim);
  XDR_REQUIRE(batch > 0 && n_steps >= 0, "xdr_train_steps: batch=%lld n_steps=%d", (long long)batch, n_steps);
  if (n_steps == 0) return XDR_OK;
  XDR_REQUIRE(user_tab && item_tab && user && item_a && user_dst && item_dst && out8 && steps_ws,
              "xdr_train_steps: null pointer");
  XDR_REQUIRE(!pairwise || item_b, "xdr_train_steps: pairwise needs the negative-item ids");
  XDR_REQUIRE(pairwise || loss_kind == XDR_LOSS_NONE || label, "xdr_train_steps: label is required for this loss kind");
  XDR_REQUIRE(pairwise || (loss_kind >= XDR_LOSS_MSE && loss_kind <= XDR_LOSS_NONE), "xdr_train_steps: bad loss_kind");
  XDR_REQUIRE(aligned16(user_tab) && aligned16(item_tab) && aligned16(user_dst) && aligned16(item_dst),
              "xdr_train_steps: tables must be 16-byte aligned");
  XDR_REQUIRE(step_stride >= batch, "xdr_train_steps: step_stride=%lld < batch", (long long)step_stride);
  // TMA bulk copies of the id tiles need 16-byte aligned sources: even batch/stride and aligned base pointers
  XDR_REQUIRE((batch % 4) == 0 && (step_stride % 4) == 0 && aligned16(user) && aligned16(item_a) &&
                  (!pairwise || aligned16(item_b)) && (label == nullptr || aligned16(label)),
              "xdr_train_steps: batch and step_stride must be multiples of 4 and the id arrays 16-byte aligned");
  XDR_REQUIRE(steps_ws_bytes >= xdr_steps_workspace_bytes(n_steps), "xdr_train_steps: steps_ws too small (%zu < %zu)",
              steps_ws_bytes, xdr_steps_workspace_bytes(n_steps));
  StepsPlan plan;
  if (!plan_steps(batch, dim / 4, pairwise != 0, &plan)) {
    set_error("xdr_train_steps: batch=%lld dim=%d does not fit the persistent kernel's shared-memory stages; "
              "use the per-step entry points",
              (long long)batch, dim);
    return XDR_ERR_UNSUPPORTED;
  }
  StepsArgs a{};
  a.user_tab = user_tab; a.item_tab = item_tab; a.n_users = n_users; a.n_items = n_items; a.nv = dim / 4;
  a.user = user; a.item_a = item_a; a.item_b = item_b; a.label = label; a.step_stride = step_stride; a.batch = batch;
  a.n_steps = n_steps; a.loss_kind = loss_kind; a.gamma = gamma; a.reg_weight = reg_weight; a.out8 = out8;
  a.grad_loss = grad_loss; a.scale = scale; a.user_dst = user_dst; a.item_dst = item_dst; a.slice = plan.slice; a.oob = oob;
  const size_t counters = (((size_t)n_steps * sizeof(unsigned int)) + 255) & ~(size_t)255;
  a.arrive = reinterpret_cast<unsigned int*>(steps_ws);
  a.partials = reinterpret_cast<float4*>(reinterpret_cast<char*>(steps_ws) + counters);
  cudaStream_t s = (cudaStream_t)stream;
  XDR_CUDA_OK(cudaMemsetAsync(steps_ws, 0, counters, s));
  int rc = XDR_OK;
  XDR_DISPATCH_VEC(a.nv, (rc = pairwise ? launch_steps<VEC, true>(a, plan, s) : launch_steps<VEC, false>(a, plan, s)));
  if (rc != XDR_OK) return rc;
  XDR_LAUNCH_OK();
  return XDR_OK;
}

}  // extern "C"
