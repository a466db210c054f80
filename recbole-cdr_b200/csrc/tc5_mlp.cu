// tc5_mlp.cu -- A4 on the 5th-generation tensor cores: the EMCDR map step (gather -> Linear -> tanh -> Linear -> MSE against
// the target rows -> whole backward -> scatter-add) as ONE kernel whose six products per 128-row tile all run on tcgen05.mma
// with accumulators in tensor memory.
//
// Reference: EMCDR.calculate_map_loss emcdr.py:156-168 with the mapping of emcdr.py:58-64,86-93 (Linear(D, 128) + Tanh +
// Linear(128, D); the target embedding is NOT detached).  Same contract as xdr_fused_mlp_step / xdr_tc_mlp_step for
// in_mode 0, head 0, two layers [D, 128, D], D % 16 == 0, D <= 64 (the yaml stack at every embedding size up to 64).
//
// Per tile of 128 batch rows (thread = row = TMEM lane in every epilogue):
//   G1  Z1 = X  W1^T      [128 x 128]   then  H = tanh(Z1 + b1)                          -> tile Hs
//   G2  Y  = H  W2^T      [128 x D]     then  dY = gs (Y + b2 - T),  loss,  dT -= dY      -> tile dYs
//   G3  dH = dY W2        [128 x 128]   then  dZ1 = dH (1 - H^2)                          -> tile Hs (in place)
//   G5  dW2^T += H^T dY   [128 x D]     accumulator RESIDENT in tensor memory over all tiles of the CTA
//   G4  dX = dZ1 W1       [128 x D]     then  scatter-add into the source table's gradient rows
//   G6  dW1 += dZ1^T X    [128 x D]     accumulator resident in tensor memory
// Operands are bf16 hi / lo planes (x ~= hi + lo) and every product is three kind::f16 MMAs (bf16x3, ~2^-16 relative per
// product, fp32 accumulation) -- see tc5.cuh.  Every tile and both weight matrices are stored ONCE, row-block-major
// (tc5::RowBlock16), and used through their K-major view by the forward / input-gradient products and through their
// MN-major view by the weight-gradient products and by the products with the transposed weights: no transposed copies.
// Shared memory at D = 64: 4 x 16 KB weights + 32 KB X + 64 KB H + 32 KB dY = 194 KB.  Tensor memory: 320 of 512 columns.
// HBM traffic is the algorithmic 1032 B per row (ids + source and target rows read + both gradient rows written).
//
// STATUS: written in a session without GPU access on this repository's reading of the tcgen05 descriptors (tc5.cuh;
// scripts/ubench_tcgen05.cu is the hardware experiment for that reading).  Logic checked under the CPU emulator against the
// oracle; NOT yet executed on hardware; opt-in only (config `xdr_fused_mlp: 'tc5'`), gpu tests marked `unvalidated`.
#include "mlp_args.cuh"
#include "tc5.cuh"
#include "tc_tile.cuh"

namespace xdr {

constexpr int kM5Rows = 128, kM5Hidden = 128, kM5MaxDim = 64, kM5TmemCols = 512;
// 512 threads: thread = (row r = tid % 128, column group cg = tid / 128).  The four threads of a row live in warps w, w + 4, w + 8,
// w + 12 -- the warps that may read the row's TMEM lane (lanes 32 (w % 4) ..) -- and share an epilogue's 16-column chunks
// (chunk c16 belongs to group c16 % 4).  Measured with 128 threads: 45 % of the stall samples were instruction fetches of the
// long unrolled epilogues, 6 % warp occupancy; sixteen warps run the same code four times shorter.
constexpr int kM5Threads = 512, kM5Groups = kM5Threads / kM5Rows;
constexpr int kM5ColZ1 = 0, kM5ColY = 128, kM5ColDW1 = 192, kM5ColDW2 = 256;   // tensor-memory columns

struct M5Smem {
  int b1, b2, dbs, W1h, W1l, W2h, W2l, Xh, Xl, Hh, Hl, Yh, Yl, total;   // byte offsets
};
__host__ __device__ inline M5Smem m5_smem_layout(int D) {
  M5Smem s{};
  int p = 128;                                  // [0, 8): mbarrier, [8, 12): tensor-memory base address
  s.b1 = p; p += kM5Hidden * 4;
  s.b2 = p; p += kM5MaxDim * 4;
  s.dbs = p; p += (kM5Hidden + kM5MaxDim) * 4;
  p = (p + 127) & ~127;
  const int wbytes = kM5Hidden * D * 2, xbytes = kM5Rows * D * 2, hbytes = kM5Rows * kM5Hidden * 2;
  s.W1h = p; p += wbytes; s.W1l = p; p += wbytes;
  s.W2h = p; p += wbytes; s.W2l = p; p += wbytes;
  s.Xh = p; p += xbytes; s.Xl = p; p += xbytes;
  s.Hh = p; p += hbytes; s.Hl = p; p += hbytes;
  s.Yh = p; p += xbytes; s.Yl = p; p += xbytes;
  s.total = p;
  return s;
}

#if defined(__CUDACC__) || defined(XDR_EMU)

// [R][C] row-major fp32 in global memory -> bf16 hi / lo row-block-major tile (all threads of the CTA)
__device__ __forceinline__ void m5_stage_matrix(const float* __restrict__ W, int R, int C, unsigned char* hi, unsigned char* lo) {
  const tc5::RowBlock16 t{R, C};
  const int c8n = C >> 3, n = R * c8n;
  constexpr int kU = 8;   // loads of eight 32-byte pieces are in flight before the first conversion
  for (int e0 = threadIdx.x; e0 < n; e0 += kU * kM5Threads) {
    float4 v0[kU], v1[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int e = e0 + u * kM5Threads;
      v0[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      v1[u] = v0[u];
      if (e < n) {
        const int r = e / c8n, c8 = e - r * c8n;
        v0[u] = __ldg(reinterpret_cast<const float4*>(W + (size_t)r * C + 8 * c8));
        v1[u] = __ldg(reinterpret_cast<const float4*>(W + (size_t)r * C + 8 * c8 + 4));
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int e = e0 + u * kM5Threads;
      if (e < n) {
        const int r = e / c8n, c8 = e - r * c8n;
        tc5::store_split8(hi, lo, t.chunk_offset(r, c8), v0[u], v1[u]);
      }
    }
  }
}

// Column sums over the 32 lanes of a warp of 16 values per lane in 16 shuffles (instead of 16 x 5): every exchange halves the
// columns a lane keeps.  Returns the sum of column `col` (the same in both lanes of a pair; the even lane publishes it).
__device__ __forceinline__ float m5_warp_colsum16(const float (&v)[16], int lane, int& col) {
  float a8[8], b4[4], c2[2];
  const bool h16 = (lane & 16) != 0, h8 = (lane & 8) != 0, h4 = (lane & 4) != 0, h2 = (lane & 2) != 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) a8[j] = (h16 ? v[j + 8] : v[j]) + __shfl_xor_sync(0xffffffffu, h16 ? v[j] : v[j + 8], 16);
#pragma unroll
  for (int j = 0; j < 4; ++j) b4[j] = (h8 ? a8[j + 4] : a8[j]) + __shfl_xor_sync(0xffffffffu, h8 ? a8[j] : a8[j + 4], 8);
#pragma unroll
  for (int j = 0; j < 2; ++j) c2[j] = (h4 ? b4[j + 2] : b4[j]) + __shfl_xor_sync(0xffffffffu, h4 ? b4[j] : b4[j + 2], 4);
  float d = (h2 ? c2[1] : c2[0]) + __shfl_xor_sync(0xffffffffu, h2 ? c2[0] : c2[1], 2);
  d += __shfl_xor_sync(0xffffffffu, d, 1);
  col = (h16 ? 8 : 0) + (h8 ? 4 : 0) + (h4 ? 2 : 0) + (h2 ? 1 : 0);
  return d;
}

__device__ __forceinline__ float m5_bf16_pair_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float m5_bf16_pair_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

__global__ void __launch_bounds__(kM5Threads, 1) tc5_mlp_kernel(MlpArgs a, Workspace ws) {
  // backward == 2, the "correction" pass of an eager step (see xdr.h): the gradients were already accumulated with an upstream
  // gradient of 1 by a backward == 1 launch at forward time; this launch adds (g - 1) times the same -- and is over before it
  // allocates anything when g == 1, which is what loss.backward() passes
  float g_up = (a.grad_loss ? __ldg(a.grad_loss) : 1.0f);
  if (a.backward == 2) {
    g_up -= 1.0f;
    if (g_up == 0.f) return;
  }
  XDR_DYN_SMEM_ALIGNED(unsigned char, smem_m5, 128);
  __shared__ float red_smem[kM5Threads / 32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rt = tid & (kM5Rows - 1), cg = tid >> 7;   // this thread's tile row (= TMEM lane) and column group
  const int D = a.dim, c8n = D >> 3;
  const M5Smem lay = m5_smem_layout(D);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_m5);
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(smem_m5 + 8);
  float* b1 = reinterpret_cast<float*>(smem_m5 + lay.b1);
  float* b2 = reinterpret_cast<float*>(smem_m5 + lay.b2);
  float* dbs = reinterpret_cast<float*>(smem_m5 + lay.dbs);      // [0, 128): db1, [128, 128 + D): db2
  unsigned char *W1h = smem_m5 + lay.W1h, *W1l = smem_m5 + lay.W1l, *W2h = smem_m5 + lay.W2h, *W2l = smem_m5 + lay.W2l;
  unsigned char *Xh = smem_m5 + lay.Xh, *Xl = smem_m5 + lay.Xl, *Hh = smem_m5 + lay.Hh, *Hl = smem_m5 + lay.Hl;
  unsigned char *Yh = smem_m5 + lay.Yh, *Yl = smem_m5 + lay.Yl;
  const tc5::RowBlock16 tW1{kM5Hidden, D}, tW2{D, kM5Hidden}, tX{kM5Rows, D}, tH{kM5Rows, kM5Hidden};

  // ---- resident parameters, barrier, tensor memory ---------------------------------------------------------------------------
  m5_stage_matrix(a.W[0], kM5Hidden, D, W1h, W1l);
  m5_stage_matrix(a.W[1], D, kM5Hidden, W2h, W2l);
  for (int n = tid; n < kM5Hidden; n += kM5Threads) b1[n] = a.b[0] ? a.b[0][n] : 0.f;
  for (int n = tid; n < D; n += kM5Threads) b2[n] = a.b[1] ? a.b[1][n] : 0.f;
  for (int n = tid; n < kM5Hidden + kM5MaxDim; n += kM5Threads) dbs[n] = 0.f;
  if (tid == 0) {
    tc5::mbar_init(bar, 1);
    tc5::mbar_init_fence();
  }
  if (warp == 0) tc5::tmem_alloc(tmem_base_smem, kM5TmemCols);
  tc5::fence_proxy_async();
  tc5::fence_before_sync();
  __syncthreads();
  tc5::fence_after_sync();
  const uint32_t tmem = *tmem_base_smem;
  const uint32_t my_lanes = (uint32_t)((warp & 3) * 32) << 16;   // this warp's 32 TMEM lanes = tile rows 32 (warp % 4) ..

  const uint32_t w1h = tc5::smem_u32(W1h), w1l = tc5::smem_u32(W1l), w2h = tc5::smem_u32(W2h), w2l = tc5::smem_u32(W2l);
  const uint32_t xh = tc5::smem_u32(Xh), xl = tc5::smem_u32(Xl), hh = tc5::smem_u32(Hh), hl = tc5::smem_u32(Hl);
  const uint32_t yh = tc5::smem_u32(Yh), yl = tc5::smem_u32(Yl);
  uint32_t phase = 0;
  // one thread issues, everybody waits for completion (commit -> mbarrier) before touching what the products read or wrote
  auto wait_mma = [&]() {
    tc5::mbar_wait(bar, phase);
    phase ^= 1u;
    tc5::fence_after_sync();
    __syncwarp();
  };
  // end of an epilogue: TMEM reads retired, shared-memory writes visible to the tensor core, everybody done
  auto epilogue_done = [&]() {
    tc5::fence_proxy_async();
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
  };

  const float gs = g_up * 2.0f / ((float)a.batch * (float)D);
  float loss_acc[1] = {0.f};
  const int64_t n_tiles = (a.batch + kM5Rows - 1) / kM5Rows;
  bool first = true;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row = tile * kM5Rows + rt;         // this thread's batch row in every epilogue
    int64_t my_id = -1;
    if (row < a.batch) {
      my_id = a.idx_u[row];
      if ((uint64_t)my_id >= (uint64_t)a.n_u) {
        if (a.oob) *a.oob = 1;
        my_id = -1;
      }
    }
    // ---- gather the source rows -> X tile (rows past the batch / bad ids are zero).  All of a thread's loads are issued before
    // the first conversion (8 threads per row, 16 rows per pass), and the thread's own target row -- needed two products later
    // -- is requested here too, so that its latency hides behind G1 / G2 instead of stalling the loss epilogue
    float4 trow[4];   // the 16 target columns of chunk c16 = cg (D = 64: one chunk per group; smaller D: the high groups idle)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      trow[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (my_id >= 0 && cg * 16 < D) trow[q] = ld_row4(a.T + my_id * D, 4 * cg + q);
    }
    constexpr int kGatherPasses = kM5Rows * (kM5MaxDim / 8) / kM5Threads;   // 2 at D = 64
    {
      float4 g0[kGatherPasses], g1[kGatherPasses];
#pragma unroll
      for (int p = 0; p < kGatherPasses; ++p) {
        const int e = tid + p * kM5Threads;
        g0[p] = make_float4(0.f, 0.f, 0.f, 0.f);
        g1[p] = g0[p];
        if (e < kM5Rows * c8n) {
          const int r = e / c8n, c8 = e - r * c8n;
          const int64_t gr = tile * kM5Rows + r;
          if (gr < a.batch) {
            const int64_t id = a.idx_u[gr];
            if ((uint64_t)id < (uint64_t)a.n_u) {
              g0[p] = ld_row4(a.Au + id * D, 2 * c8);
              g1[p] = ld_row4(a.Au + id * D, 2 * c8 + 1);
            }
          }
        }
      }
#pragma unroll
      for (int p = 0; p < kGatherPasses; ++p) {
        const int e = tid + p * kM5Threads;
        if (e < kM5Rows * c8n) {
          const int r = e / c8n, c8 = e - r * c8n;
          tc5::store_split8(Xh, Xl, tX.chunk_offset(r, c8), g0[p], g1[p]);
        }
      }
    }
    epilogue_done();

    // ---- G1: Z1 = X W1^T ----------------------------------------------------------------------------------------------------------
    if (tid == 0) {
      tc5::mma_bf16x3(tmem + kM5ColZ1, xh, xl, tX.as_k_major(), w1h, w1l, tW1.as_k_major(),
                      tc5::make_idesc_bf16(kM5Rows, kM5Hidden, false, false), D, false);
      tc5::commit(bar);
    }
    wait_mma();
    // H = act(Z1 + b1) -> Hs
#pragma unroll 1
    for (int c16 = cg; c16 < kM5Hidden / 16; c16 += kM5Groups) {
      uint32_t r[16];
      tc5::tmem_ld16(tmem + my_lanes + kM5ColZ1 + c16 * 16, r);
      tc5::tmem_ld_wait();
      float h[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) h[j] = act_apply(__uint_as_float(r[j]) + b1[c16 * 16 + j], a.hidden_act);
      tc5::store_split8(Hh, Hl, tH.chunk_offset(rt, 2 * c16), make_float4(h[0], h[1], h[2], h[3]), make_float4(h[4], h[5], h[6], h[7]));
      tc5::store_split8(Hh, Hl, tH.chunk_offset(rt, 2 * c16 + 1), make_float4(h[8], h[9], h[10], h[11]),
                        make_float4(h[12], h[13], h[14], h[15]));
    }
    epilogue_done();

    // ---- G2: Y = H W2^T -----------------------------------------------------------------------------------------------------------
    if (tid == 0) {
      tc5::mma_bf16x3(tmem + kM5ColY, hh, hl, tH.as_k_major(), w2h, w2l, tW2.as_k_major(),
                      tc5::make_idesc_bf16(kM5Rows, D, false, false), kM5Hidden, false);
      tc5::commit(bar);
    }
    wait_mma();
    // loss head: d = Y + b2 - T[id]; loss += d^2; dY = gs d; the target rows get -dY (emcdr.py:156-168: target not detached)
    if (cg * 16 < D) {   // one 16-column chunk per column group (warp-uniform: a warp's threads share cg)
      const int c16 = cg;
      uint32_t r[16];
      tc5::tmem_ld16(tmem + my_lanes + kM5ColY + c16 * 16, r);
      tc5::tmem_ld_wait();
      float g[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 t = trow[q];
        const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float d = 0.f;
          if (row < a.batch) d = __uint_as_float(r[4 * q + j]) + b2[c16 * 16 + 4 * q + j] - tv[j];
          loss_acc[0] += d * d;
          g[4 * q + j] = gs * d;
        }
        if (a.backward && my_id >= 0)
          red_add4(a.dT + my_id * D, 4 * c16 + q, scale4(-a.scale, make_float4(g[4 * q], g[4 * q + 1], g[4 * q + 2], g[4 * q + 3])));
      }
      if (a.backward) {
        tc5::store_split8(Yh, Yl, tX.chunk_offset(rt, 2 * c16), make_float4(g[0], g[1], g[2], g[3]), make_float4(g[4], g[5], g[6], g[7]));
        tc5::store_split8(Yh, Yl, tX.chunk_offset(rt, 2 * c16 + 1), make_float4(g[8], g[9], g[10], g[11]),
                          make_float4(g[12], g[13], g[14], g[15]));
        {   // db2: column sums over the tile's rows
          int col;
          const float s = m5_warp_colsum16(g, lane, col);
          if ((lane & 1) == 0) atomicAdd(&dbs[kM5Hidden + c16 * 16 + col], s);
        }
      }
    }
    epilogue_done();
    if (!a.backward) continue;

    // ---- G3: dH = dY W2;  G5: dW2^T += H^T dY ----------------------------------------------------------------------------------------
    if (tid == 0) {
      tc5::mma_bf16x3(tmem + kM5ColZ1, yh, yl, tX.as_k_major(), w2h, w2l, tW2.as_mn_major(),
                      tc5::make_idesc_bf16(kM5Rows, kM5Hidden, false, true), D, false);
      tc5::mma_bf16x3(tmem + kM5ColDW2, hh, hl, tH.as_mn_major(), yh, yl, tX.as_mn_major(),
                      tc5::make_idesc_bf16(kM5Hidden, D, true, true), kM5Rows, !first);
      tc5::commit(bar);
    }
    wait_mma();
    // dZ1 = dH * act'(H), written over H in place (each thread rewrites the chunks of its own row)
#pragma unroll 1
    for (int c16 = cg; c16 < kM5Hidden / 16; c16 += kM5Groups) {
      uint32_t r[16];
      tc5::tmem_ld16(tmem + my_lanes + kM5ColZ1 + c16 * 16, r);
      tc5::tmem_ld_wait();
      float z[16];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int off = tH.chunk_offset(rt, 2 * c16 + half);
        const uint4 ph = *reinterpret_cast<const uint4*>(Hh + off), pl = *reinterpret_cast<const uint4*>(Hl + off);
        const uint32_t hw[4] = {ph.x, ph.y, ph.z, ph.w}, lw[4] = {pl.x, pl.y, pl.z, pl.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float h0 = m5_bf16_pair_lo(hw[q]) + m5_bf16_pair_lo(lw[q]), h1 = m5_bf16_pair_hi(hw[q]) + m5_bf16_pair_hi(lw[q]);
          z[8 * half + 2 * q] = __uint_as_float(r[8 * half + 2 * q]) * act_grad(h0, a.hidden_act);
          z[8 * half + 2 * q + 1] = __uint_as_float(r[8 * half + 2 * q + 1]) * act_grad(h1, a.hidden_act);
        }
      }
      tc5::store_split8(Hh, Hl, tH.chunk_offset(rt, 2 * c16), make_float4(z[0], z[1], z[2], z[3]), make_float4(z[4], z[5], z[6], z[7]));
      tc5::store_split8(Hh, Hl, tH.chunk_offset(rt, 2 * c16 + 1), make_float4(z[8], z[9], z[10], z[11]),
                        make_float4(z[12], z[13], z[14], z[15]));
      {     // db1
        int col;
        const float s = m5_warp_colsum16(z, lane, col);
        if ((lane & 1) == 0) atomicAdd(&dbs[c16 * 16 + col], s);
      }
    }
    epilogue_done();

    // ---- G4: dX = dZ1 W1;  G6: dW1 += dZ1^T X -----------------------------------------------------------------------------------------
    if (tid == 0) {
      tc5::mma_bf16x3(tmem + kM5ColY, hh, hl, tH.as_k_major(), w1h, w1l, tW1.as_mn_major(),
                      tc5::make_idesc_bf16(kM5Rows, D, false, true), kM5Hidden, false);
      tc5::mma_bf16x3(tmem + kM5ColDW1, hh, hl, tH.as_mn_major(), xh, xl, tX.as_mn_major(),
                      tc5::make_idesc_bf16(kM5Hidden, D, true, true), kM5Rows, !first);
      tc5::commit(bar);
    }
    wait_mma();
#pragma unroll 1
    for (int c16 = cg; c16 < D / 16; c16 += kM5Groups) {
      uint32_t r[16];
      tc5::tmem_ld16(tmem + my_lanes + kM5ColY + c16 * 16, r);
      tc5::tmem_ld_wait();
      if (my_id >= 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          red_add4(a.dAu + my_id * D, 4 * c16 + q,
                   scale4(a.scale, make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                                               __uint_as_float(r[4 * q + 3]))));
      }
      __syncwarp();
    }
    epilogue_done();
    first = false;
  }

  // ---- flush the weight / bias gradients (thread = hidden unit = TMEM lane) -------------------------------------------------------------
  if (a.backward) {
#pragma unroll 1
    for (int c16 = cg; c16 < D / 16; c16 += kM5Groups) {
      uint32_t r[16];
      tc5::tmem_ld16(tmem + my_lanes + kM5ColDW1 + c16 * 16, r);
      tc5::tmem_ld_wait();
      if (a.dW[0])   // dW1 is [128][D] row-major: this thread's 16 columns are contiguous -> four 128-bit reductions
#pragma unroll
        for (int q = 0; q < 4; ++q)
          red_add4(a.dW[0] + (size_t)rt * D, 4 * c16 + q,
                   make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                               __uint_as_float(r[4 * q + 3])));
      tc5::tmem_ld16(tmem + my_lanes + kM5ColDW2 + c16 * 16, r);
      tc5::tmem_ld_wait();
      if (a.dW[1])
#pragma unroll
        for (int j = 0; j < 16; ++j) atomicAdd(&a.dW[1][(size_t)(c16 * 16 + j) * kM5Hidden + rt], __uint_as_float(r[j]));
    }
    if (a.db[0] && tid < kM5Hidden) atomicAdd(&a.db[0][tid], dbs[tid]);
    if (a.db[1] && tid < D) atomicAdd(&a.db[1][tid], dbs[kM5Hidden + tid]);
  }
  tc5::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc5::tmem_dealloc(tmem, kM5TmemCols);

  const double denom = (double)a.batch * (double)D;
  float* out8 = a.out8;
  grid_reduce_last_block<1>(loss_acc, ws, red_smem, [=](double* tot) {
    out8[0] = (float)(tot[0] / denom);
    for (int i = 1; i < 8; ++i) out8[i] = 0.f;
  });
}

#endif  // __CUDACC__ || XDR_EMU

static bool m5_stack_ok(int n_layers, const int* dims) {
  return n_layers == 2 && dims != nullptr && dims[1] == kM5Hidden && dims[0] == dims[2] && dims[0] >= 16 && dims[0] <= kM5MaxDim &&
         dims[0] % 16 == 0;
}

}  // namespace xdr

using namespace xdr;

extern "C" {

int xdr_tc5_mlp_supported(int n_layers, const int* dims_host) { return m5_stack_ok(n_layers, dims_host) ? 1 : 0; }

int xdr_tc5_mlp_step(int n_layers, const int* dims_host, const float* const* W_host, const float* const* b_host,
                     float* const* dW_host, float* const* db_host, int hidden_act, int in_mode, int head, const float* Au,
                     const float* Bu, const float* Ai, const float* Bi, const float* T, int64_t n_u, int64_t n_i, int dim,
                     const int64_t* idx_u, const int64_t* idx_i, const float* label, int64_t batch, int backward,
                     const float* grad_loss, float scale, float* dAu, float* dBu, float* dAi, float* dBi, float* dT,
                     float* prob, float* out8, void* ws, int32_t* oob, xdr_stream_t stream) {
  (void)Bu; (void)Ai; (void)Bi; (void)n_i; (void)idx_i; (void)label; (void)dBu; (void)dAi; (void)dBi; (void)prob;
  XDR_REQUIRE(m5_stack_ok(n_layers, dims_host), "xdr_tc5_mlp_step: unsupported layer stack (needs [D, 128, D], D %% 16 == 0, D <= 64)");
  XDR_REQUIRE(in_mode == 0 && head == 0, "xdr_tc5_mlp_step: only the single-table input with the MSE head (the EMCDR map step)");
  XDR_REQUIRE(dim == dims_host[0] && batch > 0, "xdr_tc5_mlp_step: bad dim/batch");
  XDR_REQUIRE(W_host && idx_u && out8 && ws && Au && T, "xdr_tc5_mlp_step: null pointer");
  XDR_REQUIRE(backward >= 0 && backward <= 2, "xdr_tc5_mlp_step: backward must be 0, 1 or 2 (correction pass)");
  XDR_REQUIRE(backward != 2 || grad_loss, "xdr_tc5_mlp_step: the correction pass needs grad_loss");
  XDR_REQUIRE(!backward || (dAu && dT), "xdr_tc5_mlp_step: null destination");
  XDR_REQUIRE(aligned16(Au) && aligned16(T) && (!backward || (aligned16(dAu) && aligned16(dT))),
              "xdr_tc5_mlp_step: tables must be 16-byte aligned");
  MlpArgs a{};
  a.n_layers = 2;
  for (int l = 0; l <= 2; ++l) a.dims[l] = dims_host[l];
  for (int l = 0; l < 2; ++l) {
    XDR_REQUIRE(W_host[l] && aligned16(W_host[l]), "xdr_tc5_mlp_step: null or misaligned weight");
    a.W[l] = W_host[l];
    a.b[l] = b_host ? b_host[l] : nullptr;
    a.dW[l] = (backward && dW_host) ? dW_host[l] : nullptr;
    a.db[l] = (backward && db_host) ? db_host[l] : nullptr;
  }
  a.hidden_act = hidden_act; a.last_act = XDR_ACT_NONE; a.in_mode = 0; a.head = 0;
  a.Au = Au; a.T = T; a.n_u = n_u; a.dim = dim; a.idx_u = idx_u; a.batch = batch; a.backward = backward;
  a.grad_loss = grad_loss; a.scale = scale; a.dAu = dAu; a.dT = dT; a.out8 = out8; a.oob = oob; a.tile_rows = kM5Rows;
  const size_t smem = (size_t)m5_smem_layout(dim).total;
  XDR_CUDA_OK(cudaFuncSetAttribute(tc5_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t n_tiles = (batch + kM5Rows - 1) / kM5Rows;
  int grid = sm_count();
  if (grid > n_tiles) grid = (int)n_tiles;
  XDR_LAUNCH((tc5_mlp_kernel), grid, kM5Threads, smem, (cudaStream_t)stream, a, Workspace(ws));
  XDR_LAUNCH_OK();
  return XDR_OK;
}

}  // extern "C"
