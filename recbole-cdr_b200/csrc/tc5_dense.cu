// tc5_dense.cu -- the dense-layer family (xdr_dense_fwd / xdr_dense_bwd_input / xdr_dense_bwd_weight, include/xdr.h) on the
// 5th-generation tensor cores: tcgen05.mma kind::f16 on THREE bf16 operand planes (x ~= hi + mid + lo, 24 mantissa bits) and the
// six products that matter (tc5.cuh mma_bf16x6; fp32 accumulation in tensor memory): fp32-faithful, so that ReLU masks and with
// them whole gradient rows agree with the fp32 reference (bf16x3 did not: round 2, GPU call 12).
//
// These are the GEMMs of the mapping MLP (emcdr.py:86-93), the CoNet cross-stitch units (conet.py:118-138), the NeuMF towers
// (recbole MLPLayers, dtcdr.py:61-67) and the MLPs of the further models: a long M (the batch, 8 192 .. 32 768 rows) against a
// short N / K (16 .. 256).  One CTA owns a 128-row tile of the batch, so the MMA's M is the batch and nothing is padded:
//   forward      Y  = act(X W^T + b + m (X2 W2^T)): ONE accumulator over the concatenated reduction [X | m X2] [W | W2]^T,
//                K-chunks of 64 staged by all threads (fp32 -> bf16 hi / lo, row-block-major tiles), issued by one thread
//   input grad   dX = m (dZ W): the dZ tile is the K-major A operand, W (row-major [N, K]) is read through the MN-major view of
//                its tile -- no transposed copy
//   weight grad  dW += (m dZ)^T X: both operands through their MN-major views (the reduction index is the batch row); a CTA
//                keeps its [N, K] accumulator RESIDENT in tensor memory over all its row tiles and flushes once (RED.128)
// Operands are small enough for two or three CTAs per SM (64 .. 96 KB of shared memory, <= 256 TMEM columns), which is what
// hides the staging latency: there is deliberately no intra-CTA pipeline.
//
// Layouts and descriptors: tc5.cuh (validated on a B200, profiles/r2_ubench_tcgen05.txt; the same building blocks drive
// tc5_mlp.cu).  Shapes this engine takes: M >= 128, N % 16 == 0, 16 <= N <= 128, K % 16 == 0, 16 <= K <= 256, 16-byte
// aligned contiguous operands; everything else stays on the fp32 FMA kernels of dense.cu.
#include "tc5.cuh"
#include <stdlib.h>
#include "tc5_dense.cuh"

namespace xdr {

constexpr int kD5Rows = 128, kD5Threads = 256;

__host__ __device__ inline int d5_chunk(int K) { return K % 64 == 0 ? 64 : (K % 48 == 0 ? 48 : (K % 32 == 0 ? 32 : 16)); }
__host__ __device__ inline uint32_t d5_tmem_cols(int n) { return n <= 32 ? 32u : (n <= 64 ? 64u : (n <= 128 ? 128u : 256u)); }

#if defined(__CUDACC__) || defined(XDR_EMU)

__device__ __forceinline__ float d5_act(float v, int act) {
  switch (act) {
    case XDR_ACT_RELU: return v > 0.f ? v : 0.f;
    case XDR_ACT_TANH: return tanhf(v);
    case XDR_ACT_SIGMOID: return sigmoidf_(v);
    default: return v;
  }
}

// rows [row0, row0 + t.rows) x columns [col0, col0 + cols) of the row-major fp32 matrix P (leading dimension ld, n_rows rows)
// -> the three bf16 planes (hi at `dst`, mid and lo `plane` bytes further each) of the row-block-major tile t (tile columns
// [0, cols); cols % 8 == 0).  Rows past n_rows and rows the mask switches off are zeros.  Four 32-byte pieces per thread are in
// flight before the first conversion; the rows were prefetched into L2 at kernel start (d5_prefetch), so these are L2 hits.
__device__ __forceinline__ void d5_stage(const float* __restrict__ P, int64_t ld, int64_t row0, int64_t n_rows, int col0,
                                         int cols, const tc5::RowBlock16& t, unsigned char* dst, int plane,
                                         const int64_t* __restrict__ mask_ids, int64_t mask_lt) {
  const int c8n = cols >> 3, n = t.rows * c8n;
  constexpr int kU = 4;
#pragma unroll 1
  for (int e0 = threadIdx.x; e0 < n; e0 += kU * kD5Threads) {
    float4 v0[kU], v1[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int e = e0 + u * kD5Threads;
      v0[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      v1[u] = v0[u];
      if (e < n) {
        const int r = e / c8n, c8 = e - r * c8n;
        const int64_t row = row0 + r;
        if (row < n_rows && (mask_ids == nullptr || mask_ids[row] < mask_lt)) {
          const float* p = P + row * ld + col0 + 8 * c8;
          v0[u] = ld_row4(p, 0);
          v1[u] = ld_row4(p, 1);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int e = e0 + u * kD5Threads;
      if (e < n) {
        const int r = e / c8n, c8 = e - r * c8n;
        tc5::store_split8_3(dst, dst + plane, dst + 2 * plane, t.chunk_offset(r, c8), v0[u], v1[u]);
      }
    }
  }
}

// asks L2 for rows [row0, row0 + rows) of P (row_bytes each, 128-byte lines): a CTA's whole input tile is on its way before
// the first staging pass needs it (one CTA per SM at these batch sizes: there is nobody else to hide the DRAM latency)
__device__ __forceinline__ void d5_prefetch(const float* __restrict__ P, int64_t ld, int64_t row0, int64_t n_rows, int rows,
                                            int row_bytes) {
#ifndef XDR_EMU
  const int lines = (row_bytes + 127) >> 7;
  for (int e = threadIdx.x; e < rows * lines; e += kD5Threads) {
    const int r = e / lines, l = e - r * lines;
    if (row0 + r < n_rows) {
      const char* p = reinterpret_cast<const char*>(P + (row0 + r) * ld) + 128 * l;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    }
  }
#endif
}

struct D5Sync {
  uint64_t* bar;
  uint32_t phase;
  // operands written by the generic proxy -> visible to the tensor core; everybody arrives
  __device__ __forceinline__ void operands_ready() {
    tc5::fence_proxy_async();
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
  }
  // every MMA committed so far has completed (operands may be overwritten, the accumulator may be read)
  __device__ __forceinline__ void wait_mma() {
    tc5::mbar_wait(bar, phase);
    phase ^= 1u;
    tc5::fence_after_sync();
    __syncwarp();
  }
};

// ---- forward -------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kD5Threads) tc5_dense_fwd_kernel(const float* __restrict__ X, const float* __restrict__ W,
                                                                   const float* __restrict__ bias, const float* __restrict__ X2,
                                                                   const float* __restrict__ W2, const int64_t* __restrict__ mask_ids,
                                                                   int64_t mask_lt, int act, float* __restrict__ Y, int64_t M, int N,
                                                                   int K) {
  XDR_DYN_SMEM_ALIGNED(unsigned char, smem_d5, 128);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KC = d5_chunk(K);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_d5);
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(smem_d5 + 8);
  const int xplane = kD5Rows * KC * 2, wplane = N * KC * 2;
  unsigned char* Xs = smem_d5 + 128;
  unsigned char* Ws = Xs + 3 * xplane;
  const tc5::RowBlock16 tX{kD5Rows, KC}, tW{N, KC};
  const int64_t row0 = (int64_t)blockIdx.x * kD5Rows;
  const uint32_t cols = d5_tmem_cols(N);
  d5_prefetch(X, K, row0, M, kD5Rows, K * 4);
  if (X2 != nullptr) d5_prefetch(X2, K, row0, M, kD5Rows, K * 4);
  if (tid == 0) {
    tc5::mbar_init(bar, 1);
    tc5::mbar_init_fence();
  }
  if (warp == 0) tc5::tmem_alloc(tmem_base_smem, cols);
  tc5::fence_before_sync();
  __syncthreads();
  tc5::fence_after_sync();
  const uint32_t tmem = *tmem_base_smem;
  D5Sync sy{bar, 0u};
  const uint32_t xs = tc5::smem_u32(Xs), ws = tc5::smem_u32(Ws);
  const uint32_t idesc = tc5::make_idesc_bf16(kD5Rows, N, false, false);
  const int n_prod = X2 != nullptr ? 2 : 1;
  bool acc = false;
#pragma unroll 1
  for (int p = 0; p < n_prod; ++p) {
    const float* Xp = p == 0 ? X : X2;
    const float* Wp = p == 0 ? W : W2;
#pragma unroll 1
    for (int kc = 0; kc < K; kc += KC) {
      d5_stage(Xp, K, row0, M, kc, KC, tX, Xs, xplane, p == 0 ? nullptr : mask_ids, mask_lt);
      d5_stage(Wp, K, 0, N, kc, KC, tW, Ws, wplane, nullptr, 0);
      sy.operands_ready();
      if (tid == 0) {
        tc5::mma_bf16x6(tmem, xs, xplane, tX.as_k_major(), ws, wplane, tW.as_k_major(), idesc, KC, acc);
        tc5::commit(bar);
      }
      acc = true;
      sy.wait_mma();
    }
  }
  // epilogue: thread = (row = TMEM lane, column group): + bias, activation, 64-byte pieces of the output row
  const int rt = (warp & 3) * 32 + lane, cg = warp >> 2;
  const int64_t row = row0 + rt;
  const uint32_t my_lanes = (uint32_t)((warp & 3) * 32) << 16;
#pragma unroll 1
  for (int c16 = cg; c16 < N / 16; c16 += kD5Threads / kD5Rows) {
    uint32_t r[16];
    tc5::tmem_ld16(tmem + my_lanes + c16 * 16, r);
    tc5::tmem_ld_wait();
    if (row < M) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int n = c16 * 16 + 4 * q + j;
          v[j] = d5_act(__uint_as_float(r[4 * q + j]) + (bias ? __ldg(bias + n) : 0.f), act);
        }
        st4(Y + row * N, 4 * c16 + q, make_float4(v[0], v[1], v[2], v[3]));
      }
    }
    __syncwarp();
  }
  tc5::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc5::tmem_dealloc(tmem, cols);
}

// ---- input gradient: dX[m, k] (=|+=) mask[m] * sum_n dZ[m, n] W[n, k] ----------------------------------------------------------
__host__ __device__ inline int d5_piece(int K) { return K <= 128 ? K : (K % 128 == 0 ? 128 : (K % 64 == 0 ? 64 : 16)); }

__global__ void __launch_bounds__(kD5Threads) tc5_dense_bwd_input_kernel(const float* __restrict__ dZ, const float* __restrict__ W,
                                                                         const int64_t* __restrict__ mask_ids, int64_t mask_lt,
                                                                         float* __restrict__ dX, int64_t M, int N, int K,
                                                                         int accumulate) {
  XDR_DYN_SMEM_ALIGNED(unsigned char, smem_d5, 128);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NP = d5_piece(K);   // output columns per product (the MMA's N)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_d5);
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(smem_d5 + 8);
  const int zplane = kD5Rows * N * 2, wplane = N * NP * 2;
  unsigned char* Zs = smem_d5 + 128;
  unsigned char* Ws = Zs + 3 * zplane;
  const tc5::RowBlock16 tZ{kD5Rows, N}, tW{N, NP};
  const int64_t row0 = (int64_t)blockIdx.x * kD5Rows;
  const uint32_t cols = d5_tmem_cols(NP);
  d5_prefetch(dZ, N, row0, M, kD5Rows, N * 4);
  if (tid == 0) {
    tc5::mbar_init(bar, 1);
    tc5::mbar_init_fence();
  }
  if (warp == 0) tc5::tmem_alloc(tmem_base_smem, cols);
  tc5::fence_before_sync();
  __syncthreads();
  tc5::fence_after_sync();
  const uint32_t tmem = *tmem_base_smem;
  D5Sync sy{bar, 0u};
  const uint32_t zs = tc5::smem_u32(Zs), ws = tc5::smem_u32(Ws);
  const uint32_t idesc = tc5::make_idesc_bf16(kD5Rows, NP, false, true);
  const int rt = (warp & 3) * 32 + lane, cg = warp >> 2;
  const int64_t row = row0 + rt;
  const uint32_t my_lanes = (uint32_t)((warp & 3) * 32) << 16;
  const float m = (row < M && (mask_ids == nullptr || mask_ids[row] < mask_lt)) ? 1.f : 0.f;
  d5_stage(dZ, N, row0, M, 0, N, tZ, Zs, zplane, nullptr, 0);
#pragma unroll 1
  for (int k0 = 0; k0 < K; k0 += NP) {
    d5_stage(W, K, 0, N, k0, NP, tW, Ws, wplane, nullptr, 0);
    sy.operands_ready();   // (also: every thread has finished the previous piece's epilogue reads of tensor memory)
    if (tid == 0) {
      tc5::mma_bf16x6(tmem, zs, zplane, tZ.as_k_major(), ws, wplane, tW.as_mn_major(), idesc, N, false);
      tc5::commit(bar);
    }
    sy.wait_mma();
#pragma unroll 1
    for (int c16 = cg; c16 < NP / 16; c16 += kD5Threads / kD5Rows) {
      uint32_t r[16];
      tc5::tmem_ld16(tmem + my_lanes + c16 * 16, r);
      tc5::tmem_ld_wait();
      if (row < M) {
        float* out = dX + row * K + k0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 v = make_float4(m * __uint_as_float(r[4 * q]), m * __uint_as_float(r[4 * q + 1]), m * __uint_as_float(r[4 * q + 2]),
                                 m * __uint_as_float(r[4 * q + 3]));
          if (accumulate) {
            const float4 o = ld_row4(out, 4 * c16 + q);
            v = make_float4(v.x + o.x, v.y + o.y, v.z + o.z, v.w + o.w);
          }
          st4(out, 4 * c16 + q, v);
        }
      }
      __syncwarp();
    }
    tc5::fence_before_sync();   // this piece's tensor-memory reads are ordered before the barrier of the next piece
  }
  tc5::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc5::tmem_dealloc(tmem, cols);
}

// ---- weight gradient: dW[n, k] += sum_m mask[m] dZ[m, n] X[m, k];  db[n] += sum_m mask[m] dZ[m, n] ------------------------------
__global__ void __launch_bounds__(kD5Threads) tc5_dense_bwd_weight_kernel(const float* __restrict__ dZ, const float* __restrict__ X,
                                                                          const int64_t* __restrict__ mask_ids, int64_t mask_lt,
                                                                          float* __restrict__ dW, float* __restrict__ db, int64_t M,
                                                                          int N, int K) {
  XDR_DYN_SMEM_ALIGNED(unsigned char, smem_d5, 128);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NP = d5_piece(K);   // columns of X (= of dW) per product; the [128, K] accumulator spans all pieces
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_d5);
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(smem_d5 + 8);
  float* dbs = reinterpret_cast<float*>(smem_d5 + 128);          // [128] column sums of the masked dZ rows this CTA saw
  const int zplane = kD5Rows * kD5Rows * 2, xplane = kD5Rows * NP * 2;
  unsigned char* Zs = smem_d5 + 128 + 512;                        // the dZ tile is [128 batch rows x 128 columns]: columns >= N stay zero
  unsigned char* Xs = Zs + 3 * zplane;
  const tc5::RowBlock16 tZ{kD5Rows, kD5Rows}, tX{kD5Rows, NP};
  const uint32_t cols = d5_tmem_cols(K);
  for (int i = tid; i < kD5Rows; i += kD5Threads) dbs[i] = 0.f;
  for (int i = tid; i < 3 * zplane / 16; i += kD5Threads) reinterpret_cast<float4*>(Zs)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid == 0) {
    tc5::mbar_init(bar, 1);
    tc5::mbar_init_fence();
  }
  if (warp == 0) tc5::tmem_alloc(tmem_base_smem, cols);
  tc5::fence_before_sync();
  __syncthreads();
  tc5::fence_after_sync();
  const uint32_t tmem = *tmem_base_smem;
  D5Sync sy{bar, 0u};
  const uint32_t zs = tc5::smem_u32(Zs), xs = tc5::smem_u32(Xs);
  const uint32_t idesc = tc5::make_idesc_bf16(kD5Rows, NP, true, true);
  const int64_t n_tiles = (M + kD5Rows - 1) / kD5Rows;
  // column sums for db: a thread always stages the same 8-column piece of the dZ rows (kD5Threads % (N / 8) == 0)
  const int c8n = N >> 3;
  float colsum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  bool first = true;
#pragma unroll 1
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * kD5Rows;
    d5_prefetch(X, K, row0, M, kD5Rows, K * 4);
    // dZ rows (masked) -> columns [0, N) of the [128 x 128] tile; the same values feed db
#pragma unroll 1
    for (int e = tid; e < kD5Rows * c8n; e += kD5Threads) {
      const int r = e / c8n, c8 = e - r * c8n;
      const int64_t row = row0 + r;
      float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
      if (row < M && (mask_ids == nullptr || mask_ids[row] < mask_lt)) {
        v0 = ld_row4(dZ + row * N + 8 * c8, 0);
        v1 = ld_row4(dZ + row * N + 8 * c8, 1);
      }
      colsum[0] += v0.x; colsum[1] += v0.y; colsum[2] += v0.z; colsum[3] += v0.w;
      colsum[4] += v1.x; colsum[5] += v1.y; colsum[6] += v1.z; colsum[7] += v1.w;
      tc5::store_split8_3(Zs, Zs + zplane, Zs + 2 * zplane, tZ.chunk_offset(r, c8), v0, v1);
    }
#pragma unroll 1
    for (int k0 = 0; k0 < K; k0 += NP) {
      d5_stage(X, K, row0, M, k0, NP, tX, Xs, xplane, nullptr, 0);
      sy.operands_ready();
      if (tid == 0) {
        tc5::mma_bf16x6(tmem + k0, zs, zplane, tZ.as_mn_major(), xs, xplane, tX.as_mn_major(), idesc, kD5Rows, !first);
        tc5::commit(bar);
      }
      sy.wait_mma();   // the X piece may be overwritten (and, after the last piece, the dZ tile)
    }
    first = false;
  }
  // flush: thread = (output row n = TMEM lane, column group) -> 128-bit reductions into dW (rows >= N are padding)
  const int n = (warp & 3) * 32 + lane, cg = warp >> 2;
  const uint32_t my_lanes = (uint32_t)((warp & 3) * 32) << 16;
  if (!first) {
#pragma unroll 1
    for (int c16 = cg; c16 < K / 16; c16 += kD5Threads / kD5Rows) {
      uint32_t r[16];
      tc5::tmem_ld16(tmem + my_lanes + c16 * 16, r);
      tc5::tmem_ld_wait();
      if (n < N) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          red_add4(dW + (size_t)n * K, 4 * c16 + q,
                   make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                               __uint_as_float(r[4 * q + 3])));
      }
      __syncwarp();
    }
  }
  if (db != nullptr) {
    const int c8 = tid % c8n;   // the 8-column piece this thread staged in every pass (tid + i * 256 keeps e % c8n)
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&dbs[8 * c8 + j], colsum[j]);
    __syncthreads();
    if (tid < N) atomicAdd(&db[tid], dbs[tid]);
  }
  tc5::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc5::tmem_dealloc(tmem, cols);
}

#endif  // __CUDACC__ || XDR_EMU

// 0 (default): fp32 FMA kernels (dense.cu) for every shape; 1: tcgen05 for every shape this file takes; 2: tcgen05 for the
// calls it measured FASTER on a B200 (profiles/r2_dense_engines_fwd2_tiles.jsonl: forward and input gradient of wide layers --
// CoNet's layer 0, 256(+256) -> 64: 47.4 vs 51.5 us and 18.7 vs 25.0 us; it loses or ties on narrow layers and on every weight
// gradient), fp32 tiles for the rest.  Per call both engines are bound by the dependent chain stage -> product -> epilogue of
// ONE 128-row tile per SM, not by the tensor pipe, so the LIBRARY default stays 0 until the engine is pipelined across tiles.
// Callers choose per call (ops.dense_engine): CoNet's stacked BOTH pass (32768-row launches) runs on engine 1 by default -- as
// a whole step it measured 1005 us against 1060 (engine 2) and 1132 (engine 0), profiles/r2_conet_stacked.md.
static int env_dense_engine() {   // XDR_DENSE_ENGINE=1|2 in the environment switches the engine on without a call (test sweeps)
  const char* e = getenv("XDR_DENSE_ENGINE");
  return (e && (e[0] == '1' || e[0] == '2')) ? e[0] - '0' : 0;
}
static int g_dense_engine = env_dense_engine();

static bool d5_shape_ok(int64_t M, int N, int K) {
  return M >= kD5Rows && N % 16 == 0 && N >= 16 && N <= 128 && K % 16 == 0 && K >= 16 && K <= 256 && (kD5Threads % (N / 8)) == 0;
}
// does the engine setting send this call to tcgen05?  (what: 0 forward, 1 input gradient, 2 weight gradient)
static bool d5_wanted(int what, int N, int K) {
  if (g_dense_engine == 1) return true;
  return g_dense_engine == 2 && what != 2 && K >= 192 && N >= 64;
}

bool tc5_dense_fwd_ok(const float* X, const float* W, const float* X2, const float* W2, const float* Y, int64_t M, int N, int K) {
  return d5_wanted(0, N, K) && d5_shape_ok(M, N, K) && aligned16(X) && aligned16(W) && aligned16(Y) && aligned16(X2) && aligned16(W2);
}
bool tc5_dense_bwd_input_ok(const float* dZ, const float* W, const float* dX, int64_t M, int N, int K) {
  return d5_wanted(1, N, K) && d5_shape_ok(M, N, K) && aligned16(dZ) && aligned16(W) && aligned16(dX);
}
bool tc5_dense_bwd_weight_ok(const float* dZ, const float* X, const float* dW, int64_t M, int N, int K) {
  return d5_wanted(2, N, K) && d5_shape_ok(M, N, K) && aligned16(dZ) && aligned16(X) && aligned16(dW);
}

// raises a kernel's dynamic shared-memory limit when a launch needs more than any launch before it (once per size, not per call)
#ifndef XDR_EMU
#define D5_SMEM_LIMIT(kern, smem)                                                                        \
  do {                                                                                                   \
    static size_t limit__ = 0;                                                                           \
    if ((smem) > limit__) {                                                                              \
      XDR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem))); \
      limit__ = (smem);                                                                                  \
    }                                                                                                    \
  } while (0)
#else
#define D5_SMEM_LIMIT(kern, smem) do { } while (0)
#endif

int tc5_dense_fwd(const float* X, const float* W, const float* bias, const float* X2, const float* W2, const int64_t* mask_ids,
                  int64_t mask_lt, int act, float* Y, int64_t M, int N, int K, cudaStream_t stream) {
  const int KC = d5_chunk(K);
  const size_t smem = 128 + (size_t)(kD5Rows + N) * KC * 6;
  D5_SMEM_LIMIT(tc5_dense_fwd_kernel, smem);
  const int grid = (int)((M + kD5Rows - 1) / kD5Rows);
  XDR_LAUNCH((tc5_dense_fwd_kernel), grid, kD5Threads, smem, stream, X, W, bias, X2, W2, mask_ids, mask_lt, act, Y, M, N, K);
  XDR_LAUNCH_OK();
  return XDR_OK;
}

int tc5_dense_bwd_input(const float* dZ, const float* W, const int64_t* mask_ids, int64_t mask_lt, float* dX, int64_t M, int N,
                        int K, int accumulate, cudaStream_t stream) {
  const int NP = d5_piece(K);
  const size_t smem = 128 + (size_t)kD5Rows * N * 6 + (size_t)N * NP * 6;
  D5_SMEM_LIMIT(tc5_dense_bwd_input_kernel, smem);
  const int grid = (int)((M + kD5Rows - 1) / kD5Rows);
  XDR_LAUNCH((tc5_dense_bwd_input_kernel), grid, kD5Threads, smem, stream, dZ, W, mask_ids, mask_lt, dX, M, N, K, accumulate);
  XDR_LAUNCH_OK();
  return XDR_OK;
}

int tc5_dense_bwd_weight(const float* dZ, const float* X, const int64_t* mask_ids, int64_t mask_lt, float* dW, float* db, int64_t M,
                         int N, int K, cudaStream_t stream) {
  const size_t smem = 128 + 512 + (size_t)kD5Rows * kD5Rows * 6 + (size_t)kD5Rows * d5_piece(K) * 6;
  D5_SMEM_LIMIT(tc5_dense_bwd_weight_kernel, smem);
  const int64_t n_tiles = (M + kD5Rows - 1) / kD5Rows;
  // one row tile per CTA up to the SM count (every CTA flushes an [N, K] accumulator with 128-bit reductions: 64 KB at most)
  int64_t grid = n_tiles;
  if (grid > sm_count()) grid = sm_count();
  if (grid < 1) grid = 1;
  XDR_LAUNCH((tc5_dense_bwd_weight_kernel), (int)grid, kD5Threads, smem, stream, dZ, X, mask_ids, mask_lt, dW, db, M, N, K);
  XDR_LAUNCH_OK();
  return XDR_OK;
}

}  // namespace xdr

extern "C" {

// 1: dense layers whose shapes qualify run on tcgen05 (bf16x6); 2: only the calls tcgen05 measured faster (forward / input
// gradient of wide layers); 0 (default): always the fp32 FMA kernels.  Returns the previous setting.
int xdr_set_dense_engine(int engine) {
  const int prev = xdr::g_dense_engine;
  xdr::g_dense_engine = (engine == 1 || engine == 2) ? engine : 0;
  return prev;
}

}  // extern "C"
