// tc_tile.cuh -- tensor-core building blocks for row-tile kernels whose operands live in shared memory as plain fp32.
//
// Every product is computed as 3xTF32: x = hi + lo with hi = tf32(x), lo = tf32(x - hi), and
//     a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi          (the dropped a_lo*b_lo term is ~2^-22 relative)
// accumulated in fp32 by mma.sync.m16n8k8 (SASS HMMA.1688.F32.TF32), so the results agree with the fp32 FMA reference
// to ~1e-6 relative -- inside the 1e-4 parity bar of the dense rows with a wide margin, which a single
// TF32 pass (~5e-4 per product) would not give for the gradients.
//
// Fragment ownership of mma.m16n8k8 (g = lane >> 2, t = lane & 3):
//     A (16x8, row):  a0 = (g, t)   a1 = (g+8, t)   a2 = (g, t+4)   a3 = (g+8, t+4)
//     B (8x8,  col):  b0 = (k = t, n = g)           b1 = (k = t+4, n = g)
//     C (16x8):       c0 = (g, 2t)  c1 = (g, 2t+1)  c2 = (g+8, 2t)  c3 = (g+8, 2t+1)
// Shared-memory operands are row-major with a leading dimension ld = width + 4 (ld / 4 odd), so the (g*ld + t) access
// of an untransposed operand touches 32 distinct banks; transposed operands ((t*ld + g)) see 2-way conflicts.
//
// Engine modes (compile-time, -DXDR_TC_MODE=n; build a second library with XDR_LIB_NAME / XDR_BUILD_DIR to compare on hardware):
//   0  3xTF32, mma.m16n8k8          the default described above
//   1  bf16x3, mma.m16n8k16         x = hi + lo with hi = bf16(x), lo = bf16(x - hi); a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi
//                                   (~2^-17 relative per product) -- half the MMA instructions per unit of K
//   2  one TF32 pass                NOT parity-grade (~5e-4 per product); a diagnostic for how tensor-bound a kernel is
// m16n8k16 bf16 fragments (two bf16 per register, low half = lower k):
//     A: a0 = (g, 2t..2t+1)  a1 = (g+8, 2t..)  a2 = (g, 2t+8..)  a3 = (g+8, 2t+8..)      B: b0 = (k = 2t..2t+1, n = g)  b1 = (k = 2t+8.., n = g)
#pragma once
#include "xdr_common.cuh"

#ifndef XDR_TC_MODE
#define XDR_TC_MODE 0
#endif

namespace xdr {

#if defined(__CUDACC__) || defined(XDR_EMU)

constexpr int kTcThreads = 256;
constexpr int kTcWarps = kTcThreads / 32;
constexpr int kTcMode = XDR_TC_MODE;
constexpr int kTcKStep = kTcMode == 1 ? 16 : 8;   // reduction elements one MMA group consumes

__device__ __forceinline__ uint32_t cvt_tf32(float x) {
#ifdef XDR_EMU
  return emu::to_tf32(x);
#else
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
#endif
}

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = cvt_tf32(x);
  lo = cvt_tf32(x - __uint_as_float(hi));
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
#ifdef XDR_EMU
  emu::mma_m16n8k8_tf32(c, a, b);
#else
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
#endif
}

// c += a * b in 3xTF32 (small terms first)
__device__ __forceinline__ void mma_3xtf32(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                           const uint32_t (&bh)[2], const uint32_t (&bl)[2]) {
  mma_tf32(c, al, bh);
  mma_tf32(c, ah, bl);
  mma_tf32(c, ah, bh);
}

// ---- bf16x3 ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bf16_bits(float x) {  // round-to-nearest-even to bf16, returned in the low 16 bits
  uint32_t u = __float_as_uint(x);
  if ((u & 0x7f800000u) == 0x7f800000u) return u >> 16;  // inf / nan
  u += 0x7fffu + ((u >> 16) & 1u);
  return u >> 16;
}
// (x0, x1) -> packed bf16 pairs hi and lo with x ~= hi + lo; element 0 in the low half
__device__ __forceinline__ void split_bf16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const uint32_t h0 = bf16_bits(x0), h1 = bf16_bits(x1);
  const float r0 = x0 - __uint_as_float(h0 << 16), r1 = x1 - __uint_as_float(h1 << 16);
  hi = h0 | (h1 << 16);
  lo = bf16_bits(r0) | (bf16_bits(r1) << 16);
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
#ifdef XDR_EMU
  emu::mma_m16n8k16_bf16(c, a, b);
#else
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
#endif
}

// One MMA group of the selected engine: c += a * b over kTcKStep reduction elements, operands already split.
__device__ __forceinline__ void mma_group(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                          const uint32_t (&bh)[2], const uint32_t (&bl)[2]) {
  if (kTcMode == 1) {
    mma_bf16(c, al, bh);
    mma_bf16(c, ah, bl);
    mma_bf16(c, ah, bh);
  } else if (kTcMode == 2) {
    mma_tf32(c, ah, bh);
  } else {
    mma_3xtf32(c, ah, al, bh, bl);
  }
}

// Operand fragments of one MMA group from shared memory.  `row_ptr(i)` addresses are formed by the callers:
//   A fragment of a ROW-MAJOR operand  (element (m, k) at P[m*ld + k]), rows m0+g / m0+g+8, reduction k0..k0+kTcKStep
//   B fragment of a K-CONTIGUOUS operand (B(k, n) = P[n*ld + k]),        column n0+g
//   and their transposed forms (reduction index strided by ld).
// kmax bounds the reduction index (elements >= kmax read as zero: K = 8 tails in bf16 mode).
__device__ __forceinline__ void frag_a_rowmajor(const float* __restrict__ P, int ld, int m, int k0, int kmax, int t,
                                                uint32_t (&h)[4], uint32_t (&l)[4]) {
  if (kTcMode == 1) {
    const float* p0 = P + m * ld + k0 + 2 * t;
    const float* p1 = p0 + 8 * ld;
    const bool lo_ok = k0 + 2 * t < kmax, hi_ok = k0 + 2 * t + 8 < kmax;
    const float2 z = make_float2(0.f, 0.f);
    const float2 v0 = lo_ok ? *reinterpret_cast<const float2*>(p0) : z, v1 = lo_ok ? *reinterpret_cast<const float2*>(p1) : z;
    const float2 v2 = hi_ok ? *reinterpret_cast<const float2*>(p0 + 8) : z, v3 = hi_ok ? *reinterpret_cast<const float2*>(p1 + 8) : z;
    split_bf16x2(v0.x, v0.y, h[0], l[0]);
    split_bf16x2(v1.x, v1.y, h[1], l[1]);
    split_bf16x2(v2.x, v2.y, h[2], l[2]);
    split_bf16x2(v3.x, v3.y, h[3], l[3]);
  } else {
    const float* ap = P + m * ld + k0 + t;
    split_tf32(ap[0], h[0], l[0]);
    split_tf32(ap[8 * ld], h[1], l[1]);
    split_tf32(ap[4], h[2], l[2]);
    split_tf32(ap[8 * ld + 4], h[3], l[3]);
  }
}
// B(k, n) with n fixed: contiguous == true -> P[n*ld + k]; false -> P[k*ld + n]
__device__ __forceinline__ void frag_b(const float* __restrict__ P, int ld, int n, int k0, int kmax, int t, bool contiguous,
                                       uint32_t (&h)[2], uint32_t (&l)[2]) {
  if (kTcMode == 1) {
    const int ka = k0 + 2 * t, kb = ka + 8;
    float x0 = 0.f, x1 = 0.f, y0 = 0.f, y1 = 0.f;
    if (contiguous) {
      const float* p = P + n * ld;
      if (ka < kmax) { const float2 v = *reinterpret_cast<const float2*>(p + ka); x0 = v.x; x1 = v.y; }
      if (kb < kmax) { const float2 v = *reinterpret_cast<const float2*>(p + kb); y0 = v.x; y1 = v.y; }
    } else {
      if (ka < kmax) { x0 = P[ka * ld + n]; x1 = P[(ka + 1) * ld + n]; }
      if (kb < kmax) { y0 = P[kb * ld + n]; y1 = P[(kb + 1) * ld + n]; }
    }
    split_bf16x2(x0, x1, h[0], l[0]);
    split_bf16x2(y0, y1, h[1], l[1]);
  } else {
    float b0, b1;
    if (contiguous) {
      const float* bp = P + n * ld + k0 + t;
      b0 = bp[0];
      b1 = bp[4];
    } else {
      const float* bp = P + (k0 + t) * ld + n;
      b0 = bp[0];
      b1 = bp[4 * ld];
    }
    split_tf32(b0, h[0], l[0]);
    split_tf32(b1, h[1], l[1]);
  }
}

__device__ __forceinline__ float act_apply(float y, int act) {
  if (act == XDR_ACT_RELU) return y > 0.f ? y : 0.f;
  if (act == XDR_ACT_TANH) return tanhf(y);
  if (act == XDR_ACT_SIGMOID) return sigmoidf_(y);
  return y;
}
// derivative of the activation expressed through its OUTPUT y
__device__ __forceinline__ float act_grad(float y, int act) {
  if (act == XDR_ACT_RELU) return y > 0.f ? 1.f : 0.f;
  if (act == XDR_ACT_TANH) return 1.f - y * y;
  if (act == XDR_ACT_SIGMOID) return (1.f - y) * y;
  return 1.f;
}

// acc += A[TR x K] * B restricted to the output tiles this warp owns; A row-major in shared memory ((m, k) at A[m*lda + k]).
//   BT == false:  B(k, n) = Bsm[n*ldb + k]   (nn.Linear.weight [N][K]:  C = A W^T, the forward of a layer)
//   BT == true :  B(k, n) = Bsm[k*ldb + n]   (C = A W with W [K][N]:     the input gradient dX = dZ W)
// Warp w owns the 16-row tile (w % MT) and the 8-column tiles (w / MT) + j*G, j < MAXNT (MT = TR/16, G = kTcWarps/MT).
// N % 8 == 0, K % 8 == 0, N <= 8*G*MAXNT.  The accumulators belong to the caller, so a product can be split over K chunks.
template <int TR, int MAXNT, bool BT>
__device__ __forceinline__ void tile_mma_acc(float (&acc)[MAXNT][4], const float* __restrict__ A, int lda,
                                             const float* __restrict__ Bsm, int ldb, int N, int K) {
  constexpr int MT = TR / 16, G = kTcWarps / MT;
  static_assert(TR % 16 == 0 && kTcWarps % MT == 0, "row tile must split into whole 16-row MMA tiles over the warps");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int m0 = (warp % MT) * 16, grp = warp / MT;
  const int ntiles = N >> 3;
  if (grp >= ntiles) return;
  for (int k0 = 0; k0 < K; k0 += kTcKStep) {
    uint32_t ah[4], al[4];
    frag_a_rowmajor(A, lda, m0 + g, k0, K, t, ah, al);
#pragma unroll
    for (int j = 0; j < MAXNT; ++j) {
      const int nt = grp + j * G;
      if (nt < ntiles) {  // warp-uniform
        uint32_t bh[2], bl[2];
        frag_b(Bsm, ldb, nt * 8 + g, k0, K, t, !BT, bh, bl);
        mma_group(acc[j], ah, al, bh, bl);
      }
    }
  }
}

// Two products that share the A operand: acc1 += A * B1, acc2 += A * B2 (same N, K, BT).  The A fragments are loaded and
// split once per reduction step for both.  do2 (warp-uniform) == false leaves acc2 untouched.
template <int TR, int MAXNT, bool BT>
__device__ __forceinline__ void tile_mma_acc2(float (&acc1)[MAXNT][4], float (&acc2)[MAXNT][4], const float* __restrict__ A,
                                              int lda, const float* __restrict__ B1, int ldb1, const float* __restrict__ B2,
                                              int ldb2, int N, int K, bool do2) {
  constexpr int MT = TR / 16, G = kTcWarps / MT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int m0 = (warp % MT) * 16, grp = warp / MT;
  const int ntiles = N >> 3;
  if (grp >= ntiles) return;
  for (int k0 = 0; k0 < K; k0 += kTcKStep) {
    uint32_t ah[4], al[4];
    frag_a_rowmajor(A, lda, m0 + g, k0, K, t, ah, al);
#pragma unroll
    for (int j = 0; j < MAXNT; ++j) {
      const int nt = grp + j * G;
      if (nt < ntiles) {  // warp-uniform
        uint32_t bh[2], bl[2];
        frag_b(B1, ldb1, nt * 8 + g, k0, K, t, !BT, bh, bl);
        mma_group(acc1[j], ah, al, bh, bl);
        if (do2) {
          frag_b(B2, ldb2, nt * 8 + g, k0, K, t, !BT, bh, bl);
          mma_group(acc2[j], ah, al, bh, bl);
        }
      }
    }
  }
}

template <int MAXNT>
__device__ __forceinline__ void tile_acc_zero(float (&acc)[MAXNT][4]) {
#pragma unroll
  for (int j = 0; j < MAXNT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
}

// Visits the accumulator pairs of this warp: fn(j, row, col, half) with acc[j][2*half], acc[j][2*half + 1] holding
// C[row][col], C[row][col + 1] (col even).  Same ownership as tile_mma_acc.
template <int TR, int MAXNT, typename Fn>
__device__ __forceinline__ void tile_acc_visit(int N, Fn fn) {
  constexpr int MT = TR / 16, G = kTcWarps / MT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int m0 = (warp % MT) * 16, grp = warp / MT;
  const int ntiles = N >> 3;
#pragma unroll
  for (int j = 0; j < MAXNT; ++j) {
    const int nt = grp + j * G;
    if (nt < ntiles) {
      fn(j, m0 + g, nt * 8 + 2 * t, 0);
      fn(j, m0 + g + 8, nt * 8 + 2 * t, 1);
    }
  }
}

// C[TR x N] = A[TR x K] * B for one CTA; epi(row, col, v0, v1) receives C[row][col], C[row][col+1] (col even).
template <int TR, int MAXNT, bool BT, typename Epi>
__device__ __forceinline__ void tile_gemm(const float* __restrict__ A, int lda, const float* __restrict__ Bsm, int ldb, int N,
                                          int K, Epi epi) {
  float acc[MAXNT][4];
  tile_acc_zero(acc);
  tile_mma_acc<TR, MAXNT, BT>(acc, A, lda, Bsm, ldb, N, K);
  tile_acc_visit<TR, MAXNT>(N, [&](int j, int row, int col, int half) { epi(row, col, acc[j][2 * half], acc[j][2 * half + 1]); });
}

// Two products over the same output tile: C1 = A1 * B1, C2 = A2 * B2 (same N, K and BT); epi(row, col, v0, v1, c0, c1).
// The cross-stitch unit of CoNet (conet.py:118-138): v = own tower, c = cross term, combined per row by the overlap mask.
// do2 (warp-uniform): false skips the second product for this warp's row tile (its C2 is then zero) -- CoNet row tiles
// without an overlapped row.
template <int TR, int MAXNT, bool BT, typename Epi>
__device__ __forceinline__ void tile_gemm2(const float* __restrict__ A1, int lda1, const float* __restrict__ B1, int ldb1,
                                           const float* __restrict__ A2, int lda2, const float* __restrict__ B2, int ldb2,
                                           int N, int K, bool do2, Epi epi) {
  float acc[MAXNT][4], acc2[MAXNT][4];
  tile_acc_zero(acc);
  tile_acc_zero(acc2);
  tile_mma_acc<TR, MAXNT, BT>(acc, A1, lda1, B1, ldb1, N, K);
  if (do2) tile_mma_acc<TR, MAXNT, BT>(acc2, A2, lda2, B2, ldb2, N, K);
  tile_acc_visit<TR, MAXNT>(N, [&](int j, int row, int col, int half) {
    epi(row, col, acc[j][2 * half], acc[j][2 * half + 1], acc2[j][2 * half], acc2[j][2 * half + 1]);
  });
}

// Picks the smallest MAXNT instantiation that covers N output columns.
template <int TR, bool BT, typename Epi>
__device__ __forceinline__ void tile_gemm_any(const float* A, int lda, const float* Bsm, int ldb, int N, int K, Epi epi) {
  constexpr int G = kTcWarps / (TR / 16);
  const int per_warp = ((N >> 3) + G - 1) / G;
  if (per_warp <= 1) tile_gemm<TR, 1, BT>(A, lda, Bsm, ldb, N, K, epi);
  else if (per_warp <= 2) tile_gemm<TR, 2, BT>(A, lda, Bsm, ldb, N, K, epi);
  else if (per_warp <= 4) tile_gemm<TR, 4, BT>(A, lda, Bsm, ldb, N, K, epi);
  else if (per_warp <= 8) tile_gemm<TR, 8, BT>(A, lda, Bsm, ldb, N, K, epi);
  else tile_gemm<TR, 16, BT>(A, lda, Bsm, ldb, N, K, epi);
}

template <int TR, bool BT, typename Epi>
__device__ __forceinline__ void tile_gemm2_any(const float* A1, int lda1, const float* B1, int ldb1, const float* A2, int lda2,
                                               const float* B2, int ldb2, int N, int K, bool do2, Epi epi) {
  constexpr int G = kTcWarps / (TR / 16);
  const int per_warp = ((N >> 3) + G - 1) / G;
  if (per_warp <= 1) tile_gemm2<TR, 1, BT>(A1, lda1, B1, ldb1, A2, lda2, B2, ldb2, N, K, do2, epi);
  else if (per_warp <= 2) tile_gemm2<TR, 2, BT>(A1, lda1, B1, ldb1, A2, lda2, B2, ldb2, N, K, do2, epi);
  else if (per_warp <= 4) tile_gemm2<TR, 4, BT>(A1, lda1, B1, ldb1, A2, lda2, B2, ldb2, N, K, do2, epi);
  else tile_gemm2<TR, 8, BT>(A1, lda1, B1, ldb1, A2, lda2, B2, ldb2, N, K, do2, epi);
}

// Number of 16x8 tiles of a [dout][din] weight gradient and how many of them each warp owns (tile i -> warp i % kTcWarps).
__host__ __device__ inline int dw_tiles(int dout, int din) { return ((dout + 15) >> 4) * (din >> 3); }
__host__ __device__ inline int dw_tiles_per_warp(int dout, int din) { return (dw_tiles(dout, din) + kTcWarps - 1) / kTcWarps; }

// acc += dZ^T X restricted to the tiles this warp owns:  dW[n][k] += sum_r dZ[r][n] * X[r][k],  r < TR.
// dZ [TR][ldz] (dout columns), X [TR][ldx] (din columns), both row-major in shared memory.  The accumulators are MMA C
// fragments that live in registers across all row tiles of the CTA; dw_flush() adds them to global memory once.
// rmask (optional, [TR]): dZ row r is multiplied by rmask[r] (the overlap mask of CoNet's cross parameters).  For a K-chunked
// layer pass the chunk slab as X with din = chunk width; dw_flush() then places the chunk inside the full weight.
template <int TR, int MAXT>
__device__ __forceinline__ void dw_accum(float (&acc)[MAXT][4], const float* __restrict__ dZ, int ldz,
                                         const float* __restrict__ X, int ldx, int dout, int din,
                                         const float* __restrict__ rmask = nullptr) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int ktiles = din >> 3, total = dw_tiles(dout, din);
#pragma unroll
  for (int j = 0; j < MAXT; ++j) {
    const int tile = warp + j * kTcWarps;
    if (tile < total) {  // warp-uniform
      const int n0 = (tile / ktiles) * 16, k0 = (tile % ktiles) * 8;
      const bool lo_ok = n0 + g < dout, hi_ok = n0 + g + 8 < dout;
#pragma unroll 2
      for (int r0 = 0; r0 < TR; r0 += kTcKStep) {
        if (rmask != nullptr) {  // a reduction step whose rows are all masked out contributes nothing (same for every lane)
          float any = 0.f;
#pragma unroll
          for (int q = 0; q < kTcKStep; ++q) any += rmask[r0 + q];
          if (any == 0.f) continue;
        }
        uint32_t ah[4], al[4], bh[2], bl[2];
        if (kTcMode == 1) {
          // A(m = output n, k = batch row): a0 = rows r0+2t, +1 at column n0+g; a1 at n0+g+8; a2 / a3 eight rows further
          const int ra = r0 + 2 * t, rb = ra + 8;
          const float* za = dZ + ra * ldz + n0 + g;
          const float* zb = dZ + rb * ldz + n0 + g;
          const float ma0 = rmask ? rmask[ra] : 1.f, ma1 = rmask ? rmask[ra + 1] : 1.f;
          const float mb0 = rmask ? rmask[rb] : 1.f, mb1 = rmask ? rmask[rb + 1] : 1.f;
          split_bf16x2(lo_ok ? ma0 * za[0] : 0.f, lo_ok ? ma1 * za[ldz] : 0.f, ah[0], al[0]);
          split_bf16x2(hi_ok ? ma0 * za[8] : 0.f, hi_ok ? ma1 * za[ldz + 8] : 0.f, ah[1], al[1]);
          split_bf16x2(lo_ok ? mb0 * zb[0] : 0.f, lo_ok ? mb1 * zb[ldz] : 0.f, ah[2], al[2]);
          split_bf16x2(hi_ok ? mb0 * zb[8] : 0.f, hi_ok ? mb1 * zb[ldz + 8] : 0.f, ah[3], al[3]);
          frag_b(X, ldx, k0 + g, r0, TR, t, false, bh, bl);
        } else {
          const float* zp = dZ + (r0 + t) * ldz + n0 + g;
          const float mk0 = rmask ? rmask[r0 + t] : 1.f, mk1 = rmask ? rmask[r0 + t + 4] : 1.f;
          split_tf32(lo_ok ? mk0 * zp[0] : 0.f, ah[0], al[0]);
          split_tf32(hi_ok ? mk0 * zp[8] : 0.f, ah[1], al[1]);
          split_tf32(lo_ok ? mk1 * zp[4 * ldz] : 0.f, ah[2], al[2]);
          split_tf32(hi_ok ? mk1 * zp[4 * ldz + 8] : 0.f, ah[3], al[3]);
          frag_b(X, ldx, k0 + g, r0, TR, t, false, bh, bl);
        }
        mma_group(acc[j], ah, al, bh, bl);
      }
    }
  }
}

template <int MAXT>
__device__ __forceinline__ void dw_flush(const float (&acc)[MAXT][4], float* __restrict__ dW, int dout, int din, int ldw = -1,
                                         int k_off = 0) {
  if (dW == nullptr) return;
  if (ldw < 0) ldw = din;
  dW += k_off;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int ktiles = din >> 3, total = dw_tiles(dout, din);
#pragma unroll
  for (int j = 0; j < MAXT; ++j) {
    const int tile = warp + j * kTcWarps;
    if (tile < total) {
      const int n0 = (tile / ktiles) * 16, k = (tile % ktiles) * 8 + 2 * t;
      if (n0 + g < dout) {
        atomicAdd(&dW[(size_t)(n0 + g) * ldw + k], acc[j][0]);
        atomicAdd(&dW[(size_t)(n0 + g) * ldw + k + 1], acc[j][1]);
      }
      if (n0 + g + 8 < dout) {
        atomicAdd(&dW[(size_t)(n0 + g + 8) * ldw + k], acc[j][2]);
        atomicAdd(&dW[(size_t)(n0 + g + 8) * ldw + k + 1], acc[j][3]);
      }
    }
  }
}

#endif  // __CUDACC__ || XDR_EMU

}  // namespace xdr
