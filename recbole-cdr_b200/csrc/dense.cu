// dense.cu -- A4/A7/A14: the small dense layers between gather and score, fp32 on CUDA cores.
//
//   xdr_dense_fwd        Y  = act(X W^T + b + mask * (X2 W2^T))     EMCDR mapping (emcdr.py:86-93), CoNet cross-stitch
//                                                                   units (conet.py:118-138), recbole MLPLayers (dtcdr.py:61-67)
//   xdr_act_bwd          dZ = dY * act'(Y)
//   xdr_dense_bwd_input  dX (+)= mask * (dZ W)
//   xdr_dense_bwd_weight dW += (mask * dZ)^T X,  db += colsum(dZ)
//
// These are GEMMs with a long M (batch) and short N/K (8..256).  One 64x64 output tile per CTA, 4x4 register
// micro-tile per thread, K staged through shared memory in slabs of 16.  fp32 FMA keeps the loss within the
// 1e-4 contract without any split-precision trick; the tcgen05 engine (tc5_dense.cu) takes over for the shapes it
// takes (M >= 128, N and K multiples of 16): xdr_set_dense_engine(0) keeps everything here.
#include <stdlib.h>
#include "xdr_common.cuh"
#include "tc5_dense.cuh"

namespace xdr {

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;
constexpr int kDenseThreads = 256;
static int env_flag_fwd2() { const char* e = getenv("XDR_DENSE_FWD2"); return e ? (e[0] == '0' ? 0 : (e[0] == '2' ? 2 : 1)) : 1; }
static int g_dense_fwd2 = env_flag_fwd2();   // XDR_DENSE_FWD2=0 in the environment keeps the 64 x 64 single-buffered forward tiles

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case XDR_ACT_RELU: return v > 0.f ? v : 0.f;
    case XDR_ACT_TANH: return tanhf(v);
    case XDR_ACT_SIGMOID: return sigmoidf_(v);
    default: return v;
  }
}

__device__ __forceinline__ void micro_fma(float (&acc)[4][4], const float (*As)[BM + PAD], const float (*Bs)[BN + PAD],
                                          int ty, int tx) {
#pragma unroll
  for (int kk = 0; kk < BK; ++kk) {
    const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
    const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
}

struct MaskCtx {
  const int64_t* ids;
  int64_t lt;
};

// Stage a [rows x BK] slab of a row-major matrix P[R, ld] (rows r0.., reduction columns k0..) TRANSPOSED into
// S[kk][r]: thread t covers row t/4, columns (t%4)*4..+3.
__device__ __forceinline__ void stage_T(float (*S)[BM + PAD], const float* __restrict__ P, int64_t R, int ld, int64_t r0,
                                        int k0, int kmax, bool row_on) {
  const int r = threadIdx.x >> 2, kq = (threadIdx.x & 3) * 4;
  const int64_t row = r0 + r;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (row < R && row_on) {
    const float* p = P + row * ld + k0 + kq;
    if (k0 + kq + 3 < kmax && ((ld & 3) == 0) && aligned16_dev(P)) {
      const float4 q = *reinterpret_cast<const float4*>(p);
      v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (k0 + kq + i < kmax) v[i] = p[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) S[kq + i][r] = v[i];
}

// Stage a [BK x cols] slab of a row-major matrix P[R, ld] (reduction rows k0.., columns c0..) as S[kk][c]:
// thread t covers row t/16, columns (t%16)*4..+3.
__device__ __forceinline__ void stage_N(float (*S)[BN + PAD], const float* __restrict__ P, int64_t R, int ld, int64_t k0,
                                        int c0, int cmax, const MaskCtx* mc) {
  const int kk = threadIdx.x >> 4, cq = (threadIdx.x & 15) * 4;
  const int64_t row = k0 + kk;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (row < R && (mc == nullptr || mc->ids == nullptr || mc->ids[row] < mc->lt)) {
    const float* p = P + row * ld + c0 + cq;
    if (c0 + cq + 3 < cmax && ((ld & 3) == 0) && aligned16_dev(P)) {
      const float4 q = *reinterpret_cast<const float4*>(p);
      v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (c0 + cq + i < cmax) v[i] = p[i];
    }
  }
  *reinterpret_cast<float4*>(&S[kk][cq]) = make_float4(v[0], v[1], v[2], v[3]);
}


// ---- forward: Y[M,N] = act(X[M,K] W[N,K]^T + b + mask * X2[M,K] W2[N,K]^T) -------------------------------
__global__ void __launch_bounds__(kDenseThreads)
    dense_fwd_kernel(const float* __restrict__ X, const float* __restrict__ W, const float* __restrict__ bias,
                     const float* __restrict__ X2, const float* __restrict__ W2, const int64_t* __restrict__ mask_ids,
                     int64_t mask_lt, int act, float* __restrict__ Y, int64_t M, int N, int K) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  float acc[4][4] = {}, acc2[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += BK) {
    stage_T(As, X, M, K, m0, k0, K, true);
    stage_T(Bs, W, N, K, n0, k0, K, true);
    __syncthreads();
    micro_fma(acc, As, Bs, ty, tx);
    __syncthreads();
  }
  if (X2 != nullptr) {
    for (int k0 = 0; k0 < K; k0 += BK) {
      stage_T(As, X2, M, K, m0, k0, K, true);
      stage_T(Bs, W2, N, K, n0, k0, K, true);
      __syncthreads();
      micro_fma(acc2, As, Bs, ty, tx);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const float mk = (X2 != nullptr && (mask_ids == nullptr || mask_ids[m] < mask_lt)) ? 1.f : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (X2 != nullptr) v += mk * acc2[i][j];
      Y[m * N + n] = apply_act(v, act);
    }
  }
}

// ---- forward, register-blocked and double-buffered (round 2): the same product as dense_fwd_kernel for 16-byte aligned
// operands with K % 4 == 0.  128 x BN output tile per CTA, 256 threads; thread (ty, tx) owns RM = 128 / (1024 / BN) rows x 4
// columns (8 x 4 at BN = 64: 32 FMAs per 3 shared-memory loads), the next K slab is prefetched into registers while the
// current one is multiplied, shared memory is double-buffered (one barrier per slab), the slabs are written transposed
// without bank conflicts (a warp covers 32 consecutive rows of one 4-column group).  The cross-stitch term continues the
// same accumulators: masked-out rows of X2 are staged as zeros.  Used for narrow layers (N <= 32), where it is 1.3-1.4x the
// 64 x 64 tiles; for N >= 64 the grid of 128-row tiles is too small to fill the SMs (see xdr_dense_fwd).
constexpr int BM2 = 128;
template <int BN2>
__global__ void __launch_bounds__(kDenseThreads)
    dense_fwd2_kernel(const float* __restrict__ X, const float* __restrict__ W, const float* __restrict__ bias,
                      const float* __restrict__ X2, const float* __restrict__ W2, const int64_t* __restrict__ mask_ids,
                      int64_t mask_lt, int act, float* __restrict__ Y, int64_t M, int N, int K) {
  constexpr int TX = BN2 / 4, TY = kDenseThreads / TX, RM = BM2 / TY;   // 16 x 16 threads, 8 rows each at BN2 = 64
  constexpr int LDA = BM2 + 4, LDB = BN2 + 4;
  constexpr int A_LOADS = BM2 * BK / 4 / kDenseThreads;                  // float4 loads per thread and slab: 2
  constexpr int B_ROWS_PER_PASS = kDenseThreads / (BK / 4);              // 64 rows of W per pass of all threads
  constexpr int B_LOADS = (BN2 + B_ROWS_PER_PASS - 1) / B_ROWS_PER_PASS; // 1
  __shared__ __align__(16) float As[2][BK][LDA];
  __shared__ __align__(16) float Bs[2][BK][LDB];
  const int tid = threadIdx.x, ty = tid / TX, tx = tid % TX;
  const int64_t m0 = (int64_t)blockIdx.x * BM2;
  const int n0 = blockIdx.y * BN2;
  // staging roles: the 64 threads tid % 64 of a pass cover 64 consecutive rows; tid / 64 picks the 4-column group of the slab
  const int lr = tid & 63, kq = (tid >> 6) * 4;
  const int slabs1 = (K + BK - 1) / BK, slabs = X2 != nullptr ? 2 * slabs1 : slabs1;
  float acc[RM][4] = {};
  float4 pa[A_LOADS], pb[B_LOADS];
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  auto fetch = [&](int sl) {   // slab sl of [X | X2] and [W | W2] -> registers
    const bool second = sl >= slabs1;
    const float* __restrict__ P = second ? X2 : X;
    const float* __restrict__ Q = second ? W2 : W;
    const int k = (second ? sl - slabs1 : sl) * BK + kq;
#pragma unroll
    for (int i = 0; i < A_LOADS; ++i) {
      const int64_t m = m0 + lr + 64 * i;
      bool on = m < M && k < K;
      if (on && second && mask_ids != nullptr) on = mask_ids[m] < mask_lt;
      pa[i] = on ? __ldg(reinterpret_cast<const float4*>(P + m * K + k)) : z4;
    }
#pragma unroll
    for (int i = 0; i < B_LOADS; ++i) {
      const int n = n0 + lr + 64 * i;
      pb[i] = (lr + 64 * i < BN2 && n < N && k < K) ? __ldg(reinterpret_cast<const float4*>(Q + (int64_t)n * K + k)) : z4;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_LOADS; ++i) {
      As[buf][kq + 0][lr + 64 * i] = pa[i].x;
      As[buf][kq + 1][lr + 64 * i] = pa[i].y;
      As[buf][kq + 2][lr + 64 * i] = pa[i].z;
      As[buf][kq + 3][lr + 64 * i] = pa[i].w;
    }
#pragma unroll
    for (int i = 0; i < B_LOADS; ++i) {
      if (lr + 64 * i < BN2) {
        Bs[buf][kq + 0][lr + 64 * i] = pb[i].x;
        Bs[buf][kq + 1][lr + 64 * i] = pb[i].y;
        Bs[buf][kq + 2][lr + 64 * i] = pb[i].z;
        Bs[buf][kq + 3][lr + 64 * i] = pb[i].w;
      }
    }
  };
  fetch(0);
  stash(0);
  __syncthreads();
  for (int sl = 0; sl < slabs; ++sl) {
    const int buf = sl & 1;
    if (sl + 1 < slabs) fetch(sl + 1);   // global loads in flight while this slab is multiplied
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[RM];
      if constexpr (RM % 4 == 0) {
#pragma unroll
        for (int i = 0; i < RM; i += 4) {
          const float4 a = *reinterpret_cast<const float4*>(&As[buf][kk][ty * RM + i]);
          av[i] = a.x; av[i + 1] = a.y; av[i + 2] = a.z; av[i + 3] = a.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < RM; ++i) av[i] = As[buf][kk][ty * RM + i];
      }
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (sl + 1 < slabs) stash(buf ^ 1);   // (the other buffer: its last readers passed the barrier of the previous slab)
    __syncthreads();
  }
  const int n = n0 + tx * 4;
  if (n >= N) return;
  float bb[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) bb[j] = (bias != nullptr && n + j < N) ? bias[n + j] : 0.f;
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    const int64_t m = m0 + ty * RM + i;
    if (m >= M) continue;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = apply_act(acc[i][j] + bb[j], act);
    if (n + 3 < N && (N & 3) == 0) {
      *reinterpret_cast<float4*>(Y + m * N + n) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n + j < N) Y[m * N + n + j] = v[j];
    }
  }
}

// ---- backward wrt input: dX[M,K] (+)= mask * dZ[M,N] W[N,K] -----------------------------------------------
__global__ void __launch_bounds__(kDenseThreads)
    dense_bwd_input_kernel(const float* __restrict__ dZ, const float* __restrict__ W, const int64_t* __restrict__ mask_ids,
                           int64_t mask_lt, float* __restrict__ dX, int64_t M, int N, int K, int accumulate) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int c0 = blockIdx.y * BN;  // output column (k) tile
  float acc[4][4] = {};
  for (int n0 = 0; n0 < N; n0 += BK) {
    stage_T(As, dZ, M, N, m0, n0, N, true);
    stage_N(Bs, W, N, K, n0, c0, K, nullptr);
    __syncthreads();
    micro_fma(acc, As, Bs, ty, tx);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const float mk = (mask_ids == nullptr || mask_ids[m] < mask_lt) ? 1.f : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = c0 + tx * 4 + j;
      if (k >= K) continue;
      const float v = mk * acc[i][j];
      if (accumulate) dX[m * K + k] += v; else dX[m * K + k] = v;
    }
  }
}

// ---- backward wrt weight: dW[N,K] += (mask*dZ)[M,N]^T X[M,K]; db[N] += colsum(dZ) -------------------------------
// grid = (N tiles, K tiles, M chunks); each CTA reduces kChunkM rows and adds its tile with fp32 atomics.
constexpr int kChunkM = 128;
__global__ void __launch_bounds__(kDenseThreads)
    dense_bwd_weight_kernel(const float* __restrict__ dZ, const float* __restrict__ X, const int64_t* __restrict__ mask_ids,
                            int64_t mask_lt, float* __restrict__ dW, float* __restrict__ db, int64_t M, int N, int K) {
  __shared__ __align__(16) float As[BK][BM + PAD];  // [m][n] slab of dZ
  __shared__ __align__(16) float Bs[BK][BN + PAD];  // [m][k] slab of X
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const int n0 = blockIdx.x * BM;
  const int c0 = blockIdx.y * BN;
  const int64_t mbeg = (int64_t)blockIdx.z * kChunkM;
  const int64_t mend = (mbeg + kChunkM < M) ? mbeg + kChunkM : M;
  MaskCtx mc{mask_ids, mask_lt};
  float acc[4][4] = {};
  float bsum = 0.f;  // threads 0..63 of CTAs with blockIdx.y == 0 own db[n0 + threadIdx.x]
  for (int64_t m0 = mbeg; m0 < mend; m0 += BK) {
    stage_N(As, dZ, mend, N, m0, n0, N, &mc);
    stage_N(Bs, X, mend, K, m0, c0, K, nullptr);
    __syncthreads();
    micro_fma(acc, As, Bs, ty, tx);
    if (db != nullptr && blockIdx.y == 0 && threadIdx.x < BM) {
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) bsum += As[kk][threadIdx.x];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = c0 + tx * 4 + j;
      if (k < K) atomicAdd(&dW[(int64_t)n * K + k], acc[i][j]);
    }
  }
  if (db != nullptr && blockIdx.y == 0 && threadIdx.x < BM && n0 + (int)threadIdx.x < N)
    atomicAdd(&db[n0 + threadIdx.x], bsum);
}

__global__ void act_bwd_kernel(const float* __restrict__ Y, const float* __restrict__ dY, int act, float* __restrict__ dZ,
                               int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float y = Y[i], g = dY[i];
    float d;
    switch (act) {
      case XDR_ACT_RELU: d = y > 0.f ? g : 0.f; break;
      case XDR_ACT_TANH: d = g * (1.f - y * y); break;
      case XDR_ACT_SIGMOID: d = g * (1.f - y) * y; break;
      default: d = g; break;
    }
    dZ[i] = d;
  }
}

// ---- MSE between dense rows Y[k,:] and gathered target rows T[idx[k],:] ----------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256) mse_rows_fwd_kernel(const float* __restrict__ Y, const float* __restrict__ T,
                                                           int64_t n_rows, int nv, const int64_t* __restrict__ idx,
                                                           int64_t n_idx, float* out8, Workspace ws, int32_t* oob) {
  __shared__ float smem[8];
  const int sub = threadIdx.x & (kLanesPerRow - 1);
  const int64_t group = ((int64_t)blockIdx.x * 256 + threadIdx.x) / kLanesPerRow;
  const int64_t n_groups = (int64_t)gridDim.x * 256 / kLanesPerRow;
  float acc[1] = {0.f};
  for (int64_t k = group; k < n_idx; k += n_groups) {
    const int64_t r = idx[k];
    const bool ok = (uint64_t)r < (uint64_t)n_rows;
    if (!ok && oob && sub == 0) *oob = 1;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int c = sub + j * kLanesPerRow;
      if (c >= nv) continue;
      const float4 y = ldg_row4(Y + k * (int64_t)nv * 4, c);
      const float4 t = ok ? ldg_row4(T + r * (int64_t)nv * 4, c) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 d = sub4(y, t);
      acc[0] += dot4(d, d);
    }
  }
  const double denom = (double)n_idx * (double)(nv * 4);
  grid_reduce_last_block<1>(acc, ws, smem, [=](double* tot) {
    out8[0] = (float)(tot[0] / denom);
    for (int i = 1; i < 8; ++i) out8[i] = 0.f;
  });
}

template <int VEC>
__global__ void __launch_bounds__(256) mse_rows_bwd_kernel(const float* __restrict__ Y, const float* T, int64_t n_rows,
                                                           int nv, const int64_t* __restrict__ idx, int64_t n_idx,
                                                           const float* __restrict__ grad_loss, float scale,
                                                           float* __restrict__ dY, float* tgt_dst) {
  const int sub = threadIdx.x & (kLanesPerRow - 1);
  const int64_t group = ((int64_t)blockIdx.x * 256 + threadIdx.x) / kLanesPerRow;
  const int64_t n_groups = (int64_t)gridDim.x * 256 / kLanesPerRow;
  const float g = (grad_loss ? __ldg(grad_loss) : 1.f) * 2.f / ((float)n_idx * (float)(nv * 4));
  for (int64_t k = group; k < n_idx; k += n_groups) {
    const int64_t r = idx[k];
    const bool ok = (uint64_t)r < (uint64_t)n_rows;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int c = sub + j * kLanesPerRow;
      if (c >= nv) continue;
      const float4 y = ldg_row4(Y + k * (int64_t)nv * 4, c);
      const float4 t = ok ? ld_row4(T + r * (int64_t)nv * 4, c) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 d = scale4(g, sub4(y, t));
      st4(dY + k * (int64_t)nv * 4, c, d);
      if (ok && tgt_dst) red_add4(tgt_dst + r * (int64_t)nv * 4, c, scale4(-scale, d));
    }
  }
}

// ---- BCE on a logit column ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bce_logit_fwd_kernel(const float* __restrict__ logit, const float* __restrict__ label,
                                                            int64_t n, float* __restrict__ prob, float* out8, Workspace ws) {
  __shared__ float smem[8];
  float acc[1] = {0.f};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float p = sigmoidf_(logit[i]), y = label[i];
    prob[i] = p;
    const float lp = fmaxf(logf(p), -100.f), lq = fmaxf(logf(1.f - p), -100.f);
    acc[0] += -(y * lp + (1.f - y) * lq);
  }
  const double denom = (double)n;
  grid_reduce_last_block<1>(acc, ws, smem, [=](double* tot) {
    out8[0] = (float)(tot[0] / denom);
    for (int i = 1; i < 8; ++i) out8[i] = 0.f;
  });
}

__global__ void bce_logit_bwd_kernel(const float* __restrict__ prob, const float* __restrict__ label, int64_t n,
                                     const float* __restrict__ grad_loss, float* __restrict__ dlogit) {
  const float g = (grad_loss ? __ldg(grad_loss) : 1.f) / (float)n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float p = prob[i], y = label[i];
    const float pq = p * (1.f - p);
    dlogit[i] = g * (p - y) / fmaxf(pq, 1e-12f) * pq;
  }
}

// ---- sum of the Frobenius norms of a few small matrices: CoNet's regulariser sum_l ||H_l||_F (conet.py:198-201) ---------
// torch runs it as a reduction + an add per matrix forward and four element-wise kernels per matrix backward (28 launches of
// 2-7 us in a CoNet step); here: one launch each way.  Forward is ONE CTA of 1024 threads going through the matrices in order
// (20 k elements at the yaml stack; first version: 256 threads, one scalar load per round = 64 dependent rounds, 37 us cold
// under ncu), so the sums have a fixed order.
constexpr int kMaxFrob = 8;
struct FrobArgs {
  const float* mat[kMaxFrob];
  float* dst[kMaxFrob];
  int64_t count[kMaxFrob];
  int n;
};

constexpr int kFrobThreads = 1024;
__global__ void __launch_bounds__(kFrobThreads) frob_sum_fwd_kernel(FrobArgs a, float* __restrict__ norms, float* __restrict__ out) {
  __shared__ float smem[kFrobThreads / 32];
  float total = 0.f;
  for (int l = 0; l < a.n; ++l) {
    const float* __restrict__ m = a.mat[l];
    const int64_t n = a.count[l];
    float acc[1] = {0.f};
    if (aligned16_dev(m)) {   // four independent 128-bit loads per thread and round: a 256 x 64 matrix is ONE round of loads
      const int64_t n4 = n / 4;
      for (int64_t i = threadIdx.x; i < n4; i += 4 * kFrobThreads) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int64_t j = i + (int64_t)u * kFrobThreads;
          v[u] = j < n4 ? ldg_row4(m, (int)j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[0] += dot4(v[u], v[u]);
      }
      for (int64_t i = n4 * 4 + threadIdx.x; i < n; i += kFrobThreads) acc[0] += m[i] * m[i];
    } else {
      for (int64_t i = threadIdx.x; i < n; i += kFrobThreads) acc[0] += m[i] * m[i];
    }
    block_sum<1>(acc, smem);
    if (threadIdx.x == 0) {
      const float nrm = sqrtf(acc[0]);
      norms[l] = nrm;
      total += nrm;
    }
  }
  if (threadIdx.x == 0) out[0] = total;
}

// d(sum_l ||H_l||_F)/dH_l = H_l / ||H_l||_F (torch: 0 where the norm is 0), times the upstream gradient
__global__ void __launch_bounds__(256) frob_sum_bwd_kernel(FrobArgs a, const float* __restrict__ norms,
                                                           const float* __restrict__ grad_loss) {
  const float g = grad_loss ? __ldg(grad_loss) : 1.f;
  for (int l = 0; l < a.n; ++l) {
    const float nrm = norms[l];
    const float sc = nrm > 0.f ? g / nrm : 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.count[l]; i += (int64_t)gridDim.x * blockDim.x)
      a.dst[l][i] = sc * a.mat[l][i];
  }
}

// ---- select + dot: score[b] = ((sel_ids[b] < n_overlap) ? mapped[b,:] : tgt_tab[sel_ids[b],:]) . other_tab[other_ids[b],:]
template <int VEC>
__global__ void __launch_bounds__(256) select_dot_kernel(const float* __restrict__ mapped, const float* __restrict__ tgt_tab,
                                                         int64_t n_sel_rows, const int64_t* __restrict__ sel_ids,
                                                         int64_t n_overlap, const float* __restrict__ other_tab,
                                                         int64_t n_other_rows, const int64_t* __restrict__ other_ids,
                                                         int nv, int64_t batch, float* __restrict__ score, int32_t* oob) {
  const int sub = threadIdx.x & (kLanesPerRow - 1);
  const int64_t group = ((int64_t)blockIdx.x * 256 + threadIdx.x) / kLanesPerRow;
  const int64_t n_groups = (int64_t)gridDim.x * 256 / kLanesPerRow;
  const int64_t warp_first = group - (group % kRowsPerWarp);
  for (int64_t base = warp_first; base < batch; base += n_groups) {
    const int64_t b = base + (group % kRowsPerWarp);
    const bool live = b < batch;
    const int64_t s = live ? sel_ids[b] : 0, o = live ? other_ids[b] : 0;
    const bool oks = live && (uint64_t)s < (uint64_t)n_sel_rows, oko = live && (uint64_t)o < (uint64_t)n_other_rows;
    if (live && oob && sub == 0 && (!oks || !oko)) *oob = 1;
    const bool use_map = s < n_overlap;
    float d = 0.f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int c = sub + j * kLanesPerRow;
      if (c >= nv || !oks || !oko) continue;
      const float4 e = use_map ? ldg_row4(mapped + b * (int64_t)nv * 4, c) : ldg_row4(tgt_tab + s * (int64_t)nv * 4, c);
      d += dot4(e, ldg_row4(other_tab + o * (int64_t)nv * 4, c));
    }
    d = group8_sum(d);
    if (live && sub == 0) score[b] = d;
  }
}

static inline int ew_grid(int64_t n, int threads) {
  int64_t b = (n + threads - 1) / threads;
  int64_t cap = (int64_t)sm_count() * 8;
  if (cap > kMaxBlocks) cap = kMaxBlocks;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace xdr

using namespace xdr;

extern "C" {

int xdr_dense_fwd(const float* X, const float* W, const float* bias, const float* X2, const float* W2,
                  const int64_t* mask_ids, int64_t mask_lt, int act, float* Y, int64_t M, int N, int K,
                  xdr_stream_t stream) {
  XDR_REQUIRE(M >= 0 && N > 0 && K > 0, "xdr_dense_fwd: bad shape M=%lld N=%d K=%d", (long long)M, N, K);
  if (M == 0) return XDR_OK;
  XDR_REQUIRE(X && W && Y, "xdr_dense_fwd: null pointer");
  XDR_REQUIRE((X2 == nullptr) == (W2 == nullptr), "xdr_dense_fwd: X2 and W2 must be given together");
  XDR_REQUIRE(act >= XDR_ACT_NONE && act <= XDR_ACT_SIGMOID, "xdr_dense_fwd: bad act %d", act);
  if (tc5_dense_fwd_ok(X, W, X2, W2, Y, M, N, K))   // tensor cores (tcgen05, bf16x3) for the shapes that are GEMMs
    return tc5_dense_fwd(X, W, bias, X2, W2, mask_ids, mask_lt, act, Y, M, N, K, (cudaStream_t)stream);
  // (measured, call 29: the 128-row tiles win for narrow layers -- N <= 32: 12.6 vs 16.7 us, 7.4 vs 10.5 us at the CoNet shapes --
  // and lose for N >= 64, where 128 CTAs of 8 warps leave the FMA pipe under-occupied: 55.6 vs 51.5 us; XDR_DENSE_FWD2=2 forces them)
  if (g_dense_fwd2 && (N <= 32 || g_dense_fwd2 == 2) && M >= BM2 && (K & 3) == 0 && aligned16(X) && aligned16(W) && aligned16(Y) && aligned16(X2) && aligned16(W2)) {
    // register-blocked, double-buffered tiles (128 x 64 / 128 x 32 / 128 x 16 by the layer's width)
    const int bn = N > 32 ? 64 : (N > 16 ? 32 : 16);
    dim3 grid((unsigned)((M + BM2 - 1) / BM2), (unsigned)((N + bn - 1) / bn));
    if (bn == 64)
      XDR_LAUNCH((dense_fwd2_kernel<64>), grid, kDenseThreads, 0, (cudaStream_t)stream, X, W, bias, X2, W2, mask_ids, mask_lt, act, Y, M, N, K);
    else if (bn == 32)
      XDR_LAUNCH((dense_fwd2_kernel<32>), grid, kDenseThreads, 0, (cudaStream_t)stream, X, W, bias, X2, W2, mask_ids, mask_lt, act, Y, M, N, K);
    else
      XDR_LAUNCH((dense_fwd2_kernel<16>), grid, kDenseThreads, 0, (cudaStream_t)stream, X, W, bias, X2, W2, mask_ids, mask_lt, act, Y, M, N, K);
    XDR_LAUNCH_OK();
    return XDR_OK;
  }
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((N + BN - 1) / BN));
  XDR_LAUNCH((dense_fwd_kernel), grid, kDenseThreads, 0, (cudaStream_t)stream, X, W, bias, X2, W2, mask_ids, mask_lt, act, Y, M, N, K);
  XDR_LAUNCH_OK();
  return XDR_OK;
}

int xdr_act_bwd(const float* Y, const float* dY, int act, float* dZ, int64_t count, xdr_stream_t stream) {
  XDR_REQUIRE(count >= 0, "xdr_act_bwd: negative count");
  if (count == 0) return XDR_OK;
  XDR_REQUIRE(Y && dY && dZ, "xdr_act_bwd: null pointer");
  XDR_LAUNCH((act_bwd_kernel), ew_grid(count, 256), 256, 0, (cudaStream_t)stream, Y, dY, act, dZ, count);
  XDR_LAUNCH_OK();
  return XDR_OK;
}

int xdr_dense_bwd_input(const float* dZ, const float* W, const int64_t* mask_ids, int64_t mask_lt, float* dX, int64_t M,
                        int N, int K, int accumulate, xdr_stream_t stream) {
  XDR_REQUIRE(M >= 0 && N > 0 && K > 0, "xdr_dense_bwd_input: bad shape");
  if (M == 0) return XDR_OK;
  XDR_REQUIRE(dZ && W && dX, "xdr_dense_bwd_input: null pointer");
  if (tc5_dense_bwd_input_ok(dZ, W, dX, M, N, K))
    return tc5_dense_bwd_input(dZ, W, mask_ids, mask_lt, dX, M, N, K, accumulate, (cudaStream_t)stream);
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((K + BN - 1) / BN));
  XDR_LAUNCH((dense_bwd_input_kernel), grid, kDenseThreads, 0, (cudaStream_t)stream, dZ, W, mask_ids, mask_lt, dX, M, N, K, accumulate);
  XDR_LAUNCH_OK();
  return XDR_OK;
}

int xdr_dense_bwd_weight(const float* dZ, const float* X, const int64_t* mask_ids, int64_t mask_lt, float* dW, float* db,
                         int64_t M, int N, int K, xdr_stream_t stream) {
  XDR_REQUIRE(M >= 0 && N > 0 && K > 0, "xdr_dense_bwd_weight: bad shape");
  if (M == 0) return XDR_OK;
  XDR_REQUIRE(dZ && X && dW, "xdr_dense_bwd_weight: null pointer");
  if (tc5_dense_bwd_weight_ok(dZ, X, dW, M, N, K))
    return tc5_dense_bwd_weight(dZ, X, mask_ids, mask_lt, dW, db, M, N, K, (cudaStream_t)stream);
  dim3 grid((unsigned)((N + BM - 1) / BM), (unsigned)((K + BN - 1) / BN), (unsigned)((M + kChunkM - 1) / kChunkM));
  XDR_LAUNCH((dense_bwd_weight_kernel), grid, kDenseThreads, 0, (cudaStream_t)stream, dZ, X, mask_ids, mask_lt, dW, db, M, N, K);
  XDR_LAUNCH_OK();
  return XDR_OK;
}

int xdr_mse_rows_fwd(const float* Y, const float* tgt_tab, int64_t n_rows, int dim, const int64_t* idx, int64_t n_idx,
                     float* out8, void* ws, int32_t* oob, xdr_stream_t stream) {
  XDR_REQUIRE(dim_ok(dim), "xdr_mse_rows_fwd: dim=%d must be a multiple of 4 in (0, 256]", dim);
  XDR_REQUIRE(n_idx > 0, "xdr_mse_rows_fwd: n_idx must be positive");
  XDR_REQUIRE(Y && tgt_tab && idx && out8 && ws, "xdr_mse_rows_fwd: null pointer");
  XDR_REQUIRE(aligned16(Y) && aligned16(tgt_tab), "xdr_mse_rows_fwd: 16-byte alignment");
  const int nv = dim / 4;
  const int grid = ew_grid(n_idx * kLanesPerRow, 256);
  XDR_DISPATCH_VEC(nv, (XDR_LAUNCH((mse_rows_fwd_kernel<VEC>), grid, 256, 0, (cudaStream_t)stream, Y, tgt_tab, n_rows, nv, idx, n_idx,
                                                                                         out8, Workspace(ws), oob)));
  XDR_LAUNCH_OK();
  return XDR_OK;
}

int xdr_mse_rows_bwd(const float* Y, const float* tgt_tab, int64_t n_rows, int dim, const int64_t* idx, int64_t n_idx,
                     const float* grad_loss, float scale, float* dY, float* tgt_dst, xdr_stream_t stream) {
  XDR_REQUIRE(dim_ok(dim), "xdr_mse_rows_bwd: dim=%d must be a multiple of 4 in (0, 256]", dim);
  XDR_REQUIRE(n_idx > 0, "xdr_mse_rows_bwd: n_idx must be positive");
  XDR_REQUIRE(Y && tgt_tab && idx && dY, "xdr_mse_rows_bwd: null pointer");
  XDR_REQUIRE(aligned16(Y) && aligned16(tgt_tab) && aligned16(dY) && aligned16(tgt_dst), "xdr_mse_rows_bwd: alignment");
  const int nv = dim / 4;
  const int grid = ew_grid(n_idx * kLanesPerRow, 256);
  XDR_DISPATCH_VEC(nv, (XDR_LAUNCH((mse_rows_bwd_kernel<VEC>), grid, 256, 0, (cudaStream_t)stream, Y, tgt_tab, n_rows, nv, idx, n_idx,
                                                                                         grad_loss, scale, dY, tgt_dst)));
  XDR_LAUNCH_OK();
  return XDR_OK;
}

int xdr_bce_logit_fwd(const float* logit, const float* label, int64_t count, float* prob, float* out8, void* ws,
                      xdr_stream_t stream) {
  XDR_REQUIRE(count > 0, "xdr_bce_logit_fwd: count must be positive");
  XDR_REQUIRE(logit && label && prob && out8 && ws, "xdr_bce_logit_fwd: null pointer");
  XDR_LAUNCH((bce_logit_fwd_kernel), ew_grid(count, 256), 256, 0, (cudaStream_t)stream, logit, label, count, prob, out8, Workspace(ws));
  XDR_LAUNCH_OK();
  return XDR_OK;
}

int xdr_bce_logit_bwd(const float* prob, const float* label, int64_t count, const float* grad_loss, float* dlogit,
                      xdr_stream_t stream) {
  XDR_REQUIRE(count > 0, "xdr_bce_logit_bwd: count must be positive");
  XDR_REQUIRE(prob && label && dlogit, "xdr_bce_logit_bwd: null pointer");
  XDR_LAUNCH((bce_logit_bwd_kernel), ew_grid(count, 256), 256, 0, (cudaStream_t)stream, prob, label, count, grad_loss, dlogit);
  XDR_LAUNCH_OK();
  return XDR_OK;
}

static int frob_args(const char* fn, const float* const* mats, float* const* dsts, const int64_t* counts, int n_mats,
                     FrobArgs* a) {
  XDR_REQUIRE(n_mats > 0 && n_mats <= kMaxFrob, "%s: n_mats=%d must be in [1, %d]", fn, n_mats, kMaxFrob);
  XDR_REQUIRE(mats && counts && (dsts != nullptr || a->n == 0), "%s: null pointer", fn);
  a->n = n_mats;
  for (int l = 0; l < n_mats; ++l) {
    XDR_REQUIRE(mats[l] && counts[l] > 0 && (dsts == nullptr || dsts[l]), "%s: matrix %d is null or empty", fn, l);
    a->mat[l] = mats[l];
    a->dst[l] = dsts ? dsts[l] : nullptr;
    a->count[l] = counts[l];
  }
  return XDR_OK;
}

int xdr_frob_sum_fwd(const float* const* mats_host, const int64_t* counts_host, int n_mats, float* norms, float* out,
                     xdr_stream_t stream) {
  FrobArgs a;
  a.n = 0;   // (no destinations on the way forward)
  const int rc = frob_args("xdr_frob_sum_fwd", mats_host, nullptr, counts_host, n_mats, &a);
  if (rc != XDR_OK) return rc;
  XDR_REQUIRE(norms && out, "xdr_frob_sum_fwd: null pointer");
  XDR_LAUNCH((frob_sum_fwd_kernel), 1, kFrobThreads, 0, (cudaStream_t)stream, a, norms, out);
  XDR_LAUNCH_OK();
  return XDR_OK;
}

int xdr_frob_sum_bwd(const float* const* mats_host, const int64_t* counts_host, int n_mats, const float* norms,
                     const float* grad_loss, float* const* dsts_host, xdr_stream_t stream) {
  FrobArgs a;
  a.n = 1;   // (destinations required)
  XDR_REQUIRE(dsts_host && norms, "xdr_frob_sum_bwd: null pointer");
  const int rc = frob_args("xdr_frob_sum_bwd", mats_host, dsts_host, counts_host, n_mats, &a);
  if (rc != XDR_OK) return rc;
  int64_t most = 0;
  for (int l = 0; l < n_mats; ++l) most = counts_host[l] > most ? counts_host[l] : most;
  XDR_LAUNCH((frob_sum_bwd_kernel), ew_grid(most, 256), 256, 0, (cudaStream_t)stream, a, norms, grad_loss);
  XDR_LAUNCH_OK();
  return XDR_OK;
}

int xdr_select_dot(const float* mapped, const float* tgt_tab, int64_t n_sel_rows, const int64_t* sel_ids,
                   int64_t n_overlap, const float* other_tab, int64_t n_other_rows, const int64_t* other_ids, int dim,
                   int64_t batch, float* score, int32_t* oob, xdr_stream_t stream) {
  XDR_REQUIRE(dim_ok(dim), "xdr_select_dot: dim=%d must be a multiple of 4 in (0, 256]", dim);
  XDR_REQUIRE(batch >= 0, "xdr_select_dot: negative batch");
  if (batch == 0) return XDR_OK;
  XDR_REQUIRE(mapped && tgt_tab && sel_ids && other_tab && other_ids && score, "xdr_select_dot: null pointer");
  XDR_REQUIRE(aligned16(mapped) && aligned16(tgt_tab) && aligned16(other_tab), "xdr_select_dot: 16-byte alignment");
  const int nv = dim / 4;
  const int grid = ew_grid(batch * kLanesPerRow, 256);
  XDR_DISPATCH_VEC(nv, (XDR_LAUNCH((select_dot_kernel<VEC>), grid, 256, 0, (cudaStream_t)stream, 
                           mapped, tgt_tab, n_sel_rows, sel_ids, n_overlap, other_tab, n_other_rows, other_ids, nv, batch,
                           score, oob)));
  XDR_LAUNCH_OK();
  return XDR_OK;
}

}  // extern "C"
