// tc_conet.cu -- A7-A8 on the tensor cores: one CoNet tower pass (both towers through the cross-stitch stack, BCE on the
// wanted tower) with its whole backward and the embedding scatter-add in ONE kernel.
//
// Replaces, per domain batch, CoNet.source_forward / target_forward + nn.BCELoss + their autograd backward
// (reference conet.py:105-181, 183-197):
//     x_s = [Es_u[u] | Es_i[i]],  x_t = [Et_u[u] | Et_i[i]]                                        (conet.py:106-111)
//     per layer l:  x_s' = relu(Ws_l x_s + bs_l + m * (H_l x_t)),  x_t' = relu(Wt_l x_t + bt_l + m * (H_l x_s))
//                   with the SAME H_l = crossparas[l].weight both ways and m = (id < n_overlap)     (conet.py:113-138)
//     p = sigmoid(w_out . x_want + b_out),  loss = BCE(p, label)                                   (conet.py:140, 196-197)
// The composed path (dense.cu) runs ~60 kernels per pass and writes every activation to HBM; here the only HBM traffic is
// ids + labels + the 4 gathered rows (read twice, the second time from L2) + the 4 scattered gradient rows + a
// [batch, 2*hidden_0] scratch of layer-1 input gradients (L2 resident).  Every product is a 3xTF32 mma.sync tile (tc_tile.cuh).
//
// Layer 0 (K = 2*dim up to 512, three [hidden_0, 2*dim] matrices = 196 KB at dim 128) does not fit shared memory next to
// a row tile, so it is processed in K chunks of 64 columns.  Two phases per CTA over its own 64-row tiles:
//   phase 1, per tile:   for each K chunk: gather the chunk of x_s, x_t + load the chunk of Ws_0, Wt_0, H_0 -> accumulate
//                        the four layer-0 products in registers;  tail layers, head, tail backward with the tail weights
//                        resident in shared memory and their gradients accumulated in registers over all tiles;
//                        the gradient of the layer-0 output goes to the scratch.
//   phase 2, per K chunk: load the weight chunk once, then per tile: re-gather the x chunk (L2), read the tile's scratch
//                        rows, accumulate the chunk of dWs_0, dWt_0, dH_0 in registers, form the chunk of dx_s, dx_t and
//                        scatter-add it (red.global.add.v4.f32) into the four gradient tables.
// Cross terms only matter for overlapped rows (m = 1): each CTA re-orders its rows so that the overlapped ones come first,
// and a 16-row MMA tile without any overlapped row skips its cross products (forward c_s / c_t, the m * (dZ H) half of the
// input gradients, the masked halves of dH) -- at 50 % overlap that is a quarter of the layer-0 MMAs.
// Both towers run through every layer; only the wanted tower feeds the loss, so the other tower's last-layer parameters
// receive exact zeros (PyTorch leaves their .grad untouched -- None/zero -- for the same reason).
//
// STATUS: written in a session without GPU access -- compiles for sm_100a and its logic runs under the CPU CTA emulator
// (tests/test_emu_conet.py), NOT yet executed on hardware; opt-in (config `xdr_fused_conet: True`), gpu tests `unvalidated`.
#include "tc_tile.cuh"

namespace xdr {

constexpr int kCnMaxLayers = 4;  // cross-stitch layers
constexpr int kCnTR = 64;        // batch rows per tile
constexpr int kCnKC = 64;        // layer-0 K chunk (columns of the concatenated input)
constexpr int kCnMaxHidden = 64; // widest hidden layer
constexpr int kCnNT0 = 4;        // layer-0 output tiles per warp: hidden_0 / 8 / (kTcWarps / (kCnTR / 16))
constexpr int kCnDw0 = 4;        // layer-0 chunk weight-gradient tiles per warp and matrix: ceil(64/16) * (64/8) / 8
constexpr int kCnDw1 = 4, kCnDw2 = 2, kCnDw3 = 1;  // tail layers 1..3, per matrix
constexpr int kCnMaxPerm = 2048; // rows of one CTA that can be re-ordered (overlapped rows first)

struct ConetArgs {
  int n_layers;                  // L cross-stitch layers (1..4)
  int dims[kCnMaxLayers + 1];    // dims[0] = 2*dim, then mlp_hidden_size
  const float* Ws[kCnMaxLayers]; // source_crossunit_linear[l].weight [dims[l+1], dims[l]]
  const float* bs[kCnMaxLayers];
  const float* Wt[kCnMaxLayers]; // target_crossunit_linear[l].weight
  const float* bt[kCnMaxLayers];
  const float* H[kCnMaxLayers];  // crossparas[l].weight [dims[l+1], dims[l]]
  float* dWs[kCnMaxLayers];
  float* dbs[kCnMaxLayers];
  float* dWt[kCnMaxLayers];
  float* dbt[kCnMaxLayers];
  float* dH[kCnMaxLayers];
  const float* w_out;            // wanted tower's output unit: weight [1, dims[L]], bias [1]
  const float* b_out;
  float* dw_out;
  float* db_out;
  int want;                      // 0: source tower feeds the loss, 1: target tower
  const float *Su, *Si, *Tu, *Ti;
  float *dSu, *dSi, *dTu, *dTi;
  int64_t n_u, n_i;
  int dim;
  const int64_t* user;
  const int64_t* item;
  const float* label;
  int64_t batch;
  int mask_on_item;              // 0: m = user < n_overlap (overlap_users), 1: m = item < n_overlap
  int64_t n_overlap;
  float* dz1;                    // scratch [batch][2*dims[1]]: gradient of the layer-0 pre-activations (source | target)
  int backward;
  const float* grad_loss;
  float scale;
  float* prob;                   // optional [batch]
  float* out8;
  int32_t* oob;
};

struct ConetSmem {
  int tailW[kCnMaxLayers][3];    // float offsets of Ws_l, Wt_l, H_l (l >= 1), row stride dims[l] + 4
  int bias[kCnMaxLayers][2];     // bs_l, bt_l
  int wout, mask, grow, trow, perm, bufX[2], wch[3];
  int act[kCnMaxLayers + 1][2];  // l >= 2 (l == 1 aliases bufX)
  int grd[kCnMaxLayers + 1][2];  // l >= 2
  int dz[2];                     // phase 2: the tile's scratch rows
  int total;                     // floats
};

__host__ __device__ inline ConetSmem conet_smem_layout(int L, const int* dims) {
  ConetSmem s{};
  int p = 0;
  auto up4 = [](int v) { return (v + 3) & ~3; };
  for (int l = 1; l < L; ++l)
    for (int m = 0; m < 3; ++m) { s.tailW[l][m] = p; p += dims[l + 1] * (dims[l] + 4); }
  for (int l = 0; l < L; ++l)
    for (int m = 0; m < 2; ++m) { s.bias[l][m] = p; p += up4(dims[l + 1]); }
  s.wout = p; p += up4(dims[L]) + 4;
  s.mask = p; p += kCnTR;
  s.grow = p; p += kCnTR;
  s.trow = p; p += kCnTR;        // global row of each slot of the current tile (int32, -1 = none)
  s.perm = p; p += kCnMaxPerm;   // this CTA's rows, overlapped ones first (int32)
  const int ldc = kCnKC + 4;
  for (int m = 0; m < 2; ++m) { s.bufX[m] = p; p += kCnTR * ldc; }
  for (int m = 0; m < 3; ++m) { s.wch[m] = p; p += dims[1] * ldc; }
  int p1 = p;  // phase 1 region
  for (int l = 2; l <= L; ++l)
    for (int m = 0; m < 2; ++m) {
      s.act[l][m] = p1; p1 += kCnTR * (dims[l] + 4);
      s.grd[l][m] = p1; p1 += kCnTR * (dims[l] + 4);
    }
  int p2 = p;  // phase 2 region (same base)
  for (int m = 0; m < 2; ++m) { s.dz[m] = p2; p2 += kCnTR * (dims[1] + 4); }
  s.total = p1 > p2 ? p1 : p2;
  return s;
}

#if defined(__CUDACC__) || defined(XDR_EMU)

// One float4 of the concatenated input row [user row | item row] of a tower: table pointer, id and local column.
struct CnCol {
  const float* tab;
  float* dtab;
  int64_t id;
  int col4;
  bool ok;
};
__device__ __forceinline__ CnCol conet_col(const ConetArgs& a, int tower, int64_t row, int gc) {
  CnCol c;
  const bool item = gc >= a.dim;
  c.id = item ? a.item[row] : a.user[row];
  c.col4 = (gc - (item ? a.dim : 0)) >> 2;
  c.tab = tower == 0 ? (item ? a.Si : a.Su) : (item ? a.Ti : a.Tu);
  c.dtab = tower == 0 ? (item ? a.dSi : a.dSu) : (item ? a.dTi : a.dTu);
  c.ok = (uint64_t)c.id < (uint64_t)(item ? a.n_i : a.n_u);
  return c;
}

// gathers columns [kc*KC, (kc+1)*KC) of x_s and x_t for the tile's rows (trow[r] = global row of slot r, -1 = empty slot)
// and refreshes the overlap mask
__device__ __forceinline__ void conet_gather_chunk(const ConetArgs& a, const int* __restrict__ trow, int kc, float* xs,
                                                   float* xt, float* mask) {
  constexpr int C4 = kCnKC / 4, ldc = kCnKC + 4;
  for (int e = threadIdx.x; e < 2 * kCnTR * C4; e += kTcThreads) {
    const int tower = e / (kCnTR * C4), rem = e - tower * (kCnTR * C4);
    const int r = rem / C4, c4 = rem - r * C4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const int row = trow[r];
    if (row >= 0) {
      const CnCol c = conet_col(a, tower, row, kc * kCnKC + 4 * c4);
      if (c.ok) v = ld_row4(c.tab + c.id * a.dim, c.col4);
      else if (a.oob) *a.oob = 1;
    }
    *reinterpret_cast<float4*>((tower ? xt : xs) + r * ldc + 4 * c4) = v;
  }
  for (int r = threadIdx.x; r < kCnTR; r += kTcThreads) {
    float m = 0.f;
    const int row = trow[r];
    if (row >= 0) {
      const int64_t id = a.mask_on_item ? a.item[row] : a.user[row];
      m = id < a.n_overlap ? 1.f : 0.f;
    }
    mask[r] = m;
  }
}

// true when any of the 16 rows of this warp's MMA row tile is overlapped (warp-uniform): otherwise its cross products vanish
__device__ __forceinline__ bool conet_mtile_overlaps(const float* __restrict__ mask) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = (warp % (kCnTR / 16)) * 16;
  return __any_sync(0xffffffffu, mask[m0 + (lane & 15)] != 0.f) != 0;
}

// loads columns [kc*KC, (kc+1)*KC) of Ws_0, Wt_0, H_0 ([N1][K0] each) as three [N1][KC + 4] slabs
__device__ __forceinline__ void conet_load_wchunk(const ConetArgs& a, int kc, float* const (&wch)[3]) {
  constexpr int C4 = kCnKC / 4, ldc = kCnKC + 4;
  const int N1 = a.dims[1], K0 = a.dims[0];
  for (int e = threadIdx.x; e < 3 * N1 * C4; e += kTcThreads) {
    const int m = e / (N1 * C4), rem = e - m * (N1 * C4);
    const int n = rem / C4, c4 = rem - n * C4;
    const float* W = m == 0 ? a.Ws[0] : (m == 1 ? a.Wt[0] : a.H[0]);
    *reinterpret_cast<float4*>(wch[m] + n * ldc + 4 * c4) =
        *reinterpret_cast<const float4*>(W + (size_t)n * K0 + kc * kCnKC + 4 * c4);
  }
}

__global__ void __launch_bounds__(kTcThreads, 1) tc_conet_kernel(ConetArgs a, Workspace ws) {
  XDR_DYN_SMEM(float, smem);
  __shared__ float red_smem[8];
  constexpr int TR = kCnTR, KC = kCnKC, ldc = kCnKC + 4;
  const int tid = threadIdx.x;
  const int L = a.n_layers, K0 = a.dims[0], N1 = a.dims[1], dL = a.dims[L];
  const int nkc = K0 / KC;
  const ConetSmem lay = conet_smem_layout(L, a.dims);
  float* const mask = smem + lay.mask;
  float* const grow = smem + lay.grow;
  float* const xs = smem + lay.bufX[0];
  float* const xt = smem + lay.bufX[1];
  float* const wch[3] = {smem + lay.wch[0], smem + lay.wch[1], smem + lay.wch[2]};
  float* const wout = smem + lay.wout;
  // activations of layer l's input (l = 1: aliases the x chunk buffers, which are dead once layer 0 is done)
  auto act = [&](int l, int tower) { return l == 1 ? (tower ? xt : xs) : smem + lay.act[l][tower]; };
  auto lda = [&](int l) { return a.dims[l] + 4; };
  auto grd = [&](int l, int tower) { return smem + lay.grd[l][tower]; };

  // ---- resident parameters: tail weights, all biases, the output unit -----------------------------------------------------
  for (int l = 1; l < L; ++l) {
    const int din = a.dims[l], dout = a.dims[l + 1];
    for (int m = 0; m < 3; ++m) {
      const float* W = m == 0 ? a.Ws[l] : (m == 1 ? a.Wt[l] : a.H[l]);
      float* dst = smem + lay.tailW[l][m];
      for (int e = tid; e < din * dout; e += kTcThreads) {
        const int n = e / din, k = e - n * din;
        dst[n * (din + 4) + k] = W[e];
      }
    }
  }
  for (int l = 0; l < L; ++l)
    for (int n = tid; n < a.dims[l + 1]; n += kTcThreads) {
      smem[lay.bias[l][0] + n] = a.bs[l] ? a.bs[l][n] : 0.f;
      smem[lay.bias[l][1] + n] = a.bt[l] ? a.bt[l][n] : 0.f;
    }
  for (int k = tid; k < dL; k += kTcThreads) wout[k] = a.w_out[k];

  // ---- this CTA's rows: tiles blockIdx.x, blockIdx.x + gridDim.x, ...  Slot p of that list (p = local tile * 64 + r) is
  // batch row perm[p]; overlapped rows come first, so that whole 16-row MMA tiles without any overlapped row can skip
  // their cross products (the reference computes them for every row and multiplies by zero, conet.py:128-134).
  const int64_t n_tiles = (a.batch + TR - 1) / TR;
  const int my_tiles = (int)((n_tiles - (int64_t)blockIdx.x + (int64_t)gridDim.x - 1) / (int64_t)gridDim.x);
  int* const trow = reinterpret_cast<int*>(smem + lay.trow);
  int* const perm = reinterpret_cast<int*>(smem + lay.perm);
  const bool use_perm = my_tiles * TR <= kCnMaxPerm;
  auto natural_row = [&](int p) -> int64_t {   // slot p in tile order; -1 past the batch
    const int64_t row = ((int64_t)blockIdx.x + (int64_t)(p / TR) * gridDim.x) * TR + (p % TR);
    return row < a.batch ? row : (int64_t)-1;
  };
  if (use_perm && tid < 32) {  // warp 0: stable partition (overlapped | others | empty) by ballot prefix sums
    const int P = my_tiles * TR;
    int n_ov = 0, n_valid = 0;
    for (int p0 = 0; p0 < P; p0 += 32) {
      const int64_t row = natural_row(p0 + tid);
      bool ov = false;
      if (row >= 0) ov = (a.mask_on_item ? a.item[row] : a.user[row]) < a.n_overlap;
      n_ov += __popc(__ballot_sync(0xffffffffu, ov));
      n_valid += __popc(__ballot_sync(0xffffffffu, row >= 0));
    }
    int at_ov = 0, at_rest = n_ov, at_none = n_valid;
    for (int p0 = 0; p0 < P; p0 += 32) {
      const int64_t row = natural_row(p0 + tid);
      bool ov = false;
      if (row >= 0) ov = (a.mask_on_item ? a.item[row] : a.user[row]) < a.n_overlap;
      const unsigned b_ov = __ballot_sync(0xffffffffu, ov), b_rest = __ballot_sync(0xffffffffu, row >= 0 && !ov),
                     b_none = __ballot_sync(0xffffffffu, row < 0);
      const unsigned below = (1u << tid) - 1u;
      if (ov) perm[at_ov + __popc(b_ov & below)] = (int)row;
      else if (row >= 0) perm[at_rest + __popc(b_rest & below)] = (int)row;
      else perm[at_none + __popc(b_none & below)] = -1;
      at_ov += __popc(b_ov);
      at_rest += __popc(b_rest);
      at_none += __popc(b_none);
    }
  }
  __syncthreads();
  auto load_tile_rows = [&](int ti) {  // fills trow for local tile ti (followed by a barrier at the call sites)
    for (int r = tid; r < TR; r += kTcThreads) trow[r] = use_perm ? perm[ti * TR + r] : (int)natural_row(ti * TR + r);
  };

  // weight-gradient accumulators of the tail layers (MMA C fragments, alive over all tiles of phase 1)
  float aWs1[kCnDw1][4], aWt1[kCnDw1][4], aH1[kCnDw1][4];
  float aWs2[kCnDw2][4], aWt2[kCnDw2][4], aH2[kCnDw2][4];
  float aWs3[kCnDw3][4], aWt3[kCnDw3][4], aH3[kCnDw3][4];
  tile_acc_zero(aWs1); tile_acc_zero(aWt1); tile_acc_zero(aH1);
  tile_acc_zero(aWs2); tile_acc_zero(aWt2); tile_acc_zero(aH2);
  tile_acc_zero(aWs3); tile_acc_zero(aWt3); tile_acc_zero(aH3);
  float accB[kCnMaxLayers] = {0.f, 0.f, 0.f, 0.f};  // tid < dout: dbs_l[tid];  128 <= tid < 128 + dout: dbt_l[tid - 128]
  float accOut = 0.f;                               // tid < dL: dw_out[tid];  tid == 64: db_out  (dL <= 64)

  const float g_up = (a.grad_loss ? __ldg(a.grad_loss) : 1.0f);
  float loss_acc[1] = {0.f};

  // =============================== phase 1: forward, head, tail backward ===================================================
  for (int ti = 0; ti < my_tiles; ++ti) {
    load_tile_rows(ti);
    __syncthreads();
    bool cross = true;  // does this warp's 16-row MMA tile hold an overlapped row? (set once the mask is in shared memory)
    // ---- layer 0, K-chunked: a_s = x_s Ws^T, c_s = x_t H^T, a_t = x_t Wt^T, c_t = x_s H^T -------------------------------
    {
      float a_s[kCnNT0][4], c_s[kCnNT0][4], a_t[kCnNT0][4], c_t[kCnNT0][4];
      tile_acc_zero(a_s); tile_acc_zero(c_s); tile_acc_zero(a_t); tile_acc_zero(c_t);
      for (int kc = 0; kc < nkc; ++kc) {
        conet_gather_chunk(a, trow, kc, xs, xt, mask);
        conet_load_wchunk(a, kc, wch);
        __syncthreads();
        if (kc == 0) cross = conet_mtile_overlaps(mask);
        tile_mma_acc2<TR, kCnNT0, false>(a_s, c_t, xs, ldc, wch[0], ldc, wch[2], ldc, N1, KC, cross);  // x_s: Ws and H
        tile_mma_acc2<TR, kCnNT0, false>(a_t, c_s, xt, ldc, wch[1], ldc, wch[2], ldc, N1, KC, cross);  // x_t: Wt and H
        __syncthreads();
      }
      const float* b0s = smem + lay.bias[0][0];
      const float* b0t = smem + lay.bias[0][1];
      const int ld1 = lda(1);
      tile_acc_visit<TR, kCnNT0>(N1, [&](int j, int row, int col, int h) {
        const float m = mask[row];
        *reinterpret_cast<float2*>(xs + row * ld1 + col) =
            make_float2(fmaxf(a_s[j][2 * h] + b0s[col] + m * c_s[j][2 * h], 0.f),
                        fmaxf(a_s[j][2 * h + 1] + b0s[col + 1] + m * c_s[j][2 * h + 1], 0.f));
        *reinterpret_cast<float2*>(xt + row * ld1 + col) =
            make_float2(fmaxf(a_t[j][2 * h] + b0t[col] + m * c_t[j][2 * h], 0.f),
                        fmaxf(a_t[j][2 * h + 1] + b0t[col + 1] + m * c_t[j][2 * h + 1], 0.f));
      });
    }
    __syncthreads();
    // ---- tail layers, weights resident ------------------------------------------------------------------------------------
    for (int l = 1; l < L; ++l) {
      const int din = a.dims[l], dout = a.dims[l + 1], ldi = lda(l), ldo = lda(l + 1), ldw = din + 4;
      const float* Wsl = smem + lay.tailW[l][0];
      const float* Wtl = smem + lay.tailW[l][1];
      const float* Hl = smem + lay.tailW[l][2];
      const float* bsl = smem + lay.bias[l][0];
      const float* btl = smem + lay.bias[l][1];
      float* os = act(l + 1, 0);
      float* ot = act(l + 1, 1);
      tile_gemm2_any<TR, false>(act(l, 0), ldi, Wsl, ldw, act(l, 1), ldi, Hl, ldw, dout, din, cross,
                                [&](int row, int col, float v0, float v1, float c0, float c1) {
                                  const float m = mask[row];
                                  *reinterpret_cast<float2*>(os + row * ldo + col) =
                                      make_float2(fmaxf(v0 + bsl[col] + m * c0, 0.f), fmaxf(v1 + bsl[col + 1] + m * c1, 0.f));
                                });
      tile_gemm2_any<TR, false>(act(l, 1), ldi, Wtl, ldw, act(l, 0), ldi, Hl, ldw, dout, din, cross,
                                [&](int row, int col, float v0, float v1, float c0, float c1) {
                                  const float m = mask[row];
                                  *reinterpret_cast<float2*>(ot + row * ldo + col) =
                                      make_float2(fmaxf(v0 + btl[col] + m * c0, 0.f), fmaxf(v1 + btl[col + 1] + m * c1, 0.f));
                                });
      __syncthreads();
    }
    // ---- head on the wanted tower: logit -> sigmoid -> BCE; gradient of the logit per row ---------------------------------
    const float* aw = act(L, a.want);
    const int ldL = lda(L);
    {
      const float gs = g_up / (float)a.batch;
      const float bo = a.b_out ? __ldg(a.b_out) : 0.f;
      for (int r = tid; r < TR; r += kTcThreads) {
        float g = 0.f;
        const int row = trow[r];
        if (row >= 0) {
          float z = bo;
          for (int k = 0; k < dL; ++k) z = fmaf(aw[r * ldL + k], wout[k], z);
          const float pz = sigmoidf_(z), y = a.label[row];
          loss_acc[0] += -(y * fmaxf(logf(pz), -100.f) + (1.f - y) * fmaxf(logf(1.f - pz), -100.f));
          if (a.prob) a.prob[row] = pz;
          const float pq = pz * (1.f - pz);
          g = gs * (pz - y) / fmaxf(pq, 1e-12f) * pq;
        }
        grow[r] = g;
      }
    }
    __syncthreads();
    if (!a.backward) continue;  // (the loop-top gather is behind the barrier above)
    // output unit gradients
    if (tid < dL) {
      float s = 0.f;
      for (int r = 0; r < TR; ++r) s = fmaf(grow[r], aw[r * ldL + tid], s);
      accOut += s;
    } else if (tid == kCnMaxHidden) {
      float s = 0.f;
      for (int r = 0; r < TR; ++r) s += grow[r];
      accOut += s;
    }
    // gradient of the last cross layer's pre-activations: wanted tower g * w_out * relu'(y), other tower 0
    for (int e = tid; e < TR * dL; e += kTcThreads) {
      const int r = e / dL, k = e - r * dL;
      const float gw = aw[r * ldL + k] > 0.f ? grow[r] * wout[k] : 0.f;
      if (L > 1) {
        grd(L, a.want)[r * ldL + k] = gw;
        grd(L, 1 - a.want)[r * ldL + k] = 0.f;
      } else if (trow[r] >= 0) {
        a.dz1[(size_t)trow[r] * 2 * N1 + a.want * N1 + k] = gw;
        a.dz1[(size_t)trow[r] * 2 * N1 + (1 - a.want) * N1 + k] = 0.f;
      }
    }
    __syncthreads();
    // ---- tail backward -------------------------------------------------------------------------------------------------------
    for (int l = L - 1; l >= 1; --l) {
      const int din = a.dims[l], dout = a.dims[l + 1], ldi = lda(l), ldz = lda(l + 1), ldw = din + 4;
      const float* dZs = grd(l + 1, 0);
      const float* dZt = grd(l + 1, 1);
      const float* Xs = act(l, 0);
      const float* Xt = act(l, 1);
      if (l == 1) {
        dw_accum<TR>(aWs1, dZs, ldz, Xs, ldi, dout, din);
        dw_accum<TR>(aWt1, dZt, ldz, Xt, ldi, dout, din);
        dw_accum<TR>(aH1, dZs, ldz, Xt, ldi, dout, din, mask);
        dw_accum<TR>(aH1, dZt, ldz, Xs, ldi, dout, din, mask);
      } else if (l == 2) {
        dw_accum<TR>(aWs2, dZs, ldz, Xs, ldi, dout, din);
        dw_accum<TR>(aWt2, dZt, ldz, Xt, ldi, dout, din);
        dw_accum<TR>(aH2, dZs, ldz, Xt, ldi, dout, din, mask);
        dw_accum<TR>(aH2, dZt, ldz, Xs, ldi, dout, din, mask);
      } else {
        dw_accum<TR>(aWs3, dZs, ldz, Xs, ldi, dout, din);
        dw_accum<TR>(aWt3, dZt, ldz, Xt, ldi, dout, din);
        dw_accum<TR>(aH3, dZs, ldz, Xt, ldi, dout, din, mask);
        dw_accum<TR>(aH3, dZt, ldz, Xs, ldi, dout, din, mask);
      }
      if (tid < dout) {
        float s = 0.f;
        for (int r = 0; r < TR; ++r) s += dZs[r * ldz + tid];
        accB[l] += s;
      } else if (tid >= 128 && tid - 128 < dout) {
        float s = 0.f;
        for (int r = 0; r < TR; ++r) s += dZt[r * ldz + tid - 128];
        accB[l] += s;
      }
      const float* Wsl = smem + lay.tailW[l][0];
      const float* Wtl = smem + lay.tailW[l][1];
      const float* Hl = smem + lay.tailW[l][2];
      float* gs_out = l > 1 ? grd(l, 0) : nullptr;
      float* gt_out = l > 1 ? grd(l, 1) : nullptr;
      float* dz1 = a.dz1;
      // dx_s = dZs Ws + m * (dZt H);  dx_t = dZt Wt + m * (dZs H);  then times relu'(layer input) = the previous layer's dZ
      tile_gemm2_any<TR, true>(dZs, ldz, Wsl, ldw, dZt, ldz, Hl, ldw, din, dout, cross,
                               [&](int row, int col, float v0, float v1, float c0, float c1) {
                                 const float m = mask[row];
                                 const float2 y = *reinterpret_cast<const float2*>(Xs + row * ldi + col);
                                 const float2 o = make_float2(y.x > 0.f ? v0 + m * c0 : 0.f, y.y > 0.f ? v1 + m * c1 : 0.f);
                                 if (gs_out) *reinterpret_cast<float2*>(gs_out + row * ldi + col) = o;
                                 else if (trow[row] >= 0) *reinterpret_cast<float2*>(dz1 + (size_t)trow[row] * 2 * N1 + col) = o;
                               });
      tile_gemm2_any<TR, true>(dZt, ldz, Wtl, ldw, dZs, ldz, Hl, ldw, din, dout, cross,
                               [&](int row, int col, float v0, float v1, float c0, float c1) {
                                 const float m = mask[row];
                                 const float2 y = *reinterpret_cast<const float2*>(Xt + row * ldi + col);
                                 const float2 o = make_float2(y.x > 0.f ? v0 + m * c0 : 0.f, y.y > 0.f ? v1 + m * c1 : 0.f);
                                 if (gt_out) *reinterpret_cast<float2*>(gt_out + row * ldi + col) = o;
                                 else if (trow[row] >= 0) *reinterpret_cast<float2*>(dz1 + (size_t)trow[row] * 2 * N1 + N1 + col) = o;
                               });
      __syncthreads();
    }
  }

  if (a.backward) {
    // ---- flush the tail gradients ----------------------------------------------------------------------------------------------
    for (int l = 1; l < L; ++l) {
      const int din = a.dims[l], dout = a.dims[l + 1];
      if (l == 1) { dw_flush(aWs1, a.dWs[l], dout, din); dw_flush(aWt1, a.dWt[l], dout, din); dw_flush(aH1, a.dH[l], dout, din); }
      else if (l == 2) { dw_flush(aWs2, a.dWs[l], dout, din); dw_flush(aWt2, a.dWt[l], dout, din); dw_flush(aH2, a.dH[l], dout, din); }
      else { dw_flush(aWs3, a.dWs[l], dout, din); dw_flush(aWt3, a.dWt[l], dout, din); dw_flush(aH3, a.dH[l], dout, din); }
      if (tid < dout) { if (a.dbs[l]) atomicAdd(&a.dbs[l][tid], accB[l]); }
      else if (tid >= 128 && tid - 128 < dout) { if (a.dbt[l]) atomicAdd(&a.dbt[l][tid - 128], accB[l]); }
    }
    if (tid < dL) { if (a.dw_out) atomicAdd(&a.dw_out[tid], accOut); }
    else if (tid == kCnMaxHidden) { if (a.db_out) atomicAdd(a.db_out, accOut); }
    __syncthreads();  // the scratch rows written above by other threads of this CTA are read below

    // ============================= phase 2: layer-0 backward, one K chunk at a time ==========================================
    float* const dzs = smem + lay.dz[0];
    float* const dzt = smem + lay.dz[1];
    const int ld1 = N1 + 4;
    float accB0 = 0.f;  // tid < N1: dbs_0[tid];  128 <= tid < 128 + N1: dbt_0[tid - 128]
    for (int kc = 0; kc < nkc; ++kc) {
      float a0s[kCnDw0][4], a0t[kCnDw0][4], a0h[kCnDw0][4];
      tile_acc_zero(a0s); tile_acc_zero(a0t); tile_acc_zero(a0h);
      conet_load_wchunk(a, kc, wch);
      for (int ti = 0; ti < my_tiles; ++ti) {
        load_tile_rows(ti);
        __syncthreads();
        conet_gather_chunk(a, trow, kc, xs, xt, mask);
        for (int e = tid; e < 2 * TR * (N1 / 4); e += kTcThreads) {
          const int tower = e / (TR * (N1 / 4)), rem = e - tower * (TR * (N1 / 4));
          const int r = rem / (N1 / 4), c4 = rem - r * (N1 / 4);
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (trow[r] >= 0) v = *reinterpret_cast<const float4*>(a.dz1 + (size_t)trow[r] * 2 * N1 + tower * N1 + 4 * c4);
          *reinterpret_cast<float4*>((tower ? dzt : dzs) + r * ld1 + 4 * c4) = v;
        }
        __syncthreads();
        if (kc == 0) {
          if (tid < N1) {
            float s = 0.f;
            for (int r = 0; r < TR; ++r) s += dzs[r * ld1 + tid];
            accB0 += s;
          } else if (tid >= 128 && tid - 128 < N1) {
            float s = 0.f;
            for (int r = 0; r < TR; ++r) s += dzt[r * ld1 + tid - 128];
            accB0 += s;
          }
        }
        dw_accum<TR>(a0s, dzs, ld1, xs, ldc, N1, KC);
        dw_accum<TR>(a0t, dzt, ld1, xt, ldc, N1, KC);
        dw_accum<TR>(a0h, dzs, ld1, xt, ldc, N1, KC, mask);
        dw_accum<TR>(a0h, dzt, ld1, xs, ldc, N1, KC, mask);
        __syncthreads();  // x chunk fully consumed: it is overwritten by the dx chunk
        const bool cross = conet_mtile_overlaps(mask);
        {  // dx_s = dZs Ws + m (dZt H),  dx_t = dZt Wt + m (dZs H): each dZ operand is split once for its two products
          constexpr int NTX = kCnKC / 16;
          float d_ss[NTX][4], d_sh[NTX][4], d_tt[NTX][4], d_th[NTX][4];
          tile_acc_zero(d_ss); tile_acc_zero(d_sh); tile_acc_zero(d_tt); tile_acc_zero(d_th);
          tile_mma_acc2<TR, NTX, true>(d_ss, d_sh, dzs, ld1, wch[0], ldc, wch[2], ldc, KC, N1, cross);
          tile_mma_acc2<TR, NTX, true>(d_tt, d_th, dzt, ld1, wch[1], ldc, wch[2], ldc, KC, N1, cross);
          tile_acc_visit<TR, NTX>(KC, [&](int j, int row, int col, int h) {
            const float m = mask[row];
            *reinterpret_cast<float2*>(xs + row * ldc + col) =
                make_float2(d_ss[j][2 * h] + m * d_th[j][2 * h], d_ss[j][2 * h + 1] + m * d_th[j][2 * h + 1]);
            *reinterpret_cast<float2*>(xt + row * ldc + col) =
                make_float2(d_tt[j][2 * h] + m * d_sh[j][2 * h], d_tt[j][2 * h + 1] + m * d_sh[j][2 * h + 1]);
          });
        }
        __syncthreads();
        constexpr int C4 = KC / 4;
        for (int e = tid; e < 2 * TR * C4; e += kTcThreads) {
          const int tower = e / (TR * C4), rem = e - tower * (TR * C4);
          const int r = rem / C4, c4 = rem - r * C4;
          if (trow[r] < 0) continue;
          const CnCol c = conet_col(a, tower, trow[r], kc * KC + 4 * c4);
          if (!c.ok) continue;
          red_add4(c.dtab + c.id * a.dim, c.col4,
                   scale4(a.scale, *reinterpret_cast<const float4*>((tower ? xt : xs) + r * ldc + 4 * c4)));
        }
        __syncthreads();
      }
      dw_flush(a0s, a.dWs[0], N1, KC, K0, kc * KC);
      dw_flush(a0t, a.dWt[0], N1, KC, K0, kc * KC);
      dw_flush(a0h, a.dH[0], N1, KC, K0, kc * KC);
      __syncthreads();  // the next chunk's weights replace wch
    }
    if (tid < N1) { if (a.dbs[0]) atomicAdd(&a.dbs[0][tid], accB0); }
    else if (tid >= 128 && tid - 128 < N1) { if (a.dbt[0]) atomicAdd(&a.dbt[0][tid - 128], accB0); }
  }

  // ---- loss ------------------------------------------------------------------------------------------------------------------
  const double denom = (double)a.batch;
  float* out8 = a.out8;
  grid_reduce_last_block<1>(loss_acc, ws, red_smem, [=](double* tot) {
    out8[0] = (float)(tot[0] / denom);
    for (int i = 1; i < 8; ++i) out8[i] = 0.f;
  });
}

#endif  // __CUDACC__ || XDR_EMU

static bool conet_stack_ok(int n_layers, const int* dims, int dim) {
  if (n_layers < 1 || n_layers > kCnMaxLayers || dims == nullptr) return false;
  if (dims[0] != 2 * dim || dims[0] % kCnKC != 0 || dims[0] > 512 || dim % 4 != 0) return false;
  const int cap[kCnMaxLayers] = {0, kCnDw1, kCnDw2, kCnDw3};
  for (int l = 1; l <= n_layers; ++l)
    if (dims[l] < 8 || dims[l] > kCnMaxHidden || dims[l] % 8 != 0) return false;
  for (int l = 1; l < n_layers; ++l)
    if (dw_tiles_per_warp(dims[l + 1], dims[l]) > cap[l]) return false;
  return (size_t)conet_smem_layout(n_layers, dims).total * sizeof(float) <= 224 * 1024;
}

// Validates the arguments of xdr_tc_conet_step and fills the kernel's argument block (shared with the emulator harness).
static int conet_make_args(ConetArgs* out, int n_layers, const int* dims_host, const float* const* Ws_host,
                           const float* const* bs_host, const float* const* Wt_host, const float* const* bt_host,
                           const float* const* H_host, float* const* dWs_host, float* const* dbs_host,
                           float* const* dWt_host, float* const* dbt_host, float* const* dH_host, const float* w_out,
                           const float* b_out, float* dw_out, float* db_out, int want, const float* Su, const float* Si,
                           const float* Tu, const float* Ti, int64_t n_u, int64_t n_i, int dim, const int64_t* user,
                           const int64_t* item, const float* label, int64_t batch, int mask_on_item, int64_t n_overlap,
                           int backward, const float* grad_loss, float scale, float* dSu, float* dSi, float* dTu, float* dTi,
                           float* dz1_scratch, float* prob, float* out8, void* ws, int32_t* oob) {
  XDR_REQUIRE(dim_ok(dim) && batch > 0, "xdr_tc_conet_step: bad dim/batch");
  XDR_REQUIRE(conet_stack_ok(n_layers, dims_host, dim), "xdr_tc_conet_step: unsupported layer stack");
  XDR_REQUIRE(want == 0 || want == 1, "xdr_tc_conet_step: want must be 0 (source) or 1 (target)");
  XDR_REQUIRE(Ws_host && Wt_host && H_host && w_out && Su && Si && Tu && Ti && user && item && label && out8 && ws,
              "xdr_tc_conet_step: null pointer");
  if (backward) {
    XDR_REQUIRE(dWs_host && dWt_host && dH_host && dSu && dSi && dTu && dTi && dz1_scratch,
                "xdr_tc_conet_step: null gradient destination / scratch");
  }
  ConetArgs a{};
  a.n_layers = n_layers;
  for (int l = 0; l <= n_layers; ++l) a.dims[l] = dims_host[l];
  for (int l = 0; l < n_layers; ++l) {
    XDR_REQUIRE(Ws_host[l] && Wt_host[l] && H_host[l], "xdr_tc_conet_step: null weight");
    XDR_REQUIRE(aligned16(Ws_host[l]) && aligned16(Wt_host[l]) && aligned16(H_host[l]),
                "xdr_tc_conet_step: weights must be 16-byte aligned");
    a.Ws[l] = Ws_host[l]; a.Wt[l] = Wt_host[l]; a.H[l] = H_host[l];
    a.bs[l] = bs_host ? bs_host[l] : nullptr;
    a.bt[l] = bt_host ? bt_host[l] : nullptr;
    if (backward) {
      XDR_REQUIRE(dWs_host[l] && dWt_host[l] && dH_host[l], "xdr_tc_conet_step: null weight gradient");
      a.dWs[l] = dWs_host[l]; a.dWt[l] = dWt_host[l]; a.dH[l] = dH_host[l];
      a.dbs[l] = dbs_host ? dbs_host[l] : nullptr;
      a.dbt[l] = dbt_host ? dbt_host[l] : nullptr;
    }
  }
  a.w_out = w_out; a.b_out = b_out; a.dw_out = backward ? dw_out : nullptr; a.db_out = backward ? db_out : nullptr;
  a.want = want; a.Su = Su; a.Si = Si; a.Tu = Tu; a.Ti = Ti; a.dSu = dSu; a.dSi = dSi; a.dTu = dTu; a.dTi = dTi;
  a.n_u = n_u; a.n_i = n_i; a.dim = dim; a.user = user; a.item = item; a.label = label; a.batch = batch;
  a.mask_on_item = mask_on_item; a.n_overlap = n_overlap; a.dz1 = dz1_scratch; a.backward = backward;
  a.grad_loss = grad_loss; a.scale = scale; a.prob = prob; a.out8 = out8; a.oob = oob;
  *out = a;
  return XDR_OK;
}

}  // namespace xdr

using namespace xdr;

extern "C" {

int xdr_tc_conet_supported(int n_layers, const int* dims_host, int dim) { return conet_stack_ok(n_layers, dims_host, dim) ? 1 : 0; }

size_t xdr_tc_conet_scratch_bytes(int64_t batch, int hidden0) { return (size_t)batch * 2 * hidden0 * sizeof(float); }

int xdr_tc_conet_step(int n_layers, const int* dims_host, const float* const* Ws_host, const float* const* bs_host,
                      const float* const* Wt_host, const float* const* bt_host, const float* const* H_host,
                      float* const* dWs_host, float* const* dbs_host, float* const* dWt_host, float* const* dbt_host,
                      float* const* dH_host, const float* w_out, const float* b_out, float* dw_out, float* db_out, int want,
                      const float* Su, const float* Si, const float* Tu, const float* Ti, int64_t n_u, int64_t n_i, int dim,
                      const int64_t* user, const int64_t* item, const float* label, int64_t batch, int mask_on_item,
                      int64_t n_overlap, int backward, const float* grad_loss, float scale, float* dSu, float* dSi,
                      float* dTu, float* dTi, float* dz1_scratch, float* prob, float* out8, void* ws, int32_t* oob,
                      xdr_stream_t stream) {
  ConetArgs a{};
  const int rc = conet_make_args(&a, n_layers, dims_host, Ws_host, bs_host, Wt_host, bt_host, H_host, dWs_host, dbs_host,
                                 dWt_host, dbt_host, dH_host, w_out, b_out, dw_out, db_out, want, Su, Si, Tu, Ti, n_u, n_i, dim,
                                 user, item, label, batch, mask_on_item, n_overlap, backward, grad_loss, scale, dSu, dSi, dTu,
                                 dTi, dz1_scratch, prob, out8, ws, oob);
  if (rc != XDR_OK) return rc;
  const size_t smem = (size_t)conet_smem_layout(n_layers, a.dims).total * sizeof(float);
  XDR_CUDA_OK(cudaFuncSetAttribute(tc_conet_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t n_tiles = (batch + kCnTR - 1) / kCnTR;
  int grid = sm_count();
  if (grid > n_tiles) grid = (int)n_tiles;
  XDR_LAUNCH((tc_conet_kernel), grid, kTcThreads, smem, (cudaStream_t)stream, a, Workspace(ws));
  XDR_LAUNCH_OK();
  return XDR_OK;
}

}  // extern "C"
