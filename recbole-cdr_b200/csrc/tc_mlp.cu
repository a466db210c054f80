// tc_mlp.cu -- A4 / A14-A15 on the tensor cores: gather -> small MLP -> loss head -> backward -> scatter-add, ONE kernel.
//
// Same contract as fused_mlp.cu (xdr_fused_mlp_step; EMCDR.calculate_map_loss emcdr.py:156-168 with the mapping of
// emcdr.py:58-64,86-93, and one DTCDR NeuMF term dtcdr.py:112-125,186-187), different engine: every layer product of a
// 64- (or 32-) row tile -- forward X W^T, input gradient dZ W and weight gradient dZ^T X -- is a 3xTF32 mma.sync tile
// product (tc_tile.cuh) instead of scalar shared-memory-fed FMA chains.  A persistent CTA keeps the layer weights in shared
// memory (row-major, ld = din + 4), activations and their gradients of the tile never leave shared memory, the weight
// gradients accumulate as MMA C fragments in registers over all of the CTA's tiles (one atomic flush at the end).
// HBM traffic = ids + gathered rows + scattered rows (1032 B/row for the map step, 2068 B/row for a DTCDR term).
// Layers whose output is narrower than 8 (the NeuMF output unit, 16 -> 1) run on the CUDA cores.
//
// STATUS: written in a session without GPU access -- compiles for sm_100a, NOT yet executed on hardware.  The models use
// it only on request (config `xdr_fused_mlp: 'tc'`); its parity tests are marked `unvalidated` (tests/conftest.py).
#include "mlp_args.cuh"
#include "tc_tile.cuh"

namespace xdr {

// DW0/DW1/DW2: weight-gradient tiles per warp of layers 0/1/2 (upper bounds; see tc_mlp_supported)
template <int TR, int DW0, int DW1, int DW2>
__global__ void __launch_bounds__(kTcThreads, 1) tc_mlp_kernel(MlpArgs a, Workspace ws) {
  XDR_DYN_SMEM(float, smem);
  __shared__ float red_smem[8];
  const int tid = threadIdx.x;
  const int nl = a.n_layers;
  // ---- carve shared memory ------------------------------------------------------------------------------------------
  float* Wsm[kMaxLayers];
  float* bsm[kMaxLayers];
  float* act[kMaxLayers + 1];
  float* grd[kMaxLayers + 1];
  int ldw[kMaxLayers], lda[kMaxLayers + 1];
  float* p = smem;
  for (int l = 0; l < nl; ++l) {
    ldw[l] = a.dims[l] + 4;
    Wsm[l] = p; p += a.dims[l + 1] * ldw[l];
    bsm[l] = p; p += (a.dims[l + 1] + 3) & ~3;
  }
  for (int l = 0; l <= nl; ++l) {
    lda[l] = ((a.dims[l] + 3) & ~3) + 4;
    act[l] = p; p += TR * lda[l];
    grd[l] = p; p += TR * lda[l];
  }
  float* tgt = grd[0];  // head 0: target rows [TR][lda[0]] (dims[0] == dims[nl]); dead before grd[0] is produced
  // resident weights: 128-bit loads (din % 8 == 0, pointers 16-byte aligned by the entry point).  No barrier here: the one
  // after the first tile's gather also orders these writes before their first use (every CTA owns at least one tile).
  for (int l = 0; l < nl; ++l) {
    const int din = a.dims[l], dout = a.dims[l + 1], q = din >> 2;
    for (int e = tid; e < q * dout; e += kTcThreads) {
      const int n = e / q, k4 = e - n * q;
      *reinterpret_cast<float4*>(Wsm[l] + n * ldw[l] + 4 * k4) = *(reinterpret_cast<const float4*>(a.W[l]) + e);
    }
    for (int n = tid; n < dout; n += kTcThreads) bsm[l][n] = a.b[l] ? a.b[l][n] : 0.f;
  }

  float accW0[DW0][4], accW1[DW1][4], accW2[DW2][4];
#pragma unroll
  for (int j = 0; j < DW0; ++j) accW0[j][0] = accW0[j][1] = accW0[j][2] = accW0[j][3] = 0.f;
#pragma unroll
  for (int j = 0; j < DW1; ++j) accW1[j][0] = accW1[j][1] = accW1[j][2] = accW1[j][3] = 0.f;
#pragma unroll
  for (int j = 0; j < DW2; ++j) accW2[j][0] = accW2[j][1] = accW2[j][2] = accW2[j][3] = 0.f;
  float accThin[kMaxLayers] = {0.f, 0.f, 0.f};  // thin layers (dout < 8, dout*din <= 256): thread e owns dW element e
  float accB[kMaxLayers] = {0.f, 0.f, 0.f};     // thread n < dout owns db[l][n]

  const float g_up = (a.grad_loss ? __ldg(a.grad_loss) : 1.0f);
  const int dl = a.dims[nl];
  const int nv = a.dim / 4;
  float loss_acc[1] = {0.f};
  const int64_t n_tiles = (a.batch + TR - 1) / TR;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t r0 = tile * TR;
    const int rows = (int)min((int64_t)TR, a.batch - r0);
    // ---- 1. gather the tile's input rows, row-major (rows past the batch are zero) ------------------------------------
    {
      const int per_row = (a.in_mode == 0 ? 1 : 2) * nv;  // float4 per input row
      for (int e = tid; e < TR * per_row; e += kTcThreads) {
        const int r = e / per_row, c = e - r * per_row;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < rows) {
          if (a.in_mode == 0) {
            const int64_t id = a.idx_u[r0 + r];
            if ((uint64_t)id < (uint64_t)a.n_u) v = ld_row4(a.Au + id * a.dim, c);
            else if (a.oob) *a.oob = 1;
          } else {
            const bool item = c >= nv;
            const int cc = item ? c - nv : c;
            const int64_t id = item ? a.idx_i[r0 + r] : a.idx_u[r0 + r];
            const int64_t n_rows = item ? a.n_i : a.n_u;
            if ((uint64_t)id < (uint64_t)n_rows) {
              const float4 x = ld_row4((item ? a.Ai : a.Au) + id * a.dim, cc);
              const float4 y = ld_row4((item ? a.Bi : a.Bu) + id * a.dim, cc);
              auto mx = [](float s, float t) { return (s != s || t != t) ? (s + t) : (s > t ? s : t); };
              v = make_float4(mx(x.x, y.x), mx(x.y, y.y), mx(x.z, y.z), mx(x.w, y.w));
            } else if (a.oob) {
              *a.oob = 1;
            }
          }
        }
        *reinterpret_cast<float4*>(act[0] + r * lda[0] + 4 * c) = v;
      }
      if (a.head == 0) {
        for (int e = tid; e < TR * nv; e += kTcThreads) {
          const int r = e / nv, c = e - r * nv;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (r < rows) {
            const int64_t id = a.idx_u[r0 + r];
            if ((uint64_t)id < (uint64_t)a.n_u) v = ld_row4(a.T + id * a.dim, c);
          }
          *reinterpret_cast<float4*>(tgt + r * lda[0] + 4 * c) = v;
        }
      }
    }
    __syncthreads();
    // ---- 2. forward -----------------------------------------------------------------------------------------------------
    for (int l = 0; l < nl; ++l) {
      const int din = a.dims[l], dout = a.dims[l + 1];
      const int actk = (l == nl - 1) ? a.last_act : a.hidden_act;
      float* out = act[l + 1];
      const int ldo = lda[l + 1];
      const float* bias = bsm[l];
      if ((dout & 7) == 0) {
        tile_gemm_any<TR, false>(act[l], lda[l], Wsm[l], ldw[l], dout, din, [=](int row, int col, float v0, float v1) {
          *reinterpret_cast<float2*>(out + row * ldo + col) =
              make_float2(act_apply(v0 + bias[col], actk), act_apply(v1 + bias[col + 1], actk));
        });
      } else {  // thin layer on the CUDA cores: one (row, output) pair per thread
        for (int e = tid; e < TR * dout; e += kTcThreads) {
          const int r = e / dout, n = e - r * dout;
          const float* x = act[l] + r * lda[l];
          const float* w = Wsm[l] + n * ldw[l];
          float s = bias[n];
          for (int k = 0; k < din; ++k) s = fmaf(w[k], x[k], s);
          out[r * ldo + n] = act_apply(s, actk);
        }
      }
      __syncthreads();
    }
    // ---- 3. loss head: loss partial and the gradient of the last pre-activation (last_act is none) -----------------------
    if (a.head == 0) {
      const float gs = g_up * 2.0f / ((float)a.batch * (float)dl);
      for (int e = tid; e < TR * dl; e += kTcThreads) {
        const int r = e / dl, n = e - r * dl;
        float g = 0.f;
        if (r < rows) {
          const float d = act[nl][r * lda[nl] + n] - tgt[r * lda[0] + n];
          loss_acc[0] += d * d;
          g = gs * d;
        }
        grd[nl][r * lda[nl] + n] = g;
      }
    } else {
      const float gs = g_up / (float)a.batch;
      for (int r = tid; r < TR; r += kTcThreads) {
        float g = 0.f;
        if (r < rows) {
          const float pz = sigmoidf_(act[nl][r * lda[nl]]), y = a.label[r0 + r];
          loss_acc[0] += -(y * fmaxf(logf(pz), -100.f) + (1.f - y) * fmaxf(logf(1.f - pz), -100.f));
          if (a.prob) a.prob[r0 + r] = pz;
          const float pq = pz * (1.f - pz);
          g = gs * (pz - y) / fmaxf(pq, 1e-12f) * pq;
        }
        grd[nl][r * lda[nl]] = g;
      }
    }
    __syncthreads();
    if (!a.backward) continue;
    // ---- 4. backward: grd[l+1] holds dZ of layer l (gradient of its pre-activation) ---------------------------------------
    for (int l = nl - 1; l >= 0; --l) {
      const int din = a.dims[l], dout = a.dims[l + 1];
      const float* dZ = grd[l + 1];
      const int ldz = lda[l + 1];
      const float* X = act[l];
      const int ldx = lda[l];
      const bool thin = (dout & 7) != 0;
      // weight and bias gradients
      if (!thin) {
        if (l == 0) dw_accum<TR>(accW0, dZ, ldz, X, ldx, dout, din);
        else if (l == 1) dw_accum<TR>(accW1, dZ, ldz, X, ldx, dout, din);
        else dw_accum<TR>(accW2, dZ, ldz, X, ldx, dout, din);
      } else if (tid < dout * din) {
        const int n = tid / din, k = tid - n * din;
        float s = 0.f;
        for (int r = 0; r < TR; ++r) s = fmaf(dZ[r * ldz + n], X[r * ldx + k], s);
        accThin[l] += s;
      }
      if (tid < dout) {
        float s = 0.f;
        for (int r = 0; r < TR; ++r) s += dZ[r * ldz + tid];
        accB[l] += s;
      }
      // input gradient, times the derivative of the previous layer's activation (act[l] is that layer's output)
      float* gout = grd[l];
      const int ldg = lda[l];
      const int prev_act = l > 0 ? a.hidden_act : XDR_ACT_NONE;
      if (!thin) {
        tile_gemm_any<TR, true>(dZ, ldz, Wsm[l], ldw[l], din, dout, [=](int row, int col, float v0, float v1) {
          const float2 y = *reinterpret_cast<const float2*>(X + row * ldx + col);
          *reinterpret_cast<float2*>(gout + row * ldg + col) =
              make_float2(v0 * act_grad(y.x, prev_act), v1 * act_grad(y.y, prev_act));
        });
      } else {
        for (int e = tid; e < TR * din; e += kTcThreads) {
          const int r = e / din, k = e - r * din;
          float s = 0.f;
          for (int n = 0; n < dout; ++n) s = fmaf(dZ[r * ldz + n], Wsm[l][n * ldw[l] + k], s);
          gout[r * ldg + k] = s * act_grad(X[r * ldx + k], prev_act);
        }
      }
      __syncthreads();
    }
    // ---- 5. scatter-add the row gradients -------------------------------------------------------------------------------
    {
      const int per_row = (a.in_mode == 0 ? 1 : 2) * nv;
      for (int e = tid; e < rows * per_row; e += kTcThreads) {
        const int r = e / per_row, c = e - r * per_row;
        const float4 g = scale4(a.scale, *reinterpret_cast<const float4*>(grd[0] + r * lda[0] + 4 * c));
        if (a.in_mode == 0) {
          const int64_t id = a.idx_u[r0 + r];
          if ((uint64_t)id < (uint64_t)a.n_u) red_add4(a.dAu + id * a.dim, c, g);
        } else {
          const bool item = c >= nv;
          const int cc = item ? c - nv : c;
          const int64_t id = item ? a.idx_i[r0 + r] : a.idx_u[r0 + r];
          if ((uint64_t)id >= (uint64_t)(item ? a.n_i : a.n_u)) continue;
          const float4 x = ld_row4((item ? a.Ai : a.Au) + id * a.dim, cc);
          const float4 y = ld_row4((item ? a.Bi : a.Bu) + id * a.dim, cc);
          // torch.maximum backward: gradient to the larger operand, split 0.5/0.5 on ties
          auto wa = [](float s, float t) { return s > t ? 1.f : (s == t ? 0.5f : 0.f); };
          red_add4((item ? a.dAi : a.dAu) + id * a.dim, cc,
                   make_float4(g.x * wa(x.x, y.x), g.y * wa(x.y, y.y), g.z * wa(x.z, y.z), g.w * wa(x.w, y.w)));
          red_add4((item ? a.dBi : a.dBu) + id * a.dim, cc,
                   make_float4(g.x * wa(y.x, x.x), g.y * wa(y.y, x.y), g.z * wa(y.z, x.z), g.w * wa(y.w, x.w)));
        }
      }
      if (a.head == 0 && a.last_act == XDR_ACT_NONE) {  // the target embedding is NOT detached (emcdr.py:156-168): dT = -dY
        for (int e = tid; e < rows * nv; e += kTcThreads) {
          const int r = e / nv, c = e - r * nv;
          const int64_t id = a.idx_u[r0 + r];
          if ((uint64_t)id >= (uint64_t)a.n_u) continue;
          red_add4(a.dT + id * a.dim, c, scale4(-a.scale, *reinterpret_cast<const float4*>(grd[nl] + r * lda[nl] + 4 * c)));
        }
      }
    }
    __syncthreads();
  }

  // ---- flush the weight gradients accumulated in registers ---------------------------------------------------------------
  if (a.backward) {
    for (int l = 0; l < nl; ++l) {
      const int din = a.dims[l], dout = a.dims[l + 1];
      if ((dout & 7) == 0) {
        if (l == 0) dw_flush(accW0, a.dW[0], dout, din);
        else if (l == 1) dw_flush(accW1, a.dW[1], dout, din);
        else dw_flush(accW2, a.dW[2], dout, din);
      } else if (a.dW[l] != nullptr && tid < dout * din) {
        atomicAdd(&a.dW[l][tid], accThin[l]);
      }
      if (a.db[l] != nullptr && tid < dout) atomicAdd(&a.db[l][tid], accB[l]);
    }
  }
  // ---- loss --------------------------------------------------------------------------------------------------------------
  const double denom = a.head == 0 ? (double)a.batch * (double)dl : (double)a.batch;
  float* out8 = a.out8;
  grid_reduce_last_block<1>(loss_acc, ws, red_smem, [=](double* tot) {
    out8[0] = (float)(tot[0] / denom);
    for (int i = 1; i < 8; ++i) out8[i] = 0.f;
  });
}

static size_t tc_smem_bytes(const MlpArgs& a, int tr) {
  size_t f = 0;
  for (int l = 0; l < a.n_layers; ++l) f += (size_t)a.dims[l + 1] * (a.dims[l] + 4) + ((a.dims[l + 1] + 3) & ~3);
  for (int l = 0; l <= a.n_layers; ++l) f += 2 * (size_t)tr * (((a.dims[l] + 3) & ~3) + 4);
  return f * sizeof(float);
}

constexpr int kTcDw0 = 8, kTcDw1 = 8, kTcDw2 = 1;   // weight-gradient tiles per warp the one instantiation carries
constexpr size_t kTcSmemMax = 224 * 1024;

// rows per tile: 64 when it fits and the batch fills the SMs with 64-row tiles, else 32; 0 = stack not supported
static int tc_pick_tile_rows(const MlpArgs& a, int64_t batch) {
  const bool fits64 = tc_smem_bytes(a, 64) <= kTcSmemMax, fits32 = tc_smem_bytes(a, 32) <= kTcSmemMax;
  if (fits64 && (batch <= 0 || batch >= (int64_t)64 * sm_count() || !fits32)) return 64;
  return fits32 ? 32 : 0;
}

static bool tc_stack_ok(int n_layers, const int* dims, MlpArgs* a) {
  if (n_layers < 1 || n_layers > kMaxLayers || dims == nullptr) return false;
  a->n_layers = n_layers;
  const int cap[kMaxLayers] = {kTcDw0, kTcDw1, kTcDw2};
  for (int l = 0; l <= n_layers; ++l) {
    if (dims[l] < 1 || dims[l] > 256) return false;
    a->dims[l] = dims[l];
  }
  for (int l = 0; l < n_layers; ++l) {
    const int din = dims[l], dout = dims[l + 1];
    if (din % 8 != 0) return false;                       // K steps of 8, float4 row loads
    if (dout % 8 == 0) {
      if (dw_tiles_per_warp(dout, din) > cap[l]) return false;
    } else if (dout >= 8 || dout * din > kTcThreads || l != n_layers - 1) {
      return false;                                       // thin layers: output unit only, one dW element per thread
    }
  }
  return true;
}

template <int TR>
static int launch_tc(const MlpArgs& a, void* ws, cudaStream_t s) {
  const size_t smem = tc_smem_bytes(a, TR);
  auto kern = tc_mlp_kernel<TR, kTcDw0, kTcDw1, kTcDw2>;
  XDR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t n_tiles = (a.batch + TR - 1) / TR;
  int grid = sm_count();
  if (grid > n_tiles) grid = (int)n_tiles;
  XDR_LAUNCH((kern), grid, kTcThreads, smem, s, a, Workspace(ws));
  return XDR_OK;
}

}  // namespace xdr

using namespace xdr;

extern "C" {

int xdr_tc_mlp_supported(int n_layers, const int* dims_host) {
  MlpArgs a{};
  if (!tc_stack_ok(n_layers, dims_host, &a)) return 0;
  return tc_pick_tile_rows(a, 0) != 0 ? 1 : 0;
}

int xdr_tc_mlp_step(int n_layers, const int* dims_host, const float* const* W_host, const float* const* b_host,
                    float* const* dW_host, float* const* db_host, int hidden_act, int in_mode, int head, const float* Au,
                    const float* Bu, const float* Ai, const float* Bi, const float* T, int64_t n_u, int64_t n_i, int dim,
                    const int64_t* idx_u, const int64_t* idx_i, const float* label, int64_t batch, int backward,
                    const float* grad_loss, float scale, float* dAu, float* dBu, float* dAi, float* dBi, float* dT,
                    float* prob, float* out8, void* ws, int32_t* oob, xdr_stream_t stream) {
  MlpArgs a{};
  XDR_REQUIRE(tc_stack_ok(n_layers, dims_host, &a), "xdr_tc_mlp_step: unsupported layer stack");
  XDR_REQUIRE(dim_ok(dim) && batch > 0, "xdr_tc_mlp_step: bad dim/batch");
  XDR_REQUIRE(in_mode == 0 || in_mode == 1, "xdr_tc_mlp_step: bad in_mode");
  XDR_REQUIRE(head == 0 || head == 1, "xdr_tc_mlp_step: bad head");
  XDR_REQUIRE(W_host && idx_u && out8 && ws && Au, "xdr_tc_mlp_step: null pointer");
  XDR_REQUIRE(dims_host[0] == (in_mode == 0 ? dim : 2 * dim), "xdr_tc_mlp_step: dims[0] does not match the input mode");
  XDR_REQUIRE(head == 1 ? dims_host[n_layers] == 1 : dims_host[n_layers] == dim, "xdr_tc_mlp_step: bad output width");
  XDR_REQUIRE(in_mode == 0 || (Bu && Ai && Bi && idx_i), "xdr_tc_mlp_step: max-combine input needs four tables");
  XDR_REQUIRE(head == 0 ? T != nullptr : label != nullptr, "xdr_tc_mlp_step: missing target table / labels");
  XDR_REQUIRE(head == 1 || in_mode == 0, "xdr_tc_mlp_step: the MSE head pairs with the single-table input");
  if (backward) {
    XDR_REQUIRE(dAu && (in_mode == 0 || (dBu && dAi && dBi)) && (head == 1 || dT), "xdr_tc_mlp_step: null destination");
  }
  for (int l = 0; l < n_layers; ++l) {
    XDR_REQUIRE(W_host[l] && aligned16(W_host[l]), "xdr_tc_mlp_step: null or misaligned weight");
    a.W[l] = W_host[l];
    a.b[l] = b_host ? b_host[l] : nullptr;
    a.dW[l] = (backward && dW_host) ? dW_host[l] : nullptr;
    a.db[l] = (backward && db_host) ? db_host[l] : nullptr;
  }
  a.hidden_act = hidden_act; a.last_act = XDR_ACT_NONE; a.in_mode = in_mode; a.head = head;
  a.Au = Au; a.Bu = Bu; a.Ai = Ai; a.Bi = Bi; a.T = T; a.n_u = n_u; a.n_i = n_i; a.dim = dim;
  a.idx_u = idx_u; a.idx_i = idx_i; a.label = label; a.batch = batch; a.backward = backward; a.grad_loss = grad_loss;
  a.scale = scale; a.dAu = dAu; a.dBu = dBu; a.dAi = dAi; a.dBi = dBi; a.dT = dT; a.prob = prob; a.out8 = out8; a.oob = oob;
  const int tr = tc_pick_tile_rows(a, batch);
  XDR_REQUIRE(tr != 0, "xdr_tc_mlp_step: layer stack does not fit shared memory");
  a.tile_rows = tr;
  cudaStream_t s = (cudaStream_t)stream;
  const int rc = tr == 64 ? launch_tc<64>(a, ws, s) : launch_tc<32>(a, ws, s);
  if (rc != XDR_OK) return rc;
  XDR_LAUNCH_OK();
  return XDR_OK;
}

}  // extern "C"
