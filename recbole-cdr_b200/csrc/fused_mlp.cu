// fused_mlp.cu -- A4 / A14-A15: gather -> small MLP -> loss head -> backward -> scatter-add in ONE kernel.
//
// Replaces, per batch:
//   EMCDR.calculate_map_loss  (reference emcdr.py:156-168): MSELoss(mapping(Es[idx]), Et[idx]) with
//       mapping = Linear(64,128) -> Tanh -> Linear(128,64)  (emcdr.py:86-93) or a single bias-free Linear (:58-59)
//   DTCDR.neumf_forward + BCE (dtcdr.py:112-125, 177-191): x = [max(Es_u[u],Et_u[u]) | max(Es_i[i],Et_i[i])],
//       recbole MLPLayers (Linear->ReLU per layer) -> Linear(.,1) -> sigmoid -> BCELoss
// The composed path (gather / dense / act / loss kernels, dense.cu) needs ~20 launches per tower pass and writes every
// activation to HBM; these MLPs are tiny (<= 16 K weights), so here a persistent CTA keeps W and W^T of every layer in
// shared memory, walks tiles of 32 batch rows (activations and their gradients never leave shared memory), and
// accumulates the weight gradients in REGISTERS across all its tiles (one atomic flush per CTA at the end).
// HBM traffic is then just ids + gathered rows + scattered rows:
//   map step   8 + 2*256 + 2*256 = 1032 B/row (dim 64)      DTCDR   16 + 4 + 4*256 + 4*256 = 2068 B/row
// fp32 FMA on CUDA cores: 98 kFLOP (map) / 28 kFLOP (DTCDR) per row is far below the FMA roofline at these byte rates.
// STATUS (round 1, profiles/r1_rows.md): parity-green but NOT yet faster than the graph-replayed composed path (map
// step 398 vs 98 us, DTCDR BOTH step 312 vs 227 us): one 256-thread CTA per SM running scalar shared-memory-fed FMA
// chains (2 LDS per FMA) is latency-bound on its own inner loops.  It needs 4x4 register tiles and 512+ threads; until
// then the models use it only on request (config key `xdr_fused_mlp: True`).
#include "xdr_common.cuh"
#include "mlp_args.cuh"

namespace xdr {

constexpr int kMlpThreads = 256;
constexpr int kMaxTileRows = 32;  // batch rows per tile: 32, or 16 / 8 when the weights leave less shared memory
// Activations and their gradients are kept TRANSPOSED in shared memory: actT[l][k * LD + r] with LD = tile_rows + 4, so the
// values of four consecutive batch rows at one feature are a single 16-byte LDS (the +4 pad keeps consecutive features on
// different banks).  Every inner loop then does 2 LDS per 4 FMAs with four independent accumulators.
template <int CAP>
__device__ __forceinline__ void mlp_accum_dw(float (&acc)[CAP], const float* dzT, const float* xT, int din, int dout, int tid,
                                             int tile_rows, int LD) {
#pragma unroll
  for (int i = 0; i < CAP; ++i) {
    const int e = tid + i * kMlpThreads;
    if (e < din * dout) {
      const int n = e / din, k = e - n * din;
      const float4* dz = reinterpret_cast<const float4*>(dzT + n * LD);
      const float4* x = reinterpret_cast<const float4*>(xT + k * LD);
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      for (int r4 = 0; r4 < tile_rows / 4; ++r4) {
        const float4 a = dz[r4], b = x[r4];
        s0 = fmaf(a.x, b.x, s0);
        s1 = fmaf(a.y, b.y, s1);
        s2 = fmaf(a.z, b.z, s2);
        s3 = fmaf(a.w, b.w, s3);
      }
      acc[i] += (s0 + s1) + (s2 + s3);
    }
  }
}

template <int CAP>
__device__ __forceinline__ void mlp_flush_dw(const float (&acc)[CAP], float* dW, int total, int tid) {
  if (dW == nullptr) return;
#pragma unroll
  for (int i = 0; i < CAP; ++i) {
    const int e = tid + i * kMlpThreads;
    if (e < total) atomicAdd(&dW[e], acc[i]);
  }
}

// E0/E1/E2: upper bounds on weight elements per thread of layers 0/1/2 (ceil(dout*din / 256))
template <int E0, int E1, int E2>
__global__ void __launch_bounds__(kMlpThreads, 1) fused_mlp_kernel(MlpArgs a, Workspace ws) {
  XDR_DYN_SMEM(float, smem);
  __shared__ float red_smem[8];
  const int tid = threadIdx.x;
  const int nl = a.n_layers;
  const int TR = a.tile_rows;  // multiple of 4
  const int LD = TR + 4;
  // ---- carve shared memory: per layer W [dout][din], Wt [din][dout], bias; per tile actT[l] / grdT[l] [dims[l]][LD]
  float* Wm[kMaxLayers];
  float* Wt[kMaxLayers];
  float* bs[kMaxLayers];
  float* act[kMaxLayers + 1];
  float* grd[kMaxLayers + 1];
  float* p = smem;
  for (int l = 0; l < nl; ++l) {
    const int sz = (a.dims[l] * a.dims[l + 1] + 3) & ~3;  // keep every region 16-byte aligned
    Wm[l] = p; p += sz;
    Wt[l] = p; p += sz;
    bs[l] = p; p += (a.dims[l + 1] + 3) & ~3;
  }
  for (int l = 0; l <= nl; ++l) {
    act[l] = p; p += a.dims[l] * LD;
    grd[l] = p; p += a.dims[l] * LD;
  }
  float* tgt = p;  // [TR][dims[nl]] target rows, row-major (head 0)
  for (int l = 0; l < nl; ++l) {
    const int din = a.dims[l], dout = a.dims[l + 1];
    for (int e = tid; e < din * dout; e += kMlpThreads) {
      const float w = a.W[l][e];
      const int n = e / din, k = e - n * din;
      Wm[l][e] = w;
      Wt[l][k * dout + n] = w;
    }
    for (int n = tid; n < dout; n += kMlpThreads) bs[l][n] = a.b[l] ? a.b[l][n] : 0.f;
  }
  __syncthreads();

  float accW0[E0], accW1[E1 > 0 ? E1 : 1], accW2[E2 > 0 ? E2 : 1];
  float accB[kMaxLayers] = {0.f, 0.f, 0.f};  // thread n < dout owns db[l][n] (dout <= 256)
#pragma unroll
  for (int i = 0; i < E0; ++i) accW0[i] = 0.f;
#pragma unroll
  for (int i = 0; i < (E1 > 0 ? E1 : 1); ++i) accW1[i] = 0.f;
#pragma unroll
  for (int i = 0; i < (E2 > 0 ? E2 : 1); ++i) accW2[i] = 0.f;

  const float g_up = (a.grad_loss ? __ldg(a.grad_loss) : 1.0f);
  const int d0 = a.dims[0], dl = a.dims[nl];
  const int nv = a.dim / 4;
  float loss_acc[1] = {0.f};
  const int64_t n_tiles = (a.batch + TR - 1) / TR;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t r0 = tile * TR;
    const int rows = (int)min((int64_t)TR, a.batch - r0);
    // ---- 1. gather the tile's input rows (one float4 = 4 consecutive features of one row) -> actT[0] ----------------
    {
      const int per_row = (a.in_mode == 0 ? 1 : 2) * nv;  // float4 per input row
      for (int e = tid; e < TR * per_row; e += kMlpThreads) {
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
        float* dst = act[0] + (4 * c) * LD + r;
        dst[0] = v.x; dst[LD] = v.y; dst[2 * LD] = v.z; dst[3 * LD] = v.w;
      }
      if (a.head == 0) {
        for (int e = tid; e < TR * nv; e += kMlpThreads) {
          const int r = e / nv, c = e - r * nv;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (r < rows) {
            const int64_t id = a.idx_u[r0 + r];
            if ((uint64_t)id < (uint64_t)a.n_u) v = ld_row4(a.T + id * a.dim, c);
          }
          reinterpret_cast<float4*>(tgt + r * dl)[c] = v;
        }
      }
    }
    __syncthreads();
    // ---- 2. forward: work item = (block of 4 rows, output feature n) ------------------------------------------------
    for (int l = 0; l < nl; ++l) {
      const int din = a.dims[l], dout = a.dims[l + 1];
      const int actk = (l == nl - 1) ? a.last_act : a.hidden_act;
      const float* xT = act[l];
      const float* wt = Wt[l];
      for (int e = tid; e < (TR / 4) * dout; e += kMlpThreads) {
        const int rb = e / dout, n = e - rb * dout;
        const float bias = bs[l][n];
        float4 s = make_float4(bias, bias, bias, bias);
        const float4* x4 = reinterpret_cast<const float4*>(xT + 4 * rb);
        const int ld4 = LD / 4;
#pragma unroll 4
        for (int k = 0; k < din; ++k) s = axpy4(wt[k * dout + n], x4[k * ld4], s);
        float y[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (actk == XDR_ACT_RELU) y[i] = y[i] > 0.f ? y[i] : 0.f;
          else if (actk == XDR_ACT_TANH) y[i] = tanhf(y[i]);
          else if (actk == XDR_ACT_SIGMOID) y[i] = sigmoidf_(y[i]);
        }
        *reinterpret_cast<float4*>(act[l + 1] + n * LD + 4 * rb) = make_float4(y[0], y[1], y[2], y[3]);
      }
      __syncthreads();
    }
    // ---- 3. loss head: loss partial and gradient of the last activation ---------------------------------------------
    if (a.head == 0) {
      const float gs = g_up * 2.0f / ((float)a.batch * (float)dl);
      for (int e = tid; e < TR * dl; e += kMlpThreads) {
        const int n = e / TR, r = e - n * TR;
        float g = 0.f;
        if (r < rows) {
          const float d = act[nl][n * LD + r] - tgt[r * dl + n];
          loss_acc[0] += d * d;
          g = gs * d;
        }
        grd[nl][n * LD + r] = g;
      }
    } else {
      const float gs = g_up / (float)a.batch;
      for (int r = tid; r < TR; r += kMlpThreads) {
        float g = 0.f;
        if (r < rows) {
          const float pz = sigmoidf_(act[nl][r]), y = a.label[r0 + r];
          loss_acc[0] += -(y * fmaxf(logf(pz), -100.f) + (1.f - y) * fmaxf(logf(1.f - pz), -100.f));
          if (a.prob) a.prob[r0 + r] = pz;
          const float pq = pz * (1.f - pz);
          g = gs * (pz - y) / fmaxf(pq, 1e-12f) * pq;
        }
        grd[nl][r] = g;
      }
    }
    __syncthreads();
    if (!a.backward) continue;
    // ---- 4. backward through the layers ------------------------------------------------------------------------------
    for (int l = nl - 1; l >= 0; --l) {
      const int din = a.dims[l], dout = a.dims[l + 1];
      const int actk = (l == nl - 1) ? a.last_act : a.hidden_act;
      if (actk != XDR_ACT_NONE) {  // dz = g * act'(y) in place
        for (int e = tid; e < TR * dout; e += kMlpThreads) {
          const int n = e / TR, r = e - n * TR;
          const float y = act[l + 1][n * LD + r], g = grd[l + 1][n * LD + r];
          float d = g;
          if (actk == XDR_ACT_RELU) d = y > 0.f ? g : 0.f;
          else if (actk == XDR_ACT_TANH) d = g * (1.f - y * y);
          else if (actk == XDR_ACT_SIGMOID) d = g * (1.f - y) * y;
          grd[l + 1][n * LD + r] = d;
        }
        __syncthreads();
      }
      const float* dzT = grd[l + 1];
      const float* xT = act[l];
      // weight gradient: thread owns elements e = tid + 256*i of W[l]  (n = e / din, k = e % din)
      if (l == 0) mlp_accum_dw(accW0, dzT, xT, din, dout, tid, TR, LD);
      else if (l == 1) mlp_accum_dw(accW1, dzT, xT, din, dout, tid, TR, LD);
      else mlp_accum_dw(accW2, dzT, xT, din, dout, tid, TR, LD);
      if (tid < dout) {
        float s = 0.f;
        for (int r = 0; r < TR; ++r) s += dzT[tid * LD + r];
        accB[l] += s;
      }
      // input gradient: work item = (block of 4 rows, input feature k):  gT[l][k][r] = sum_n dzT[n][r] * W[n][k]
      const float* wm = Wm[l];
      for (int e = tid; e < (TR / 4) * din; e += kMlpThreads) {
        const int rb = e / din, k = e - rb * din;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* dz4 = reinterpret_cast<const float4*>(dzT + 4 * rb);
        const int ld4 = LD / 4;
#pragma unroll 4
        for (int n = 0; n < dout; ++n) s = axpy4(wm[n * din + k], dz4[n * ld4], s);
        *reinterpret_cast<float4*>(grd[l] + k * LD + 4 * rb) = s;
      }
      __syncthreads();
    }
    // ---- 5. scatter-add the row gradients ---------------------------------------------------------------------------
    {
      const int per_row = (a.in_mode == 0 ? 1 : 2) * nv;
      for (int e = tid; e < rows * per_row; e += kMlpThreads) {
        const int r = e / per_row, c = e - r * per_row;
        const float* src = grd[0] + (4 * c) * LD + r;
        const float4 g = scale4(a.scale, make_float4(src[0], src[LD], src[2 * LD], src[3 * LD]));
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
        for (int e = tid; e < rows * nv; e += kMlpThreads) {
          const int r = e / nv, c = e - r * nv;
          const int64_t id = a.idx_u[r0 + r];
          if ((uint64_t)id >= (uint64_t)a.n_u) continue;
          const float* src = grd[nl] + (4 * c) * LD + r;
          red_add4(a.dT + id * a.dim, c, scale4(-a.scale, make_float4(src[0], src[LD], src[2 * LD], src[3 * LD])));
        }
      }
    }
    __syncthreads();
  }

  // ---- flush the weight gradients accumulated in registers -------------------------------------------------------------
  if (a.backward) {
    if (nl > 0) mlp_flush_dw(accW0, a.dW[0], a.dims[0] * a.dims[1], tid);
    if (nl > 1) mlp_flush_dw(accW1, a.dW[1], a.dims[1] * a.dims[2], tid);
    if (nl > 2) mlp_flush_dw(accW2, a.dW[2], a.dims[2] * a.dims[3], tid);
    for (int l = 0; l < nl; ++l)
      if (a.db[l] != nullptr && tid < a.dims[l + 1]) atomicAdd(&a.db[l][tid], accB[l]);
  }
  // ---- loss ------------------------------------------------------------------------------------------------------------
  const double denom = a.head == 0 ? (double)a.batch * (double)dl : (double)a.batch;
  float* out8 = a.out8;
  grid_reduce_last_block<1>(loss_acc, ws, red_smem, [=](double* tot) {
    out8[0] = (float)(tot[0] / denom);
    for (int i = 1; i < 8; ++i) out8[i] = 0.f;
  });
}

static size_t mlp_smem_bytes(const MlpArgs& a) {
  const int kTileRows = a.tile_rows, LD = a.tile_rows + 4;
  size_t f = 0;
  for (int l = 0; l < a.n_layers; ++l) f += 2 * (size_t)((a.dims[l] * a.dims[l + 1] + 3) & ~3) + ((a.dims[l + 1] + 3) & ~3);
  for (int l = 0; l <= a.n_layers; ++l) f += 2 * (size_t)a.dims[l] * LD;
  f += (size_t)kTileRows * a.dims[a.n_layers];
  return f * sizeof(float);
}

// largest tile (32, 16, 8 rows) whose activations fit next to the weights in 200 KB of shared memory
static bool pick_tile_rows(MlpArgs* a) {
  for (int tr = kMaxTileRows; tr >= 8; tr >>= 1) {
    a->tile_rows = tr;
    if (mlp_smem_bytes(*a) <= 200 * 1024) return true;
  }
  return false;
}

template <int E0, int E1, int E2>
static int launch_mlp(const MlpArgs& a, void* ws, cudaStream_t s) {
  const int kTileRows = a.tile_rows;
  const size_t smem = mlp_smem_bytes(a);
  auto kern = fused_mlp_kernel<E0, E1, E2>;
  XDR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t n_tiles = (a.batch + kTileRows - 1) / kTileRows;
  int grid = sm_count();
  if (grid > n_tiles) grid = (int)n_tiles;
  XDR_LAUNCH((kern), grid, kMlpThreads, smem, s, a, Workspace(ws));
  return XDR_OK;
}

}  // namespace xdr

using namespace xdr;

extern "C" {

// Returns 1 when xdr_fused_mlp_step takes this layer stack (else the caller composes the dense-layer kernels).
int xdr_fused_mlp_supported(int n_layers, const int* dims_host) {
  if (n_layers < 1 || n_layers > kMaxLayers || dims_host == nullptr) return 0;
  MlpArgs a{};
  a.n_layers = n_layers;
  for (int l = 0; l <= n_layers; ++l) {
    if (dims_host[l] < 1 || dims_host[l] > 256) return 0;
    a.dims[l] = dims_host[l];
  }
  if (dims_host[0] % 4 != 0) return 0;
  for (int l = 0; l < n_layers; ++l) {
    const int per_thread = (dims_host[l] * dims_host[l + 1] + kMlpThreads - 1) / kMlpThreads;
    if (per_thread > 32) return 0;
  }
  return pick_tile_rows(&a) ? 1 : 0;
}

int xdr_fused_mlp_step(int n_layers, const int* dims_host, const float* const* W_host, const float* const* b_host,
                       float* const* dW_host, float* const* db_host, int hidden_act, int in_mode, int head,
                       const float* Au, const float* Bu, const float* Ai, const float* Bi, const float* T, int64_t n_u,
                       int64_t n_i, int dim, const int64_t* idx_u, const int64_t* idx_i, const float* label,
                       int64_t batch, int backward, const float* grad_loss, float scale, float* dAu, float* dBu,
                       float* dAi, float* dBi, float* dT, float* prob, float* out8, void* ws, int32_t* oob,
                       xdr_stream_t stream) {
  XDR_REQUIRE(xdr_fused_mlp_supported(n_layers, dims_host), "xdr_fused_mlp_step: unsupported layer stack");
  XDR_REQUIRE(dim_ok(dim) && batch > 0, "xdr_fused_mlp_step: bad dim/batch");
  XDR_REQUIRE(in_mode == 0 || in_mode == 1, "xdr_fused_mlp_step: bad in_mode");
  XDR_REQUIRE(head == 0 || head == 1, "xdr_fused_mlp_step: bad head");
  XDR_REQUIRE(W_host && idx_u && out8 && ws && Au, "xdr_fused_mlp_step: null pointer");
  XDR_REQUIRE(dims_host[0] == (in_mode == 0 ? dim : 2 * dim), "xdr_fused_mlp_step: dims[0] does not match the input mode");
  XDR_REQUIRE(head == 1 ? dims_host[n_layers] == 1 : dims_host[n_layers] == dim, "xdr_fused_mlp_step: bad output width");
  XDR_REQUIRE(in_mode == 0 || (Bu && Ai && Bi && idx_i), "xdr_fused_mlp_step: max-combine input needs four tables");
  XDR_REQUIRE(head == 0 ? T != nullptr : label != nullptr, "xdr_fused_mlp_step: missing target table / labels");
  if (backward) {
    XDR_REQUIRE(dAu && (in_mode == 0 || (dBu && dAi && dBi)) && (head == 1 || dT), "xdr_fused_mlp_step: null destination");
  }
  MlpArgs a{};
  a.n_layers = n_layers;
  for (int l = 0; l <= n_layers; ++l) a.dims[l] = dims_host[l];
  for (int l = 0; l < n_layers; ++l) {
    XDR_REQUIRE(W_host[l], "xdr_fused_mlp_step: null weight");
    a.W[l] = W_host[l];
    a.b[l] = b_host ? b_host[l] : nullptr;
    a.dW[l] = (backward && dW_host) ? dW_host[l] : nullptr;
    a.db[l] = (backward && db_host) ? db_host[l] : nullptr;
  }
  a.hidden_act = hidden_act; a.last_act = XDR_ACT_NONE; a.in_mode = in_mode; a.head = head;
  a.Au = Au; a.Bu = Bu; a.Ai = Ai; a.Bi = Bi; a.T = T; a.n_u = n_u; a.n_i = n_i; a.dim = dim;
  a.idx_u = idx_u; a.idx_i = idx_i; a.label = label; a.batch = batch; a.backward = backward; a.grad_loss = grad_loss;
  pick_tile_rows(&a);
  a.scale = scale; a.dAu = dAu; a.dBu = dBu; a.dAi = dAi; a.dBi = dBi; a.dT = dT; a.prob = prob; a.out8 = out8; a.oob = oob;
  int e[3] = {0, 0, 0};
  for (int l = 0; l < n_layers; ++l) e[l] = (a.dims[l] * a.dims[l + 1] + kMlpThreads - 1) / kMlpThreads;
  cudaStream_t s = (cudaStream_t)stream;
  int rc;
  if (e[0] <= 16 && e[1] <= 2 && e[2] <= 1) rc = launch_mlp<16, 2, 1>(a, ws, s);        // DTCDR [128,32,16,1]
  else if (e[0] <= 32 && e[1] <= 32 && e[2] == 0) rc = launch_mlp<32, 32, 0>(a, ws, s);  // EMCDR map [64,128,64]
  else rc = launch_mlp<32, 32, 32>(a, ws, s);
  if (rc != XDR_OK) return rc;
  XDR_LAUNCH_OK();
  return XDR_OK;
}

}  // extern "C"
