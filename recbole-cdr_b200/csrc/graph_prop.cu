// graph_prop.cu -- A10-A12: BiTGCF's graph propagate-and-transfer, forward and backward.
//
// Replaces, per training batch and per layer (reference model/cross_domain_recommender/bitgcf.py):
//   graph_layer    :130-135   S = L.E (torch.sparse.mm), E' = E + S + E*S
//   transfer_layer :137-172   rows < n_overlap (users and items separately) become
//                             0.5*[(lam*E_own + (1-lam)*E_other) + (d_s*E_s + d_t*E_t)/(d_s + d_t + 1e-7)]
//   forward        :185-189   F.normalize(p=2, dim=1) of both domains
// The reference materialises ~14 full-size [N, D] temporaries per transfer (split / slice / cat); here one kernel
// reads the two propagated tables once and writes the transferred tables (next layer's input) and the normalised
// rows straight into their slot of the layer-combine buffer.
//
// SpMM: L is CSR over N = n_users + n_items nodes (values D^-1/2 A D^-1/2, built once on the host).  Item popularity
// is heavy-tailed, so rows are cut into work items of <= chunk nonzeros (built once per graph): a work item is owned
// by one 8-lane group (one full 128-byte line per neighbour-row request); rows with a single work item are stored
// directly, split rows are accumulated with RED.v4 into pre-zeroed output rows.
// HBM-bound: per layer-domain  nnz*(8 + 4 + 4D)  (neighbour gathers, no reuse assumed)  +  2*N*4D  (SURVEY 8 D3).
#include "xdr_common.cuh"

namespace xdr {

constexpr int kGThreads = 256;

// ---- S = L . X over work items ------------------------------------------------------------------------------------
// SHARDED: the operand X is block-cyclically row-sharded over G peer-mapped shards (column c lives on shard c mod G at
// local row c div G); the neighbour-row gathers then cross NVLink directly and the output rows are this
// rank's.  Peer rows are read with plain coherent loads: the exchange buffers are rewritten by their owners between
// launches.
template <bool SHARDED>
__device__ __forceinline__ float4 spmm_ld(const float* X, const Shards& xs, int log2g, int64_t c, int64_t row_f, int cc) {
  if (SHARDED) return ld_row4(shard_row(xs, log2g, c, row_f), cc);
  return ldg_row4(X + c * row_f, cc);
}

template <int VEC, bool SHARDED>
__global__ void __launch_bounds__(kGThreads)
    spmm_work_kernel(const int64_t* __restrict__ work_row, const int64_t* __restrict__ work_beg,
                     const int64_t* __restrict__ work_end, const uint8_t* __restrict__ work_split, int64_t n_work,
                     const int64_t* __restrict__ col, const float* __restrict__ val, const float* __restrict__ X, Shards xs,
                     int log2g, int nv, float* __restrict__ S) {
  const int sub = threadIdx.x & (kLanesPerRow - 1);
  const int64_t group = ((int64_t)blockIdx.x * kGThreads + threadIdx.x) / kLanesPerRow;
  const int64_t n_groups = (int64_t)gridDim.x * kGThreads / kLanesPerRow;
  const int64_t row_f = (int64_t)nv * 4;
  for (int64_t w = group; w < n_work; w += n_groups) {
    const int64_t r = work_row[w], beg = work_beg[w], end = work_end[w];
    float4 acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    int64_t e = beg;
    // two neighbours per iteration: 2*VEC independent 16-byte gathers in flight per lane
    for (; e + 1 < end; e += 2) {
      const int64_t c0 = col[e], c1 = col[e + 1];
      const float a0 = val[e], a1 = val[e + 1];
      float4 x0[VEC], x1[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const int cc = sub + v * kLanesPerRow;
        x0[v] = cc < nv ? spmm_ld<SHARDED>(X, xs, log2g, c0, row_f, cc) : make_float4(0.f, 0.f, 0.f, 0.f);
        x1[v] = cc < nv ? spmm_ld<SHARDED>(X, xs, log2g, c1, row_f, cc) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[v] = axpy4(a1, x1[v], axpy4(a0, x0[v], acc[v]));
    }
    if (e < end) {
      const int64_t c0 = col[e];
      const float a0 = val[e];
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const int cc = sub + v * kLanesPerRow;
        if (cc < nv) acc[v] = axpy4(a0, spmm_ld<SHARDED>(X, xs, log2g, c0, row_f, cc), acc[v]);
      }
    }
    float* out = S + r * row_f;
    const bool split = work_split[w] != 0;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int cc = sub + v * kLanesPerRow;
      if (cc >= nv) continue;
      if (split) red_add4(out, cc, acc[v]); else st4(out, cc, acc[v]);
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(kGThreads) zero_rows_kernel(const int64_t* __restrict__ rows, int64_t n, int nv, float* S) {
  const int sub = threadIdx.x & (kLanesPerRow - 1);
  const int64_t group = ((int64_t)blockIdx.x * kGThreads + threadIdx.x) / kLanesPerRow;
  const int64_t n_groups = (int64_t)gridDim.x * kGThreads / kLanesPerRow;
  for (int64_t k = group; k < n; k += n_groups) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int cc = sub + v * kLanesPerRow;
      if (cc < nv) st4(S + rows[k] * (int64_t)nv * 4, cc, make_float4(0.f, 0.f, 0.f, 0.f));
    }
  }
}

// ---- element-wise pieces of graph_layer and its backward (float4 grid-stride) --------------------------------------
// mode 0: out = E + S + E*S                    (forward epilogue)
// mode 1: out = G * (1 + E)                    (backward: dS, the operand of the second SpMM)
// mode 2: out = G * (1 + S) + T                (backward: dE from the saved S and T = L.dS)
__global__ void __launch_bounds__(kGThreads) prop_elementwise_kernel(const float4* __restrict__ A, const float4* __restrict__ Bv,
                                                                     const float4* __restrict__ C, float4* __restrict__ out,
                                                                     int64_t n4, int mode) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = A[i], b = Bv[i];
    float4 o;
    if (mode == 0) {
      o = make_float4(a.x + (b.x + a.x * b.x), a.y + (b.y + a.y * b.y), a.z + (b.z + a.z * b.z), a.w + (b.w + a.w * b.w));
    } else if (mode == 1) {
      o = make_float4(a.x * (1.f + b.x), a.y * (1.f + b.y), a.z * (1.f + b.z), a.w * (1.f + b.w));
    } else {
      const float4 c = C[i];
      o = make_float4(a.x * (1.f + b.x) + c.x, a.y * (1.f + b.y) + c.y, a.z * (1.f + b.z) + c.z, a.w * (1.f + b.w) + c.w);
    }
    out[i] = o;
  }
}

// ---- transfer_layer + F.normalize, both domains in one pass ------------------------------------------------------------
struct TransferArgs {
  int64_t n_users, n_items, n_ov_users, n_ov_items;
  float lam_s, lam_t;
  const float* deg_s;  // [N] per-node degree in the source domain (users then items), fp32 (bitgcf.py:79-82)
  const float* deg_t;  // [N]
  int nv;
};

// coefficients of the (linear) transfer at node `row`:  Es' = ass*Es + ast*Et,  Et' = ats*Es + att*Et
__device__ __forceinline__ void transfer_coeffs(const TransferArgs& t, int64_t row, float& ass, float& ast, float& ats,
                                                float& att) {
  const bool is_user = row < t.n_users;
  const bool ov = is_user ? (row < t.n_ov_users) : (row - t.n_users < t.n_ov_items);
  if (!ov) {
    ass = 1.f; ast = 0.f; ats = 0.f; att = 1.f;
    return;
  }
  const float ds = t.deg_s[row], dt = t.deg_t[row];
  const float inv = 1.0f / (ds + dt + 1e-7f);
  const float ws = ds * inv, wt = dt * inv;
  ass = 0.5f * (t.lam_s + ws);
  ast = 0.5f * ((1.f - t.lam_s) + wt);
  ats = 0.5f * ((1.f - t.lam_t) + ws);
  att = 0.5f * (t.lam_t + wt);
}

template <int VEC>
__global__ void __launch_bounds__(kGThreads)
    transfer_norm_fwd_kernel(TransferArgs t, const float* __restrict__ Ps, const float* __restrict__ Pt, float* __restrict__ Es,
                             float* __restrict__ Et, float* __restrict__ Ns, float* __restrict__ Nt, int64_t n_ld) {
  const int sub = threadIdx.x & (kLanesPerRow - 1);
  const int64_t group = ((int64_t)blockIdx.x * kGThreads + threadIdx.x) / kLanesPerRow;
  const int64_t n_groups = (int64_t)gridDim.x * kGThreads / kLanesPerRow;
  const int64_t n = t.n_users + t.n_items, row_f = (int64_t)t.nv * 4;
  const int64_t warp_first = group - (group % kRowsPerWarp);
  for (int64_t base = warp_first; base < n; base += n_groups) {
    const int64_t row = base + (group % kRowsPerWarp);
    const bool live = row < n;
    const int64_t rr = live ? row : 0;
    float ass, ast, ats, att;
    transfer_coeffs(t, rr, ass, ast, ats, att);
    float4 es[VEC], et[VEC];
    float ss = 0.f, tt = 0.f;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int cc = sub + v * kLanesPerRow;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 ps = (live && cc < t.nv) ? ldg_row4(Ps + rr * row_f, cc) : z;
      const float4 pt = (live && cc < t.nv) ? ldg_row4(Pt + rr * row_f, cc) : z;
      es[v] = axpy4(ass, ps, scale4(ast, pt));
      et[v] = axpy4(ats, ps, scale4(att, pt));
      ss += dot4(es[v], es[v]);
      tt += dot4(et[v], et[v]);
    }
    ss = group8_sum(ss);
    tt = group8_sum(tt);
    const float is = 1.0f / fmaxf(sqrtf(ss), 1e-12f), it = 1.0f / fmaxf(sqrtf(tt), 1e-12f);  // F.normalize eps
    if (!live) continue;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int cc = sub + v * kLanesPerRow;
      if (cc >= t.nv) continue;
      st4(Es + row * row_f, cc, es[v]);
      st4(Et + row * row_f, cc, et[v]);
      st4(Ns + row * n_ld, cc, scale4(is, es[v]));
      st4(Nt + row * n_ld, cc, scale4(it, et[v]));
    }
  }
}

// backward: inputs Es, Et (the transferred tables saved by forward), dNs/dNt (grads of the normalised rows, row stride
// n_ld), dEs2/dEt2 (grads flowing into the transferred tables from the next layer; may be NULL) -> dPs, dPt
template <int VEC>
__global__ void __launch_bounds__(kGThreads)
    transfer_norm_bwd_kernel(TransferArgs t, const float* __restrict__ Es, const float* __restrict__ Et,
                             const float* __restrict__ dNs, const float* __restrict__ dNt, int64_t n_ld,
                             const float* __restrict__ dEs2, const float* __restrict__ dEt2, float* __restrict__ dPs,
                             float* __restrict__ dPt) {
  const int sub = threadIdx.x & (kLanesPerRow - 1);
  const int64_t group = ((int64_t)blockIdx.x * kGThreads + threadIdx.x) / kLanesPerRow;
  const int64_t n_groups = (int64_t)gridDim.x * kGThreads / kLanesPerRow;
  const int64_t n = t.n_users + t.n_items, row_f = (int64_t)t.nv * 4;
  const int64_t warp_first = group - (group % kRowsPerWarp);
  for (int64_t base = warp_first; base < n; base += n_groups) {
    const int64_t row = base + (group % kRowsPerWarp);
    const bool live = row < n;
    const int64_t rr = live ? row : 0;
    float ass, ast, ats, att;
    transfer_coeffs(t, rr, ass, ast, ats, att);
    float4 es[VEC], et[VEC], gs[VEC], gt[VEC];
    float ss = 0.f, tt = 0.f, sg = 0.f, tg = 0.f;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int cc = sub + v * kLanesPerRow;
      const bool on = live && cc < t.nv;
      es[v] = on ? ldg_row4(Es + rr * row_f, cc) : z;
      et[v] = on ? ldg_row4(Et + rr * row_f, cc) : z;
      gs[v] = on ? ldg_row4(dNs + rr * n_ld, cc) : z;
      gt[v] = on ? ldg_row4(dNt + rr * n_ld, cc) : z;
      ss += dot4(es[v], es[v]);
      tt += dot4(et[v], et[v]);
      sg += dot4(es[v], gs[v]);
      tg += dot4(et[v], gt[v]);
    }
    ss = group8_sum(ss);
    tt = group8_sum(tt);
    sg = group8_sum(sg);
    tg = group8_sum(tg);
    // y = x / max(||x||, eps):  dx = (g - y (y.g)) / max(||x||, eps)  when ||x|| > eps, else g / eps
    const float ns = sqrtf(ss), nt = sqrtf(tt);
    const float is = 1.0f / fmaxf(ns, 1e-12f), it = 1.0f / fmaxf(nt, 1e-12f);
    const float ks = ns > 1e-12f ? sg * is * is * is : 0.f, kt = nt > 1e-12f ? tg * it * it * it : 0.f;
    if (!live) continue;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int cc = sub + v * kLanesPerRow;
      if (cc >= t.nv) continue;
      float4 des = axpy4(-ks, es[v], scale4(is, gs[v]));  // grad wrt the transferred source row from its normalised copy
      float4 det = axpy4(-kt, et[v], scale4(it, gt[v]));
      if (dEs2) {
        const float4 a = ldg_row4(dEs2 + row * row_f, cc), b = ldg_row4(dEt2 + row * row_f, cc);
        des = make_float4(des.x + a.x, des.y + a.y, des.z + a.z, des.w + a.w);
        det = make_float4(det.x + b.x, det.y + b.y, det.z + b.z, det.w + b.w);
      }
      st4(dPs + row * row_f, cc, axpy4(ass, des, scale4(ats, det)));
      st4(dPt + row * row_f, cc, axpy4(ast, des, scale4(att, det)));
    }
  }
}

static inline int rows_grid(int64_t n_rows) {
  const int64_t per_block = kGThreads / kLanesPerRow;
  int64_t blocks = (n_rows + per_block - 1) / per_block;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace xdr

using namespace xdr;

extern "C" {

static int spmm_launch(const char* fn, const int64_t* work_row, const int64_t* work_beg, const int64_t* work_end,
                       const uint8_t* work_split, int64_t n_work, const int64_t* split_rows, int64_t n_split_rows,
                       const int64_t* col, const float* val, const float* X, const Shards* xs, int log2g, int dim, float* S,
                       xdr_stream_t stream) {
  XDR_REQUIRE(dim_ok(dim), "%s: dim=%d must be a multiple of 4 in (0, 256]", fn, dim);
  XDR_REQUIRE(n_work >= 0 && n_split_rows >= 0, "%s: negative size", fn);
  if (n_work == 0) return XDR_OK;
  XDR_REQUIRE(work_row && work_beg && work_end && work_split && col && val && (X || xs) && S, "%s: null pointer", fn);
  XDR_REQUIRE(aligned16(X) && aligned16(S), "%s: X and S must be 16-byte aligned", fn);
  const int nv = dim / 4;
  cudaStream_t s = (cudaStream_t)stream;
  if (n_split_rows > 0) {
    XDR_REQUIRE(split_rows, "%s: null split_rows", fn);
    XDR_DISPATCH_VEC(nv, (XDR_LAUNCH((zero_rows_kernel<VEC>), rows_grid(n_split_rows), kGThreads, 0, s, split_rows, n_split_rows, nv, S)));
    XDR_LAUNCH_OK();
  }
  if (xs != nullptr) {
    const Shards sh = *xs;
    XDR_DISPATCH_VEC(nv, (XDR_LAUNCH((spmm_work_kernel<VEC, true>), rows_grid(n_work), kGThreads, 0, s, 
                             work_row, work_beg, work_end, work_split, n_work, col, val, nullptr, sh, log2g, nv, S)));
  } else {
    const Shards sh{};
    XDR_DISPATCH_VEC(nv, (XDR_LAUNCH((spmm_work_kernel<VEC, false>), rows_grid(n_work), kGThreads, 0, s, 
                             work_row, work_beg, work_end, work_split, n_work, col, val, X, sh, 0, nv, S)));
  }
  XDR_LAUNCH_OK();
  return XDR_OK;
}

int xdr_spmm_csr(const int64_t* work_row, const int64_t* work_beg, const int64_t* work_end, const uint8_t* work_split,
                 int64_t n_work, const int64_t* split_rows, int64_t n_split_rows, const int64_t* col, const float* val,
                 const float* X, int dim, float* S, xdr_stream_t stream) {
  return spmm_launch("xdr_spmm_csr", work_row, work_beg, work_end, work_split, n_work, split_rows, n_split_rows, col, val, X,
                     nullptr, 0, dim, S, stream);
}

int xdr_spmm_csr_sharded(const int64_t* work_row, const int64_t* work_beg, const int64_t* work_end, const uint8_t* work_split,
                         int64_t n_work, const int64_t* split_rows, int64_t n_split_rows, const int64_t* col,
                         const float* val, const float* const* x_shards, int n_shards, int dim, float* S,
                         xdr_stream_t stream) {
  XDR_REQUIRE(n_shards >= 1 && n_shards <= kMaxShards && (n_shards & (n_shards - 1)) == 0,
              "xdr_spmm_csr_sharded: n_shards=%d must be a power of two <= %d", n_shards, kMaxShards);
  XDR_REQUIRE(x_shards, "xdr_spmm_csr_sharded: null pointer");
  Shards xs{};
  int log2g = 0;
  while ((1 << log2g) < n_shards) ++log2g;
  for (int g = 0; g < n_shards; ++g) {
    XDR_REQUIRE(x_shards[g] && aligned16(x_shards[g]), "xdr_spmm_csr_sharded: null or unaligned shard %d", g);
    xs.p[g] = const_cast<float*>(x_shards[g]);
  }
  return spmm_launch("xdr_spmm_csr_sharded", work_row, work_beg, work_end, work_split, n_work, split_rows, n_split_rows, col,
                     val, nullptr, &xs, log2g, dim, S, stream);
}

int xdr_prop_elementwise(const float* A, const float* B, const float* C, float* out, int64_t count, int mode,
                         xdr_stream_t stream) {
  XDR_REQUIRE(count >= 0 && count % 4 == 0, "xdr_prop_elementwise: count must be a non-negative multiple of 4");
  if (count == 0) return XDR_OK;
  XDR_REQUIRE(mode >= 0 && mode <= 2, "xdr_prop_elementwise: bad mode %d", mode);
  XDR_REQUIRE(A && B && out && (mode != 2 || C), "xdr_prop_elementwise: null pointer");
  XDR_REQUIRE(aligned16(A) && aligned16(B) && aligned16(out) && aligned16(C), "xdr_prop_elementwise: 16-byte alignment");
  const int64_t n4 = count / 4;
  int64_t blocks = (n4 + kGThreads - 1) / kGThreads;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  XDR_LAUNCH((prop_elementwise_kernel), (int)blocks, kGThreads, 0, (cudaStream_t)stream, 
      reinterpret_cast<const float4*>(A), reinterpret_cast<const float4*>(B), reinterpret_cast<const float4*>(C),
      reinterpret_cast<float4*>(out), n4, mode);
  XDR_LAUNCH_OK();
  return XDR_OK;
}

static int check_transfer(const char* fn, int64_t n_users, int64_t n_items, int64_t n_ov_users, int64_t n_ov_items, int dim,
                          const float* deg_s, const float* deg_t, int64_t n_ld) {
  XDR_REQUIRE(dim_ok(dim), "%s: dim=%d must be a multiple of 4 in (0, 256]", fn, dim);
  XDR_REQUIRE(n_users >= 0 && n_items >= 0 && n_ov_users >= 0 && n_ov_users <= n_users && n_ov_items >= 0 &&
                  n_ov_items <= n_items,
              "%s: bad node counts", fn);
  XDR_REQUIRE(deg_s && deg_t, "%s: null degree vector", fn);
  XDR_REQUIRE(n_ld >= dim && n_ld % 4 == 0, "%s: n_ld=%lld invalid", fn, (long long)n_ld);
  return XDR_OK;
}

int xdr_transfer_norm_fwd(const float* Ps, const float* Pt, int64_t n_users, int64_t n_items, int64_t n_ov_users,
                          int64_t n_ov_items, int dim, float lam_s, float lam_t, const float* deg_s, const float* deg_t,
                          float* Es, float* Et, float* Ns, float* Nt, int64_t n_ld, xdr_stream_t stream) {
  int rc = check_transfer("xdr_transfer_norm_fwd", n_users, n_items, n_ov_users, n_ov_items, dim, deg_s, deg_t, n_ld);
  if (rc) return rc;
  if (n_users + n_items == 0) return XDR_OK;
  XDR_REQUIRE(Ps && Pt && Es && Et && Ns && Nt, "xdr_transfer_norm_fwd: null pointer");
  XDR_REQUIRE(aligned16(Ps) && aligned16(Pt) && aligned16(Es) && aligned16(Et) && aligned16(Ns) && aligned16(Nt),
              "xdr_transfer_norm_fwd: 16-byte alignment");
  TransferArgs t{n_users, n_items, n_ov_users, n_ov_items, lam_s, lam_t, deg_s, deg_t, dim / 4};
  XDR_DISPATCH_VEC(t.nv, (XDR_LAUNCH((transfer_norm_fwd_kernel<VEC>), rows_grid(n_users + n_items), kGThreads, 0, (cudaStream_t)stream, 
                             t, Ps, Pt, Es, Et, Ns, Nt, n_ld)));
  XDR_LAUNCH_OK();
  return XDR_OK;
}

int xdr_transfer_norm_bwd(const float* Es, const float* Et, const float* dNs, const float* dNt, int64_t n_ld,
                          const float* dEs2, const float* dEt2, int64_t n_users, int64_t n_items, int64_t n_ov_users,
                          int64_t n_ov_items, int dim, float lam_s, float lam_t, const float* deg_s, const float* deg_t,
                          float* dPs, float* dPt, xdr_stream_t stream) {
  int rc = check_transfer("xdr_transfer_norm_bwd", n_users, n_items, n_ov_users, n_ov_items, dim, deg_s, deg_t, n_ld);
  if (rc) return rc;
  if (n_users + n_items == 0) return XDR_OK;
  XDR_REQUIRE(Es && Et && dNs && dNt && dPs && dPt, "xdr_transfer_norm_bwd: null pointer");
  XDR_REQUIRE((dEs2 == nullptr) == (dEt2 == nullptr), "xdr_transfer_norm_bwd: dEs2 and dEt2 go together");
  TransferArgs t{n_users, n_items, n_ov_users, n_ov_items, lam_s, lam_t, deg_s, deg_t, dim / 4};
  XDR_DISPATCH_VEC(t.nv, (XDR_LAUNCH((transfer_norm_bwd_kernel<VEC>), rows_grid(n_users + n_items), kGThreads, 0, (cudaStream_t)stream, 
                             t, Es, Et, dNs, dNt, n_ld, dEs2, dEt2, dPs, dPt)));
  XDR_LAUNCH_OK();
  return XDR_OK;
}

}  // extern "C"
