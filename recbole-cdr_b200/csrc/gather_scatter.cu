// gather_scatter.cu -- A1: embedding-row gather and gradient scatter-add (plus DTCDR's max-combine).
//
// Replaces torch.nn.Embedding.__call__ / embedding_dense_backward on the RecBole-CDR hot path
// (reference emcdr.py:99-100, conet.py:106-109, dtcdr.py:113-119, cmf.py:53-73, bitgcf.py:221-224).
//
// Mapping: one row per 8-lane group, each lane moves VEC float4 columns (col = sub + 8*k), so a group touches
// whole 128-byte lines and a warp keeps 4 rows x VEC independent 16-byte requests in flight.  HBM-bound:
// algorithmic bytes per row = 8 (id) + 2 * 4*dim (read + write).
#include "xdr_common.cuh"

namespace xdr {

constexpr int kThreads = 256;

template <int VEC>
__global__ void __launch_bounds__(kThreads) gather_rows_kernel(const float* __restrict__ table, int64_t n_rows, int nv,
                                                               const int64_t* __restrict__ idx, int64_t n_idx,
                                                               float* __restrict__ out, int64_t out_ld, int32_t* oob) {
  const int sub = threadIdx.x & (kLanesPerRow - 1);
  const int64_t group = ((int64_t)blockIdx.x * kThreads + threadIdx.x) / kLanesPerRow;
  const int64_t n_groups = (int64_t)gridDim.x * kThreads / kLanesPerRow;
  for (int64_t k = group; k < n_idx; k += n_groups) {
    const int64_t r = idx[k];
    const bool ok = (uint64_t)r < (uint64_t)n_rows;
    if (!ok && oob && sub == 0) *oob = 1;
    const float* src = table + (ok ? r : 0) * (int64_t)(nv * 4);
    float4 v[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int c = sub + j * kLanesPerRow;
      v[j] = (ok && c < nv) ? ldg_row4(src, c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float* dst = out + k * out_ld;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int c = sub + j * kLanesPerRow;
      if (c < nv) st4(dst, c, v[j]);
    }
  }
}

// out[k,:] = shard[idx[k] mod G][idx[k] div G, :]: the peer-gather that pulls a chunk's rows over NVLink one chunk ahead
template <int VEC>
__global__ void __launch_bounds__(kThreads) gather_rows_sharded_kernel(Shards tab, int log2g, int64_t n_rows, int nv,
                                                                       const int64_t* __restrict__ idx, int64_t n_idx,
                                                                       int64_t idx_batch, int64_t idx_step_stride,
                                                                       float* __restrict__ out, int64_t out_ld, int32_t* oob) {
  // id k lives at idx[(k / idx_batch) * idx_step_stride + k % idx_batch]: the [K, rows, B] id block of the trainer
  auto id_at = [&](int64_t k) { return idx[(k / idx_batch) * idx_step_stride + (k % idx_batch)]; };
  const int sub = threadIdx.x & (kLanesPerRow - 1);
  const int64_t group = ((int64_t)blockIdx.x * kThreads + threadIdx.x) / kLanesPerRow;
  const int64_t n_groups = (int64_t)gridDim.x * kThreads / kLanesPerRow;
  // two rows per group iteration: 2*VEC independent 16-byte requests in flight per lane (NVLink latency is long)
  for (int64_t k = group; k < n_idx; k += 2 * n_groups) {
    const int64_t k2 = k + n_groups;
    const int64_t r0 = id_at(k), r1 = k2 < n_idx ? id_at(k2) : -1;
    const bool ok0 = (uint64_t)r0 < (uint64_t)n_rows, ok1 = (uint64_t)r1 < (uint64_t)n_rows;
    if (oob && sub == 0 && (!ok0 || (k2 < n_idx && !ok1))) *oob = 1;
    const float* s0 = shard_row(tab, log2g, ok0 ? r0 : 0, (int64_t)nv * 4);
    const float* s1 = shard_row(tab, log2g, ok1 ? r1 : 0, (int64_t)nv * 4);
    float4 v0[VEC], v1[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int c = sub + j * kLanesPerRow;
      v0[j] = (ok0 && c < nv) ? ldg_row4(s0, c) : make_float4(0.f, 0.f, 0.f, 0.f);
      v1[j] = (ok1 && c < nv) ? ldg_row4(s1, c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int c = sub + j * kLanesPerRow;
      if (c >= nv) continue;
      st4(out + k * out_ld, c, v0[j]);
      if (k2 < n_idx) st4(out + k2 * out_ld, c, v1[j]);
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(kThreads) scatter_add_rows_kernel(float* __restrict__ dst, int64_t n_rows, int nv,
                                                                    const int64_t* __restrict__ idx, int64_t n_idx,
                                                                    const float* __restrict__ rows, int64_t rows_ld,
                                                                    float scale, int32_t* oob) {
  const int sub = threadIdx.x & (kLanesPerRow - 1);
  const int64_t group = ((int64_t)blockIdx.x * kThreads + threadIdx.x) / kLanesPerRow;
  const int64_t n_groups = (int64_t)gridDim.x * kThreads / kLanesPerRow;
  for (int64_t k = group; k < n_idx; k += n_groups) {
    const int64_t r = idx[k];
    const bool ok = (uint64_t)r < (uint64_t)n_rows;
    if (!ok) {
      if (oob && sub == 0) *oob = 1;
      continue;
    }
    const float* src = rows + k * rows_ld;
    float4 v[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int c = sub + j * kLanesPerRow;
      v[j] = (c < nv) ? ldg_row4(src, c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float* d = dst + r * (int64_t)(nv * 4);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int c = sub + j * kLanesPerRow;
      if (c < nv) red_add4(d, c, scale4(scale, v[j]));
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(kThreads) gather_max2_kernel(const float* __restrict__ ta, const float* __restrict__ tb,
                                                               int64_t n_rows, int nv, const int64_t* __restrict__ idx,
                                                               int64_t n_idx, float* __restrict__ out, int64_t out_ld,
                                                               int32_t* oob) {
  const int sub = threadIdx.x & (kLanesPerRow - 1);
  const int64_t group = ((int64_t)blockIdx.x * kThreads + threadIdx.x) / kLanesPerRow;
  const int64_t n_groups = (int64_t)gridDim.x * kThreads / kLanesPerRow;
  for (int64_t k = group; k < n_idx; k += n_groups) {
    const int64_t r = idx[k];
    const bool ok = (uint64_t)r < (uint64_t)n_rows;
    if (!ok && oob && sub == 0) *oob = 1;
    const int64_t off = (ok ? r : 0) * (int64_t)(nv * 4);
    float4 a[VEC], b[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int c = sub + j * kLanesPerRow;
      const bool on = ok && c < nv;
      a[j] = on ? ldg_row4(ta + off, c) : make_float4(0.f, 0.f, 0.f, 0.f);
      b[j] = on ? ldg_row4(tb + off, c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float* dst = out + k * out_ld;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int c = sub + j * kLanesPerRow;
      // torch.maximum propagates NaN; fmaxf does not -> spell it out
      auto mx = [](float x, float y) { return (x != x || y != y) ? (x + y) : (x > y ? x : y); };
      if (c < nv) st4(dst, c, make_float4(mx(a[j].x, b[j].x), mx(a[j].y, b[j].y), mx(a[j].z, b[j].z), mx(a[j].w, b[j].w)));
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(kThreads)
    scatter_max2_bwd_kernel(const float* __restrict__ ta, const float* __restrict__ tb, int64_t n_rows, int nv,
                            const int64_t* __restrict__ idx, int64_t n_idx, const float* __restrict__ grad,
                            int64_t grad_ld, float scale, float* __restrict__ da, float* __restrict__ db, int32_t* oob) {
  const int sub = threadIdx.x & (kLanesPerRow - 1);
  const int64_t group = ((int64_t)blockIdx.x * kThreads + threadIdx.x) / kLanesPerRow;
  const int64_t n_groups = (int64_t)gridDim.x * kThreads / kLanesPerRow;
  for (int64_t k = group; k < n_idx; k += n_groups) {
    const int64_t r = idx[k];
    const bool ok = (uint64_t)r < (uint64_t)n_rows;
    if (!ok) {
      if (oob && sub == 0) *oob = 1;
      continue;
    }
    const int64_t off = r * (int64_t)(nv * 4);
    float4 a[VEC], b[VEC], g[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int c = sub + j * kLanesPerRow;
      const bool on = c < nv;
      a[j] = on ? ld_row4(ta + off, c) : make_float4(0.f, 0.f, 0.f, 0.f);
      b[j] = on ? ld_row4(tb + off, c) : make_float4(0.f, 0.f, 0.f, 0.f);
      g[j] = on ? ldg_row4(grad + k * grad_ld, c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int c = sub + j * kLanesPerRow;
      if (c >= nv) continue;
      // torch.maximum backward: grad * (a > b) + grad/2 * (a == b) for a, mirrored for b
      auto wa = [](float x, float y) { return x > y ? 1.f : (x == y ? 0.5f : 0.f); };
      float4 ga = make_float4(scale * g[j].x * wa(a[j].x, b[j].x), scale * g[j].y * wa(a[j].y, b[j].y),
                              scale * g[j].z * wa(a[j].z, b[j].z), scale * g[j].w * wa(a[j].w, b[j].w));
      float4 gb = make_float4(scale * g[j].x * wa(b[j].x, a[j].x), scale * g[j].y * wa(b[j].y, a[j].y),
                              scale * g[j].z * wa(b[j].z, a[j].z), scale * g[j].w * wa(b[j].w, a[j].w));
      red_add4(da + off, c, ga);
      red_add4(db + off, c, gb);
    }
  }
}

static inline int grid_for_rows(int64_t n_rows) {
  const int64_t groups_per_block = kThreads / kLanesPerRow;
  int64_t blocks = (n_rows + groups_per_block - 1) / groups_per_block;
  const int64_t cap = (int64_t)sm_count() * 8;  // 8 resident 256-thread CTAs per SM: one full wave, then grid-stride
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace xdr

using namespace xdr;


extern "C" {

int xdr_gather_rows(const float* table, int64_t n_rows, int dim, const int64_t* idx, int64_t n_idx, float* out,
                    int64_t out_ld, int32_t* oob, xdr_stream_t stream) {
  XDR_REQUIRE(dim_ok(dim), "xdr_gather_rows: dim=%d must be a multiple of 4 in (0, 256]", dim);
  XDR_REQUIRE(n_idx >= 0 && n_rows >= 0, "xdr_gather_rows: negative size");
  if (n_idx == 0) return XDR_OK;
  XDR_REQUIRE(table && idx && out, "xdr_gather_rows: null pointer");
  XDR_REQUIRE(out_ld >= dim && out_ld % 4 == 0, "xdr_gather_rows: out_ld=%lld must be >= dim and a multiple of 4",
              (long long)out_ld);
  XDR_REQUIRE(aligned16(table) && aligned16(out), "xdr_gather_rows: table/out must be 16-byte aligned");
  const int nv = dim / 4;
  cudaStream_t s = (cudaStream_t)stream;
  XDR_DISPATCH_VEC(nv, (XDR_LAUNCH((gather_rows_kernel<VEC>), grid_for_rows(n_idx), kThreads, 0, s, table, n_rows, nv, idx, n_idx,
                                                                                          out, out_ld, oob)));
  XDR_LAUNCH_OK();
  return XDR_OK;
}

int xdr_gather_rows_sharded(const float* const* shards, int n_shards, int64_t n_rows, int dim, const int64_t* idx,
                            int64_t n_idx, int64_t idx_batch, int64_t idx_step_stride, float* out, int64_t out_ld,
                            int32_t* oob, xdr_stream_t stream) {
  if (idx_batch <= 0) { idx_batch = n_idx > 0 ? n_idx : 1; idx_step_stride = idx_batch; }
  XDR_REQUIRE(dim_ok(dim), "xdr_gather_rows_sharded: dim=%d must be a multiple of 4 in (0, 256]", dim);
  XDR_REQUIRE(n_shards >= 1 && n_shards <= kMaxShards && (n_shards & (n_shards - 1)) == 0,
              "xdr_gather_rows_sharded: n_shards=%d must be a power of two <= %d", n_shards, kMaxShards);
  XDR_REQUIRE(n_idx >= 0 && n_rows >= 0, "xdr_gather_rows_sharded: negative size");
  if (n_idx == 0) return XDR_OK;
  XDR_REQUIRE(shards && idx && out, "xdr_gather_rows_sharded: null pointer");
  XDR_REQUIRE(out_ld >= dim && out_ld % 4 == 0 && aligned16(out), "xdr_gather_rows_sharded: bad output layout");
  Shards t{};
  int log2g = 0;
  while ((1 << log2g) < n_shards) ++log2g;
  for (int g = 0; g < n_shards; ++g) {
    XDR_REQUIRE(shards[g] && aligned16(shards[g]), "xdr_gather_rows_sharded: null or unaligned shard %d", g);
    t.p[g] = const_cast<float*>(shards[g]);
  }
  const int nv = dim / 4;
  cudaStream_t s = (cudaStream_t)stream;
  // NVLink latency is ~10x local: run many CTAs per SM (grid-stride) so that thousands of row requests are in flight
  int64_t blocks = (n_idx + 2 * (kThreads / kLanesPerRow) - 1) / (2 * (kThreads / kLanesPerRow));
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  XDR_DISPATCH_VEC(nv, (XDR_LAUNCH((gather_rows_sharded_kernel<VEC>), (int)blocks, kThreads, 0, s, t, log2g, n_rows, nv, idx, n_idx, idx_batch,
                                                                                         idx_step_stride, out, out_ld, oob)));
  XDR_LAUNCH_OK();
  return XDR_OK;
}

int xdr_scatter_add_rows(float* dst, int64_t n_rows, int dim, const int64_t* idx, int64_t n_idx, const float* rows,
                         int64_t rows_ld, float scale, int32_t* oob, xdr_stream_t stream) {
  XDR_REQUIRE(dim_ok(dim), "xdr_scatter_add_rows: dim=%d must be a multiple of 4 in (0, 256]", dim);
  XDR_REQUIRE(n_idx >= 0 && n_rows >= 0, "xdr_scatter_add_rows: negative size");
  if (n_idx == 0) return XDR_OK;
  XDR_REQUIRE(dst && idx && rows, "xdr_scatter_add_rows: null pointer");
  XDR_REQUIRE(rows_ld >= dim && rows_ld % 4 == 0, "xdr_scatter_add_rows: rows_ld=%lld invalid", (long long)rows_ld);
  XDR_REQUIRE(aligned16(dst) && aligned16(rows), "xdr_scatter_add_rows: dst/rows must be 16-byte aligned");
  const int nv = dim / 4;
  cudaStream_t s = (cudaStream_t)stream;
  XDR_DISPATCH_VEC(nv, (XDR_LAUNCH((scatter_add_rows_kernel<VEC>), grid_for_rows(n_idx), kThreads, 0, s, 
                           dst, n_rows, nv, idx, n_idx, rows, rows_ld, scale, oob)));
  XDR_LAUNCH_OK();
  return XDR_OK;
}

int xdr_gather_max2(const float* table_a, const float* table_b, int64_t n_rows, int dim, const int64_t* idx,
                    int64_t n_idx, float* out, int64_t out_ld, int32_t* oob, xdr_stream_t stream) {
  XDR_REQUIRE(dim_ok(dim), "xdr_gather_max2: dim=%d must be a multiple of 4 in (0, 256]", dim);
  XDR_REQUIRE(n_idx >= 0 && n_rows >= 0, "xdr_gather_max2: negative size");
  if (n_idx == 0) return XDR_OK;
  XDR_REQUIRE(table_a && table_b && idx && out, "xdr_gather_max2: null pointer");
  XDR_REQUIRE(out_ld >= dim && out_ld % 4 == 0, "xdr_gather_max2: out_ld invalid");
  XDR_REQUIRE(aligned16(table_a) && aligned16(table_b) && aligned16(out), "xdr_gather_max2: 16-byte alignment");
  const int nv = dim / 4;
  cudaStream_t s = (cudaStream_t)stream;
  XDR_DISPATCH_VEC(nv, (XDR_LAUNCH((gather_max2_kernel<VEC>), grid_for_rows(n_idx), kThreads, 0, s, table_a, table_b, n_rows, nv,
                                                                                          idx, n_idx, out, out_ld, oob)));
  XDR_LAUNCH_OK();
  return XDR_OK;
}

int xdr_scatter_max2_bwd(const float* table_a, const float* table_b, int64_t n_rows, int dim, const int64_t* idx,
                         int64_t n_idx, const float* grad_rows, int64_t grad_ld, float scale, float* dst_a,
                         float* dst_b, int32_t* oob, xdr_stream_t stream) {
  XDR_REQUIRE(dim_ok(dim), "xdr_scatter_max2_bwd: dim=%d must be a multiple of 4 in (0, 256]", dim);
  XDR_REQUIRE(n_idx >= 0 && n_rows >= 0, "xdr_scatter_max2_bwd: negative size");
  if (n_idx == 0) return XDR_OK;
  XDR_REQUIRE(table_a && table_b && idx && grad_rows && dst_a && dst_b, "xdr_scatter_max2_bwd: null pointer");
  XDR_REQUIRE(grad_ld >= dim && grad_ld % 4 == 0, "xdr_scatter_max2_bwd: grad_ld invalid");
  XDR_REQUIRE(aligned16(table_a) && aligned16(table_b) && aligned16(grad_rows) && aligned16(dst_a) && aligned16(dst_b),
              "xdr_scatter_max2_bwd: 16-byte alignment");
  const int nv = dim / 4;
  cudaStream_t s = (cudaStream_t)stream;
  XDR_DISPATCH_VEC(nv, (XDR_LAUNCH((scatter_max2_bwd_kernel<VEC>), grid_for_rows(n_idx), kThreads, 0, s, 
                           table_a, table_b, n_rows, nv, idx, n_idx, grad_rows, grad_ld, scale, dst_a, dst_b, oob)));
  XDR_LAUNCH_OK();
  return XDR_OK;
}

}  // extern "C"
