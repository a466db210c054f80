// sparse_optim.cu -- section 8 F1: row-sparse optimizer step over the rows a batch touched.
//
// The reference trains with dense torch optimizers (recbole Trainer._build_optimizer [recbole-1.0.1], `learner: adam`,
// overall.yaml:20-21): every step zero-fills a dense [N, D] gradient per table, index_adds the batch into it and then reads
// and writes EVERY row of the weights and of the optimizer state -- >= 95 % of the reference step at 10^6-row tables
// (SURVEY section 8 A17).  Here the step kernels scatter-add the batch's gradient rows into a gradient table that is all
// zeros outside the touched rows; this kernel then visits only the batch's ids:
//     first visitor of a row this step (atomicMax on a per-row step stamp) owns it: reads the row's summed gradient,
//     updates the state and the weights, and writes the gradient row back to zero -- so there is no dense zero_grad either.
// Per touched row: 8 B id + 4 B stamp + rows read/written (SGD: G r/w + W r/w = 4 x 4D; Adagrad 6 x 4D; Adam 8 x 4D).
//   XDR_OPT_SGD       w -= lr * g                                     == torch.optim.SGD (no momentum / weight decay)
//   XDR_OPT_ADAGRAD   s += g*g;  w -= lr * g / (sqrt(s) + eps)        == torch.optim.Adagrad (lr_decay 0): a zero gradient
//                                                                        changes nothing there either, so row-sparse == dense
//   XDR_OPT_LAZY_ADAM m += (1-b1)(g-m); v += (1-b2)(g*g-v);
//                     w -= lr*sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v)+eps)  == torch.optim.SparseAdam (touched rows only;
//                                                                        dense Adam also moves untouched rows -- different)
#include <math.h>

#include "xdr_common.cuh"

namespace xdr {

constexpr int kOptThreads = 256;

template <int KIND>
__global__ void __launch_bounds__(kOptThreads)
    sparse_optim_rows_kernel(float* __restrict__ W, float* __restrict__ G, float* __restrict__ S1, float* __restrict__ S2,
                             int* __restrict__ stamp, const int64_t* __restrict__ ids, int64_t n, int64_t n_rows, int dim,
                             int step_id, float lr, float eps, float one_minus_beta1, float one_minus_beta2,
                             float adam_step_size,
                             int32_t* __restrict__ oob) {
  const int lane = threadIdx.x & 31, sub = lane & (kLanesPerRow - 1), grp = lane >> 3;
  const int nv = dim >> 2;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i0 = warp0 * kRowsPerWarp; i0 < n; i0 += n_warps * kRowsPerWarp) {  // warp-uniform trip count
    const int64_t i = i0 + grp;
    int64_t id = -1;
    if (i < n) {
      id = ids[i];
      if ((uint64_t)id >= (uint64_t)n_rows) {
        if (oob && sub == 0) *oob = 1;
        id = -1;
      }
    }
    int owner = 0;
    if (sub == 0 && id >= 0) owner = atomicMax(&stamp[id], step_id) < step_id ? 1 : 0;  // first visitor this step
    owner = __shfl_sync(0xffffffffu, owner, lane & ~(kLanesPerRow - 1));
    if (!owner) continue;
    float* g_row = G + id * dim;
    float* w_row = W + id * dim;
    for (int c = sub; c < nv; c += kLanesPerRow) {
      const float4 g = ld_row4(g_row, c);
      float4 w = ld_row4(w_row, c);
      if (KIND == XDR_OPT_SGD) {
        w = axpy4(-lr, g, w);
      } else if (KIND == XDR_OPT_ADAGRAD) {
        float4 s = ld_row4(S1 + id * dim, c);
        s = make_float4(fmaf(g.x, g.x, s.x), fmaf(g.y, g.y, s.y), fmaf(g.z, g.z, s.z), fmaf(g.w, g.w, s.w));
        st4(S1 + id * dim, c, s);
        w.x -= lr * g.x / (sqrtf(s.x) + eps);
        w.y -= lr * g.y / (sqrtf(s.y) + eps);
        w.z -= lr * g.z / (sqrtf(s.z) + eps);
        w.w -= lr * g.w / (sqrtf(s.w) + eps);
      } else {
        float4 m = ld_row4(S1 + id * dim, c), v = ld_row4(S2 + id * dim, c);
        const float a1 = one_minus_beta1, a2 = one_minus_beta2;  // formed in double on the host, as torch does
        m = make_float4(m.x + a1 * (g.x - m.x), m.y + a1 * (g.y - m.y), m.z + a1 * (g.z - m.z), m.w + a1 * (g.w - m.w));
        v = make_float4(v.x + a2 * (g.x * g.x - v.x), v.y + a2 * (g.y * g.y - v.y), v.z + a2 * (g.z * g.z - v.z),
                        v.w + a2 * (g.w * g.w - v.w));
        st4(S1 + id * dim, c, m);
        st4(S2 + id * dim, c, v);
        w.x -= adam_step_size * m.x / (sqrtf(v.x) + eps);
        w.y -= adam_step_size * m.y / (sqrtf(v.y) + eps);
        w.z -= adam_step_size * m.z / (sqrtf(v.z) + eps);
        w.w -= adam_step_size * m.w / (sqrtf(v.w) + eps);
      }
      st4(w_row, c, w);
      st4(g_row, c, make_float4(0.f, 0.f, 0.f, 0.f));  // leave the gradient table clean for the next step
    }
  }
}

static int optim_check(int kind, const float* W, const float* G, const float* S1, const float* S2, const int* stamp,
                       const int64_t* ids, int64_t n, int64_t n_rows, int dim, int step_id) {
  XDR_REQUIRE(kind == XDR_OPT_SGD || kind == XDR_OPT_ADAGRAD || kind == XDR_OPT_LAZY_ADAM, "xdr_sparse_optim_rows: bad kind %d", kind);
  XDR_REQUIRE(dim_ok(dim) && n >= 0 && n_rows > 0, "xdr_sparse_optim_rows: bad dim/count");
  XDR_REQUIRE(step_id > 0, "xdr_sparse_optim_rows: step_id must start at 1 and grow (the stamp table starts at 0)");
  if (n == 0) return XDR_OK;
  XDR_REQUIRE(W && G && stamp && ids, "xdr_sparse_optim_rows: null pointer");
  XDR_REQUIRE(kind == XDR_OPT_SGD || S1, "xdr_sparse_optim_rows: missing state table");
  XDR_REQUIRE(kind != XDR_OPT_LAZY_ADAM || S2, "xdr_sparse_optim_rows: lazy Adam needs two state tables");
  XDR_REQUIRE(aligned16(W) && aligned16(G) && (!S1 || aligned16(S1)) && (!S2 || aligned16(S2)),
              "xdr_sparse_optim_rows: tables must be 16-byte aligned");
  return XDR_OK;
}

// bias-corrected Adam step size of torch.optim.SparseAdam: lr * sqrt(1 - b2^t) / (1 - b1^t)
static float adam_step_size(float lr, double beta1, double beta2, int64_t t) {
  const double bc1 = 1.0 - pow(beta1, (double)t), bc2 = 1.0 - pow(beta2, (double)t);
  return (float)((double)lr * sqrt(bc2) / bc1);
}

}  // namespace xdr

using namespace xdr;

extern "C" {

int xdr_sparse_optim_rows(int kind, float* W, float* G, float* S1, float* S2, int32_t* stamp, const int64_t* ids, int64_t n,
                          int64_t n_rows, int dim, int step_id, int64_t adam_t, float lr, float eps, double beta1, double beta2,
                          int32_t* oob, xdr_stream_t stream) {
  const int rc = optim_check(kind, W, G, S1, S2, stamp, ids, n, n_rows, dim, step_id);
  if (rc != XDR_OK || n == 0) return rc;
  XDR_REQUIRE(kind != XDR_OPT_LAZY_ADAM || adam_t > 0, "xdr_sparse_optim_rows: adam_t (optimizer step count) must be >= 1");
  const float ss = kind == XDR_OPT_LAZY_ADAM ? adam_step_size(lr, beta1, beta2, adam_t) : 0.f;
  const float a1 = (float)(1.0 - beta1), a2 = (float)(1.0 - beta2);
  const int64_t warps = (n + kRowsPerWarp - 1) / kRowsPerWarp;
  int64_t blocks = (warps * 32 + kOptThreads - 1) / kOptThreads;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  cudaStream_t s = (cudaStream_t)stream;
  if (kind == XDR_OPT_SGD)
    XDR_LAUNCH((sparse_optim_rows_kernel<XDR_OPT_SGD>), (unsigned)blocks, kOptThreads, 0, s, W, G, S1, S2, stamp, ids, n, n_rows, dim, step_id, lr, eps, a1, a2, ss, oob);
  else if (kind == XDR_OPT_ADAGRAD)
    XDR_LAUNCH((sparse_optim_rows_kernel<XDR_OPT_ADAGRAD>), (unsigned)blocks, kOptThreads, 0, s, W, G, S1, S2, stamp, ids, n, n_rows, dim, step_id, lr, eps, a1, a2, ss, oob);
  else
    XDR_LAUNCH((sparse_optim_rows_kernel<XDR_OPT_LAZY_ADAM>), (unsigned)blocks, kOptThreads, 0, s, W, G, S1, S2, stamp, ids, n, n_rows, dim, step_id, lr, eps, a1, a2, ss, oob);
  XDR_LAUNCH_OK();
  return XDR_OK;
}

}  // extern "C"
