// pair_score.cu -- A2/A3/A13/A16: fused gather -> dot-product score -> loss (+EmbLoss), and its backward
// (re-gather -> per-row gradient -> vector-atomic scatter-add).
//
// Replaces, per batch:
//   EMCDR.source_forward/target_forward + calculate_*_loss  (reference emcdr.py:98-108, 110-154)
//   CMF.forward + per-domain loss term                     (cmf.py:75-98)
//   the two halves of BiTGCF.calculate_loss                (bitgcf.py:221-247)
// and the autograd backward of those (embedding_dense_backward = index_add into a dense [N, D] grad).
//
// The reference performs 6 row gathers per BPR interaction and materialises 6 [B, D] temporaries; here the
// 3 rows are read once in forward (registers only) and once more in backward (L2-resident: a batch's rows are
// ~6 MB against 126 MB of L2), and gradients go straight to their destination rows with REDG.E.ADD.F32x4.
// EmbLoss couples the whole batch (d||E||_F/dE_r = E_r/||E||_F), hence the fwd -> bwd two-kernel structure.
//
// HBM roofline: algorithmic bytes per BPR interaction = 3*8 (ids) + 3*4*dim (gather) + 3*4*dim (scatter)
//             = 1560 B at dim 64; pointwise = 2*8 + 4 + 4*4*dim = 1044 B at dim 64  (SURVEY.md section 8 D3).
#include "xdr_common.cuh"

namespace xdr {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

struct PairArgs {
  const float* user_tab;
  const float* item_tab;
  int64_t n_users, n_items;
  int nv;  // float4s per row
  const int64_t* user;
  const int64_t* item_a;  // positive item (BPR) / the item (pointwise)
  const int64_t* item_b;  // negative item (BPR) / unused
  const float* label;     // pointwise only
  int64_t batch;
  int loss_kind;  // pointwise only
  float gamma, reg_weight;
  float* score_a;  // pos score / raw score
  float* score_b;  // neg score
  float* out8;
  // backward only
  const float* grad_loss;
  float scale;
  float* user_dst;
  float* item_dst;
};

__device__ __forceinline__ float bce_term(float p, float y) {
  // torch.nn.BCELoss: -(y*max(log p, -100) + (1-y)*max(log(1-p), -100))
  const float lp = fmaxf(logf(p), -100.f), lq = fmaxf(logf(1.f - p), -100.f);
  return -(y * lp + (1.f - y) * lq);
}

// ---------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------
template <int VEC, bool PAIRWISE>
__global__ void __launch_bounds__(kThreads) score_fwd_kernel(PairArgs a, Workspace ws, int32_t* oob) {
  __shared__ float smem[3 * kWarps];
  const int sub = threadIdx.x & (kLanesPerRow - 1);
  const int64_t group = ((int64_t)blockIdx.x * kThreads + threadIdx.x) / kLanesPerRow;
  const int64_t n_groups = (int64_t)gridDim.x * kThreads / kLanesPerRow;
  const int64_t row_f = (int64_t)a.nv * 4;
  float acc[3] = {0.f, 0.f, 0.f};  // data-loss sum, sum ||u||^2, sum ||i||^2
  // warp-uniform trip count: every lane of a warp runs the same number of iterations (shuffles inside)
  const int64_t warp_first = group - (group % kRowsPerWarp);
  for (int64_t base = warp_first; base < a.batch; base += n_groups) {
    const int64_t b = base + (group % kRowsPerWarp);
    const bool live = b < a.batch;
    int64_t u = 0, ia = 0, ib = 0;
    if (live) {
      u = a.user[b];
      ia = a.item_a[b];
      if (PAIRWISE) ib = a.item_b[b];
    }
    const bool oku = live && (uint64_t)u < (uint64_t)a.n_users;
    const bool oka = live && (uint64_t)ia < (uint64_t)a.n_items;
    const bool okb = PAIRWISE && live && (uint64_t)ib < (uint64_t)a.n_items;
    if (live && oob && sub == 0 && (!oku || !oka || (PAIRWISE && !okb))) *oob = 1;
    const float* pu = a.user_tab + (oku ? u : 0) * row_f;
    const float* pa = a.item_tab + (oka ? ia : 0) * row_f;
    const float* pb = a.item_tab + (okb ? ib : 0) * row_f;
    float4 ru[VEC], ra[VEC], rb[VEC];
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int c = sub + j * kLanesPerRow;
      const bool on = c < a.nv;
      ru[j] = (oku && on) ? ldg_row4(pu, c) : z;
      ra[j] = (oka && on) ? ldg_row4(pa, c) : z;
      if (PAIRWISE) rb[j] = (okb && on) ? ldg_row4(pb, c) : z;
    }
    float da = 0.f, db = 0.f, uu = 0.f, aa = 0.f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      da += dot4(ru[j], ra[j]);
      if (PAIRWISE) db += dot4(ru[j], rb[j]);
      uu += dot4(ru[j], ru[j]);
      aa += dot4(ra[j], ra[j]);
    }
    da = group8_sum(da);
    if (PAIRWISE) db = group8_sum(db);
    uu = group8_sum(uu);
    aa = group8_sum(aa);
    if (live && sub == 0) {
      a.score_a[b] = da;
      float term;
      if (PAIRWISE) {
        a.score_b[b] = db;
        term = -logf(a.gamma + sigmoidf_(da - db));
      } else if (a.loss_kind == XDR_LOSS_MSE) {
        const float d = da - a.label[b];
        term = d * d;
      } else if (a.loss_kind == XDR_LOSS_BCE_SIGMOID) {
        term = bce_term(sigmoidf_(da), a.label[b]);
      } else {
        term = 0.f;
      }
      acc[0] += term;
      acc[1] += uu;
      acc[2] += aa;
    }
  }
  const double n_batch = (double)a.batch;
  const float rw = a.reg_weight;
  float* out8 = a.out8;
  grid_reduce_last_block<3>(acc, ws, smem, [=](double* tot) {
    const float data = (float)(tot[0] / n_batch);
    const float nu = (float)sqrt(tot[1]), ni = (float)sqrt(tot[2]);
    const float reg = (float)(((double)nu + (double)ni) / n_batch);
    out8[0] = data + rw * reg;
    out8[1] = data;
    out8[2] = nu;
    out8[3] = ni;
    out8[4] = reg;
    out8[5] = 0.f;
    out8[6] = 0.f;
    out8[7] = 0.f;
  });
}

// ---------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------
template <int VEC, bool PAIRWISE>
__global__ void __launch_bounds__(kThreads) score_bwd_kernel(PairArgs a) {
  const int sub = threadIdx.x & (kLanesPerRow - 1);
  const int64_t group = ((int64_t)blockIdx.x * kThreads + threadIdx.x) / kLanesPerRow;
  const int64_t n_groups = (int64_t)gridDim.x * kThreads / kLanesPerRow;
  const int64_t row_f = (int64_t)a.nv * 4;
  const float g = (a.grad_loss ? __ldg(a.grad_loss) : 1.0f) * a.scale;
  const float inv_b = 1.0f / (float)a.batch;
  const float nu = __ldg(a.out8 + 2), ni = __ldg(a.out8 + 3);
  // d(reg_weight * (||U||_F + ||I||_F)/B)/dU_r = reg_weight/(B*||U||_F) * U_r   (torch: 0 when the norm is 0)
  const float cu = (a.reg_weight != 0.f && nu > 0.f) ? g * a.reg_weight * inv_b / nu : 0.f;
  const float ci = (a.reg_weight != 0.f && ni > 0.f) ? g * a.reg_weight * inv_b / ni : 0.f;
  for (int64_t b = group; b < a.batch; b += n_groups) {
    const int64_t u = a.user[b], ia = a.item_a[b];
    const int64_t ib = PAIRWISE ? a.item_b[b] : 0;
    const bool oku = (uint64_t)u < (uint64_t)a.n_users;
    const bool oka = (uint64_t)ia < (uint64_t)a.n_items;
    const bool okb = PAIRWISE && (uint64_t)ib < (uint64_t)a.n_items;
    float c;  // g * dL_data/dscore_a  (for BPR: dscore_b = -c)
    if (PAIRWISE) {
      const float s = sigmoidf_(a.score_a[b] - a.score_b[b]);
      c = -g * inv_b * (s * (1.f - s)) / (a.gamma + s);
    } else if (a.loss_kind == XDR_LOSS_MSE) {
      c = g * inv_b * 2.f * (a.score_a[b] - a.label[b]);
    } else if (a.loss_kind == XDR_LOSS_BCE_SIGMOID) {
      const float p = sigmoidf_(a.score_a[b]), y = a.label[b];
      const float pq = p * (1.f - p);
      c = g * inv_b * (p - y) / fmaxf(pq, 1e-12f) * pq;  // BCELoss backward (eps 1e-12) x sigmoid backward
    } else {
      c = 0.f;
    }
    const float* pu = a.user_tab + (oku ? u : 0) * row_f;
    const float* pa = a.item_tab + (oka ? ia : 0) * row_f;
    const float* pb = a.item_tab + (okb ? ib : 0) * row_f;
    float4 ru[VEC], ra[VEC], rb[VEC];
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int cidx = sub + j * kLanesPerRow;
      const bool on = cidx < a.nv;
      // coherent loads: dst may alias the tables (fused SGD), so the non-coherent path is not allowed here
      ru[j] = (oku && on) ? ld_row4(pu, cidx) : z;
      ra[j] = (oka && on) ? ld_row4(pa, cidx) : z;
      if (PAIRWISE) rb[j] = (okb && on) ? ld_row4(pb, cidx) : z;
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int cidx = sub + j * kLanesPerRow;
      if (cidx >= a.nv) continue;
      if (PAIRWISE) {
        if (oku) red_add4(a.user_dst + u * row_f, cidx, axpy4(cu, ru[j], scale4(c, sub4(ra[j], rb[j]))));
        if (oka) red_add4(a.item_dst + ia * row_f, cidx, axpy4(ci, ra[j], scale4(c, ru[j])));
        if (okb) red_add4(a.item_dst + ib * row_f, cidx, scale4(-c, ru[j]));
      } else {
        if (oku) red_add4(a.user_dst + u * row_f, cidx, axpy4(cu, ru[j], scale4(c, ra[j])));
        if (oka) red_add4(a.item_dst + ia * row_f, cidx, axpy4(ci, ra[j], scale4(c, ru[j])));
      }
    }
  }
}

static inline int grid_for_batch(int64_t batch) {
  const int64_t per_block = kThreads / kLanesPerRow;
  int64_t blocks = (batch + per_block - 1) / per_block;
  int64_t cap = (int64_t)sm_count() * 8;
  if (cap > kMaxBlocks) cap = kMaxBlocks;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

static int check_common(const char* fn, const PairArgs& a, int dim, bool pairwise, bool bwd) {
  XDR_REQUIRE(dim_ok(dim), "%s: dim=%d must be a multiple of 4 in (0, 256]", fn, dim);
  XDR_REQUIRE(a.batch > 0, "%s: batch=%lld must be positive", fn, (long long)a.batch);
  XDR_REQUIRE(a.user_tab && a.item_tab && a.user && a.item_a && a.score_a && a.out8, "%s: null pointer", fn);
  XDR_REQUIRE(!pairwise || (a.item_b && a.score_b), "%s: null negative-item pointer", fn);
  XDR_REQUIRE(pairwise || a.loss_kind == XDR_LOSS_NONE || a.label, "%s: label is required for this loss kind", fn);
  XDR_REQUIRE(pairwise || (a.loss_kind >= XDR_LOSS_MSE && a.loss_kind <= XDR_LOSS_NONE), "%s: bad loss_kind %d", fn,
              a.loss_kind);
  XDR_REQUIRE(aligned16(a.user_tab) && aligned16(a.item_tab), "%s: tables must be 16-byte aligned", fn);
  if (bwd) {
    XDR_REQUIRE(a.user_dst && a.item_dst, "%s: null destination", fn);
    XDR_REQUIRE(aligned16(a.user_dst) && aligned16(a.item_dst), "%s: destinations must be 16-byte aligned", fn);
  }
  return XDR_OK;
}

}  // namespace xdr

using namespace xdr;

extern "C" {

int xdr_bpr_fwd(const float* user_tab, const float* item_tab, int64_t n_users, int64_t n_items, int dim,
                const int64_t* user, const int64_t* pos_item, const int64_t* neg_item, int64_t batch, float gamma,
                float reg_weight, float* pos_score, float* neg_score, float* out8, void* ws, int32_t* oob,
                xdr_stream_t stream) {
  PairArgs a{};
  a.user_tab = user_tab; a.item_tab = item_tab; a.n_users = n_users; a.n_items = n_items; a.nv = dim / 4;
  a.user = user; a.item_a = pos_item; a.item_b = neg_item; a.batch = batch; a.gamma = gamma; a.reg_weight = reg_weight;
  a.score_a = pos_score; a.score_b = neg_score; a.out8 = out8;
  int rc = check_common("xdr_bpr_fwd", a, dim, true, false);
  if (rc) return rc;
  XDR_REQUIRE(ws, "xdr_bpr_fwd: null workspace");
  cudaStream_t s = (cudaStream_t)stream;
  XDR_DISPATCH_VEC(a.nv, (XDR_LAUNCH((score_fwd_kernel<VEC, true>), grid_for_batch(batch), kThreads, 0, s, a, Workspace(ws), oob)));
  XDR_LAUNCH_OK();
  return XDR_OK;
}

int xdr_bpr_bwd(const float* user_tab, const float* item_tab, int64_t n_users, int64_t n_items, int dim,
                const int64_t* user, const int64_t* pos_item, const int64_t* neg_item, int64_t batch, float gamma,
                float reg_weight, const float* pos_score, const float* neg_score, const float* out8,
                const float* grad_loss, float scale, float* user_dst, float* item_dst, xdr_stream_t stream) {
  PairArgs a{};
  a.user_tab = user_tab; a.item_tab = item_tab; a.n_users = n_users; a.n_items = n_items; a.nv = dim / 4;
  a.user = user; a.item_a = pos_item; a.item_b = neg_item; a.batch = batch; a.gamma = gamma; a.reg_weight = reg_weight;
  a.score_a = const_cast<float*>(pos_score); a.score_b = const_cast<float*>(neg_score);
  a.out8 = const_cast<float*>(out8); a.grad_loss = grad_loss; a.scale = scale; a.user_dst = user_dst; a.item_dst = item_dst;
  int rc = check_common("xdr_bpr_bwd", a, dim, true, true);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  XDR_DISPATCH_VEC(a.nv, (XDR_LAUNCH((score_bwd_kernel<VEC, true>), grid_for_batch(batch), kThreads, 0, s, a)));
  XDR_LAUNCH_OK();
  return XDR_OK;
}

int xdr_point_fwd(const float* user_tab, const float* item_tab, int64_t n_users, int64_t n_items, int dim,
                  const int64_t* user, const int64_t* item, const float* label, int64_t batch, int loss_kind,
                  float reg_weight, float* score, float* out8, void* ws, int32_t* oob, xdr_stream_t stream) {
  PairArgs a{};
  a.user_tab = user_tab; a.item_tab = item_tab; a.n_users = n_users; a.n_items = n_items; a.nv = dim / 4;
  a.user = user; a.item_a = item; a.label = label; a.batch = batch; a.loss_kind = loss_kind; a.reg_weight = reg_weight;
  a.score_a = score; a.out8 = out8;
  int rc = check_common("xdr_point_fwd", a, dim, false, false);
  if (rc) return rc;
  XDR_REQUIRE(ws, "xdr_point_fwd: null workspace");
  cudaStream_t s = (cudaStream_t)stream;
  XDR_DISPATCH_VEC(a.nv, (XDR_LAUNCH((score_fwd_kernel<VEC, false>), grid_for_batch(batch), kThreads, 0, s, a, Workspace(ws), oob)));
  XDR_LAUNCH_OK();
  return XDR_OK;
}

int xdr_point_bwd(const float* user_tab, const float* item_tab, int64_t n_users, int64_t n_items, int dim,
                  const int64_t* user, const int64_t* item, const float* label, int64_t batch, int loss_kind,
                  float reg_weight, const float* score, const float* out8, const float* grad_loss, float scale,
                  float* user_dst, float* item_dst, xdr_stream_t stream) {
  PairArgs a{};
  a.user_tab = user_tab; a.item_tab = item_tab; a.n_users = n_users; a.n_items = n_items; a.nv = dim / 4;
  a.user = user; a.item_a = item; a.label = label; a.batch = batch; a.loss_kind = loss_kind; a.reg_weight = reg_weight;
  a.score_a = const_cast<float*>(score); a.out8 = const_cast<float*>(out8); a.grad_loss = grad_loss; a.scale = scale;
  a.user_dst = user_dst; a.item_dst = item_dst;
  int rc = check_common("xdr_point_bwd", a, dim, false, true);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  XDR_DISPATCH_VEC(a.nv, (XDR_LAUNCH((score_bwd_kernel<VEC, false>), grid_for_batch(batch), kThreads, 0, s, a)));
  XDR_LAUNCH_OK();
  return XDR_OK;
}

}  // extern "C"
