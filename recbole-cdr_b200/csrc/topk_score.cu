// topk_score.cu -- section 8 F2: full-sort scoring fused with history masking and a streaming top-k.
//
// Replaces, for a block of evaluation users, full_sort_predict's dense score matrix (reference emcdr.py:208-233,
// cmf.py:107-112: torch.matmul(user_e, all_item_e.T) -> [B, n_items] fp32) PLUS what recbole's full-sort evaluation does
// with it next [recbole-1.0.1 FullSortEvalDataLoader / Collector]: scores[:, 0] = -inf (the PAD item), scores[u, history of u]
// = -inf, torch.topk(scores, k).  The [B, n_items] matrix (4 MB per user at 10^6 items) is never written: a CTA owns 64 users
// and a contiguous slice of the items, streams the slice through shared memory 64 rows at a time, scores each 64 x 64 block
// on the tensor cores (3xTF32 mma.sync, fp32-equivalent, tc_tile.cuh), masks, and keeps each user's k best in shared memory.
// A second kernel merges the per-slice lists (one warp per user, one lane per slice).
// Order: score descending, ties by ascending item id (torch.topk leaves tie order unspecified).
// Algorithmic bytes: the item table once per 64-user block (n_items * 4D) + B * 4D + B * k * 12;  2 * B * n_items * D FLOP.
#include "tc_tile.cuh"

namespace xdr {

constexpr int kTkUsers = 64;   // users per CTA (the MMA row tile)
constexpr int kTkItems = 64;   // items per streamed chunk
constexpr int kTkMaxK = 128;
constexpr int kTkMaxSplits = 32;

struct TopkArgs {
  const float* U;          // [B][D] user-side vectors (already gathered / mapped)
  const float* I;          // [n_items][D] item table
  int64_t B, n_items;
  int D, k;
  int64_t first_item;      // items [first_item, n_items) are candidates (1 skips the PAD row)
  const int64_t* hist_ptr; // [B + 1] CSR offsets into hist_ids, or NULL (no masking)
  const int64_t* hist_ids; // per user ascending item ids to exclude
  int splits;
  int64_t split_items;     // items per slice (multiple of kTkItems)
  float* part_score;       // [B][splits][k]
  int64_t* part_id;        // [B][splits][k]
};

#if defined(__CUDACC__) || defined(XDR_EMU)

// lower_bound over a user's ascending history: true iff `id` is in it
__device__ __forceinline__ bool in_history(const int64_t* __restrict__ h, int64_t n, int64_t id) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (h[mid] < id) lo = mid + 1;
    else hi = mid;
  }
  return lo < n && h[lo] == id;
}

__global__ void __launch_bounds__(kTcThreads, 1) topk_score_kernel(TopkArgs a) {
  XDR_DYN_SMEM(float, smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int D = a.D, ld = D + 4, k = a.k;
  float* Us = smem;                                   // [64][ld]
  float* Is = Us + kTkUsers * ld;                     // [64][ld]
  float* S = Is + kTkItems * ld;                      // [64][64 + 4]
  constexpr int lds = kTkItems + 4;
  float* lsc = S + kTkUsers * lds;                    // [64][k] scores, descending
  int* lid = reinterpret_cast<int*>(lsc + kTkUsers * k);  // [64][k] item ids
  int* cnt = lid + kTkUsers * k;                      // [64]
  const int64_t u0 = (int64_t)blockIdx.x * kTkUsers;
  const int sp = blockIdx.y;
  const int64_t lo = a.first_item + (int64_t)sp * a.split_items;
  const int64_t hi = min(a.n_items, lo + a.split_items);
  const int nv = D >> 2;

  for (int e = tid; e < kTkUsers * nv; e += kTcThreads) {
    const int r = e / nv, c = e - r * nv;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (u0 + r < a.B) v = ld_row4(a.U + (u0 + r) * D, c);
    *reinterpret_cast<float4*>(Us + r * ld + 4 * c) = v;
  }
  for (int r = tid; r < kTkUsers; r += kTcThreads) cnt[r] = 0;
  __syncthreads();

  for (int64_t i0 = lo; i0 < hi; i0 += kTkItems) {
    for (int e = tid; e < kTkItems * nv; e += kTcThreads) {
      const int r = e / nv, c = e - r * nv;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i0 + r < hi) v = ldg_row4(a.I + (i0 + r) * D, c);
      *reinterpret_cast<float4*>(Is + r * ld + 4 * c) = v;
    }
    __syncthreads();
    tile_gemm<kTkUsers, kTkItems / 16, false>(Us, ld, Is, ld, kTkItems, D, [&](int row, int col, float v0, float v1) {
      *reinterpret_cast<float2*>(S + row * lds + col) = make_float2(v0, v1);
    });
    __syncthreads();
    // ---- selection: warp w owns users 8w .. 8w+7; lanes cover the chunk's items (lane, lane + 32) ---------------------
    for (int uu = 0; uu < kTkUsers / kTcWarps; ++uu) {
      const int u = warp * (kTkUsers / kTcWarps) + uu;
      const int64_t row = u0 + u;
      if (row >= a.B) continue;  // warp-uniform
      const int64_t* h = nullptr;
      int64_t hn = 0;
      if (a.hist_ptr) {
        const int64_t hb = a.hist_ptr[row];
        hn = a.hist_ptr[row + 1] - hb;
        h = a.hist_ids + hb;
      }
      float* usc = lsc + u * k;
      int* uid = lid + u * k;
      int n = cnt[u];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int j = lane + 32 * half;
        const int64_t item = i0 + j;
        float s = S[u * lds + j];
        bool ok = item < hi && !(s != s);
        if (ok && hn > 0 && in_history(h, hn, item)) ok = false;
        const float thr = n == k ? usc[k - 1] : -INFINITY;
        unsigned bits = __ballot_sync(0xffffffffu, ok && s > thr);
        while (bits) {  // ascending item order; every lane follows the same sequence
          const int src = __ffs(bits) - 1;
          bits &= bits - 1;
          const float cs = __shfl_sync(0xffffffffu, s, src);
          const int cid = (int)(i0 + src + 32 * half);
          if (n == k && !(cs > usc[k - 1])) continue;  // the threshold moved since the ballot
          // position: entries with a score >= cs stay in front (equal scores carry smaller ids: items ascend)
          int ge = 0;
          for (int q = lane; q < n; q += 32) ge += usc[q] >= cs ? 1 : 0;
          const int p = (int)warp_sum((float)ge);
          const int new_n = n < k ? n + 1 : k;
          // shift [p, new_n - 1) one to the right: read, sync, write
          float sh_s[kTkMaxK / 32];
          int sh_i[kTkMaxK / 32];
#pragma unroll
          for (int t = 0; t < kTkMaxK / 32; ++t) {
            const int q = lane + 32 * t;
            if (q > p && q < new_n) { sh_s[t] = usc[q - 1]; sh_i[t] = uid[q - 1]; }
          }
          __syncwarp();
#pragma unroll
          for (int t = 0; t < kTkMaxK / 32; ++t) {
            const int q = lane + 32 * t;
            if (q > p && q < new_n) { usc[q] = sh_s[t]; uid[q] = sh_i[t]; }
          }
          if (lane == 0) { usc[p] = cs; uid[p] = cid; }
          __syncwarp();
          n = new_n;
        }
      }
      if (lane == 0) cnt[u] = n;
    }
    __syncthreads();
  }
  // ---- publish this slice's lists ---------------------------------------------------------------------------------------------
  for (int e = tid; e < kTkUsers * k; e += kTcThreads) {
    const int u = e / k, q = e - u * k;
    const int64_t row = u0 + u;
    if (row >= a.B) continue;
    const bool have = q < cnt[u];
    const size_t o = ((size_t)row * a.splits + sp) * k + q;
    a.part_score[o] = have ? lsc[u * k + q] : -INFINITY;
    a.part_id[o] = have ? (int64_t)lid[u * k + q] : (int64_t)-1;
  }
}

// one warp per user, lane s walks slice s's sorted list: k rounds of a warp arg-best over the slice heads
__global__ void __launch_bounds__(kTcThreads) topk_merge_kernel(const float* __restrict__ part_score,
                                                                 const int64_t* __restrict__ part_id, int64_t B, int splits,
                                                                 int k, float* __restrict__ out_score,
                                                                 int64_t* __restrict__ out_id) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= B) return;  // warp-uniform
  int head = 0;
  const size_t base = ((size_t)row * splits + lane) * k;
  for (int q = 0; q < k; ++q) {
    float s = -INFINITY;
    int64_t id = -1;
    if (lane < splits && head < k) { s = part_score[base + head]; id = part_id[base + head]; }
    int who = id >= 0 ? lane : -1;
    // arg-best: higher score, then lower id; empty heads (id < 0) lose
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float s2 = __shfl_xor_sync(0xffffffffu, s, o);
      const int64_t id2 = __shfl_xor_sync(0xffffffffu, id, o);
      const int who2 = __shfl_xor_sync(0xffffffffu, who, o);
      const bool take = who2 >= 0 && (who < 0 || s2 > s || (s2 == s && id2 < id));
      if (take) { s = s2; id = id2; who = who2; }
    }
    if (lane == 0) { out_score[row * k + q] = who >= 0 ? s : -INFINITY; out_id[row * k + q] = who >= 0 ? id : (int64_t)-1; }
    if (lane == who) ++head;
  }
}

#endif  // __CUDACC__ || XDR_EMU

static size_t topk_smem_bytes(int D, int k) {
  return sizeof(float) * ((size_t)2 * kTkUsers * (D + 4) + (size_t)kTkUsers * (kTkItems + 4) + (size_t)kTkUsers * k) +
         sizeof(int) * ((size_t)kTkUsers * k + kTkUsers);
}

static int topk_plan(int64_t B, int64_t n_cand, int* splits, int64_t* split_items) {
  const int64_t user_tiles = (B + kTkUsers - 1) / kTkUsers;
  int64_t want = (2 * (int64_t)sm_count() + user_tiles - 1) / user_tiles;  // ~2 CTAs per SM in total
  const int64_t chunks = (n_cand + kTkItems - 1) / kTkItems;
  if (want > chunks) want = chunks;
  if (want > kTkMaxSplits) want = kTkMaxSplits;
  if (want < 1) want = 1;
  const int64_t per = ((chunks + want - 1) / want) * kTkItems;
  *split_items = per;
  *splits = (int)((n_cand + per - 1) / per);
  return XDR_OK;
}

}  // namespace xdr

using namespace xdr;

extern "C" {

// bytes of scratch for the per-slice lists: [B][splits <= 32][k] (fp32 score + int64 id)
size_t xdr_topk_workspace_bytes(int64_t batch, int k) {
  return (size_t)batch * kTkMaxSplits * k * (sizeof(float) + sizeof(int64_t)) + 16;
}

int xdr_full_sort_topk(const float* user_vecs, int64_t batch, const float* item_tab, int64_t n_items, int dim,
                       int64_t first_item, const int64_t* hist_ptr, const int64_t* hist_ids, int k, float* out_score,
                       int64_t* out_id, void* topk_ws, size_t topk_ws_bytes, xdr_stream_t stream) {
  XDR_REQUIRE(dim > 0 && dim % 8 == 0 && dim <= 256, "xdr_full_sort_topk: dim=%d must be a multiple of 8 in (0, 256]", dim);
  XDR_REQUIRE(batch >= 0 && n_items > 0 && first_item >= 0 && first_item < n_items, "xdr_full_sort_topk: bad sizes");
  XDR_REQUIRE(k >= 1 && k <= kTkMaxK, "xdr_full_sort_topk: k=%d must be in [1, %d]", k, kTkMaxK);
  if (batch == 0) return XDR_OK;
  XDR_REQUIRE(user_vecs && item_tab && out_score && out_id && topk_ws, "xdr_full_sort_topk: null pointer");
  XDR_REQUIRE((hist_ptr == nullptr) == (hist_ids == nullptr), "xdr_full_sort_topk: hist_ptr and hist_ids go together");
  XDR_REQUIRE(aligned16(user_vecs) && aligned16(item_tab), "xdr_full_sort_topk: operands must be 16-byte aligned");
  XDR_REQUIRE(topk_ws_bytes >= xdr_topk_workspace_bytes(batch, k), "xdr_full_sort_topk: workspace too small");
  XDR_REQUIRE(topk_smem_bytes(dim, k) <= 224 * 1024, "xdr_full_sort_topk: dim/k do not fit shared memory");
  TopkArgs a{};
  a.U = user_vecs; a.I = item_tab; a.B = batch; a.n_items = n_items; a.D = dim; a.k = k; a.first_item = first_item;
  a.hist_ptr = hist_ptr; a.hist_ids = hist_ids;
  topk_plan(batch, n_items - first_item, &a.splits, &a.split_items);
  a.part_score = reinterpret_cast<float*>(topk_ws);
  a.part_id = reinterpret_cast<int64_t*>(reinterpret_cast<char*>(topk_ws) +
                                         (((size_t)batch * kTkMaxSplits * k * sizeof(float) + 15) & ~(size_t)15));
  const size_t smem = topk_smem_bytes(dim, k);
  XDR_CUDA_OK(cudaFuncSetAttribute(topk_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaStream_t s = (cudaStream_t)stream;
  dim3 grid((unsigned)((batch + kTkUsers - 1) / kTkUsers), (unsigned)a.splits);
  XDR_LAUNCH((topk_score_kernel), grid, kTcThreads, smem, s, a);
  XDR_LAUNCH_OK();
  const int64_t warps_per_block = kTcThreads / 32;
  XDR_LAUNCH((topk_merge_kernel), (unsigned)((batch + warps_per_block - 1) / warps_per_block), kTcThreads, 0, s, a.part_score,
             a.part_id, batch, a.splits, k, out_score, out_id);
  XDR_LAUNCH_OK();
  return XDR_OK;
}

}  // extern "C"
