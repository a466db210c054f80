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
#define XDR_TC5_SELFTEST_IMPL 1
#include "tc5.cuh"
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

// ---- the same scoring on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in tensor memory) ----------
// One CTA = 128 users x a slice of the items.  Warp roles, each looping over the slice's 64-item chunks on its own:
//   warps 5-7  loaders : item rows -> hi / lo TF32 planes in the canonical K-major layout (double-buffered), arrive `full`
//   warp  4    issuer  : one thread: wait `full` + `tempty`, 3 x (D/8) MMAs D[128 x 64] = U I^T (3xTF32), commit -> `empty`, `tfull`
//   warps 0-3  epilogue: thread = user row = TMEM lane: wait `tfull`, tcgen05.ld 16 columns at a time, history mask, keep the
//                        row's k best in a shared-memory list (transposed: [k][128], conflict-free), arrive `tempty`
// The user tile's hi / lo planes are written once.  Shared memory: 2 x 128 x 4D (users) + 4 x 64 x 4D (items) + lists.
constexpr int kT5Users = 128, kT5Items = 64, kT5Threads = 256, kT5LoaderThreads = 96, kT5TmemCols = 128, kT5MaxDim = 64;

__global__ void __launch_bounds__(kT5Threads, 1) topk_score5_kernel(TopkArgs a) {
  XDR_DYN_SMEM_ALIGNED(unsigned char, smem5, 128);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = a.D, k = a.k, nv = D >> 2;
  const tc5::KMajor layA{kT5Users}, layB{kT5Items};
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem5);          // full[2], empty[2], tfull[2], tempty[2]
  uint64_t *full = bars, *empty = bars + 2, *tfull = bars + 4, *tempty = bars + 6;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(smem5 + 64);
  unsigned char* A_hi = smem5 + 128;
  unsigned char* A_lo = A_hi + layA.bytes(D);
  unsigned char* B_hi[2] = {A_lo + layA.bytes(D), A_lo + layA.bytes(D) + 2 * layB.bytes(D)};
  unsigned char* B_lo[2] = {B_hi[0] + layB.bytes(D), B_hi[1] + layB.bytes(D)};
  float* lsc = reinterpret_cast<float*>(B_lo[1] + layB.bytes(D));   // [k][128] scores, descending along k
  int* lid = reinterpret_cast<int*>(lsc + k * kT5Users);            // [k][128] item ids
  const int64_t u0 = (int64_t)blockIdx.x * kT5Users;
  const int sp = blockIdx.y;
  const int64_t lo = a.first_item + (int64_t)sp * a.split_items;
  const int64_t hi = min(a.n_items, lo + a.split_items);
  const int n_chunks = hi > lo ? (int)((hi - lo + kT5Items - 1) / kT5Items) : 0;

  for (int e = tid; e < kT5Users * nv; e += kT5Threads) {
    const int r = e / nv, c = e - r * nv;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (u0 + r < a.B) v = ld_row4(a.U + (u0 + r) * D, c);
    tc5::store_split4(A_hi, A_lo, layA, r, c, v);
  }
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      tc5::mbar_init(&full[i], kT5LoaderThreads);
      tc5::mbar_init(&empty[i], 1);
      tc5::mbar_init(&tfull[i], 1);
      tc5::mbar_init(&tempty[i], kT5Users);
    }
    tc5::mbar_init_fence();
  }
  if (warp == 4) tc5::tmem_alloc(tmem_base_smem, kT5TmemCols);
  tc5::fence_proxy_async();
  tc5::fence_before_sync();
  __syncthreads();
  tc5::fence_after_sync();
  const uint32_t tmem = *tmem_base_smem;

  if (warp >= 5) {
    // ===== loaders =====
    const int lt = tid - 5 * 32;
    for (int i = 0; i < n_chunks; ++i) {
      const int st = i & 1, ph = (i >> 1) & 1;
      tc5::mbar_wait(&empty[st], ph ^ 1);
      const int64_t i0 = lo + (int64_t)i * kT5Items;
      for (int e = lt; e < kT5Items * nv; e += kT5LoaderThreads) {
        const int r = e / nv, c = e - r * nv;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i0 + r < hi) v = ldg_row4(a.I + (i0 + r) * D, c);
        tc5::store_split4(B_hi[st], B_lo[st], layB, r, c, v);
      }
      tc5::fence_proxy_async();
      tc5::mbar_arrive(&full[st]);
    }
  } else if (warp == 4) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      const uint32_t idesc = tc5::make_idesc_tf32(kT5Users, kT5Items, false, false);
      const uint32_t a_hi = tc5::smem_u32(A_hi), a_lo = tc5::smem_u32(A_lo);
      for (int i = 0; i < n_chunks; ++i) {
        const int st = i & 1, ph = (i >> 1) & 1;
        tc5::mbar_wait(&full[st], ph);
        tc5::mbar_wait(&tempty[st], ph ^ 1);
        tc5::fence_after_sync();
        tc5::mma_3xtf32(tmem + st * kT5Items, a_hi, a_lo, layA, tc5::smem_u32(B_hi[st]), tc5::smem_u32(B_lo[st]), layB, idesc, D,
                        false);
        tc5::commit(&empty[st]);   // the item planes may be overwritten once these MMAs have read them
        tc5::commit(&tfull[st]);   // ... and the accumulator is complete
      }
    }
  } else {
    // ===== epilogue / selection: thread = user row =====
    const int u = tid;  // 0..127
    const int64_t row = u0 + u;
    const bool live = row < a.B;
    const int64_t* h = nullptr;
    int64_t hn = 0;
    if (live && a.hist_ptr) {
      const int64_t hb = a.hist_ptr[row];
      hn = a.hist_ptr[row + 1] - hb;
      h = a.hist_ids + hb;
    }
    int n = 0;
    float thr = -INFINITY;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int i = 0; i < n_chunks; ++i) {
      const int st = i & 1, ph = (i >> 1) & 1;
      const int64_t i0 = lo + (int64_t)i * kT5Items;
      tc5::mbar_wait(&tfull[st], ph);
      tc5::fence_after_sync();
#pragma unroll 1
      for (int c16 = 0; c16 < kT5Items / 16; ++c16) {
        uint32_t r[16];
        __syncwarp();  // the per-row insertions below diverge; tcgen05.ld is warp-collective
        tc5::tmem_ld16(tmem + lane_base + st * kT5Items + c16 * 16, r);
        tc5::tmem_ld_wait();
        if (!live) continue;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float sc = __uint_as_float(r[j]);
          const int64_t item = i0 + c16 * 16 + j;
          if (!(sc > thr) || item >= hi) continue;           // also drops NaN
          if (hn > 0 && in_history(h, hn, item)) continue;
          int q = n < k ? n : k - 1;
          while (q > 0 && lsc[(q - 1) * kT5Users + u] < sc) {  // equal scores stay in front: they carry smaller ids
            lsc[q * kT5Users + u] = lsc[(q - 1) * kT5Users + u];
            lid[q * kT5Users + u] = lid[(q - 1) * kT5Users + u];
            --q;
          }
          lsc[q * kT5Users + u] = sc;
          lid[q * kT5Users + u] = (int)item;
          if (n < k) ++n;
          thr = n == k ? lsc[(k - 1) * kT5Users + u] : -INFINITY;
        }
      }
      __syncwarp();
      tc5::fence_before_sync();
      tc5::mbar_arrive(&tempty[st]);
    }
    if (live) {
      for (int q = 0; q < k; ++q) {
        const size_t o = ((size_t)row * a.splits + sp) * k + q;
        a.part_score[o] = q < n ? lsc[q * kT5Users + u] : -INFINITY;
        a.part_id[o] = q < n ? (int64_t)lid[q * kT5Users + u] : (int64_t)-1;
      }
    }
  }
  tc5::fence_before_sync();
  __syncthreads();
  if (warp == 4) tc5::tmem_dealloc(tmem, kT5TmemCols);
}

#endif  // __CUDACC__ || XDR_EMU

static size_t topk_smem_bytes(int D, int k) {
  return sizeof(float) * ((size_t)2 * kTkUsers * (D + 4) + (size_t)kTkUsers * (kTkItems + 4) + (size_t)kTkUsers * k) +
         sizeof(int) * ((size_t)kTkUsers * k + kTkUsers);
}

static int topk_plan(int64_t B, int64_t n_cand, int* splits, int64_t* split_items, int users_per_cta = kTkUsers) {
  const int64_t user_tiles = (B + users_per_cta - 1) / users_per_cta;
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

static size_t topk5_smem_bytes(int D, int k) {
  return 128 + (size_t)2 * kT5Users * D * 4 + (size_t)4 * kT5Items * D * 4 + (size_t)kT5Users * k * 8;
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

// The same contract on tcgen05 (UTCHMMA / TMEM): dim % 8 == 0 and dim <= 64.
int xdr_full_sort_topk_tc5(const float* user_vecs, int64_t batch, const float* item_tab, int64_t n_items, int dim,
                           int64_t first_item, const int64_t* hist_ptr, const int64_t* hist_ids, int k, float* out_score,
                           int64_t* out_id, void* topk_ws, size_t topk_ws_bytes, xdr_stream_t stream) {
  XDR_REQUIRE(dim > 0 && dim % 8 == 0 && dim <= kT5MaxDim, "xdr_full_sort_topk_tc5: dim=%d must be a multiple of 8 in (0, %d]", dim, kT5MaxDim);
  XDR_REQUIRE(batch >= 0 && n_items > 0 && first_item >= 0 && first_item < n_items, "xdr_full_sort_topk_tc5: bad sizes");
  XDR_REQUIRE(k >= 1 && k <= kTkMaxK, "xdr_full_sort_topk_tc5: k=%d must be in [1, %d]", k, kTkMaxK);
  if (batch == 0) return XDR_OK;
  XDR_REQUIRE(user_vecs && item_tab && out_score && out_id && topk_ws, "xdr_full_sort_topk_tc5: null pointer");
  XDR_REQUIRE((hist_ptr == nullptr) == (hist_ids == nullptr), "xdr_full_sort_topk_tc5: hist_ptr and hist_ids go together");
  XDR_REQUIRE(aligned16(user_vecs) && aligned16(item_tab), "xdr_full_sort_topk_tc5: operands must be 16-byte aligned");
  XDR_REQUIRE(topk_ws_bytes >= xdr_topk_workspace_bytes(batch, k), "xdr_full_sort_topk_tc5: workspace too small");
  const size_t smem = topk5_smem_bytes(dim, k);
  XDR_REQUIRE(smem <= 224 * 1024, "xdr_full_sort_topk_tc5: dim/k do not fit shared memory");
  TopkArgs a{};
  a.U = user_vecs; a.I = item_tab; a.B = batch; a.n_items = n_items; a.D = dim; a.k = k; a.first_item = first_item;
  a.hist_ptr = hist_ptr; a.hist_ids = hist_ids;
  topk_plan(batch, n_items - first_item, &a.splits, &a.split_items, kT5Users);
  a.part_score = reinterpret_cast<float*>(topk_ws);
  a.part_id = reinterpret_cast<int64_t*>(reinterpret_cast<char*>(topk_ws) +
                                         (((size_t)batch * kTkMaxSplits * k * sizeof(float) + 15) & ~(size_t)15));
  XDR_CUDA_OK(cudaFuncSetAttribute(topk_score5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaStream_t s = (cudaStream_t)stream;
  dim3 grid((unsigned)((batch + kT5Users - 1) / kT5Users), (unsigned)a.splits);
  XDR_LAUNCH((topk_score5_kernel), grid, kT5Threads, smem, s, a);
  XDR_LAUNCH_OK();
  const int64_t warps_per_block = kTcThreads / 32;
  XDR_LAUNCH((topk_merge_kernel), (unsigned)((batch + warps_per_block - 1) / warps_per_block), kTcThreads, 0, s, a.part_score,
             a.part_id, batch, a.splits, k, out_score, out_id);
  XDR_LAUNCH_OK();
  return XDR_OK;
}

// Self-test of the tcgen05 building blocks: D[128, N] = A[128, K] B[N, K]^T (3xTF32) with each operand staged K-major
// (a_mn / b_mn = 0) or MN-major (= 1) in shared memory.  N % 16 == 0, N <= 256, K % 8 == 0, both planes of both operands
// must fit shared memory.  Device pointers, row-major fp32.
static int tc5_selftest(const char* who, const float* A, const float* B, int N, int K, int a_mn, int b_mn, int fmt, float* D,
                        xdr_stream_t stream) {
  const int kq = fmt ? 16 : 8, esz = fmt ? 2 : 4;
  XDR_REQUIRE(A && B && D, "%s: null pointer", who);
  XDR_REQUIRE(a_mn >= 0 && a_mn <= (fmt ? 2 : 1) && (b_mn == 0 || b_mn == 1), "%s: bad operand major", who);
  XDR_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0 && K >= kq && K % kq == 0, "%s: bad shape N=%d K=%d", who, N, K);
  XDR_REQUIRE(aligned16(A) && aligned16(B), "%s: operands must be 16-byte aligned", who);
  const size_t smem = 128 + (size_t)2 * 128 * K * esz + (size_t)2 * N * K * esz;
  XDR_REQUIRE(smem <= 224 * 1024, "%s: operands do not fit shared memory", who);
  XDR_CUDA_OK(cudaFuncSetAttribute(tc5::selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  XDR_LAUNCH((tc5::selftest_kernel), 1, 128, smem, (cudaStream_t)stream, A, B, N, K, a_mn, b_mn, fmt, D);
  XDR_LAUNCH_OK();
  return XDR_OK;
}

int xdr_tc5_selftest(const float* A, const float* B, int N, int K, int a_mn, int b_mn, float* D, xdr_stream_t stream) {
  // Measured on a B200 (profiles/r2_ubench_tcgen05.txt): with SWIZZLE_NONE, kind::tf32 reproduces the product for K-major
  // operands only -- neither assignment of the two stride fields does for MN-major 32-bit operands (kind::f16 takes both
  // majors).  Nothing in the library stages TF32 operands MN-major; the request is refused instead of answering wrongly.
  if (a_mn != 0 || b_mn != 0) {
    set_error("xdr_tc5_selftest: MN-major TF32 operands are not supported with the SWIZZLE_NONE layouts (measured on B200); "
              "use xdr_tc5_selftest_bf16 for MN-major operands");
    return XDR_ERR_UNSUPPORTED;
  }
  return tc5_selftest("xdr_tc5_selftest", A, B, N, K, a_mn, b_mn, 0, D, stream);
}

// The same product as bf16x3 on kind::f16 (bf16 hi / lo planes, K % 16 == 0): the operand format planned for the tcgen05
// training kernels.
int xdr_tc5_selftest_bf16(const float* A, const float* B, int N, int K, int a_mn, int b_mn, float* D, xdr_stream_t stream) {
  return tc5_selftest("xdr_tc5_selftest_bf16", A, B, N, K, a_mn, b_mn, 1, D, stream);
}

}  // extern "C"
