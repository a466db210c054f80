// emu_kernels.cpp -- TEST INFRASTRUCTURE ONLY.  The kernel sources of libxdr compile with g++ against the CTA emulator
// (cuda_emu.h, -DXDR_EMU): every .cu listed in tests/emu_util.py becomes a translation unit of tests/emu/_build/
// libxdr_emu.so, whose xdr_* entry points are the REAL host entry points (argument checks, launch geometry) with
// XDR_LAUNCH running the CTAs on the CPU -- they take host pointers.  This file adds the two host hooks the library expects
// (set_error, sm_count) and a few direct kernel drivers (forced tile sizes) for tests/test_emu_*.py.
#define XDR_EMU 1
#include "../../recbole-cdr_b200/csrc/fused_mlp.cu"
#include "../../recbole-cdr_b200/csrc/tc_mlp.cu"
#include "../../recbole-cdr_b200/csrc/tc_conet.cu"
#include "../../recbole-cdr_b200/csrc/sparse_optim.cu"
#include "../../recbole-cdr_b200/csrc/tc5.cuh"

#include <cstdarg>

namespace xdr {
static int g_sms = 4;
static char g_err[512];
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int sm_count() { return g_sms; }
}  // namespace xdr

using namespace xdr;

static void fill_mlp_args(MlpArgs& a, int n_layers, const int* dims, const float* const* W, const float* const* b,
                          float* const* dW, float* const* db, int hidden_act, int in_mode, int head, const float* Au,
                          const float* Bu, const float* Ai, const float* Bi, const float* T, int64_t n_u, int64_t n_i, int dim,
                          const int64_t* idx_u, const int64_t* idx_i, const float* label, int64_t batch, int backward,
                          const float* grad_loss, float scale, float* dAu, float* dBu, float* dAi, float* dBi, float* dT,
                          float* prob, float* out8, int32_t* oob) {
  a.n_layers = n_layers;
  for (int l = 0; l <= n_layers; ++l) a.dims[l] = dims[l];
  for (int l = 0; l < n_layers; ++l) {
    a.W[l] = W[l];
    a.b[l] = b ? b[l] : nullptr;
    a.dW[l] = (backward && dW) ? dW[l] : nullptr;
    a.db[l] = (backward && db) ? db[l] : nullptr;
  }
  a.hidden_act = hidden_act; a.last_act = XDR_ACT_NONE; a.in_mode = in_mode; a.head = head;
  a.Au = Au; a.Bu = Bu; a.Ai = Ai; a.Bi = Bi; a.T = T; a.n_u = n_u; a.n_i = n_i; a.dim = dim;
  a.idx_u = idx_u; a.idx_i = idx_i; a.label = label; a.batch = batch; a.backward = backward; a.grad_loss = grad_loss;
  a.scale = scale; a.dAu = dAu; a.dBu = dBu; a.dAi = dAi; a.dBi = dBi; a.dT = dT; a.prob = prob; a.out8 = out8; a.oob = oob;
}

extern "C" {

void emu_config(int sms, uint64_t schedule_seed) {
  g_sms = sms;
  emu::set_schedule_seed(schedule_seed);
}
const char* emu_last_error() { return g_err; }
// several persistent launches ("GPUs") resident together: queue them between begin and run
// work counters since the last reset: [mma.sync tf32, mma.sync bf16, tcgen05.mma tf32, row-load bytes, row-RED bytes, CTA barriers,
// tcgen05.mma f16]
void emu_counters(uint64_t* out6, int reset) {
  emu::Counters& c = emu::counters();
  out6[0] = c.mma_tf32; out6[1] = c.mma_bf16; out6[2] = c.umma_tf32; out6[3] = c.row_load_bytes; out6[4] = c.row_red_bytes;
  out6[5] = c.cta_barriers;
  out6[6] = c.umma_bf16;
  if (reset) c = emu::Counters();
}
void emu_group_begin() { emu::group_begin(); }
void emu_group_run() { emu::group_run(); }

// impl 0: fused_mlp_kernel (fp32 FMA, validated on hardware -- run here to validate the emulator itself)
// impl 1: tc_mlp_kernel (tensor-core tiles)
int emu_mlp_step(int impl, int n_layers, const int* dims, const float* const* W, const float* const* b, float* const* dW,
                 float* const* db, int hidden_act, int in_mode, int head, const float* Au, const float* Bu, const float* Ai,
                 const float* Bi, const float* T, int64_t n_u, int64_t n_i, int dim, const int64_t* idx_u,
                 const int64_t* idx_i, const float* label, int64_t batch, int backward, const float* grad_loss, float scale,
                 float* dAu, float* dBu, float* dAi, float* dBi, float* dT, float* prob, float* out8, void* ws, int32_t* oob,
                 int force_tile_rows) {
  MlpArgs a{};
  fill_mlp_args(a, n_layers, dims, W, b, dW, db, hidden_act, in_mode, head, Au, Bu, Ai, Bi, T, n_u, n_i, dim, idx_u, idx_i,
                label, batch, backward, grad_loss, scale, dAu, dBu, dAi, dBi, dT, prob, out8, oob);
  if (impl == 0) {
    if (!pick_tile_rows(&a)) return -3;
    if (force_tile_rows) a.tile_rows = force_tile_rows;
    const size_t smem = mlp_smem_bytes(a);
    const int64_t n_tiles = (a.batch + a.tile_rows - 1) / a.tile_rows;
    const unsigned grid = (unsigned)std::min<int64_t>(g_sms, n_tiles);
    emu::launch(grid, kMlpThreads, smem, [&] { fused_mlp_kernel<32, 32, 32>(a, Workspace(ws)); });
    return 0;
  }
  MlpArgs chk{};
  if (!tc_stack_ok(n_layers, dims, &chk)) return -3;
  int tr = tc_pick_tile_rows(a, batch);
  if (force_tile_rows) tr = force_tile_rows;
  if (tr == 0) return -3;
  a.tile_rows = tr;
  const size_t smem = tc_smem_bytes(a, tr);
  const int64_t n_tiles = (a.batch + tr - 1) / tr;
  const unsigned grid = (unsigned)std::min<int64_t>(g_sms, n_tiles);
  if (tr == 64) emu::launch(grid, kTcThreads, smem, [&] { tc_mlp_kernel<64, kTcDw0, kTcDw1, kTcDw2>(a, Workspace(ws)); });
  else emu::launch(grid, kTcThreads, smem, [&] { tc_mlp_kernel<32, kTcDw0, kTcDw1, kTcDw2>(a, Workspace(ws)); });
  return 0;
}

int emu_tc_mlp_supported(int n_layers, const int* dims) {
  MlpArgs a{};
  if (!tc_stack_ok(n_layers, dims, &a)) return 0;
  return tc_pick_tile_rows(a, 0) != 0;
}

int emu_conet_supported(int n_layers, const int* dims, int dim) { return conet_stack_ok(n_layers, dims, dim) ? 1 : 0; }

int emu_conet_step(int n_layers, const int* dims, const float* const* Ws, const float* const* bs, const float* const* Wt,
                   const float* const* bt, const float* const* H, float* const* dWs, float* const* dbs, float* const* dWt,
                   float* const* dbt, float* const* dH, const float* w_out, const float* b_out, float* dw_out, float* db_out,
                   int want, const float* Su, const float* Si, const float* Tu, const float* Ti, int64_t n_u, int64_t n_i,
                   int dim, const int64_t* user, const int64_t* item, const float* label, int64_t batch, int mask_on_item,
                   int64_t n_overlap, int backward, const float* grad_loss, float scale, float* dSu, float* dSi, float* dTu,
                   float* dTi, float* dz1, float* prob, float* out8, void* ws, int32_t* oob) {
  ConetArgs a{};
  const int rc = conet_make_args(&a, n_layers, dims, Ws, bs, Wt, bt, H, dWs, dbs, dWt, dbt, dH, w_out, b_out, dw_out, db_out,
                                 want, Su, Si, Tu, Ti, n_u, n_i, dim, user, item, label, batch, mask_on_item, n_overlap,
                                 backward, grad_loss, scale, dSu, dSi, dTu, dTi, dz1, prob, out8, ws, oob);
  if (rc != XDR_OK) return rc;
  const size_t smem = (size_t)conet_smem_layout(n_layers, a.dims).total * sizeof(float);
  const int64_t n_tiles = (batch + kCnTR - 1) / kCnTR;
  const unsigned grid = (unsigned)std::min<int64_t>(g_sms, n_tiles);
  emu::launch(grid, kTcThreads, smem, [&] { tc_conet_kernel(a, Workspace(ws)); });
  return 0;
}


// ---- a deliberately breakable tcgen05 kernel: proves that the emulator's asynchronous model catches protocol mistakes -------------
// D[128 x 16] = A[128 x 16] B[16 x 16]^T (bf16 K-major planes, one kind::f16 MMA).  fault: 0 none; 1 tensor memory read without
// waiting for the commit; 2 operand stores not followed by fence.proxy.async; 3 tcgen05.ld results used before tcgen05.wait::ld;
// 4 the A plane refilled (stores + fence) right after the issue, before the MMA has run.
static void async_probe_kernel(const float* A, const float* B, float* D, int fault) {
  XDR_DYN_SMEM_ALIGNED(unsigned char, sm, 128);
  const int tid = threadIdx.x, warp = tid >> 5;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm);
  uint32_t* tbase = reinterpret_cast<uint32_t*>(sm + 8);
  unsigned char *ah = sm + 128, *al = ah + 128 * 16 * 2, *bh = al + 128 * 16 * 2, *bl = bh + 16 * 16 * 2;
  const tc5::KMajor16 ka{128}, kb{16};
  if (fault == 2) tc5::fence_proxy_async();    // a fence BEFORE the stores does not publish them
  for (int e = tid; e < 128 * 2; e += 128)
    tc5::store_split8(ah, al, ka.chunk_offset(e / 2, e % 2), *reinterpret_cast<const float4*>(A + (e / 2) * 16 + 8 * (e % 2)),
                      *reinterpret_cast<const float4*>(A + (e / 2) * 16 + 8 * (e % 2) + 4));
  for (int e = tid; e < 16 * 2; e += 128)
    tc5::store_split8(bh, bl, kb.chunk_offset(e / 2, e % 2), *reinterpret_cast<const float4*>(B + (e / 2) * 16 + 8 * (e % 2)),
                      *reinterpret_cast<const float4*>(B + (e / 2) * 16 + 8 * (e % 2) + 4));
  if (tid == 0) { tc5::mbar_init(bar, 1); tc5::mbar_init_fence(); }
  if (warp == 0) tc5::tmem_alloc(tbase, 32);
  if (fault != 2) tc5::fence_proxy_async();
  tc5::fence_before_sync();
  __syncthreads();
  tc5::fence_after_sync();
  const uint32_t tmem = *tbase;
  if (tid == 0) {
    tc5::mma_bf16x3(tmem, tc5::smem_u32(ah), tc5::smem_u32(al), ka, tc5::smem_u32(bh), tc5::smem_u32(bl), kb,
                    tc5::make_idesc_bf16(128, 16, false, false), 16, false);
    tc5::commit(bar);
    if (fault == 4) {   // the buffer is refilled (stores + fence, as a real reuse would do) without waiting for the MMA that reads it
      std::memset(ah, 0, 128 * 16 * 2);
      tc5::fence_proxy_async();
    }
  }
  if (fault != 1) tc5::mbar_wait(bar, 0);
  tc5::fence_after_sync();
  uint32_t r[16];
  tc5::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), r);
  if (fault != 3) tc5::tmem_ld_wait();
  for (int j = 0; j < 16; ++j) D[tid * 16 + j] = __uint_as_float(r[j]);
  tc5::tmem_ld_wait();
  if (fault == 1) tc5::mbar_wait(bar, 0);
  tc5::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc5::tmem_dealloc(tmem, 32);
}

int emu_async_probe(const float* A, const float* B, float* D, int fault) {
  const size_t smem = 128 + 2 * 128 * 16 * 2 + 2 * 16 * 16 * 2;
  emu::launch(1, 128, smem, [&] { async_probe_kernel(A, B, D, fault); });
  return 0;
}

}  // extern "C"
