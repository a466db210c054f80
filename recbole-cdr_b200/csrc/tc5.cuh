// tc5.cuh -- tcgen05 (5th-generation tensor core) building blocks for kernels that write their own operands to shared
// memory: SWIZZLE_NONE canonical layouts + hand-built matrix descriptors, TMEM allocation, single-thread MMA issue with
// mbarrier completion, TMEM -> register loads.  kind::tf32 with fp32 accumulation; products that must be fp32-equivalent
// are issued as three MMAs on hi / lo operand copies (3xTF32, as in tc_tile.cuh).
//
// Layouts (units of 16 bytes = 4 tf32; cute/atom/mma_traits_sm100.hpp "make_umma_desc"):
//   K-major operand of R rows:  element (r, k) at  (k/4)*LBO + (r/8)*SBO + (r%8)*16 + (k%4)*4   with LBO = R*16, SBO = 128
//   -- a core matrix is 8 rows x 16 bytes, stored contiguously (128 B); one MMA consumes K = 8 = two 16-byte K chunks.
// Descriptor (cute/arch/mma_sm100_desc.hpp): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version 1 [46,48) | layout [61,64) = 0.
// Instruction descriptor: F32 accumulate (1 << 4) | TF32 A (2 << 7) | TF32 B (2 << 10) | a_major [15] | b_major [16] | N>>3 [17,23) | M>>4 [24,29).
//
// STATUS: written in round 1 without GPU access as the repository's READING of the interface; confirmed on a B200 in round 2 by
// scripts/ubench_tcgen05.cu (profiles/r2_ubench_tcgen05.txt: kind::tf32 with both operands K-major and every kind::f16 major
// combination PASS with LBO / SBO as read here; kind::tf32 does NOT take MN-major operands -- the kernels that need an MN-major
// view use bf16 planes) and by the GPU parity tests of every kernel built on it.  The CPU emulator (tests/emu) models the same
// reading; what its tests prove is the logic AROUND the MMAs (pipelines, barriers, epilogues).
#pragma once
#include "xdr_common.cuh"

namespace xdr {
namespace tc5 {

#if defined(__CUDACC__) || defined(XDR_EMU)

// ---- operand layout -------------------------------------------------------------------------------------------------------
struct KMajor {
  int rows;  // R (multiple of 8)
  __host__ __device__ int lbo() const { return rows * 16; }
  __host__ __device__ int sbo() const { return 128; }
  __host__ __device__ int bytes(int k) const { return rows * k * 4; }           // footprint of an R x k operand
  __host__ __device__ int chunk_offset(int r, int k4) const {                   // byte offset of the 16-byte chunk (r, 4*k4..)
    return k4 * lbo() + (r >> 3) * sbo() + (r & 7) * 16;
  }
  __host__ __device__ int k_step_bytes() const { return 2 * lbo(); }            // advance of the start address per MMA (K = 8)
  __host__ __device__ bool mn_major() const { return false; }
};

// MN-major operand of R rows (R % 4 == 0) and K reduction elements (K % 8 == 0): contiguous along the row index --
// what a weight gradient dZ^T X needs, whose reduction index is the batch row.
//   element (r, k) at  (k/8)*LBO + (r/4)*SBO + (k%8)*16 + (r%4)*4   with SBO = 128, LBO = (R/4)*128
//   -- a core matrix is 8 K-rows x 16 bytes (4 consecutive r); one MMA consumes K = 8 = one LBO step.
struct MNMajor {
  int rows;
  __host__ __device__ int lbo() const { return (rows >> 2) * 128; }
  __host__ __device__ int sbo() const { return 128; }
  __host__ __device__ int bytes(int k) const { return rows * k * 4; }
  __host__ __device__ int chunk_offset(int r4, int k) const {                   // 16-byte chunk of rows 4*r4.. at reduction k
    return (k >> 3) * lbo() + r4 * sbo() + (k & 7) * 16;
  }
  __host__ __device__ int k_step_bytes() const { return lbo(); }
  __host__ __device__ bool mn_major() const { return true; }
};

// 16-bit operands (bf16 on kind::f16): the same canonical layouts with 8 elements per 16-byte chunk; one MMA consumes
// K = 16 = two 16-byte K chunks (K-major) or two groups of eight reduction rows (MN-major).  A bf16 hi plane + a bf16 lo
// plane (x ~= hi + lo, bf16x3) together take 4 bytes per element -- the footprint of the fp32 data itself, half of the TF32
// hi / lo pair -- at twice the MMA rate: the operand format of the tcgen05 training kernels.
//   K-major : element (r, k) at (k/8)*LBO + (r/8)*SBO + (r%8)*16 + (k%8)*2      LBO = R*16,        SBO = 128
//   MN-major: element (r, k) at (k/8)*LBO + (r/8)*SBO + (k%8)*16 + (r%8)*2      LBO = (R/8)*128,   SBO = 128
struct KMajor16 {
  int rows;  // R (multiple of 8)
  __host__ __device__ int lbo() const { return rows * 16; }
  __host__ __device__ int sbo() const { return 128; }
  __host__ __device__ int bytes(int k) const { return rows * k * 2; }
  __host__ __device__ int chunk_offset(int r, int k8) const { return k8 * lbo() + (r >> 3) * sbo() + (r & 7) * 16; }
  __host__ __device__ int k_step_bytes() const { return 2 * lbo(); }            // K = 16 per MMA
  __host__ __device__ bool mn_major() const { return false; }
};
struct MNMajor16 {
  int rows;  // R (multiple of 8)
  __host__ __device__ int lbo() const { return (rows >> 3) * 128; }
  __host__ __device__ int sbo() const { return 128; }
  __host__ __device__ int bytes(int k) const { return rows * k * 2; }
  __host__ __device__ int chunk_offset(int r8, int k) const { return (k >> 3) * lbo() + r8 * sbo() + (k & 7) * 16; }
  __host__ __device__ int k_step_bytes() const { return 2 * lbo(); }            // K = 16 = two groups of eight reduction rows
  __host__ __device__ bool mn_major() const { return true; }
};

// ONE tile, two views.  The physical layout of MNMajor16 over X[k = row][n = column] stores the 8 x 8 core matrices row-block
// by row-block.  Read with the two stride fields exchanged in meaning it is also a K-major operand (m = row, k = column):
// the next 16 bytes of K are the next core matrix (LBO = 128), the next eight rows a whole row of core matrices further
// (SBO = (C/8)*128).  So an activation tile written once feeds both the forward / input-gradient product (K-major A) and the
// weight-gradient product dZ^T X (MN-major B) -- no second, transposed copy in shared memory.  Needs both strides to be free
// parameters of the descriptor: the "Kb" rows of scripts/ubench_tcgen05.cu check that on hardware.
struct RowBlock16 {
  int rows, cols;  // tile shape, both multiples of 8; bf16 elements
  __host__ __device__ int bytes() const { return rows * cols * 2; }
  __host__ __device__ int chunk_offset(int r, int c8) const { return (r >> 3) * (cols >> 3) * 128 + c8 * 128 + (r & 7) * 16; }
  struct View {
    int lbo_, sbo_, kstep_;
    bool mn_;
    __host__ __device__ int lbo() const { return lbo_; }
    __host__ __device__ int sbo() const { return sbo_; }
    __host__ __device__ int k_step_bytes() const { return kstep_; }
    __host__ __device__ bool mn_major() const { return mn_; }
  };
  __host__ __device__ View as_k_major() const { return View{128, (cols >> 3) * 128, 256, false}; }               // (m = row, k = col)
  __host__ __device__ View as_mn_major() const { return View{(cols >> 3) * 128, 128, 2 * (cols >> 3) * 128, true}; }  // (n = col, k = row)
};

// XDR_TC5_SWAP (compile-time, -DXDR_TC5_SWAP=n): which descriptor field carries which stride is the one part of this reading
// that only hardware can settle (scripts/ubench_tcgen05.cu prints the answer).  bit 0 exchanges the two fields for K-major
// operands, bit 1 for MN-major operands; scripts/r2_gpu_session.sh rebuilds with the other settings when the self-test of the
// default reading fails, so the answer costs no extra GPU round trip.
#ifndef XDR_TC5_SWAP
#define XDR_TC5_SWAP 0
#endif
__host__ __device__ inline uint64_t make_desc(uint32_t smem_addr, int lbo, int sbo, bool mn_major = false) {
  if ((XDR_TC5_SWAP >> (mn_major ? 1 : 0)) & 1) {
    const int t = lbo;
    lbo = sbo;
    sbo = t;
  }
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__host__ __device__ inline uint32_t make_idesc_tf32(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__host__ __device__ inline uint32_t make_idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {   // kind::f16, BF16 x BF16 -> F32
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- PTX wrappers (emulator twins under XDR_EMU) --------------------------------------------------------------------------
#ifdef XDR_EMU
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return emu::smem_addr(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { emu::mbar_init(bar, count); }
__device__ __forceinline__ void mbar_init_fence() {}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { emu::mbar_arrive(bar); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { emu::mbar_wait(bar, parity); }
__device__ __forceinline__ void fence_proxy_async() { emu::proxy_fence(); }
__device__ __forceinline__ void fence_before_sync() {}
__device__ __forceinline__ void fence_after_sync() {}
// warp-collective on the hardware (.sync.aligned: the whole warp must execute it together, one allocation per warp): every
// lane joins a warp rendezvous -- a lane that never arrives shows up as a deadlock -- and lane 0 acts for the warp
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  emu::warp_barrier();
  if ((threadIdx.x & 31) == 0) *dst_smem = emu::tmem_alloc(ncols);
  emu::warp_barrier();
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  emu::warp_barrier();
  if ((threadIdx.x & 31) == 0) emu::tmem_dealloc(taddr, ncols);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, bool accumulate) {
  emu::umma_tf32(tmem_d, da, db, idesc, accumulate);
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, bool accumulate) {
  emu::umma_bf16(tmem_d, da, db, idesc, accumulate);
}
__device__ __forceinline__ void commit(uint64_t* bar) { emu::tc_commit(bar); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  emu::warp_barrier();   // .sync.aligned: all 32 lanes of the warp execute the load together
  emu::tmem_ld(taddr, r, 16);
}
__device__ __forceinline__ void tmem_ld_wait() { emu::tmem_ld_wait(); }
#else
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  const uint32_t a = smem_u32(bar);
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
    if (!done && ++spins > (1u << 22)) __trap();  // a broken hand-off becomes an error, never a hung GPU
  }
}
// generic-proxy writes to shared memory -> visible to the tensor core (async proxy)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// one full warp; ncols a power of two >= 32; the base address lands in *dst_smem
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  __syncwarp();  // .sync.aligned: the warp must be converged
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  __syncwarp();
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
// issued by ONE thread: D[tmem] (+)= A[smem desc] * B[smem desc]^T, K = 8
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, bool accumulate) {
  const uint32_t acc = accumulate ? 1u : 0u, zero = 0u;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(zero)
      : "memory");
}
// the 16-bit twin: kind::f16 (the instruction descriptor says BF16), K = 16
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, bool accumulate) {
  const uint32_t acc = accumulate ? 1u : 0u, zero = 0u;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(zero)
      : "memory");
}
// arrives on `bar` once every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// warp w of a warpgroup reads TMEM lanes 32*(w%4)..+31 (taddr carries that lane base in its upper half): 16 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  __syncwarp();  // .sync.aligned: callers' per-row branches must have reconverged
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
#endif

// round to TF32 (10 mantissa bits, nearest, ties away) -- the hi part of the 3xTF32 split
__device__ __forceinline__ float tf32_round(float x) {
#ifdef XDR_EMU
  return emu::as_float(emu::to_tf32(x));
#else
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
#endif
}
// stores a float4 of 4 consecutive K elements of row r as its hi and lo TF32 parts into two K-major operand planes
__device__ __forceinline__ void store_split4(unsigned char* hi_plane, unsigned char* lo_plane, const KMajor& lay, int r, int k4,
                                             float4 v) {
  const float4 h = make_float4(tf32_round(v.x), tf32_round(v.y), tf32_round(v.z), tf32_round(v.w));
  const float4 l = make_float4(tf32_round(v.x - h.x), tf32_round(v.y - h.y), tf32_round(v.z - h.z), tf32_round(v.w - h.w));
  const int off = lay.chunk_offset(r, k4);
  *reinterpret_cast<float4*>(hi_plane + off) = h;
  *reinterpret_cast<float4*>(lo_plane + off) = l;
}

// the MN-major twin of store_split4: v = 4 consecutive ROWS (4*r4 ..) at reduction index k
__device__ __forceinline__ void store_split4_mn(unsigned char* hi_plane, unsigned char* lo_plane, const MNMajor& lay, int r4, int k,
                                                float4 v) {
  const float4 h = make_float4(tf32_round(v.x), tf32_round(v.y), tf32_round(v.z), tf32_round(v.w));
  const float4 l = make_float4(tf32_round(v.x - h.x), tf32_round(v.y - h.y), tf32_round(v.z - h.z), tf32_round(v.w - h.w));
  const int off = lay.chunk_offset(r4, k);
  *reinterpret_cast<float4*>(hi_plane + off) = h;
  *reinterpret_cast<float4*>(lo_plane + off) = l;
}

// D (+)= A * B^T over K (multiple of 8) in 3xTF32: A / B given as hi and lo planes, each K-major or MN-major (the layout
// objects say which; the instruction descriptor must carry the matching major bits).  ONE thread calls it.
template <typename LayA, typename LayB>
__device__ __forceinline__ void mma_3xtf32(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, const LayA& la, uint32_t b_hi,
                                           uint32_t b_lo, const LayB& lb, uint32_t idesc, int K, bool accumulate) {
  for (int ks = 0; ks < K / 8; ++ks) {
    const uint32_t ao = ks * la.k_step_bytes(), bo = ks * lb.k_step_bytes();
    const uint64_t dah = make_desc(a_hi + ao, la.lbo(), la.sbo(), la.mn_major()),
                   dal = make_desc(a_lo + ao, la.lbo(), la.sbo(), la.mn_major());
    const uint64_t dbh = make_desc(b_hi + bo, lb.lbo(), lb.sbo(), lb.mn_major()),
                   dbl = make_desc(b_lo + bo, lb.lbo(), lb.sbo(), lb.mn_major());
    mma_tf32(tmem_d, dal, dbh, idesc, accumulate || ks > 0);   // small terms first
    mma_tf32(tmem_d, dah, dbl, idesc, true);
    mma_tf32(tmem_d, dah, dbh, idesc, true);
  }
}

// ---- bf16x3 -----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bf16_rn(float x) {   // round-to-nearest-even, result in the low 16 bits
  uint32_t u = __float_as_uint(x);
  if ((u & 0x7f800000u) == 0x7f800000u) return u >> 16;
  u += 0x7fffu + ((u >> 16) & 1u);
  return u >> 16;
}
#ifndef XDR_EMU
// two fp32 -> one packed bf16x2 word (x0 in the low half), round-to-nearest-even: ONE conversion instruction (F2FP.BF16) where
// the integer form above costs ~6 per element -- the split was a third of the tcgen05 kernels' epilogue instructions
__device__ __forceinline__ uint32_t bf16x2_rn(float x0, float x1) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x1), "f"(x0));
  return r;
}
#else
__device__ __forceinline__ uint32_t bf16x2_rn(float x0, float x1) { return bf16_rn(x0) | (bf16_rn(x1) << 16); }
#endif
__device__ __forceinline__ void split_bf16_pair(float x0, float x1, uint32_t& hi, uint32_t& lo) {   // element 0 in the low half
  hi = bf16x2_rn(x0, x1);
  lo = bf16x2_rn(x0 - __uint_as_float(hi << 16), x1 - __uint_as_float(hi & 0xffff0000u));
}
// eight values that are consecutive along the operand's contiguous direction -> one 16-byte chunk in the hi plane and one in
// the lo plane.  K-major: (row r, K elements 8*k8 ..);  MN-major: (rows 8*r8 .., reduction index k) -- pass the matching offset.
__device__ __forceinline__ void store_split8(unsigned char* hi_plane, unsigned char* lo_plane, int chunk_off, float4 v0, float4 v1) {
  uint4 h, l;
  split_bf16_pair(v0.x, v0.y, h.x, l.x);
  split_bf16_pair(v0.z, v0.w, h.y, l.y);
  split_bf16_pair(v1.x, v1.y, h.z, l.z);
  split_bf16_pair(v1.z, v1.w, h.w, l.w);
  *reinterpret_cast<uint4*>(hi_plane + chunk_off) = h;
  *reinterpret_cast<uint4*>(lo_plane + chunk_off) = l;
}
// D (+)= A * B^T over K (multiple of 16) as bf16x3 (a_lo b_hi + a_hi b_lo + a_hi b_hi, ~2^-17 relative per product).  ONE thread.
template <typename LayA, typename LayB>
__device__ __forceinline__ void mma_bf16x3(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, const LayA& la, uint32_t b_hi,
                                           uint32_t b_lo, const LayB& lb, uint32_t idesc, int K, bool accumulate) {
  for (int ks = 0; ks < K / 16; ++ks) {
    const uint32_t ao = ks * la.k_step_bytes(), bo = ks * lb.k_step_bytes();
    const uint64_t dah = make_desc(a_hi + ao, la.lbo(), la.sbo(), la.mn_major()),
                   dal = make_desc(a_lo + ao, la.lbo(), la.sbo(), la.mn_major());
    const uint64_t dbh = make_desc(b_hi + bo, lb.lbo(), lb.sbo(), lb.mn_major()),
                   dbl = make_desc(b_lo + bo, lb.lbo(), lb.sbo(), lb.mn_major());
    mma_bf16(tmem_d, dal, dbh, idesc, accumulate || ks > 0);
    mma_bf16(tmem_d, dah, dbl, idesc, true);
    mma_bf16(tmem_d, dah, dbh, idesc, true);
  }
}

// ---- bf16x6: three bf16 planes (x ~= hi + mid + lo, 24 mantissa bits) and the six products whose weight is >= 2^-16 of the
// leading one (hi hi, hi mid, mid hi, hi lo, lo hi, mid mid; what is dropped is below 2^-24) -- fp32-faithful results, for the
// layers whose outputs go through a ReLU: with bf16x3 (~2^-16 per product) a pre-activation within 1e-5 of zero gets another
// sub-gradient than the fp32 reference, which showed up as whole rows of differing gradients (round 2, GPU call 12).
__device__ __forceinline__ void split_bf16_pair3(float x0, float x1, uint32_t& hi, uint32_t& mid, uint32_t& lo) {
  hi = bf16x2_rn(x0, x1);
  const float r0 = x0 - __uint_as_float(hi << 16), r1 = x1 - __uint_as_float(hi & 0xffff0000u);
  mid = bf16x2_rn(r0, r1);
  lo = bf16x2_rn(r0 - __uint_as_float(mid << 16), r1 - __uint_as_float(mid & 0xffff0000u));
}
__device__ __forceinline__ void store_split8_3(unsigned char* hi_plane, unsigned char* mid_plane, unsigned char* lo_plane,
                                               int chunk_off, float4 v0, float4 v1) {
  uint4 h, m, l;
  split_bf16_pair3(v0.x, v0.y, h.x, m.x, l.x);
  split_bf16_pair3(v0.z, v0.w, h.y, m.y, l.y);
  split_bf16_pair3(v1.x, v1.y, h.z, m.z, l.z);
  split_bf16_pair3(v1.z, v1.w, h.w, m.w, l.w);
  *reinterpret_cast<uint4*>(hi_plane + chunk_off) = h;
  *reinterpret_cast<uint4*>(mid_plane + chunk_off) = m;
  *reinterpret_cast<uint4*>(lo_plane + chunk_off) = l;
}
// D (+)= A * B^T over K (multiple of 16); a0 / b0 = shared-memory address of the hi plane, the mid and lo planes follow at
// a_plane / b_plane bytes.  ONE thread.  Small terms first.
template <typename LayA, typename LayB>
__device__ __forceinline__ void mma_bf16x6(uint32_t tmem_d, uint32_t a0, uint32_t a_plane, const LayA& la, uint32_t b0,
                                           uint32_t b_plane, const LayB& lb, uint32_t idesc, int K, bool accumulate) {
  for (int ks = 0; ks < K / 16; ++ks) {
    const uint32_t ao = a0 + ks * la.k_step_bytes(), bo = b0 + ks * lb.k_step_bytes();
    uint64_t da[3], db[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      da[i] = make_desc(ao + i * a_plane, la.lbo(), la.sbo(), la.mn_major());
      db[i] = make_desc(bo + i * b_plane, lb.lbo(), lb.sbo(), lb.mn_major());
    }
    mma_bf16(tmem_d, da[1], db[1], idesc, accumulate || ks > 0);   // mid mid
    mma_bf16(tmem_d, da[2], db[0], idesc, true);                   // lo  hi
    mma_bf16(tmem_d, da[0], db[2], idesc, true);                   // hi  lo
    mma_bf16(tmem_d, da[1], db[0], idesc, true);                   // mid hi
    mma_bf16(tmem_d, da[0], db[1], idesc, true);                   // hi  mid
    mma_bf16(tmem_d, da[0], db[0], idesc, true);                   // hi  hi
  }
}

// ---- self-test: D[128 x N] = A[128 x K] * B[N x K]^T from row-major operands in global memory, one CTA of 128 threads, with
// A and B staged K-major or MN-major.  Runs under the emulator and, on hardware, checks the descriptor reading through the
// library itself (xdr_tc5_selftest).
// fmt 0: 3xTF32 on fp32 hi / lo planes (K % 8 == 0);  fmt 1: bf16x3 on bf16 hi / lo planes (K % 16 == 0).
// (a kernel definition: emitted only in the translation unit that defines XDR_TC5_SELFTEST_IMPL -- topk_score.cu)
#ifdef XDR_TC5_SELFTEST_IMPL
__global__ void __launch_bounds__(128, 1) selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, int N, int K,
                                                          int a_mn, int b_mn, int fmt, float* __restrict__ D) {
  XDR_DYN_SMEM_ALIGNED(unsigned char, smem_t5, 128);
  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr int M = 128;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_t5);
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(smem_t5 + 8);
  const int esz = fmt ? 2 : 4;
  unsigned char* a_hi = smem_t5 + 128;
  unsigned char* a_lo = a_hi + M * K * esz;
  unsigned char* b_hi = a_lo + M * K * esz;
  unsigned char* b_lo = b_hi + N * K * esz;
  const KMajor ka{M}, kb{N};
  const MNMajor ma{M}, mb{N};
  const KMajor16 ka16{M}, kb16{N};
  const MNMajor16 ma16{M}, mb16{N};
  // eight consecutive values along an operand's contiguous direction (X is row-major [R][K])
  auto along_k = [&](const float* X, int r, int k8, float4& v0, float4& v1) {
    v0 = *reinterpret_cast<const float4*>(X + r * K + 8 * k8);
    v1 = *reinterpret_cast<const float4*>(X + r * K + 8 * k8 + 4);
  };
  auto along_rows = [&](const float* X, int r8, int k, float4& v0, float4& v1) {
    const float* p = X + (size_t)(8 * r8) * K + k;
    v0 = make_float4(p[0], p[K], p[2 * K], p[3 * K]);
    v1 = make_float4(p[4 * K], p[5 * K], p[6 * K], p[7 * K]);
  };
  if (fmt) {
    for (int e = tid; e < M * K / 8; e += 128) {
      float4 v0, v1;
      if (a_mn == 2) { along_k(A, e / (K / 8), e % (K / 8), v0, v1); store_split8(a_hi, a_lo, RowBlock16{M, K}.chunk_offset(e / (K / 8), e % (K / 8)), v0, v1); }
      else if (!a_mn) { along_k(A, e / (K / 8), e % (K / 8), v0, v1); store_split8(a_hi, a_lo, ka16.chunk_offset(e / (K / 8), e % (K / 8)), v0, v1); }
      else { along_rows(A, e / K, e % K, v0, v1); store_split8(a_hi, a_lo, ma16.chunk_offset(e / K, e % K), v0, v1); }
    }
    for (int e = tid; e < N * K / 8; e += 128) {
      float4 v0, v1;
      if (!b_mn) { along_k(B, e / (K / 8), e % (K / 8), v0, v1); store_split8(b_hi, b_lo, kb16.chunk_offset(e / (K / 8), e % (K / 8)), v0, v1); }
      else { along_rows(B, e / K, e % K, v0, v1); store_split8(b_hi, b_lo, mb16.chunk_offset(e / K, e % K), v0, v1); }
    }
  }
  for (int e = tid; !fmt && e < M * K / 4; e += 128) {
    if (!a_mn) {
      const int r = e / (K / 4), k4 = e % (K / 4);
      store_split4(a_hi, a_lo, ka, r, k4, *reinterpret_cast<const float4*>(A + r * K + 4 * k4));
    } else {
      const int r4 = e / K, k = e % K;
      store_split4_mn(a_hi, a_lo, ma, r4, k, make_float4(A[(4 * r4) * K + k], A[(4 * r4 + 1) * K + k], A[(4 * r4 + 2) * K + k],
                                                        A[(4 * r4 + 3) * K + k]));
    }
  }
  for (int e = tid; !fmt && e < N * K / 4; e += 128) {
    if (!b_mn) {
      const int r = e / (K / 4), k4 = e % (K / 4);
      store_split4(b_hi, b_lo, kb, r, k4, *reinterpret_cast<const float4*>(B + r * K + 4 * k4));
    } else {
      const int r4 = e / K, k = e % K;
      store_split4_mn(b_hi, b_lo, mb, r4, k, make_float4(B[(4 * r4) * K + k], B[(4 * r4 + 1) * K + k], B[(4 * r4 + 2) * K + k],
                                                        B[(4 * r4 + 3) * K + k]));
    }
  }
  const uint32_t cols = N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : 256));
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init_fence();
  }
  if (warp == 0) tmem_alloc(tmem_base_smem, cols);
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_base_smem;
  if (tid == 0) {
    const uint32_t idesc = fmt ? make_idesc_bf16(M, N, a_mn == 1, b_mn != 0) : make_idesc_tf32(M, N, a_mn != 0, b_mn != 0);
    const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
    if (fmt && a_mn == 2) {
      const RowBlock16::View va = RowBlock16{M, K}.as_k_major();
      if (!b_mn) mma_bf16x3(tmem, ah, al, va, bh, bl, kb16, idesc, K, false);
      else mma_bf16x3(tmem, ah, al, va, bh, bl, mb16, idesc, K, false);
    } else if (fmt) {
      if (!a_mn && !b_mn) mma_bf16x3(tmem, ah, al, ka16, bh, bl, kb16, idesc, K, false);
      else if (!a_mn && b_mn) mma_bf16x3(tmem, ah, al, ka16, bh, bl, mb16, idesc, K, false);
      else if (a_mn && !b_mn) mma_bf16x3(tmem, ah, al, ma16, bh, bl, kb16, idesc, K, false);
      else mma_bf16x3(tmem, ah, al, ma16, bh, bl, mb16, idesc, K, false);
    } else if (!a_mn && !b_mn) mma_3xtf32(tmem, ah, al, ka, bh, bl, kb, idesc, K, false);
    else if (!a_mn && b_mn) mma_3xtf32(tmem, ah, al, ka, bh, bl, mb, idesc, K, false);
    else if (a_mn && !b_mn) mma_3xtf32(tmem, ah, al, ma, bh, bl, kb, idesc, K, false);
    else mma_3xtf32(tmem, ah, al, ma, bh, bl, mb, idesc, K, false);
    commit(bar);
  }
  mbar_wait(bar, 0);
  fence_after_sync();
  for (int c = 0; c < N; c += 16) {
    uint32_t r[16];
    __syncwarp();  // the spin above exits lane by lane; tcgen05.ld is warp-collective
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
    tmem_ld_wait();
    for (int j = 0; j < 16 && c + j < N; ++j) D[(size_t)tid * N + c + j] = __uint_as_float(r[j]);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, cols);
}
#endif  // XDR_TC5_SELFTEST_IMPL

#endif  // __CUDACC__ || XDR_EMU

}  // namespace tc5
}  // namespace xdr
