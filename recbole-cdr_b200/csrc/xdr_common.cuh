// xdr_common.cuh -- device/host helpers shared by every kernel file of libxdr (sm_100a only).
#pragma once
#ifdef XDR_EMU
#include "cuda_emu.h"  // tests/emu: CPU CTA emulator, test infrastructure only (never part of libxdr.so)
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include <stdio.h>
#ifndef XDR_EMU
#include <tuple>
#include <utility>
#endif
#include "../../include/xdr.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libxdr is written for sm_100a (B200) only"
#endif

namespace xdr {

// ---------------------------------------------------------------------------------------------------
// host side: error reporting
// ---------------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);  // defined in xdr_api.cu (thread-local buffer)

#define XDR_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      xdr::set_error(__VA_ARGS__);        \
      return XDR_ERR_INVALID;             \
    }                                     \
  } while (0)

#define XDR_CUDA_OK(expr)                                                                  \
  do {                                                                                     \
    cudaError_t e__ = (expr);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      xdr::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
      return XDR_ERR_CUDA;                                                                 \
    }                                                                                      \
  } while (0)

#define XDR_LAUNCH_OK() XDR_CUDA_OK(cudaGetLastError())

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline bool dim_ok(int dim) { return dim > 0 && dim <= 256 && (dim % 4) == 0; }

int sm_count();  // cached per process, defined in xdr_api.cu

#ifndef XDR_EMU
template <typename Tuple, size_t... I>
static inline void fill_arg_ptrs(void** ptrs, Tuple& t, std::index_sequence<I...>) {
  ((ptrs[I] = static_cast<void*>(&std::get<I>(t))), ...);
}
#endif

// ---------------------------------------------------------------------------------------------------
// workspace layout (xdr_workspace_bytes()): [0,64) tickets (uint32), then fp32 partial slots
// ---------------------------------------------------------------------------------------------------
constexpr int kMaxBlocks = 2048;       // upper bound on gridDim.x of any reducing kernel
constexpr int kPartialsPerBlock = 4;   // fp32 slots per block
constexpr size_t kWsTicketBytes = 64;
constexpr size_t kWsBytes = kWsTicketBytes + sizeof(float) * kPartialsPerBlock * kMaxBlocks;

struct Workspace {
  unsigned int* ticket;
  float* partials;
  __host__ __device__ explicit Workspace(void* ws)
      : ticket(reinterpret_cast<unsigned int*>(ws)),
        partials(reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + kWsTicketBytes)) {}
};

// ---------------------------------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------------------------------
#if defined(__CUDACC__) || defined(XDR_EMU)

constexpr int kMaxShards = 8;

// A table row-sharded over 2^log2g GPUs: global row r lives on shard (r mod G) at local row (r div G).  Block-cyclic so
// that the three id ranges of the joint layout (overlapped / target-only / source-only) and Zipf-hot low ids spread
// evenly.  Shard pointers are local or peer-mapped (CUDA IPC over NVLink) device pointers; G = 1 is the plain table.
struct Shards {
  float* p[kMaxShards];
};
__device__ __forceinline__ float* shard_row(const Shards& t, int log2g, int64_t row, int64_t row_f) {
  return t.p[row & ((1 << log2g) - 1)] + (row >> log2g) * row_f;
}


// Row geometry: a row of `dim` floats is nv = dim/4 float4s.  An interaction (or row) is owned by a group of
// 8 consecutive lanes; lane `sub` of the group owns float4 columns sub, sub+8, ... (VEC of them).  One 8-lane
// slice of a row is one full 128-byte line, so every LDG.128 / RED.128 of a group is a single-line request.
constexpr int kLanesPerRow = 8;
constexpr int kRowsPerWarp = 32 / kLanesPerRow;

// 128-bit read-only gather that does not allocate in L1 (rows are touched once per kernel).
__device__ __forceinline__ float4 ldg_row4(const float* __restrict__ row, int col4) {
#ifdef XDR_EMU
  emu::counters().row_load_bytes += 16;
  return *(reinterpret_cast<const float4*>(row) + col4);
#else
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(reinterpret_cast<const float4*>(row) + col4));
  return v;
#endif
}

// 128-bit coherent load (tables that may be written by a concurrent scatter in the same launch).
__device__ __forceinline__ float4 ld_row4(const float* row, int col4) {
#ifdef XDR_EMU
  emu::counters().row_load_bytes += 16;
#endif
  return *(reinterpret_cast<const float4*>(row) + col4);
}

// fp32 x4 reduction to global memory: REDG.E.ADD.F32x4 on sm_100a.
__device__ __forceinline__ void red_add4(float* row, int col4, float4 v) {
#ifdef XDR_EMU
  emu::counters().row_red_bytes += 16;
  float* q = row + 4 * col4;
  q[0] += v.x; q[1] += v.y; q[2] += v.z; q[3] += v.w;
#else
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(reinterpret_cast<float4*>(row) + col4), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
#endif
}

__device__ __forceinline__ void st4(float* row, int col4, float4 v) {
  *(reinterpret_cast<float4*>(row) + col4) = v;
}

__device__ __forceinline__ bool aligned16_dev(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ float4 axpy4(float a, float4 x, float4 y) {
  return make_float4(fmaf(a, x.x, y.x), fmaf(a, x.y, y.y), fmaf(a, x.z, y.z), fmaf(a, x.w, y.w));
}
__device__ __forceinline__ float4 scale4(float a, float4 x) { return make_float4(a * x.x, a * x.y, a * x.z, a * x.w); }
__device__ __forceinline__ float4 sub4(float4 a, float4 b) {
  return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
}

// sum over the 8 lanes of a row group (all 32 lanes must participate)
__device__ __forceinline__ float group8_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// torch.sigmoid in fp32: 1 / (1 + exp(-x)) with the accurate expf
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// Block-level sum of NV values per thread -> thread 0 of the block holds the result in v[].
// smem: at least NV * (blockDim.x / 32) floats.
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) smem[i * nwarp + warp] = v[i];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float x = lane < nwarp ? smem[i * nwarp + lane] : 0.f;
      v[i] = warp_sum(x);
    }
  }
  __syncthreads();
}

// Deterministic grid reduction: every block publishes NV partials; the last block to arrive (ticket) sums all
// gridDim.x partial sets in a fixed order in fp64 and calls `fin(double sums[NV])` from its thread 0, then resets
// the ticket so the workspace is clean for the next launch on the same stream.
template <int NV, typename Fin>
__device__ __forceinline__ void grid_reduce_last_block(float (&v)[NV], Workspace ws, float* smem, Fin fin) {
  static_assert(NV <= kPartialsPerBlock, "too many partials");
  __shared__ bool is_last;
  block_sum<NV>(v, smem);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) ws.partials[(size_t)blockIdx.x * kPartialsPerBlock + i] = v[i];
    __threadfence();
    unsigned int t = atomicAdd(ws.ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.0;
  for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] += (double)__ldcg(&ws.partials[(size_t)b * kPartialsPerBlock + i]);
  }
  // block reduce in fp64 (fixed order: lane tree, then warps in index order)
  __shared__ double dsm[NV][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    acc[i] = warp_sum(acc[i]);
    if (lane == 0) dsm[i][warp] = acc[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      tot[i] = 0.0;
      for (int w = 0; w < nwarp; ++w) tot[i] += dsm[i][w];
    }
    fin(tot);
    *ws.ticket = 0u;
  }
}

#endif  // __CUDACC__ || XDR_EMU

// Kernel launch.  CUDA: kernel<<<grid, block, smem, stream>>>(args...).  Emulator (tests/emu, -DXDR_EMU): the CTAs run one
// after another on the CPU.  `kernel` is passed parenthesised so that template-ids with commas survive the preprocessor.
#ifdef XDR_EMU
#define XDR_LAUNCH(kernel, grid, block, smem, stream, ...) emu::launch(grid, block, smem, [&] { kernel(__VA_ARGS__); })
#else
#define XDR_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#endif

// Launch of a persistent kernel whose CTAs hand data to each other (all CTAs must be resident at once, or the ones that are
// would poll for the ones that are not, forever).  CUDA: xdr::coop_launch -- (1) the dynamic shared-memory limit of the kernel is
// raised once per kernel and device, not per call; (2) occupancy x #SMs >= grid is checked before the launch
// (XDR_ERR_UNSUPPORTED, never a grid that cannot fit); (3) the launch itself is cudaLaunchCooperativeKernel, for which the
// driver guarantees co-residency of the whole grid (it is started only when every CTA can be scheduled, whatever else is
// running on the device), unless xdr_set_coop_launch(0) asked for plain launches.  Emulator: all CTAs under one fiber scheduler.
#ifdef XDR_EMU
#define XDR_LAUNCH_COOP(kernel, grid, block, smem, stream, ...) \
  emu::launch(grid, block, smem, [=] { kernel(__VA_ARGS__); }, /*concurrent=*/true)  /* by value: may run deferred */
#else
int coop_prepare(const void* kern, int grid, int block, size_t smem, const char* name);  // xdr_api.cu (cached per kernel+device)
bool coop_enabled();                                                                     // xdr_api.cu
template <typename... KArgs, typename... Args>
static inline int coop_launch(const char* name, void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t s,
                              Args... args) {
  const int rc = coop_prepare(reinterpret_cast<const void*>(kern), grid, block, smem, name);
  if (rc != XDR_OK) return rc;
  if (coop_enabled()) {
    std::tuple<KArgs...> held(static_cast<KArgs>(args)...);
    void* ptrs[sizeof...(KArgs)];
    fill_arg_ptrs(ptrs, held, std::index_sequence_for<KArgs...>{});
    XDR_CUDA_OK(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kern), dim3(grid), dim3(block), ptrs, smem, s));
  } else {
    kern<<<grid, block, smem, s>>>(args...);
  }
  return XDR_OK;
}
#define XDR_LAUNCH_COOP(kernel, grid, block, smem, stream, ...)                                         \
  do {                                                                                                  \
    const int rc__ = xdr::coop_launch(#kernel, kernel, grid, block, smem, stream, __VA_ARGS__);         \
    if (rc__ != XDR_OK) return rc__;                                                                    \
  } while (0)
#endif

// Dynamic shared memory of the CTA as `type* name` (CUDA: the extern __shared__ array; emulator: the CTA's heap block).
#ifdef XDR_EMU
#define XDR_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::dyn_smem())
#define XDR_DYN_SMEM_ALIGNED(type, name, align) type* name = reinterpret_cast<type*>(emu::dyn_smem())
#else
#define XDR_DYN_SMEM(type, name) extern __shared__ __align__(16) type name[]
#define XDR_DYN_SMEM_ALIGNED(type, name, align) extern __shared__ __align__(align) type name[]
#endif

// Instantiate CALL with `constexpr int VEC` = float4 columns per lane for a row of nv float4s (8 lanes per row).
#define XDR_DISPATCH_VEC(nv, CALL)                                   \
  do {                                                               \
    const int vec__ = ((nv) + xdr::kLanesPerRow - 1) / xdr::kLanesPerRow;      \
    switch (vec__) {                                                 \
      case 1: { constexpr int VEC = 1; CALL; } break;                \
      case 2: { constexpr int VEC = 2; CALL; } break;                \
      case 3: { constexpr int VEC = 3; CALL; } break;                \
      case 4: { constexpr int VEC = 4; CALL; } break;                \
      case 5: case 6: { constexpr int VEC = 6; CALL; } break;        \
      default: { constexpr int VEC = 8; CALL; } break;               \
    }                                                                \
  } while (0)

}  // namespace xdr
