// neg_sample.cu -- A18: uniform negative draw with per-user rejection, on the device.
//
// Replaces CrossDomainSourceSampler._uni_sampling + AbstractSampler.sample_by_key_ids
// (reference sampler/crossdomain_sampler.py:220-221, 139-176) and recbole's target-domain Sampler (uniform over
// [1, item_num)) [recbole-1.0.1]: draw uniformly from the domain's valid item ids, redraw every position whose draw is
// in the user's used-item set, output laid out as `num` blocks of len(key_ids) (np.tile(key_ids, num), :166).
// The reference loops in Python over NumPy draws and Python sets; here one thread owns one output position:
// Philox4x32-10 counter-based draws (counter = position, attempt, stream; key = seed), candidate -> joint id by the
// A0 layout arithmetic, membership test by binary search in the user's sorted CSR row.  Integer work: results are
// bit-identical to oracle/sampler_oracle.py.  Latency-bound (a few dependent loads per position), not on the HBM roofline.
#include "xdr_common.cuh"

namespace xdr {

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
  c[0] = n0;
  c[1] = lo1;
  c[2] = n2;
  c[3] = lo0;
}

__device__ __forceinline__ uint64_t philox64(uint64_t pos, uint32_t attempt, uint32_t stream_id, uint64_t seed) {
  uint32_t c[4] = {(uint32_t)pos, (uint32_t)(pos >> 32), attempt, stream_id};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return (uint64_t)c[0] | ((uint64_t)c[1] << 32);
}

__global__ void __launch_bounds__(256)
    neg_sample_uniform_kernel(const int64_t* __restrict__ key_ids, int64_t n_keys_in, int64_t total,
                              const int64_t* __restrict__ rowptr, const int64_t* __restrict__ col, int64_t n_rows,
                              int64_t n_overlap, int64_t n_gap, uint64_t n_valid, uint64_t seed, uint32_t stream_id,
                              int max_attempts, int64_t* __restrict__ out, int32_t* status) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t u = key_ids[i % n_keys_in];  // np.tile(key_ids, num)
    if ((uint64_t)u >= (uint64_t)n_rows) {     // the reference turns this IndexError into ValueError('user_id ... not exist')
      if (status) atomicOr(status, 2);
      out[i] = 0;
      continue;
    }
    const int64_t lo = rowptr[u], hi = rowptr[u + 1];
    int64_t id = 0;
    int attempt = 0;
    for (;;) {
      const uint64_t r = philox64((uint64_t)i, (uint32_t)attempt, stream_id, seed);
      const int64_t k1 = (int64_t)__umul64hi(r, n_valid) + 1;
      id = k1 < n_overlap ? k1 : k1 + n_gap;
      // binary search in the user's sorted used-item row
      int64_t a = lo, b = hi;
      while (a < b) {
        const int64_t m = (a + b) >> 1;
        if (col[m] < id) a = m + 1; else b = m;
      }
      if (!(a < hi && col[a] == id)) break;
      if (++attempt >= max_attempts) {
        if (status) atomicOr(status, 1);
        break;
      }
    }
    out[i] = id;
  }
}

}  // namespace xdr

using namespace xdr;

extern "C" {

int xdr_neg_sample_uniform(const int64_t* key_ids, int64_t n_keys, int num, const int64_t* used_rowptr,
                           const int64_t* used_col, int64_t n_rows, int64_t n_overlap, int64_t n_gap, int64_t n_valid,
                           uint64_t seed, uint32_t stream_id, int max_attempts, int64_t* out, int32_t* status,
                           xdr_stream_t stream) {
  XDR_REQUIRE(n_keys >= 0 && num >= 0, "xdr_neg_sample_uniform: negative size");
  const int64_t total = n_keys * (int64_t)num;
  if (total == 0) return XDR_OK;
  XDR_REQUIRE(key_ids && used_rowptr && out, "xdr_neg_sample_uniform: null pointer");
  XDR_REQUIRE(n_valid > 0 && n_overlap >= 1 && n_gap >= 0, "xdr_neg_sample_uniform: empty candidate set");
  XDR_REQUIRE(max_attempts > 0, "xdr_neg_sample_uniform: max_attempts must be positive");
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  XDR_LAUNCH((neg_sample_uniform_kernel), (int)blocks, 256, 0, (cudaStream_t)stream, 
      key_ids, n_keys, total, used_rowptr, used_col, n_rows, n_overlap, n_gap, (uint64_t)n_valid, seed, stream_id,
      max_attempts, out, status);
  XDR_LAUNCH_OK();
  return XDR_OK;
}

}  // extern "C"
