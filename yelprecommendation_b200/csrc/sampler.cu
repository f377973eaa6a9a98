// On-device negative sampler (sm_100a) — SURVEY.md §8(f)2.
// Replaces MFDataset._negative_sampling (reference data/datasets/mf_dataset.py:18-22): one negative per training
// interaction, uniform over the items that are not in the user's `pos_items`. The reference draws one sample at a
// time from NumPy's global Mersenne Twister inside `__getitem__`; here every triple owns a counter-based
// Philox4x32-10 stream (key = seed, counter = (t_lo, t_hi, block, 0)), so the result depends only on
// (seed, global triple index) — reproducible under any sharding of the triples across GPUs. Same distribution,
// different stream; oracle/yr_oracle.c restates this stream and the kernel is bit-exact against it.
#include "common.cuh"
#include "philox.cuh"

namespace yr {

__global__ void __launch_bounds__(256)
sample_negatives_kernel(const int64_t* __restrict__ uid, int64_t n, const int32_t* __restrict__ pos_ptr,
                        const int32_t* __restrict__ pos_idx, int64_t num_users, uint32_t nI, uint64_t seed,
                        uint64_t offset, int max_blocks, int64_t* __restrict__ neg_out, int32_t* err) {
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  const uint32_t thresh = (0u - nI) % nI;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t u = uid[t];
    if (u < 0 || u >= num_users) {
      if (err) atomicExch(err, 1);
      neg_out[t] = -1;
      continue;
    }
    const uint64_t idx = offset + (uint64_t)t;
    const int32_t lo = __ldg(pos_ptr + u), hi = __ldg(pos_ptr + u + 1);
    int64_t found = -1;
    for (int blk = 0; blk < max_blocks && found < 0; ++blk) {
      uint32_t w[4];
      philox4x32_10((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)blk, 0u, k0, k1, w);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (found >= 0) break;
        const uint64_t m = (uint64_t)w[q] * nI;
        if ((uint32_t)m < thresh) continue;           // Lemire: keep the map exactly uniform
        const int32_t cand = (int32_t)(m >> 32);
        int32_t a = lo, b = hi;
        while (a < b) {
          const int32_t mid = (a + b) >> 1;
          if (__ldg(pos_idx + mid) < cand) a = mid + 1; else b = mid;
        }
        if (a < hi && __ldg(pos_idx + a) == cand) continue;
        found = cand;
      }
    }
    if (found < 0 && err) atomicExch(err, 2);          // the reference would loop forever here
    neg_out[t] = found;
  }
}

}  // namespace yr

using namespace yr;

extern "C" int yr_sample_negatives(const int64_t* uid, int64_t n, const int32_t* pos_ptr, const int32_t* pos_idx,
                                   int64_t num_users, int64_t num_items, uint64_t seed, uint64_t offset,
                                   int max_blocks, int64_t* neg_out, int32_t* err, yr_stream stream) {
  if (!uid || !pos_ptr || !pos_idx || !neg_out || n < 0 || num_users <= 0 || num_items <= 0 || num_items >= (1LL << 31) ||
      max_blocks <= 0)
    return YR_ERR_BAD_ARG;
  if (n == 0) return YR_OK;
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)yr_sm_count() * 8;
  if (blocks > cap) blocks = cap;
  sample_negatives_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      uid, n, pos_ptr, pos_idx, num_users, (uint32_t)num_items, seed, offset, max_blocks, neg_out, err);
  YR_CHECK_LAUNCH();
  return YR_OK;
}
