// Synthetic interaction graph at BASELINE config 5 scale, generated on the device (SURVEY.md 8(d) "Scaled (config 5)":
// 10 M users x 2 M items, ~500 M interactions, counter-based Philox stream, same skew model as the Yelp-shape
// generator: log-normal user activity, Zipf item popularity). The reference has no generator (it reads Yelp's
// review.json); this is the input producer of the scaled benchmark and of its parity tests, not a port of anything.
//
// Output: the USER side of the bipartite graph as CSR — for every user the SORTED, DUPLICATE-FREE list of item ids.
// One warp per user: draws -> shared-memory bitonic sort -> unique. Two passes with the same counters (Philox is
// counter-based): count (rowptr == NULL) and fill. A function of (seed, user id) only, hence identical on every rank.
#include "common.cuh"
#include "philox.cuh"

namespace yr {

constexpr int kSynthMaxDeg = 1024;
constexpr int kSynthWarps = 8;

struct SynthParams {
  uint64_t seed;
  int64_t nU, nI;
  double mu, sigma;         // log-normal activity: draws = exp(mu + sigma z)
  double zipf_c, zipf_p;    // item rank k = floor(((c * U + 1) ^ p)), c = (nI + 1)^(1-alpha) - 1, p = 1 / (1 - alpha)
  int32_t min_deg, max_deg;
  uint64_t perm_a, perm_c;  // item id = (rank * a + c) mod nI — popularity is not a function of the id order
};

__device__ __forceinline__ double u01(uint32_t hi, uint32_t lo) {      // (0, 1), 53 bits
  const uint64_t x = ((uint64_t)hi << 32) | lo;
  return ((double)(x >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ int synth_degree(const SynthParams& P, int64_t u) {
  uint32_t w[4];
  philox4x32_10((uint32_t)u, (uint32_t)(u >> 32), 0xD6E8FEB8u, 1u, (uint32_t)P.seed, (uint32_t)(P.seed >> 32), w);
  const double a = u01(w[0], w[1]), b = u01(w[2], w[3]);
  const double z = sqrt(-2.0 * log(a)) * cos(6.283185307179586 * b);   // Box-Muller
  double d = exp(P.mu + P.sigma * z);
  if (d < (double)P.min_deg) d = (double)P.min_deg;
  if (d > (double)P.max_deg) d = (double)P.max_deg;
  return (int)d;
}

__global__ void __launch_bounds__(32 * kSynthWarps)
synth_user_rows_kernel(SynthParams P, const int32_t* __restrict__ rowptr, int32_t* __restrict__ cnt_out,
                       int32_t* __restrict__ items_out) {
  __shared__ int32_t buf[kSynthWarps][kSynthMaxDeg];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int32_t* a = buf[wib];
  const int64_t nw = (int64_t)gridDim.x * kSynthWarps;
  for (int64_t u = (int64_t)blockIdx.x * kSynthWarps + wib; u < P.nU; u += nw) {
    const int d = synth_degree(P, u);
    int Pw = 32;
    while (Pw < d) Pw <<= 1;
    // draws: Philox block j gives 2 items
    for (int j = lane; j < Pw / 2; j += 32) {
      uint32_t w[4];
      philox4x32_10((uint32_t)u, (uint32_t)(u >> 32), (uint32_t)j, 2u, (uint32_t)P.seed, (uint32_t)(P.seed >> 32), w);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int e = 2 * j + q;
        int32_t item = 0x7fffffff;                                     // padding sorts last
        if (e < d) {
          const double x = u01(w[2 * q], w[2 * q + 1]);
          double k = floor(pow(P.zipf_c * x + 1.0, P.zipf_p));          // rank in [1, nI]
          if (k < 1.0) k = 1.0;
          if (k > (double)P.nI) k = (double)P.nI;
          const uint64_t r = (uint64_t)k - 1u;
          item = (int32_t)((r * P.perm_a + P.perm_c) % (uint64_t)P.nI);
        }
        a[e] = item;
      }
    }
    __syncwarp();
    for (int k = 2; k <= Pw; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = lane; i < Pw; i += 32) {
          const int q = i ^ j;
          if (q > i) {
            const int32_t x = a[i], y = a[q];
            if ((x > y) == ((i & k) == 0)) { a[i] = y; a[q] = x; }
          }
        }
        __syncwarp();
      }
    // unique (sorted): keep the first of every run
    int kept = 0;
    const int64_t base = rowptr ? (int64_t)rowptr[u] : 0;
    for (int i0 = 0; i0 < d; i0 += 32) {
      const int i = i0 + lane;
      const bool keep = i < d && (i == 0 || a[i] != a[i - 1]);
      const unsigned m = __ballot_sync(kFull, keep);
      if (keep && rowptr) items_out[base + kept + __popc(m & ((1u << lane) - 1u))] = a[i];
      kept += __popc(m);
    }
    if (!rowptr && lane == 0) cnt_out[u] = kept;
    __syncwarp();
  }
}

// val[e] = (dinv[row] * 1) * dinv[col] for a binary adjacency, dinv = 1 / sqrt(deg) in fp32 — the association and the
// rounding of data/datasets/ngcf_data_pipeline.py:34-42 (deg as an exact integer count < 2^24).
__global__ void __launch_bounds__(256)
lap_binary_values_kernel(const int32_t* __restrict__ rowptr, int64_t n_rows, const int32_t* __restrict__ row_node,
                         const int32_t* __restrict__ col_node, const int32_t* __restrict__ deg, float* __restrict__ val) {
  const int lane = threadIdx.x & 31;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rows; r += nw) {
    const float dr = __fdiv_rn(1.f, __fsqrt_rn((float)deg[row_node ? row_node[r] : r]));
    for (int32_t k = rowptr[r] + lane; k < rowptr[r + 1]; k += 32) {
      const float dc = __fdiv_rn(1.f, __fsqrt_rn((float)deg[col_node[k]]));
      val[k] = __fmul_rn(__fmul_rn(dr, 1.f), dc);
    }
  }
}

}  // namespace yr

using namespace yr;

static uint64_t gcd64(uint64_t a, uint64_t b) { while (b) { const uint64_t t = a % b; a = b; b = t; } return a; }

extern "C" int yr_synth_user_rows(uint64_t seed, int64_t num_users, int64_t num_items, double zipf_alpha, double mu,
                                  double sigma, int32_t min_deg, int32_t max_deg, const int32_t* rowptr, int32_t* cnt_out,
                                  int32_t* items_out, yr_stream stream) {
  if (num_users <= 0 || num_items <= 0 || num_items >= (1LL << 31) || min_deg < 1 || max_deg < min_deg ||
      max_deg > kSynthMaxDeg || zipf_alpha < 0.0 || zipf_alpha >= 1.0)
    return YR_ERR_BAD_ARG;
  if (rowptr ? !items_out : !cnt_out) return YR_ERR_BAD_ARG;
  SynthParams P;
  P.seed = seed; P.nU = num_users; P.nI = num_items; P.mu = mu; P.sigma = sigma;
  P.zipf_p = 1.0 / (1.0 - zipf_alpha);
  P.zipf_c = pow((double)num_items + 1.0, 1.0 - zipf_alpha) - 1.0;
  P.min_deg = min_deg; P.max_deg = max_deg;
  uint64_t a = 2654435761ull % (uint64_t)num_items;
  if (a == 0) a = 1;
  while (gcd64(a, (uint64_t)num_items) != 1) ++a;
  P.perm_a = a; P.perm_c = (0x9E3779B97F4A7C15ull ^ seed) % (uint64_t)num_items;
  int64_t blocks = (num_users + kSynthWarps - 1) / kSynthWarps;
  const int64_t cap = (int64_t)yr_sm_count() * 8;
  if (blocks > cap) blocks = cap;
  synth_user_rows_kernel<<<(unsigned)blocks, 32 * kSynthWarps, 0, (cudaStream_t)stream>>>(P, rowptr, cnt_out, items_out);
  YR_CHECK_LAUNCH();
  return YR_OK;
}

extern "C" int yr_laplacian_binary_values(const int32_t* rowptr, int64_t n_rows, const int32_t* row_node,
                                          const int32_t* col_node, const int32_t* deg, float* val, yr_stream stream) {
  if (!rowptr || !col_node || !deg || !val || n_rows < 0) return YR_ERR_BAD_ARG;
  if (n_rows == 0) return YR_OK;
  int64_t blocks = (n_rows * 32 + 255) / 256;
  const int64_t cap = (int64_t)yr_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  lap_binary_values_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(rowptr, n_rows, row_node, col_node, deg, val);
  YR_CHECK_LAUNCH();
  return YR_OK;
}
