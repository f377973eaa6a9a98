// Per-user random split on the device (sm_100a) — SURVEY.md 8(f)3, second half.
// Replaces MFDataPipeline.split (reference data/datasets/mf_data_pipeline.py:18-52): for every user
//     train_test_split(user_df, test_size=.2, random_state=seed), then train_test_split(rest, test_size=.25, random_state=seed)
// (sklearn ShuffleSplit: n_test = ceil(test_size * n), a permutation from np.random.RandomState(seed).permutation(n), test =
// perm[:n_test], train = perm[n_test:]). Every call builds a FRESH RandomState(seed), so the permutation is a function of n
// alone: one table row per list length. RandomState(int) is Mersenne Twister MT19937 seeded with init_genrand; legacy
// permutation(n) = Fisher-Yates from the top, j = random_interval(i): 32-bit draws masked to the smallest 2^k - 1 >= i,
// rejected while > i (numpy/random/src/distributions: random_interval; legacy shuffle in mtrand.pyx).
//   yr_split_perm_tables : the MT19937 output stream once (one thread), then one thread per length n walks it
//   yr_split_per_user    : one warp per user scatters its items into the train / valid / test lists
// Bit-identical to data/synthetic.py::split_per_user, which is pinned to the reference's own split (tests/test_host_logic.py).
#include "common.cuh"

namespace yr {

__global__ void mt19937_stream_kernel(uint32_t seed, uint32_t* __restrict__ out, int n_out) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  constexpr int N = 624, M = 397;
  uint32_t mt[N];
  mt[0] = seed;
  for (int i = 1; i < N; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
  int produced = 0;
  while (produced < n_out) {
    for (int k = 0; k < N; ++k) {                       // regenerate the whole block (genrand's lazy twist)
      const uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % N] & 0x7fffffffu);
      mt[k] = mt[(k + M) % N] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    for (int k = 0; k < N && produced < n_out; ++k) {
      uint32_t y = mt[k];
      y ^= (y >> 11);
      y ^= (y << 7) & 0x9d2c5680u;
      y ^= (y << 15) & 0xefc60000u;
      y ^= (y >> 18);
      out[produced++] = y;
    }
  }
}

// perm[n * ld .. + n) = RandomState(seed).permutation(n); *err = 1 if the stream was too short (never with the sizing
// yr_split_ws_bytes uses: the expected number of draws is < 2 n, the stream holds 4 n + 1024)
__global__ void __launch_bounds__(128)
perm_tables_kernel(const uint32_t* __restrict__ stream, int n_stream, int max_n, int ld, int32_t* __restrict__ perm, int32_t* err) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < 1 || n > max_n) return;
  int32_t* a = perm + (int64_t)n * ld;
  for (int i = 0; i < n; ++i) a[i] = i;
  int pos = 0;
  for (int i = n - 1; i >= 1; --i) {
    uint32_t mask = (uint32_t)i;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    uint32_t v;
    do {
      if (pos >= n_stream) { atomicExch(err, 1); return; }
      v = stream[pos++] & mask;
    } while (v > (uint32_t)i);
    const int32_t t = a[i]; a[i] = a[v]; a[v] = t;
  }
}

__device__ __forceinline__ void split_sizes(int n, int& n_train, int& n_valid, int& n_test) {
  n_test = (int)ceil(0.2 * (double)n);                  // sklearn _validate_shuffle_split: ceil(test_size * n_samples)
  const int n_rest = n - n_test;
  n_valid = (int)ceil(0.25 * (double)n_rest);
  n_train = n_rest - n_valid;
}

__global__ void __launch_bounds__(256)
split_count_kernel(const int32_t* __restrict__ ptr, int64_t n_users, int32_t* __restrict__ n_tr, int32_t* __restrict__ n_va,
                   int32_t* __restrict__ n_te) {
  for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < n_users; u += (int64_t)gridDim.x * blockDim.x) {
    const int n = ptr[u + 1] - ptr[u];
    int a = 0, b = 0, c = 0;
    if (n > 0) split_sizes(n, a, b, c);
    n_tr[u] = a; n_va[u] = b; n_te[u] = c;
  }
}

__global__ void __launch_bounds__(256)
split_scatter_kernel(const int32_t* __restrict__ ptr, const int64_t* __restrict__ items, int64_t n_users,
                     const int32_t* __restrict__ perm, int ld, const int32_t* __restrict__ tr_ptr,
                     const int32_t* __restrict__ va_ptr, const int32_t* __restrict__ te_ptr, int64_t* __restrict__ tr,
                     int64_t* __restrict__ va, int64_t* __restrict__ te) {
  const int lane = threadIdx.x & 31;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t u = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; u < n_users; u += nw) {
    const int s = ptr[u], n = ptr[u + 1] - s;
    if (n <= 0) continue;
    int n_train, n_valid, n_test;
    split_sizes(n, n_train, n_valid, n_test);
    const int n_rest = n - n_test;
    const int32_t* p1 = perm + (int64_t)n * ld;          // first split: test = p1[:n_test], rest = p1[n_test:]
    const int32_t* p2 = perm + (int64_t)n_rest * ld;     // second split of the rest (a fresh RandomState(seed)): valid = rest[p2[:n_valid]]
    for (int k = lane; k < n_test; k += 32) te[te_ptr[u] + k] = items[s + p1[k]];
    for (int k = lane; k < n_valid; k += 32) va[va_ptr[u] + k] = items[s + p1[n_test + p2[k]]];
    for (int k = lane; k < n_train; k += 32) tr[tr_ptr[u] + k] = items[s + p1[n_test + p2[n_valid + k]]];
  }
}

}  // namespace yr

using namespace yr;

static size_t split_ws_layout(int max_n, void* base, uint32_t** stream, int32_t** perm, int* n_stream, int* ld) {
  const int ns = 4 * max_n + 1024, l = (max_n + 3) / 4 * 4;
  const size_t o_stream = 0, o_perm = ((size_t)ns * 4 + 255) / 256 * 256;
  const size_t total = o_perm + (size_t)(max_n + 1) * l * 4;
  if (base) { *stream = (uint32_t*)((char*)base + o_stream); *perm = (int32_t*)((char*)base + o_perm); }
  if (n_stream) *n_stream = ns;
  if (ld) *ld = l;
  return total;
}

extern "C" size_t yr_split_ws_bytes(int max_list_len) {
  if (max_list_len < 1) return 0;
  return split_ws_layout(max_list_len, nullptr, nullptr, nullptr, nullptr, nullptr);
}

extern "C" int yr_split_sizes(const int32_t* ptr, int64_t num_users, int32_t* n_train, int32_t* n_valid, int32_t* n_test,
                              yr_stream stream) {
  if (!ptr || !n_train || !n_valid || !n_test || num_users < 0) return YR_ERR_BAD_ARG;
  if (num_users == 0) return YR_OK;
  int64_t blocks = (num_users + 255) / 256;
  if (blocks > (int64_t)yr_sm_count() * 8) blocks = (int64_t)yr_sm_count() * 8;
  split_count_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(ptr, num_users, n_train, n_valid, n_test);
  YR_CHECK_LAUNCH();
  return YR_OK;
}

extern "C" int yr_split_per_user(const int32_t* ptr, const int64_t* items, int64_t num_users, int max_list_len, uint32_t seed,
                                 const int32_t* train_ptr, const int32_t* valid_ptr, const int32_t* test_ptr,
                                 int64_t* train_items, int64_t* valid_items, int64_t* test_items, void* ws, size_t ws_bytes,
                                 int32_t* err, yr_stream stream) {
  if (!ptr || !items || !train_ptr || !valid_ptr || !test_ptr || !train_items || !valid_items || !test_items || !ws || !err ||
      num_users < 0 || max_list_len < 1)
    return YR_ERR_BAD_ARG;
  uint32_t* strm; int32_t* perm; int n_stream, ld;
  if (ws_bytes < split_ws_layout(max_list_len, ws, &strm, &perm, &n_stream, &ld)) return YR_ERR_WORKSPACE;
  if (num_users == 0) return YR_OK;
  cudaStream_t s = (cudaStream_t)stream;
  mt19937_stream_kernel<<<1, 1, 0, s>>>(seed, strm, n_stream);
  YR_CHECK_LAUNCH();
  perm_tables_kernel<<<(max_list_len + 128) / 128, 128, 0, s>>>(strm, n_stream, max_list_len, ld, perm, err);
  YR_CHECK_LAUNCH();
  int64_t blocks = (num_users * 32 + 255) / 256;
  if (blocks > (int64_t)yr_sm_count() * 16) blocks = (int64_t)yr_sm_count() * 16;
  split_scatter_kernel<<<(unsigned)blocks, 256, 0, s>>>(ptr, items, num_users, perm, ld, train_ptr, valid_ptr, test_ptr,
                                                       train_items, valid_items, test_items);
  YR_CHECK_LAUNCH();
  return YR_OK;
}
