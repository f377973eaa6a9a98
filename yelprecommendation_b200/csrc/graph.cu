// Device-side Laplacian / CSR builder (sm_100a) — SURVEY.md §8(f)3.
// Replaces NGCFDataPipeline._set_laplacian_matrix (reference data/datasets/ngcf_data_pipeline.py:19-44), which
// materialises two dense (U+I)^2 float32 arrays (19.4 GB each at Yelp shape) on the host:
//     A[u, U+i] = A[U+i, u] = mean rating of (u, i)   (pivot_table default aggfunc, :23-24; zeros dropped by to_sparse)
//     deg = A.sum(axis=0)                              (:34, fp32, accumulated row by row)
//     L = (D^-1/2 A) D^-1/2                            (:40-42, fp32, this association)
// Here: interactions (COO, any order, duplicates allowed) -> CSR of L with sorted columns, entirely on the device:
//   count -> scan -> fill (64-bit keys col<<32 | rating bits) -> per-row bitonic sort (warp / shared memory / global
//   scratch by row length) -> duplicate runs collapsed to their mean (double, like pandas), zeros dropped -> scan ->
//   emit -> degrees as sequential fp32 sums in column order (= the reference's row-by-row accumulation, by symmetry)
//   -> scale. Bit-exact against the host restatement data/graph.py::build_laplacian (tests/test_builders.py).
#include "common.cuh"

namespace yr {

constexpr int kScanTile = 1024;        // 256 threads x 4
constexpr int kSortSmallCap = 1024;    // shared-memory bitonic, 256 threads
constexpr int kSortLargeCap = 16384;   // shared-memory bitonic, 1024 threads (128 KB)
constexpr unsigned long long kPadKey = ~0ull;

__global__ void __launch_bounds__(256)
lap_count_kernel(const int64_t* __restrict__ user, const int64_t* __restrict__ item, int64_t nnz, int64_t nU,
                 int64_t nI, int32_t* cnt, int32_t* err) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t u = user[e], i = item[e];
    if (u < 0 || u >= nU || i < 0 || i >= nI) { if (err) atomicExch(err, 1); continue; }
    atomicAdd(cnt + u, 1);
    atomicAdd(cnt + nU + i, 1);
  }
}

// ---- exclusive scan of int32 (three small kernels; n up to 2^31) -------------------------------------------
__global__ void __launch_bounds__(256)
scan_tile_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out, int64_t n, int32_t* __restrict__ tile_sums) {
  __shared__ int32_t warp_tot[8];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + threadIdx.x * 4;
  int32_t v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = (base + j < n) ? in[base + j] : 0;
  const int32_t mine = v[0] + v[1] + v[2] + v[3];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int32_t inc = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int32_t t = __shfl_up_sync(kFull, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  int32_t wbase = 0;
  for (int k = 0; k < w; ++k) wbase += warp_tot[k];
  int32_t run = wbase + inc - mine;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (base + j < n) out[base + j] = run;
    run += v[j];
  }
  if (threadIdx.x == 255) tile_sums[blockIdx.x] = wbase + inc;
}

__global__ void __launch_bounds__(1024)
scan_sums_kernel(int32_t* tile_sums, int64_t n_tiles) {      // single block: exclusive scan in place
  __shared__ int32_t warp_tot[32];
  __shared__ int32_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int64_t b = 0; b < n_tiles; b += 1024) {
    const int64_t i = b + threadIdx.x;
    const int32_t x = (i < n_tiles) ? tile_sums[i] : 0;
    int32_t inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t t = __shfl_up_sync(kFull, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    int32_t wbase = 0;
    for (int k = 0; k < w; ++k) wbase += warp_tot[k];
    const int32_t carry = carry_s;
    if (i < n_tiles) tile_sums[i] = carry + wbase + inc - x;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + wbase + inc;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
scan_add_kernel(int32_t* __restrict__ out, int64_t n, const int32_t* __restrict__ tile_sums) {
  const int64_t base = (int64_t)blockIdx.x * kScanTile + threadIdx.x * 4;
  const int32_t add = tile_sums[blockIdx.x];
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (base + j < n) out[base + j] += add;
}

static int exclusive_scan(const int32_t* in, int32_t* out, int64_t n, int32_t* tile_sums, cudaStream_t s) {
  const int64_t n_tiles = (n + kScanTile - 1) / kScanTile;
  scan_tile_kernel<<<(unsigned)n_tiles, 256, 0, s>>>(in, out, n, tile_sums);
  YR_CHECK_LAUNCH();
  scan_sums_kernel<<<1, 1024, 0, s>>>(tile_sums, n_tiles);
  YR_CHECK_LAUNCH();
  scan_add_kernel<<<(unsigned)n_tiles, 256, 0, s>>>(out, n, tile_sums);
  YR_CHECK_LAUNCH();
  return YR_OK;
}

// ---- fill: every interaction takes a slot in row u (column U+i) and in row U+i (column u) --------------------
__global__ void __launch_bounds__(256)
lap_fill_kernel(const int64_t* __restrict__ user, const int64_t* __restrict__ item, const float* __restrict__ rating,
                int64_t nnz, int64_t nU, int64_t nI, const int32_t* __restrict__ ptr, int32_t* cursor,
                unsigned long long* __restrict__ keys) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t u = user[e], i = item[e];
    if (u < 0 || u >= nU || i < 0 || i >= nI) continue;
    const unsigned long long vb = (unsigned long long)__float_as_uint(rating[e]);
    const int32_t a = atomicAdd(cursor + u, 1), b = atomicAdd(cursor + nU + i, 1);
    keys[(int64_t)ptr[u] + a] = ((unsigned long long)(nU + i) << 32) | vb;
    keys[(int64_t)ptr[nU + i] + b] = ((unsigned long long)u << 32) | vb;
  }
}

// ---- per-row sort ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t pow2_ceil(int64_t x) { int64_t p = 1; while (p < x) p <<= 1; return p; }

// rows of <= 32 entries: one warp per row, bitonic network on registers; longer rows are queued by size class
__global__ void __launch_bounds__(256)
lap_sort_warp_kernel(const int32_t* __restrict__ ptr, int64_t n_rows, unsigned long long* __restrict__ keys,
                     int32_t* listA, int32_t* listB, int32_t* listC, int64_t* scratchC_off, int32_t* counters) {
  const int lane = threadIdx.x & 31;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rows; r += nw) {
    const int32_t s = ptr[r], len = ptr[r + 1] - s;
    if (len <= 1) continue;
    if (len > 32) {
      if (lane == 0) {
        if (len <= kSortSmallCap) listA[atomicAdd(counters + 0, 1)] = (int32_t)r;
        else if (len <= kSortLargeCap) listB[atomicAdd(counters + 1, 1)] = (int32_t)r;
        else {
          const int k = atomicAdd(counters + 2, 1);
          listC[k] = (int32_t)r;
          scratchC_off[k] = (int64_t)atomicAdd(reinterpret_cast<unsigned long long*>(scratchC_off - 1),
                                               (unsigned long long)pow2_ceil(len));
        }
      }
      continue;
    }
    unsigned long long x = (lane < len) ? keys[(int64_t)s + lane] : kPadKey;
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
      for (int j = k >> 1; j > 0; j >>= 1) {
        const unsigned long long y = __shfl_xor_sync(kFull, x, j);
        const bool up = ((lane & k) == 0), lower = ((lane & j) == 0);
        x = (lower == up) ? (x < y ? x : y) : (x > y ? x : y);
      }
    if (lane < len) keys[(int64_t)s + lane] = x;
  }
}

template <typename Ptr>
__device__ __forceinline__ void bitonic_block(Ptr a, int64_t P) {
  for (int64_t k = 2; k <= P; k <<= 1)
    for (int64_t j = k >> 1; j > 0; j >>= 1) {
      for (int64_t i = threadIdx.x; i < P; i += blockDim.x) {
        const int64_t p = i ^ j;
        if (p > i) {
          const unsigned long long x = a[i], y = a[p];
          if ((x > y) == ((i & k) == 0)) { a[i] = y; a[p] = x; }
        }
      }
      __syncthreads();
    }
}

// rows of 33 .. CAP entries: one CTA per row, bitonic network in shared memory
__global__ void lap_sort_smem_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ list,
                                     const int32_t* __restrict__ count, unsigned long long* __restrict__ keys) {
  extern __shared__ unsigned long long sk[];
  const int n_list = *count;
  for (int li = blockIdx.x; li < n_list; li += gridDim.x) {
    const int32_t r = list[li], s = ptr[r], len = ptr[r + 1] - s;
    const int64_t P = pow2_ceil(len);
    for (int64_t i = threadIdx.x; i < P; i += blockDim.x) sk[i] = (i < len) ? keys[(int64_t)s + i] : kPadKey;
    __syncthreads();
    bitonic_block(sk, P);
    for (int64_t i = threadIdx.x; i < len; i += blockDim.x) keys[(int64_t)s + i] = sk[i];
    __syncthreads();
  }
}

// rows longer than kSortLargeCap: one CTA per row on a power-of-two padded slice of global scratch
__global__ void __launch_bounds__(1024)
lap_sort_global_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ list, const int64_t* __restrict__ off,
                       const int32_t* __restrict__ count, unsigned long long* __restrict__ keys,
                       unsigned long long* __restrict__ scratch) {
  const int n_list = *count;
  for (int li = blockIdx.x; li < n_list; li += gridDim.x) {
    const int32_t r = list[li], s = ptr[r], len = ptr[r + 1] - s;
    const int64_t P = pow2_ceil(len);
    unsigned long long* a = scratch + off[li];
    for (int64_t i = threadIdx.x; i < P; i += blockDim.x) a[i] = (i < len) ? keys[(int64_t)s + i] : kPadKey;
    __syncthreads();
    bitonic_block(a, P);
    for (int64_t i = threadIdx.x; i < len; i += blockDim.x) keys[(int64_t)s + i] = a[i];
    __syncthreads();
  }
}

// ---- duplicates -> mean, zeros dropped. EMIT = false: count the surviving entries of each row ------------------
template <bool EMIT>
__global__ void __launch_bounds__(256)
lap_unique_kernel(const int32_t* __restrict__ ptr, int64_t n_rows, const unsigned long long* __restrict__ keys,
                  int32_t* __restrict__ cnt_out, const int32_t* __restrict__ rowptr, int32_t* __restrict__ col,
                  float* __restrict__ val) {
  const int lane = threadIdx.x & 31;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rows; r += nw) {
    const int64_t s = ptr[r];
    const int32_t len = ptr[r + 1] - (int32_t)s;
    int32_t kept = 0;
    for (int32_t i0 = 0; i0 < len; i0 += 32) {
      const int32_t i = i0 + lane;
      bool keep = false;
      uint32_t c = 0;
      float mean = 0.f;
      if (i < len) {
        const unsigned long long k = keys[s + i];
        c = (uint32_t)(k >> 32);
        const bool head = (i == 0) || ((uint32_t)(keys[s + i - 1] >> 32) != c);
        if (head) {                          // runs of duplicates are short: the head walks its run
          double sum = (double)__uint_as_float((uint32_t)k);
          int n = 1;
          for (int32_t q = i + 1; q < len; ++q) {
            const unsigned long long kq = keys[s + q];
            if ((uint32_t)(kq >> 32) != c) break;
            sum += (double)__uint_as_float((uint32_t)kq);
            ++n;
          }
          mean = (float)(sum / (double)n);   // pandas: float64 mean, then stored into the float32 matrix
          keep = (mean != 0.f);
        }
      }
      const unsigned m = __ballot_sync(kFull, keep);
      if (EMIT && keep) {
        const int64_t o = (int64_t)rowptr[r] + kept + __popc(m & ((1u << lane) - 1u));
        col[o] = (int32_t)c;
        val[o] = mean;
      }
      kept += __popc(m);
    }
    if (!EMIT && lane == 0) cnt_out[r] = kept;
  }
}

// deg[r] = fp32 sum of the row in column order; dinv = 1 / sqrt(deg) (numpy: float32 sqrt, then float32 divide)
__global__ void __launch_bounds__(256)
lap_degree_kernel(const int32_t* __restrict__ rowptr, const float* __restrict__ val, int64_t n_rows,
                  float* __restrict__ dinv) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * blockDim.x) {
    float d = 0.f;
    for (int32_t k = rowptr[r]; k < rowptr[r + 1]; ++k) d = __fadd_rn(d, val[k]);
    dinv[r] = __fdiv_rn(1.f, __fsqrt_rn(d));
  }
}

__global__ void __launch_bounds__(256)
lap_scale_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n_rows,
                 const float* __restrict__ dinv, float* __restrict__ val) {
  const int lane = threadIdx.x & 31;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rows; r += nw) {
    const float dr = dinv[r];
    for (int32_t k = rowptr[r] + lane; k < rowptr[r + 1]; k += 32)
      val[k] = __fmul_rn(__fmul_rn(dr, val[k]), dinv[col[k]]);
  }
}

struct LapWs {
  int32_t* cnt;        // [N + 1]
  int32_t* ptr;        // [N + 1]
  int32_t* cursor;     // [N + 1]  (also cnt2)
  int32_t* tile_sums;  // [(N + 1) / 1024 + 1]
  int32_t* listA; int32_t* listB; int32_t* listC;   // [N] each
  int64_t* scratch_off;                              // [1 + N]: [0] = running total, [1..] per listC entry
  int32_t* counters;   // [4]
  float* dinv;         // [N]
  unsigned long long* keys;     // [2 nnz]
  unsigned long long* scratch;  // [4 nnz] (power-of-two padded slices of the very long rows)
};

static size_t lap_ws_layout(int64_t nnz, int64_t N, void* base, LapWs* w) {
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
  const size_t o_cnt = take((size_t)(N + 1) * 4), o_ptr = take((size_t)(N + 1) * 4), o_cur = take((size_t)(N + 1) * 4);
  const size_t o_ts = take((size_t)((N + 1) / kScanTile + 2) * 4);
  const size_t o_a = take((size_t)N * 4), o_b = take((size_t)N * 4), o_c = take((size_t)N * 4);
  const size_t o_so = take((size_t)(N + 1) * 8), o_ctr = take(64), o_dinv = take((size_t)N * 4);
  const size_t o_keys = take((size_t)2 * nnz * 8), o_scr = take((size_t)4 * nnz * 8);
  if (base && w) {
    char* b = (char*)base;
    w->cnt = (int32_t*)(b + o_cnt); w->ptr = (int32_t*)(b + o_ptr); w->cursor = (int32_t*)(b + o_cur);
    w->tile_sums = (int32_t*)(b + o_ts);
    w->listA = (int32_t*)(b + o_a); w->listB = (int32_t*)(b + o_b); w->listC = (int32_t*)(b + o_c);
    w->scratch_off = (int64_t*)(b + o_so); w->counters = (int32_t*)(b + o_ctr); w->dinv = (float*)(b + o_dinv);
    w->keys = (unsigned long long*)(b + o_keys); w->scratch = (unsigned long long*)(b + o_scr);
  }
  return off;
}

}  // namespace yr

using namespace yr;

extern "C" size_t yr_laplacian_ws_bytes(int64_t nnz, int64_t num_users, int64_t num_items) {
  if (nnz < 0 || num_users <= 0 || num_items <= 0) return 0;
  return lap_ws_layout(nnz, num_users + num_items, nullptr, nullptr);
}

extern "C" int yr_laplacian_build(const int64_t* user, const int64_t* item, const float* rating, int64_t nnz,
                                  int64_t num_users, int64_t num_items, int32_t* rowptr, int32_t* col, float* val,
                                  void* ws, size_t ws_bytes, int32_t* err, yr_stream stream) {
  if (!user || !item || !rating || !rowptr || !col || !val || !ws || nnz <= 0 || num_users <= 0 || num_items <= 0)
    return YR_ERR_BAD_ARG;
  const int64_t N = num_users + num_items;
  if (2 * nnz >= (1LL << 31) || N >= (1LL << 31)) return YR_ERR_BAD_DIM;
  LapWs w;
  if (ws_bytes < lap_ws_layout(nnz, N, ws, &w)) return YR_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned sm = (unsigned)yr_sm_count();
  const unsigned g_e = (unsigned)((nnz + 255) / 256 < (int64_t)sm * 16 ? (nnz + 255) / 256 : (int64_t)sm * 16);
  const unsigned g_r = (unsigned)((N * 32 + 255) / 256 < (int64_t)sm * 16 ? (N * 32 + 255) / 256 : (int64_t)sm * 16);

  YR_CUDA(cudaMemsetAsync(w.cnt, 0, (size_t)(N + 1) * 4, s));
  YR_CUDA(cudaMemsetAsync(w.cursor, 0, (size_t)(N + 1) * 4, s));
  YR_CUDA(cudaMemsetAsync(w.counters, 0, 64, s));
  YR_CUDA(cudaMemsetAsync(w.scratch_off, 0, 8, s));
  lap_count_kernel<<<g_e, 256, 0, s>>>(user, item, nnz, num_users, num_items, w.cnt, err);
  YR_CHECK_LAUNCH();
  int rc = exclusive_scan(w.cnt, w.ptr, N + 1, w.tile_sums, s);
  if (rc) return rc;
  lap_fill_kernel<<<g_e, 256, 0, s>>>(user, item, rating, nnz, num_users, num_items, w.ptr, w.cursor, w.keys);
  YR_CHECK_LAUNCH();
  lap_sort_warp_kernel<<<g_r, 256, 0, s>>>(w.ptr, N, w.keys, w.listA, w.listB, w.listC, w.scratch_off + 1, w.counters);
  YR_CHECK_LAUNCH();
  static yr::AttrOnce attr;
  { int rc_ = attr.set(lap_sort_smem_kernel, kSortLargeCap * 8); if (rc_) return rc_; }
  lap_sort_smem_kernel<<<sm * 8, 256, kSortSmallCap * 8, s>>>(w.ptr, w.listA, w.counters + 0, w.keys);
  YR_CHECK_LAUNCH();
  lap_sort_smem_kernel<<<sm, 1024, kSortLargeCap * 8, s>>>(w.ptr, w.listB, w.counters + 1, w.keys);
  YR_CHECK_LAUNCH();
  lap_sort_global_kernel<<<sm, 1024, 0, s>>>(w.ptr, w.listC, w.scratch_off + 1, w.counters + 2, w.keys, w.scratch);
  YR_CHECK_LAUNCH();
  // surviving entries per row -> final rowptr -> emit
  int32_t* cnt2 = w.cursor;
  YR_CUDA(cudaMemsetAsync(cnt2, 0, (size_t)(N + 1) * 4, s));
  lap_unique_kernel<false><<<g_r, 256, 0, s>>>(w.ptr, N, w.keys, cnt2, nullptr, nullptr, nullptr);
  YR_CHECK_LAUNCH();
  rc = exclusive_scan(cnt2, rowptr, N + 1, w.tile_sums, s);
  if (rc) return rc;
  lap_unique_kernel<true><<<g_r, 256, 0, s>>>(w.ptr, N, w.keys, nullptr, rowptr, col, val);
  YR_CHECK_LAUNCH();
  lap_degree_kernel<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(rowptr, val, N, w.dinv);
  YR_CHECK_LAUNCH();
  lap_scale_kernel<<<g_r, 256, 0, s>>>(rowptr, col, N, w.dinv, val);
  YR_CHECK_LAUNCH();
  return YR_OK;
}
