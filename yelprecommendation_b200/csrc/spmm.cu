// CSR SpMM for the NGCF propagation (sm_100a): Y = A X / Y += A X, replacing torch.sparse.mm(L, E)
// (reference models/ngcf.py:64,67) and its transpose in the backward pass.
//
// Work decomposition (yr_csr plan): one CHUNK = <= YR_SPMM_CHUNK consecutive non-zeros of one row. A group of
// LPR = d/4 lanes owns a chunk (d = 64: half a warp, two chunks per warp), every lane carries one float4 of the
// row, so a neighbour row is ONE 256-byte request per group and 8 of them are kept in flight per group. Chunk
// descriptors are one 16-byte load. fma chain in CSR order inside a chunk; chunks of split rows store a partial
// and bump the row's split_count; the LAST chunk to arrive (atomic counter, threadfence) adds the row's partials left to
// right in the same launch, so the sum order is fixed whatever the arrival order (the oracle restates it -> bit-exact).
//
// At Yelp shape X (17.85 MB) is L2-resident; the 3.12 M gathered rows are 800 MB of L2->SM traffic per SpMM
// against 61 MB of algorithmic HBM traffic, so this kernel lives on L2 latency / bandwidth, not on HBM.
#include <stdlib.h>
#include "common.cuh"

namespace yr {

template <int D> struct SpmmCfg {
  static constexpr int kVec = D / 4;                       // float4 per row
  static constexpr int LPR = kVec >= 32 ? 32 : kVec;       // lanes per row (group width)
  static constexpr int VPT = kVec / LPR;                   // float4 per lane
  static constexpr int CPW = 32 / LPR;                     // chunks per warp
};

__device__ __forceinline__ void fma4(float4& acc, float a, const float4& x) {
  acc.x = fmaf(a, x.x, acc.x); acc.y = fmaf(a, x.y, acc.y);
  acc.z = fmaf(a, x.z, acc.z); acc.w = fmaf(a, x.w, acc.w);
}

// Last arriver of a split row: left-to-right sum of the row's chunk partials (fixed order whatever the arrival order).
// Kept out of line: it runs for a handful of rows per launch and must not cost the gather loop its registers.
template <int D, bool ACC>
__device__ __noinline__ void spmm_finish_split(const yr_csr& A, float* __restrict__ Y, int row, int split_idx, int sl) {
  using C = SpmmCfg<D>;
  constexpr int VPT = C::VPT;
  const int p0 = A.split_ptr[split_idx], p1 = A.split_ptr[split_idx + 1];
  const float4* P4 = reinterpret_cast<const float4*>(A.partials);
  float4 tot[VPT];
#pragma unroll
  for (int v = 0; v < VPT; ++v) tot[v] = __ldcg(P4 + (int64_t)p0 * C::kVec + sl * VPT + v);
  int p = p0 + 1;
  for (; p + 8 <= p1; p += 8) {
    float4 x[8][VPT];
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
      for (int v = 0; v < VPT; ++v) x[q][v] = __ldcg(P4 + (int64_t)(p + q) * C::kVec + sl * VPT + v);
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
      for (int v = 0; v < VPT; ++v) {
        tot[v].x += x[q][v].x; tot[v].y += x[q][v].y; tot[v].z += x[q][v].z; tot[v].w += x[q][v].w;
      }
  }
  for (; p < p1; ++p) {
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
      const float4 x = __ldcg(P4 + (int64_t)p * C::kVec + sl * VPT + v);
      tot[v].x += x.x; tot[v].y += x.y; tot[v].z += x.z; tot[v].w += x.w;
    }
  }
  float4* y4 = reinterpret_cast<float4*>(Y + (int64_t)row * D);
#pragma unroll
  for (int v = 0; v < VPT; ++v) {
    if (ACC) {
      const float4 y = y4[sl * VPT + v];
      tot[v].x = y.x + tot[v].x; tot[v].y = y.y + tot[v].y; tot[v].z = y.z + tot[v].z; tot[v].w = y.w + tot[v].w;
    }
    y4[sl * VPT + v] = tot[v];
  }
}

// THREADS / MINB: CTA size and CTAs per SM the register allocation is bounded for; U: neighbour rows in flight per group;
// DIRECT: every lane reads the column index / value of a non-zero itself (one broadcast L1 access per group) instead of
// one coalesced read per LPR non-zeros followed by shuffles. The summation order is the same for every variant.
template <int D, bool ACC, int THREADS, int MINB, int U, bool DIRECT>
__global__ void __launch_bounds__(THREADS, MINB)
spmm_chunk_kernel(yr_csr A, const float* __restrict__ X, float* __restrict__ Y, const int32_t* __restrict__ row_flag) {
  using C = SpmmCfg<D>;
  constexpr int LPR = C::LPR, VPT = C::VPT, CPW = C::CPW;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, sl = lane % LPR;
  const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int4* __restrict__ desc = reinterpret_cast<const int4*>(A.chunk_desc);
  const float4* __restrict__ X4 = reinterpret_cast<const float4*>(X);

  for (int cb = gwarp * CPW; cb < A.n_chunks; cb += nwarps * CPW) {
    const int c = cb + sub;
    bool c_valid = c < A.n_chunks;
    int4 dsc = make_int4(0, 0, 0, -1);
    if (c < A.n_chunks) dsc = __ldg(desc + c);
    // row subset (last layer of a BPR step): chunks of unflagged rows behave like padding; all chunks of a row share
    // the flag, so the split-row arrival count stays consistent
    if (row_flag && c < A.n_chunks && !__ldg(row_flag + dsc.x)) { dsc = make_int4(0, 0, 0, -1); c_valid = false; }
    const int row = dsc.x, s = dsc.y, len = dsc.z & 0xff, slot = dsc.w;
    const int split_idx = (dsc.z & 0x7fffffff) >> 8;   // index into split_row / split_ptr / split_count (split chunks only)
    const bool big = dsc.z < 0;                  // hub row: its partials are summed by spmm_finish_big_kernel afterwards
    float4 acc[VPT];
#pragma unroll
    for (int v = 0; v < VPT; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ACC && slot < 0 && len >= 0 && c_valid) {
#pragma unroll
      for (int v = 0; v < VPT; ++v) acc[v] = reinterpret_cast<const float4*>(Y + (int64_t)row * D)[sl * VPT + v];
    }
    if constexpr (DIRECT) {
      for (int j0 = 0; j0 < len; j0 += U) {
        float4 x[U][VPT];
        float a[U];
        int cq[U];
#pragma unroll
        for (int q = 0; q < U; ++q) {
          const bool in = j0 + q < len;
          cq[q] = in ? __ldg(A.col + s + j0 + q) : -1;
          a[q] = in ? __ldg(A.val + s + j0 + q) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < U; ++q) {
          if (cq[q] >= 0) {
#pragma unroll
            for (int v = 0; v < VPT; ++v) x[q][v] = __ldg(X4 + (int64_t)cq[q] * C::kVec + sl * VPT + v);
          }
        }
#pragma unroll
        for (int q = 0; q < U; ++q) {
          if (cq[q] >= 0) {
#pragma unroll
            for (int v = 0; v < VPT; ++v) fma4(acc[v], a[q], x[q][v]);
          }
        }
      }
    } else {
      int maxlen = len;
#pragma unroll
      for (int o = LPR; o < 32; o <<= 1) maxlen = max(maxlen, __shfl_xor_sync(kFull, maxlen, o));
      for (int j0 = 0; j0 < maxlen; j0 += LPR) {
        const int j = j0 + sl;
        const int cc = (j < len) ? __ldg(A.col + s + j) : -1;
        const float aa = (j < len) ? __ldg(A.val + s + j) : 0.f;
#pragma unroll
        for (int t = 0; t < LPR; t += U) {
          if (j0 + t >= maxlen) break;
          float4 x[U][VPT];
          float a[U];
          int cq[U];
#pragma unroll
          for (int q = 0; q < U; ++q) {
            cq[q] = __shfl_sync(kFull, cc, t + q, LPR);
            a[q] = __shfl_sync(kFull, aa, t + q, LPR);
          }
#pragma unroll
          for (int q = 0; q < U; ++q) {
            if (cq[q] >= 0) {
#pragma unroll
              for (int v = 0; v < VPT; ++v) x[q][v] = __ldg(X4 + (int64_t)cq[q] * C::kVec + sl * VPT + v);
            }
          }
#pragma unroll
          for (int q = 0; q < U; ++q) {
            if (cq[q] >= 0) {
#pragma unroll
              for (int v = 0; v < VPT; ++v) fma4(acc[v], a[q], x[q][v]);
            }
          }
        }
      }
    }
    if (c_valid) {
      float4* dst = (slot < 0) ? reinterpret_cast<float4*>(Y + (int64_t)row * D)
                               : reinterpret_cast<float4*>(A.partials + (int64_t)slot * D);
#pragma unroll
      for (int v = 0; v < VPT; ++v) dst[sl * VPT + v] = acc[v];
    }
    // ---- split rows: the chunk that arrives LAST sums the row's partials left to right (deterministic order).
    // Long rows' chunks are scheduled first (plan order), so this tail work overlaps the bulk of the kernel.
    const bool is_split = (c_valid) && slot >= 0 && !big;
    if (__any_sync(kFull, is_split)) {
      if (is_split) __threadfence();                       // my part of the partial is visible device-wide
      __syncwarp();
      int last = 0;
      if (is_split && sl == 0) {
        const int nparts = A.split_ptr[split_idx + 1] - A.split_ptr[split_idx];
        last = (atomicAdd(A.split_count + split_idx, 1) == nparts - 1) ? 1 : 0;
      }
      last = __shfl_sync(kFull, last, 0, LPR);
      if (last) {
        __threadfence();
        spmm_finish_split<D, ACC>(A, Y, row, split_idx, sl);
        if (sl == 0) A.split_count[split_idx] = 0;         // re-arm for the next call
      }
    }
  }
}

// Hub rows (more than YR_SPMM_BIG_CHUNKS chunks — a config-5 item row has up to ~28,000): one CTA per row sums the chunk
// partials in the SAME left-to-right order as spmm_finish_split, but the loads are decoupled from the dependent adds: all
// threads stage a 32 KB tile of partials in shared memory, the first D/4 threads add them. One warp walking the partials eight
// at a time took 3.5 ms for the largest row — a serial tail longer than a rank's whole row-panel SpMM on 8 GPUs.
constexpr int kBigBytes = 32 * 1024;     // static shared memory of the staging tile
template <int D, bool ACC>
__global__ void __launch_bounds__(256)
spmm_finish_big_kernel(yr_csr A, float* __restrict__ Y, const int32_t* __restrict__ row_flag) {
  constexpr int kVec = D / 4;
  constexpr int kBigTile = kBigBytes / (D * 4);                 // partials per tile: 128 at d = 64, 64 at d = 128
  __shared__ float4 tile[kBigTile * kVec];
  const int sidx = A.big_split_idx[blockIdx.x];
  const int row = A.split_row[sidx];
  if (row_flag && !row_flag[row]) return;
  const int p0 = A.split_ptr[sidx], p1 = A.split_ptr[sidx + 1];
  const float4* P4 = reinterpret_cast<const float4*>(A.partials);
  float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
  bool first = true;
  for (int t0 = p0; t0 < p1; t0 += kBigTile) {
    const int nt = min(kBigTile, p1 - t0);
    for (int i = threadIdx.x; i < nt * kVec; i += blockDim.x) tile[i] = __ldcg(P4 + (int64_t)t0 * kVec + i);
    __syncthreads();
    if (threadIdx.x < kVec) {
      for (int q = 0; q < nt; ++q) {
        const float4 x = tile[q * kVec + threadIdx.x];
        if (first) { tot = x; first = false; }
        else { tot.x += x.x; tot.y += x.y; tot.z += x.z; tot.w += x.w; }
      }
    }
    __syncthreads();
  }
  if (threadIdx.x < kVec) {
    float4* y4 = reinterpret_cast<float4*>(Y + (int64_t)row * D) + threadIdx.x;
    if (ACC) {
      const float4 y = *y4;
      tot.x = y.x + tot.x; tot.y = y.y + tot.y; tot.z = y.z + tot.z; tot.w = y.w + tot.w;
    }
    *y4 = tot;
  }
}

template <int D, int THREADS, int MINB, int U, bool DIRECT>
static int launch_spmm_v(const yr_csr* A, const float* X, float* Y, int accumulate, cudaStream_t s, const int32_t* row_flag) {
  using C = SpmmCfg<D>;
  const int wpb = THREADS / 32;
  int64_t blocks = ((int64_t)A->n_chunks + (int64_t)wpb * C::CPW - 1) / ((int64_t)wpb * C::CPW);
  const int64_t cap = (int64_t)yr_sm_count() * 8 * 32;
  if (blocks > cap) blocks = cap;
  if (THREADS == 1024) {           // reserve_sms: one persistent CTA per SM (1,024 threads x >= 33 registers exclude a second one)
    int64_t sms = (int64_t)yr_sm_count() - A->reserve_sms;
    if (sms < 1) sms = 1;
    if (blocks > sms) blocks = sms;
  }
  if (accumulate) spmm_chunk_kernel<D, true, THREADS, MINB, U, DIRECT><<<(unsigned)blocks, THREADS, 0, s>>>(*A, X, Y, row_flag);
  else spmm_chunk_kernel<D, false, THREADS, MINB, U, DIRECT><<<(unsigned)blocks, THREADS, 0, s>>>(*A, X, Y, row_flag);
  YR_CHECK_LAUNCH();
  if (A->n_big_rows > 0) {
    if (!A->big_split_idx) return YR_ERR_BAD_ARG;
    if (accumulate) spmm_finish_big_kernel<D, true><<<(unsigned)A->n_big_rows, 256, 0, s>>>(*A, Y, row_flag);
    else spmm_finish_big_kernel<D, false><<<(unsigned)A->n_big_rows, 256, 0, s>>>(*A, Y, row_flag);
    YR_CHECK_LAUNCH();
  }
  return YR_OK;
}

// YR_SPMM_VARIANT (experiments; the default is the measured best, profiles/README.md): same arithmetic in every variant.
static int spmm_variant() {
  const char* e = getenv("YR_SPMM_VARIANT");      // read per launch: a getenv is nanoseconds next to a kernel launch
  return (e && *e) ? atoi(e) : -1;
}

template <int D>
static int launch_spmm(const yr_csr* A, const float* X, float* Y, int accumulate, cudaStream_t s,
                       const int32_t* row_flag = nullptr) {
  // Default = the measured best of the sweeps in profiles/r02_spmm_sweep*.txt (B200, Yelp-shape graph, sorted plan): 128-thread
  // CTAs bounded to 48 registers (10 CTAs = 40 warps per SM), 4 neighbour rows in flight per group, direct index loads:
  // 53.8 us at d = 64 (14.9 TB/s of gathers) / 98.9 us at d = 128 (16.2 TB/s) against a measured L2 -> SM gather ceiling of
  // 17.6 TB/s (scripts/l2_gather_bench.cu); the round-1 shape (variant 0, row-order plan) took 85.5 / 216 us.
  if (A->reserve_sms > 0) return launch_spmm_v<D, 1024, 1, 4, true>(A, X, Y, accumulate, s, row_flag);
  switch (spmm_variant()) {
    case 0: return launch_spmm_v<D, 256, (D <= 64 ? 3 : 2), 8, false>(A, X, Y, accumulate, s, row_flag);
    case 1: return launch_spmm_v<D, 128, 8, 4, false>(A, X, Y, accumulate, s, row_flag);
    case 2: return launch_spmm_v<D, 128, 8, 4, true>(A, X, Y, accumulate, s, row_flag);
    case 3: return launch_spmm_v<D, 128, 8, 6, true>(A, X, Y, accumulate, s, row_flag);
    case 4: return launch_spmm_v<D, 64, 16, 4, true>(A, X, Y, accumulate, s, row_flag);
    default: return launch_spmm_v<D, 128, 10, 4, true>(A, X, Y, accumulate, s, row_flag);
  }
}

}  // namespace yr

using namespace yr;

extern "C" int yr_spmm_plan_size_h(const int32_t* rowptr_h, int64_t n_rows, int32_t* n_chunks_h,
                                   int32_t* n_split_rows_h, int32_t* n_partials_h) {
  if (!rowptr_h || n_rows < 0 || !n_chunks_h || !n_split_rows_h || !n_partials_h) return YR_ERR_BAD_ARG;
  int64_t chunks = 0, split = 0, parts = 0;
  for (int64_t r = 0; r < n_rows; ++r) {
    const int64_t len = rowptr_h[r + 1] - rowptr_h[r];
    const int64_t c = len <= YR_SPMM_CHUNK ? 1 : (len + YR_SPMM_CHUNK - 1) / YR_SPMM_CHUNK;
    chunks += c;
    if (c > 1) { ++split; parts += c; }
  }
  if (chunks >= (1LL << 31)) return YR_ERR_BAD_DIM;
  *n_chunks_h = (int32_t)chunks; *n_split_rows_h = (int32_t)split; *n_partials_h = (int32_t)parts;
  return YR_OK;
}

// Long rows' chunks first (they are the longest work items), then one chunk per short row.
extern "C" int yr_spmm_plan_fill_h(const int32_t* rowptr_h, int64_t n_rows, int32_t* chunk_desc_h,
                                   int32_t* split_row_h, int32_t* split_ptr_h) {
  if (!rowptr_h || n_rows < 0 || !chunk_desc_h || !split_ptr_h) return YR_ERR_BAD_ARG;
  int64_t c = 0, sr = 0, slot = 0;
  split_ptr_h[0] = 0;
  for (int64_t r = 0; r < n_rows; ++r) {
    const int64_t len = rowptr_h[r + 1] - rowptr_h[r];
    if (len <= YR_SPMM_CHUNK) continue;
    if (!split_row_h) return YR_ERR_BAD_ARG;
    for (int64_t s = rowptr_h[r]; s < rowptr_h[r + 1]; s += YR_SPMM_CHUNK) {
      const int64_t e = s + YR_SPMM_CHUNK < rowptr_h[r + 1] ? s + YR_SPMM_CHUNK : rowptr_h[r + 1];
      int32_t* d = chunk_desc_h + 4 * c;
      const int64_t n_ch = (len + YR_SPMM_CHUNK - 1) / YR_SPMM_CHUNK;
      d[0] = (int32_t)r; d[1] = (int32_t)s; d[3] = (int32_t)slot;
      d[2] = (int32_t)((uint32_t)(e - s) | ((uint32_t)sr << 8) | (n_ch > YR_SPMM_BIG_CHUNKS ? 0x80000000u : 0u));
      ++c; ++slot;
    }
    split_row_h[sr] = (int32_t)r;
    split_ptr_h[++sr] = (int32_t)slot;
  }
  // short rows are emitted longest first (counting sort by length), so that the chunks that share a warp / CTA have
  // equal lengths (85 -> 60 us on the Yelp-shape graph, profiles/r02_spmm_sweep.txt); YR_SPMM_PLAN_SORT=0 keeps row
  // order. Either way every row is one chain: results do not change.
  const char* e = getenv("YR_SPMM_PLAN_SORT");
  if (!(e && e[0] == '0')) {
    int64_t start[YR_SPMM_CHUNK + 2] = {0};
    for (int64_t r = 0; r < n_rows; ++r) {
      const int64_t len = rowptr_h[r + 1] - rowptr_h[r];
      if (len <= YR_SPMM_CHUNK) ++start[YR_SPMM_CHUNK - len + 1];      // bucket 0 = longest
    }
    for (int k = 0; k <= YR_SPMM_CHUNK; ++k) start[k + 1] += start[k];
    for (int64_t r = 0; r < n_rows; ++r) {
      const int64_t len = rowptr_h[r + 1] - rowptr_h[r];
      if (len > YR_SPMM_CHUNK) continue;
      int32_t* d = chunk_desc_h + 4 * (c + start[YR_SPMM_CHUNK - len]++);
      d[0] = (int32_t)r; d[1] = rowptr_h[r]; d[2] = (int32_t)len; d[3] = -1;
    }
    return YR_OK;
  }
  for (int64_t r = 0; r < n_rows; ++r) {
    const int64_t len = rowptr_h[r + 1] - rowptr_h[r];
    if (len > YR_SPMM_CHUNK) continue;
    int32_t* d = chunk_desc_h + 4 * c;
    d[0] = (int32_t)r; d[1] = rowptr_h[r]; d[2] = (int32_t)len; d[3] = -1;
    ++c;
  }
  return YR_OK;
}

extern "C" int yr_spmm_plan_big_h(const int32_t* rowptr_h, int64_t n_rows, int32_t* n_big_h, int32_t* big_split_idx_h) {
  if (!rowptr_h || n_rows < 0 || !n_big_h) return YR_ERR_BAD_ARG;
  int64_t sr = 0, nb = 0;
  for (int64_t r = 0; r < n_rows; ++r) {
    const int64_t len = rowptr_h[r + 1] - rowptr_h[r];
    if (len <= YR_SPMM_CHUNK) continue;
    if ((len + YR_SPMM_CHUNK - 1) / YR_SPMM_CHUNK > YR_SPMM_BIG_CHUNKS) {
      if (big_split_idx_h) big_split_idx_h[nb] = (int32_t)sr;
      ++nb;
    }
    ++sr;
  }
  if (sr >= (1LL << 23)) return YR_ERR_BAD_DIM;        // split-row index shares a descriptor word with the length and the flag
  *n_big_h = (int32_t)nb;
  return YR_OK;
}

int yr_spmm_csr_rows(const yr_csr* A, int d, const float* X, float* Y, const int32_t* row_flag, cudaStream_t s) {
  int rc = yr_csr_ok(A);
  if (rc) return rc;
  if (!X || !Y) return YR_ERR_BAD_ARG;
  if (A->n_rows == 0 || A->n_chunks == 0) return YR_OK;
  switch (d) {
    case 32: return launch_spmm<32>(A, X, Y, 0, s, row_flag);
    case 64: return launch_spmm<64>(A, X, Y, 0, s, row_flag);
    case 128: return launch_spmm<128>(A, X, Y, 0, s, row_flag);
    case 256: return launch_spmm<256>(A, X, Y, 0, s, row_flag);
    default: return YR_ERR_BAD_DIM;
  }
}

extern "C" int yr_spmm_csr(const yr_csr* A, int d, const float* X, float* Y, int accumulate, yr_stream stream) {
  int rc = yr_csr_ok(A);
  if (rc) return rc;
  if (!X || !Y) return YR_ERR_BAD_ARG;
  if (A->n_rows == 0 || A->n_chunks == 0) return YR_OK;
  cudaStream_t s = (cudaStream_t)stream;
  switch (d) {
    case 32: return launch_spmm<32>(A, X, Y, accumulate, s);
    case 64: return launch_spmm<64>(A, X, Y, accumulate, s);
    case 128: return launch_spmm<128>(A, X, Y, accumulate, s);
    case 256: return launch_spmm<256>(A, X, Y, accumulate, s);
    default: return YR_ERR_BAD_DIM;
  }
}
