// NGCF dense transforms on the 5th-gen tensor cores (sm_100a), forward:
//     E_next = leaky_relu( [LE+E | E*LE] . [W1 | W2]^T )          (reference models/ngcf.py:64-72)
// as tcgen05.mma.kind::tf32 with the 3xTF32 split (x = hi + lo, hi = top 11 significand bits):
//     A.B ~= A_lo.B_hi + A_hi.B_lo + A_hi.B_hi      (relative error ~2^-21 per product, fp32 accumulation in TMEM)
// which keeps the layer inside the 1e-5 parity bar while the 2*N*d*2d flops leave the FP32 pipe.
//
// Per CTA (persistent over 128-row tiles, one CTA per SM), 13 warps:
//   warps 0-7 : loaders   — float4 reads of E, LE one 32-column slab at a time (two slabs in flight per thread, the next tile's
//                           rows asked for in L2 by one bulk prefetch), S = LE+E, P = E*LE, split hi/lo, stores into the ring
//                           stage in the K-major 128B-swizzled layout the UMMA descriptor expects
//   warps 8-11: epilogue  — TMEM lane = row: tcgen05.ld of the two accumulators, LeakyReLU, rows leave through a per-warp
//                           shared-memory transpose as whole 256-byte segments
//   warp  12  : TMEM allocator + MMA issuer (one elected lane): 12 MMAs (M = 128, N = d, K = 8) per ring stage
// W1/W2 are nn.Linear weights [out x in] = K-major B operands as stored. Two TMEM accumulator stages let the epilogue of
// tile t overlap the loads + MMAs of tile t+1. (Round 1's kernel handed a whole tile from the loaders to the MMA warp at once.)
#include <stdlib.h>
#include "tc_common.cuh"

namespace yr {

constexpr int kFwdThreads = 416;   // 8 loader warps + 4 epilogue warps + 1 MMA/TMEM warp
constexpr int kFwdTM = 128;

// ---------------------------------------------------------------------------------------------------------------------
// Ring version (round 2), d = 64 and d = 128. The K dimension (2d columns of [S | P]) is cut into 32-float slabs (one
// 128-byte swizzle row each); a ring STAGE holds one operand slab pair:
//     A: X_hi, X_lo  [128 rows x 32]   X = S = LE + E  or  X = P = E * LE, columns 32 j .. 32 j + 31          2 x 16 KB
//     B: W_X[:, 32 j ..] hi, lo  [d x 32]   — resident in shared memory at d = 64 (64 KB for all slabs); at d = 128 the
//        256 KB of split weights do not fit, so the slab travels with the stage (cp.async.bulk from a pre-split,
//        pre-swizzled workspace that ngcf_split_weights_kernel fills once per launch; L2-resident)          2 x 16 KB
// and feeds 12 MMAs (M = 128, N = d, K = 8; lo.hi, hi.lo, hi.hi). Loaders run up to kNst stages ahead of the MMA warp
// with two slabs of global loads in flight per thread, so staging, MMAs and the epilogue of the previous tile overlap
// inside a tile as well as across tiles.
// ---------------------------------------------------------------------------------------------------------------------
template <int D>
struct FwdTc {
  static_assert(D == 64 || D == 128, "tensor-core forward: d in {64, 128}");
  static constexpr bool kBRes = (D == 64);
  static constexpr int kSlabs = D / 32;
  static constexpr uint32_t kABytes = 128 * 128;                 // one [128 x 32] fp32 block
  static constexpr uint32_t kBBytes = D * 128;                   // one [d x 32] fp32 block
  static constexpr uint32_t kStage = 2 * kABytes + (kBRes ? 0u : 2 * kBBytes);
  static constexpr int kNst = kBRes ? 4 : 3;
  static constexpr uint32_t kRing = kNst * kStage;
  static constexpr uint32_t kBResBytes = kBRes ? (uint32_t)(2 * kSlabs * 2) * kBBytes : 0u;   // [X][j][hi/lo]
  static constexpr uint32_t kEpiOff = kRing + kBResBytes;          // 4 epilogue warps x [32 rows x 64 floats] = 32 KB
  static constexpr uint32_t kEpiBytes = 4 * 32 * 256;
  static constexpr uint32_t kBarOff = kEpiOff + kEpiBytes;
  static constexpr size_t kSmem = (size_t)kBarOff + 256 + 1024;
  static constexpr uint32_t kTmemCols = 4 * D;                     // 2 stages x (hi.hi accumulator + small-terms accumulator)
  static constexpr size_t kWsBytes = (size_t)(2 * kSlabs) * 2 * kBBytes;                      // [j][X][hi/lo]
};

__device__ __forceinline__ void mbar_expect_tx_only(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(s2u(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_tc(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s2u(dst)),
               "l"(src), "r"(bytes), "r"(s2u(bar))
               : "memory");
}

// [W1 | W2] -> hi / lo, K-major 128B-swizzled [d x 32] blocks in the order the ring consumes them: block (j, X)
template <int D>
__global__ void __launch_bounds__(256)
ngcf_split_weights_kernel(const float* __restrict__ W1, const float* __restrict__ W2, unsigned char* __restrict__ ws) {
  using C = FwdTc<D>;
  const int total = 2 * D * (D / 4);
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int X = idx / (D * (D / 4)), rem = idx % (D * (D / 4));
    const int o = rem / (D / 4), c4 = rem % (D / 4), j = c4 >> 3, c = c4 & 7;
    float4 hi, lo;
    split4(__ldg(reinterpret_cast<const float4*>((X ? W2 : W1) + o * D) + c4), hi, lo);
    unsigned char* blk = ws + (size_t)(j * 2 + X) * 2 * C::kBBytes;
    const uint32_t off = sw_off(D, o, c);
    *reinterpret_cast<float4*>(blk + off) = hi;
    *reinterpret_cast<float4*>(blk + C::kBBytes + off) = lo;
  }
}

template <int D>
__global__ void __launch_bounds__(kFwdThreads, 1)
ngcf_dense_fwd_tc_kernel(const float* __restrict__ E, const float* __restrict__ LE, const float* __restrict__ W1,
                         const float* __restrict__ W2, const unsigned char* __restrict__ wsplit, float slope, int64_t n,
                         float* __restrict__ Eout, const int32_t* __restrict__ row_list,
                         const int32_t* __restrict__ row_count) {
  using C = FwdTc<D>;
  constexpr int kSlabs = C::kSlabs, kNst = C::kNst;
  constexpr uint32_t kABytes = C::kABytes, kBBytes = C::kBBytes, kStage = C::kStage;
  if (row_list) n = *row_count;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (s2u(smem_raw) & 1023u)) & 1023u);
  unsigned char* ring = smem;
  unsigned char* bres = smem + C::kRing;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kBarOff);
  uint64_t* full = bars; uint64_t* empty = bars + kNst; uint64_t* d_full = bars + 2 * kNst; uint64_t* d_empty = d_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + 2);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n_tiles = (n + kFwdTM - 1) / kFwdTM;
  const int64_t my_tiles = (int64_t)blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (tid == 0) {
    for (int s = 0; s < kNst; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(d_full + a, 1); mbar_init(d_empty + a, 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 12) tmem_alloc(tmem_slot, C::kTmemCols);  // two stages of (big, small) accumulators, d columns each
  if constexpr (C::kBRes) {
    // weights -> hi / lo K-major swizzled B blocks (row = output feature o, k contiguous), built once per CTA
    for (int idx = tid; idx < D * (D / 4); idx += kFwdThreads) {
      const int o = idx / (D / 4), c4 = idx % (D / 4), j = c4 >> 3, c = c4 & 7;
      const uint32_t off = sw_off(D, o, c);
      float4 hi, lo;
      split4(__ldg(reinterpret_cast<const float4*>(W1 + o * D) + c4), hi, lo);
      *reinterpret_cast<float4*>(bres + (size_t)((0 * kSlabs + j) * 2 + 0) * kBBytes + off) = hi;
      *reinterpret_cast<float4*>(bres + (size_t)((0 * kSlabs + j) * 2 + 1) * kBBytes + off) = lo;
      split4(__ldg(reinterpret_cast<const float4*>(W2 + o * D) + c4), hi, lo);
      *reinterpret_cast<float4*>(bres + (size_t)((1 * kSlabs + j) * 2 + 0) * kBBytes + off) = hi;
      *reinterpret_cast<float4*>(bres + (size_t)((1 * kSlabs + j) * 2 + 1) * kBBytes + off) = lo;
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // D = f32, A = B = tf32, K-major both, N = d, M = 128
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(D >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

  if (warp < 8) {
    // ================= loaders (256 threads) =================
    // slab j of a tile: thread t owns rows (t >> 3) + 32 i, i = 0..3, chunk c = t & 7 (a warp request = 4 rows x 128 B)
    const int c = tid & 7, rb = tid >> 3;
    const int64_t n_items = my_tiles * kSlabs;
    uint32_t s = 0, ph = 0;
    // the tile's 128 rows of E and of LE are contiguous: one thread asks for them in L2 one tile ahead (bulk prefetch), so the
    // register loads below see L2 latency instead of HBM latency (two slabs in flight per thread is all the register file
    // allows). Measured at 1.5 M x 128 (profiles/r02_dense_ring.txt): 775 us without, 655 us one tile ahead, 717 / 809 us
    // two / three tiles ahead (148 SMs x 3 tiles x 128 KB no longer survive in L2 next to the output stream: DRAM reads
    // double).
    constexpr int kPfTiles = 1;
    auto prefetch_tile = [&](int64_t ti) {
      if (row_list || ti >= my_tiles) return;
      const int64_t r0 = (blockIdx.x + ti * gridDim.x) * kFwdTM;
      const int64_t rows = (n - r0) < kFwdTM ? (n - r0) : kFwdTM;
      l2_prefetch_bulk(E + r0 * D, (uint32_t)(rows * D * 4));
      l2_prefetch_bulk(LE + r0 * D, (uint32_t)(rows * D * 4));
    };
    auto issue = [&](int64_t item, float4 (&e)[4], float4 (&le)[4]) {
      const int64_t tile = blockIdx.x + (item / kSlabs) * gridDim.x;
      const int j = (int)(item % kSlabs);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t q = tile * kFwdTM + rb + 32 * i;
        e[i] = make_float4(0.f, 0.f, 0.f, 0.f); le[i] = e[i];
        if (q < n) {
          const int64_t g = (row_list ? (int64_t)__ldg(row_list + q) : q) * (D / 4) + 8 * j + c;
          e[i] = __ldg(reinterpret_cast<const float4*>(E) + g);
          le[i] = __ldg(reinterpret_cast<const float4*>(LE) + g);
        }
      }
    };
    auto produce = [&](int j, int X, const float4 (&e)[4], const float4 (&le)[4]) {
      mbar_wait(empty + s, ph ^ 1);                      // the MMAs that read this stage have completed
      unsigned char* st = ring + (size_t)s * kStage;
      if constexpr (!C::kBRes) {
        if (tid == 0) {
          mbar_expect_tx_only(full + s, 2 * kBBytes);
          bulk_g2s_tc(st + 2 * kABytes, wsplit + (size_t)(j * 2 + X) * 2 * kBBytes, 2 * kBBytes, full + s);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 a = e[i], b = le[i];
        const float4 v = X ? make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w)
                           : make_float4(b.x + a.x, b.y + a.y, b.z + a.z, b.w + a.w);
        float4 hi, lo;
        split4(v, hi, lo);
        const uint32_t off = sw_off(kFwdTM, rb + 32 * i, c);
        *reinterpret_cast<float4*>(st + off) = hi;
        *reinterpret_cast<float4*>(st + kABytes + off) = lo;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");     // all 128 rows of the slab written
      if (tid == 0) mbar_arrive(full + s);
      if (++s == (uint32_t)kNst) { s = 0; ph ^= 1; }
    };
    auto process = [&](int64_t item, const float4 (&e)[4], const float4 (&le)[4]) {
      const int j = (int)(item % kSlabs);
      if (j == 0 && tid == 32) prefetch_tile(item / kSlabs + kPfTiles);
      produce(j, 0, e, le);
      produce(j, 1, e, le);
    };
    if (tid == 32)
      for (int ti = 0; ti < kPfTiles; ++ti) prefetch_tile(ti);
    float4 ea[4], la[4], eb[4], lb[4], ec[4], lc[4];
    if (n_items > 0) issue(0, ea, la);
    if (n_items > 1) issue(1, eb, lb);
    for (int64_t it = 0; it < n_items; it += 3) {
      if (it + 2 < n_items) issue(it + 2, ec, lc);
      process(it, ea, la);
      if (it + 1 < n_items) {
        if (it + 3 < n_items) issue(it + 3, ea, la);
        process(it + 1, eb, lb);
      }
      if (it + 2 < n_items) {
        if (it + 4 < n_items) issue(it + 4, eb, lb);
        process(it + 2, ec, lc);
      }
    }
  } else if (warp < 12) {
    // ================= epilogue: TMEM lane = row; rows leave through a per-warp shared-memory transpose =================
    // A thread holds one ROW of the accumulator (TMEM lane = row), but a row-per-thread global store touches 32 lines
    // per instruction. Each warp therefore parks its 32 rows x 64 columns in shared memory (16-byte slots XOR-ed with the
    // row: conflict-free both ways) and writes them back two whole 256-byte row segments per instruction.
    uint32_t acc = 0, aph = 0;
    const int wq = warp & 3;
    unsigned char* stg = smem + C::kEpiOff + (size_t)wq * 8192;
    for (int64_t t = 0; t < my_tiles; ++t) {
      const int64_t q0 = (blockIdx.x + t * gridDim.x) * kFwdTM + wq * 32;      // first row of this warp
      mbar_wait(d_full + acc, aph);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(wq * 32) << 16) + acc * 2 * D;
#pragma unroll 1
      for (int h = 0; h < D / 64; ++h) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float vb[32], vs[32];
          tmem_ld32(trow + 64 * h + 32 * half, vb);          // hi.hi terms
          tmem_ld32(trow + D + 64 * h + 32 * half, vs);      // lo.hi + hi.lo terms
          if (h == D / 64 - 1 && half == 1) {
            tc_fence_before();
            mbar_arrive(d_empty + acc);                      // accumulator stage free again
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float4 o;
            o.x = vb[4 * k + 0] + vs[4 * k + 0]; o.y = vb[4 * k + 1] + vs[4 * k + 1];
            o.z = vb[4 * k + 2] + vs[4 * k + 2]; o.w = vb[4 * k + 3] + vs[4 * k + 3];
            o.x = o.x > 0.f ? o.x : o.x * slope; o.y = o.y > 0.f ? o.y : o.y * slope;
            o.z = o.z > 0.f ? o.z : o.z * slope; o.w = o.w > 0.f ? o.w : o.w * slope;
            *reinterpret_cast<float4*>(stg + lane * 256 + (((8 * half + k) ^ (lane & 15)) << 4)) = o;
          }
        }
        __syncwarp();
        {
#pragma unroll 4
          for (int i = 0; i < 16; ++i) {
            const int rr = 2 * i + (lane >> 4), c4 = lane & 15;
            const int64_t q = q0 + rr;
            const float4 o = *reinterpret_cast<const float4*>(stg + rr * 256 + ((c4 ^ (rr & 15)) << 4));
            if (q < n) {
              const int64_t row = row_list ? (int64_t)__ldg(row_list + q) : q;
              reinterpret_cast<float4*>(Eout + row * D + 64 * h)[c4] = o;
            }
          }
        }
        __syncwarp();
      }
      if (++acc == 2) { acc = 0; aph ^= 1; }
    }
  } else if (lane == 0) {
    // ================= MMA issuer =================
    uint32_t s = 0, ph = 0, acc = 0, aph = 0;
    for (int64_t t = 0; t < my_tiles; ++t) {
      mbar_wait(d_empty + acc, aph ^ 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + acc * 2 * D;        // [hi.hi accumulator | small-terms accumulator]
      uint32_t first_big = 1, first_small = 1;
#pragma unroll 1
      for (int jx = 0; jx < 2 * kSlabs; ++jx) {
        const int j = jx >> 1, X = jx & 1;
        mbar_wait(full + s, ph);
        tc_fence_after();
        const uint32_t a_hi = s2u(ring + (size_t)s * kStage), a_lo = a_hi + kABytes;
        const uint32_t b_hi = C::kBRes ? s2u(bres + (size_t)((X * kSlabs + j) * 2) * kBBytes) : a_hi + 2 * kABytes;
        const uint32_t b_lo = b_hi + kBBytes;
        // X_lo.W_hi and X_hi.W_lo go to their own accumulator (the tensor core's fp32 accumulation drops the low bits of
        // a small term added to a large partial sum: measured 2x the error when all three share one accumulator);
        // X_hi.W_hi to the other; the epilogue adds the two.
        const uint32_t pa[3] = {a_lo, a_hi, a_hi};
        const uint32_t pb[3] = {b_hi, b_lo, b_hi};
#pragma unroll
        for (int p = 0; p < 3; ++p) {
          const uint64_t ad = sw128_desc(pa[p]), bd = sw128_desc(pb[p]);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (p < 2) { umma_tf32(tacc + D, ad + 2 * ks, bd + 2 * ks, idesc, first_small ? 0u : 1u); first_small = 0; }
            else       { umma_tf32(tacc, ad + 2 * ks, bd + 2 * ks, idesc, first_big ? 0u : 1u); first_big = 0; }
          }
        }
        umma_commit(empty + s);
        if (++s == (uint32_t)kNst) { s = 0; ph ^= 1; }
      }
      umma_commit(d_full + acc);
      if (++acc == 2) { acc = 0; aph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) tmem_dealloc(tmem_base, C::kTmemCols);
}

}  // namespace yr

using namespace yr;

template <int D>
static int fwd_tc_launch(const float* E, const float* LE, const float* W1, const float* W2, float slope, int64_t n,
                         float* Eout, cudaStream_t s, const int32_t* row_list, const int32_t* row_count, int64_t row_cap,
                         int reserve_sms) {
  using C = FwdTc<D>;
  static yr::AttrOnce attr;
  { int rc_ = attr.set(ngcf_dense_fwd_tc_kernel<D>, (int)C::kSmem); if (rc_) return rc_; }
  const int64_t n_tiles = ((row_list ? row_cap : n) + kFwdTM - 1) / kFwdTM;
  int64_t grid = yr_sm_count() - reserve_sms;
  if (grid > n_tiles) grid = n_tiles;
  if (grid < 1) grid = 1;
  unsigned char* ws = nullptr;
  if (!C::kBRes) {                                        // stream-ordered scratch for the split weights, freed below
    static bool pool_set[64] = {};
    int dev = 0;
    YR_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !pool_set[dev]) {         // keep freed scratch in the pool across synchronisations
      cudaMemPool_t pool;
      unsigned long long keep = 64ull << 20;
      YR_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
      YR_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
      pool_set[dev] = true;
    }
    YR_CUDA(cudaMallocAsync((void**)&ws, C::kWsBytes, s));
    ngcf_split_weights_kernel<D><<<16, 256, 0, s>>>(W1, W2, ws);
  }
  ngcf_dense_fwd_tc_kernel<D><<<(unsigned)grid, kFwdThreads, C::kSmem, s>>>(E, LE, W1, W2, ws, slope, n, Eout, row_list,
                                                                            row_count);
  cudaError_t e = cudaGetLastError();
  if (ws) cudaFreeAsync(ws, s);
  return e == cudaSuccess ? YR_OK : (int)e;
}

// internal launcher used by yr_ngcf_dense_fwd / yr_ngcf_train_step (ngcf.cu): d in {64, 128}
int yr_ngcf_dense_fwd_tc_launch_d(int d, const float* E, const float* LE, const float* W1, const float* W2, float slope,
                                  int64_t n, float* Eout, cudaStream_t s, const int32_t* row_list,
                                  const int32_t* row_count, int64_t row_cap, int reserve_sms) {
  if (d == 64) return fwd_tc_launch<64>(E, LE, W1, W2, slope, n, Eout, s, row_list, row_count, row_cap, reserve_sms);
  if (d == 128) return fwd_tc_launch<128>(E, LE, W1, W2, slope, n, Eout, s, row_list, row_count, row_cap, reserve_sms);
  return YR_ERR_BAD_ARG;
}
