// NGCF dense transforms on the 5th-gen tensor cores (sm_100a), backward of one layer on its rows
// (autograd of reference models/ngcf.py:64-72), d = 64 and d = 128:
//     dZ = G_next * leaky'(E_next)
//     [dS | dP] = dZ . [W1 | W2]                       GEMM1  (M = 128 rows, K = d)
//     T = dS + dP * E ;  G += dS + dP * LE             epilogue
//     [dW1 | dW2]^T = [S | P]^T . dZ                   GEMM2  (M = features, N = d, K = rows), accumulated in TMEM over the
//                                                              CTA's tiles, one partial per CTA (fixed-order reduce afterwards)
// Both GEMMs are tcgen05.mma.kind::tf32 with the 3xTF32 split (lo.hi + hi.lo + hi.hi, fp32 accumulation in TMEM).
//
// Operand layouts (notes/README.md, verified by notes/mn_test): GEMM1 reads dZ K-major (128B swizzle) and the weights AS
// STORED ([out][in] row-major = N contiguous) through an MN-major descriptor; GEMM2 reduces over ROWS, so both its operands
// are row-major tiles read MN-major (SWIZZLE_128B_BASE32B atoms of 32 columns x 4 rows).
//
// Round 2: a ring of operand STAGES instead of whole-tile hand-overs. A stage is 1 + XG pairs (hi, lo) of 16 KB blocks,
// XG = 1 at d = 64 (S and P side by side in one M = N = 128 block), 2 at d = 128:
//     GEMM1 stage j (K-slab of 32 output features o):  pair 0 = dZ[:, 32 j ..] K-major [128 x 32]
//                                                      pair 1 + x = W_x[32 j .., :] MN-major [32 x 128], cp.async.bulk from a
//                                                      pre-split, pre-swizzled workspace (ngcf_split_weights_bwd_kernel)
//     GEMM2 stage q (32 rows of the tile):             pair 0 = dZ[32 q .., :] MN-major [32 x d]
//                                                      pair 1 + x = S / P rows MN-major [32 x 128]
// 12 MMAs per pair 1 + x. d = 64: 3 stages of 64 KB; d = 128: 2 stages of 96 KB. Per tile 2 + 4 (d = 64) or 4 + 4 stages.
// The epilogue (4 warps, TMEM lane = row) parks 32 columns of dS and dP per warp in shared memory and re-reads them
// row-contiguous, so that E / LE / G are read and T / G written as whole 128-byte row segments (v1 read them one row per
// thread: 25 of its 85 us); it runs underneath the tile's GEMM2 MMAs.
//
// Per CTA (persistent over 128-row tiles, one CTA per SM), 416 threads: warps 0-7 loaders (one item of
// global loads in flight per thread), warps 8-11 epilogue, warp 12 TMEM allocator + MMA issuer (one elected lane).
#include <stdlib.h>
#include "tc_common.cuh"

namespace yr {

constexpr int kBwdThreads = 416;      // 8 loader warps + 4 epilogue warps + 1 MMA / TMEM warp
constexpr int kBwdTM = 128;
constexpr int kFlushTiles = 4;        // dW leaves TMEM every this many tiles (192 accumulation steps)

// byte offset of element (mn, k) in an MN-major SWIZZLE_128B_BASE32B tile: atoms of 32 mn x 4 k (512 B), atoms
// lbo apart along MN and sbo apart along K, the four 32-byte chunks of a k-row XOR-ed with k & 3
__device__ __forceinline__ uint32_t mn_off(int mn, int k, uint32_t lbo, uint32_t sbo) {
  return (uint32_t)(k >> 2) * sbo + (uint32_t)(mn >> 5) * lbo + (uint32_t)(k & 3) * 128u +
         ((((uint32_t)(mn & 31) >> 3) ^ (uint32_t)(k & 3)) << 5) + (uint32_t)(mn & 7) * 4u;
}
__device__ __forceinline__ uint64_t mn_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;      // leading byte offset: next 32-column atom
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;      // stride byte offset: next 4-row atom
  d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell)
  d |= (uint64_t)1 << 61;                           // SWIZZLE_128B_BASE32B
  return d;
}
__device__ __forceinline__ void mbar_expect_tx_b(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(s2u(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_b(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s2u(dst)),
               "l"(src), "r"(bytes), "r"(s2u(bar))
               : "memory");
}

template <int D>
struct BwdTc {
  static_assert(D == 64 || D == 128, "tensor-core backward: d in {64, 128}");
  static constexpr int kSlabs = D / 32;                      // K-slabs of GEMM1
  static constexpr int kXG = D == 128 ? 2 : 1;               // S / P as separate M = N = 128 groups, or one combined group
  static constexpr int kHalves = D / 64;                     // a 32-row GEMM2 chunk is loaded as this many items
  static constexpr uint32_t kBlk = 16384;                    // one [128 x 32] / [32 x 128] fp32 block
  static constexpr uint32_t kPair = 2 * kBlk;                // hi, lo
  static constexpr uint32_t kStage = (1 + kXG) * kPair;
  static constexpr int kNst = D == 128 ? 2 : 3;
  static constexpr uint32_t kRing = kNst * kStage;           // 192 KB
  static constexpr uint32_t kEpiOff = kRing;                 // 4 epilogue warps x [32 rows x (32 dS + 32 dP)] = 32 KB
  static constexpr uint32_t kEpiBytes = 4 * 32 * 256;
  static constexpr uint32_t kBarOff = kEpiOff + kEpiBytes;
  static constexpr size_t kSmem = (size_t)kBarOff + 256 + 1024;
  // TMEM: GEMM1 accumulators [dS | dP] (2 x 128 columns at d = 128, 128 columns at d = 64, where there is room for a
  // second stage so that the epilogue of tile t overlaps GEMM1 of tile t + 1), then the GEMM2 accumulators
  static constexpr int kAcc = D == 128 ? 1 : 2;
  static constexpr uint32_t kAccCols = D == 128 ? 256 : 128;
  static constexpr uint32_t kTmemCols = 512;
  static constexpr uint32_t kColG2 = 256;                    // first TMEM column of the GEMM2 accumulators
  static constexpr uint32_t kSboW = 2048, kSboSP = 2048, kSboZ = (D / 32) * 512, kLbo = 512;
  static constexpr int kItems = kSlabs + 4 * kHalves;        // loader items per tile
  static constexpr size_t kWsplitBytes = (size_t)kSlabs * kXG * kPair;
};

// [W1 | W2] -> hi / lo MN-major blocks in the order the ring consumes them: block j = rows 32 j .. 32 j + 31 of the weights
// (K index o), pair x = W_x (d = 128) or one pair with W1 in columns 0..63 and W2 in 64..127 (d = 64)
template <int D>
__global__ void __launch_bounds__(256)
ngcf_split_weights_bwd_kernel(const float* __restrict__ W1, const float* __restrict__ W2, unsigned char* __restrict__ ws) {
  using C = BwdTc<D>;
  const int total = 2 * D * (D / 4);
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int X = idx / (D * (D / 4)), rem = idx % (D * (D / 4));
    const int o = rem / (D / 4), c4 = rem % (D / 4), j = o >> 5, k = o & 31;
    float4 hi, lo;
    split4(__ldg(reinterpret_cast<const float4*>((X ? W2 : W1) + o * D) + c4), hi, lo);
    const int pair = C::kXG == 2 ? X : 0;
    const int mn = C::kXG == 2 ? 4 * c4 : X * D + 4 * c4;
    unsigned char* blk = ws + (size_t)(j * C::kXG + pair) * C::kPair;
    const uint32_t off = mn_off(mn, k, C::kLbo, C::kSboW);
    *reinterpret_cast<float4*>(blk + off) = hi;
    *reinterpret_cast<float4*>(blk + C::kBlk + off) = lo;
  }
}

template <int D>
__global__ void __launch_bounds__(kBwdThreads, 1)
ngcf_dense_bwd_tc_kernel(const float* __restrict__ E, const float* __restrict__ LE, const float* __restrict__ Enext,
                         const float* __restrict__ Gnext, const unsigned char* __restrict__ wsplit, float slope, int64_t n,
                         float* __restrict__ G, float* __restrict__ T, float* __restrict__ ws,
                         const int32_t* __restrict__ row_list, const int32_t* __restrict__ row_count) {
  using C = BwdTc<D>;
  constexpr int TM = kBwdTM, kNst = C::kNst, kXG = C::kXG, kSlabs = C::kSlabs, kHalves = C::kHalves, kItems = C::kItems;
  constexpr uint32_t kBlk = C::kBlk, kPair = C::kPair, kStage = C::kStage;
  if (row_list) n = *row_count;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (s2u(smem_raw) & 1023u)) & 1023u);
  unsigned char* ring = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kBarOff);
  uint64_t* full = bars; uint64_t* empty = bars + kNst; uint64_t* d1_full = bars + 2 * kNst; uint64_t* d1_empty = d1_full + 2;
  uint64_t* d2_full = d1_full + 4; uint64_t* d2_empty = d1_full + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d1_full + 6);
  constexpr int kAcc = C::kAcc;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n_tiles = (n + TM - 1) / TM;
  const int64_t my_tiles = (int64_t)blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (tid == 0) {
    for (int s = 0; s < kNst; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(d1_full + a, 1); mbar_init(d1_empty + a, 128); }
    mbar_init(d2_full, 1); mbar_init(d2_empty, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 12) tmem_alloc(tmem_slot, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // D = f32, A = B = tf32, M = 128; GEMM1: A K-major, B MN-major, N = 128; GEMM2: both MN-major, N = d
  const uint32_t idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(D >> 3) << 17) |
                          ((uint32_t)(128 >> 4) << 24);

  if (warp < 8) {
    // ================= loaders (256 threads) =================
    // item i of a tile: i < kSlabs: GEMM1 slab j = i — thread t owns rows (t >> 3) + 32 r, r = 0..3, 16-byte chunk t & 7 of
    // G_next / E_next; otherwise GEMM2 chunk q (32 rows), half h (d = 128: 16 rows each) — thread t owns two rows and one
    // 16-byte chunk of the whole row of G_next / E_next / E / LE. Eight float4 per item either way.
    const int64_t n_items = my_tiles * kItems;
    uint32_t s = 0, ph = 0;
    // (A tile-ahead bulk L2 prefetch, which pays in the forward kernel, loses here: a tile lasts ~20 us at d = 128 and the
    // five row blocks of the next tile do not survive that long in L2 — DRAM reads 4.3 -> 8.0 GB, 1.69 -> 1.92 ms at
    // 1.5 M x 128, profiles/r02_dense_ring.txt.)
    auto issue = [&](int64_t item, float4 (&x)[8]) {
      const int64_t tile = blockIdx.x + (item / kItems) * gridDim.x;
      const int it = (int)(item % kItems);
      if (it < kSlabs) {
        const int c = tid & 7, rb = tid >> 3;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int64_t q = tile * TM + rb + 32 * r;
          x[r] = make_float4(0.f, 0.f, 0.f, 0.f); x[4 + r] = x[r];
          if (q < n) {
            const int64_t g = (row_list ? (int64_t)__ldg(row_list + q) : q) * (D / 4) + 8 * it + c;
            x[r] = __ldg(reinterpret_cast<const float4*>(Gnext) + g);
            x[4 + r] = __ldg(reinterpret_cast<const float4*>(Enext) + g);
          }
        }
      } else {
        const int u = it - kSlabs, qc = u / kHalves, h = u % kHalves;
        const int c4 = tid & (D / 4 - 1), r0 = tid / (D / 4);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int64_t q = tile * TM + 32 * qc + 16 * h + r0 + (256 / (D / 4)) * r;
          x[r] = make_float4(0.f, 0.f, 0.f, 0.f); x[2 + r] = x[r]; x[4 + r] = x[r]; x[6 + r] = x[r];
          if (q < n) {
            const int64_t g = (row_list ? (int64_t)__ldg(row_list + q) : q) * (D / 4) + c4;
            x[r] = __ldg(reinterpret_cast<const float4*>(Gnext) + g);
            x[2 + r] = __ldg(reinterpret_cast<const float4*>(Enext) + g);
            x[4 + r] = __ldg(reinterpret_cast<const float4*>(E) + g);
            x[6 + r] = __ldg(reinterpret_cast<const float4*>(LE) + g);
          }
        }
      }
    };
    auto dz_of = [&](const float4& g, const float4& en) {
      float4 z;
      z.x = en.x > 0.f ? g.x : g.x * slope; z.y = en.y > 0.f ? g.y : g.y * slope;
      z.z = en.z > 0.f ? g.z : g.z * slope; z.w = en.w > 0.f ? g.w : g.w * slope;
      return z;
    };
    auto hand_over = [&]() {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");     // every loader's part of the stage is written
      if (tid == 0) mbar_arrive(full + s);
      if (++s == (uint32_t)kNst) { s = 0; ph ^= 1; }
    };
    auto process = [&](int64_t item, const float4 (&x)[8]) {
      const int it = (int)(item % kItems);
      unsigned char* st = ring + (size_t)s * kStage;
      if (it < kSlabs) {
        mbar_wait(empty + s, ph ^ 1);                    // the MMAs that read this stage have completed
        if (tid == 0) {
          mbar_expect_tx_b(full + s, kXG * kPair);
          bulk_g2s_b(st + kPair, wsplit + (size_t)it * kXG * kPair, kXG * kPair, full + s);
        }
        const int c = tid & 7, rb = tid >> 3;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          float4 hi, lo;
          split4(dz_of(x[r], x[4 + r]), hi, lo);
          const uint32_t off = sw_off(TM, rb + 32 * r, c);
          *reinterpret_cast<float4*>(st + off) = hi;
          *reinterpret_cast<float4*>(st + kBlk + off) = lo;
        }
        hand_over();
      } else {
        const int u = it - kSlabs, h = u % kHalves;
        if (h == 0) mbar_wait(empty + s, ph ^ 1);
        const int c4 = tid & (D / 4 - 1), r0 = tid / (D / 4);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int k = 16 * h + r0 + (256 / (D / 4)) * r;          // row inside the 32-row chunk = K index
          const float4 e = x[4 + r], le = x[6 + r];
          float4 hi, lo;
          split4(dz_of(x[r], x[2 + r]), hi, lo);
          const uint32_t oz = mn_off(4 * c4, k, C::kLbo, C::kSboZ);
          *reinterpret_cast<float4*>(st + oz) = hi;
          *reinterpret_cast<float4*>(st + kBlk + oz) = lo;
          split4(make_float4(le.x + e.x, le.y + e.y, le.z + e.z, le.w + e.w), hi, lo);
          const uint32_t os = mn_off(4 * c4, k, C::kLbo, C::kSboSP);
          *reinterpret_cast<float4*>(st + kPair + os) = hi;
          *reinterpret_cast<float4*>(st + kPair + kBlk + os) = lo;
          split4(make_float4(e.x * le.x, e.y * le.y, e.z * le.z, e.w * le.w), hi, lo);
          // P: its own pair (d = 128) or columns 64..127 of the combined block (d = 64)
          unsigned char* pp = st + (kXG == 2 ? 2 * kPair : kPair);
          const uint32_t op = kXG == 2 ? os : mn_off(D + 4 * c4, k, C::kLbo, C::kSboSP);
          *reinterpret_cast<float4*>(pp + op) = hi;
          *reinterpret_cast<float4*>(pp + kBlk + op) = lo;
        }
        if (h == kHalves - 1) hand_over();
      }
    };
    // one item ahead in registers; a third register set spills at the
    // 128 registers per thread that 13 warps leave (four warps share one 16 K-register scheduler partition)
    float4 xa[8], xb[8];
    if (n_items > 0) issue(0, xa);
    for (int64_t it = 0; it < n_items; it += 2) {
      if (it + 1 < n_items) issue(it + 1, xb);
      process(it, xa);
      if (it + 1 < n_items) {
        if (it + 2 < n_items) issue(it + 2, xa);
        process(it + 1, xb);
      }
    }
  } else if (warp < 12) {
    // ================= epilogue: TMEM lane = row; 32 columns of dS and dP at a time through a per-warp transpose =================
    const int wq = warp & 3;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    unsigned char* stg = smem + C::kEpiOff + (size_t)wq * 8192;
    uint32_t fph = 0;
    bool flushed = false;
    const int r = wq * 32 + lane;
    // dW accumulators -> this CTA's partial block. TMEM lane = feature (d = 64: S features in lanes 0..63 -> dW1, P features
    // -> dW2; d = 128: lanes = features, dW1^T in the first 128 columns, dW2^T in the next), column = o.
    auto drain_dw = [&]() {
#pragma unroll 1
      for (int x = 0; x < kXG; ++x) {
        float* my = ws + (size_t)blockIdx.x * 2 * D * D + (kXG == 2 ? (size_t)x * D * D + r : (size_t)(r >> 6) * D * D + (r & 63));
#pragma unroll 1
        for (int c0 = 0; c0 < D; c0 += 16) {
          float v[16];
          tmem_ld16(tmem_base + lane_base + C::kColG2 + x * 128 + c0, v);
          if (flushed) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += my[(size_t)(c0 + j) * D];
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) my[(size_t)(c0 + j) * D] = v[j];      // ws[x][o][i]: lanes write consecutive i
        }
      }
    };
    for (int64_t t = 0; t < my_tiles; ++t) {
      if (t > 0 && t % kFlushTiles == 0) {
        // the tensor core's fp32 accumulation truncates (measured: error ~ 3.5e-8 x accumulation steps), so the dW sums are
        // taken out of TMEM every kFlushTiles tiles and continued in fp32 round-to-nearest in the CTA's partial block
        mbar_wait(d2_full, fph);
        fph ^= 1;
        tc_fence_after();
        drain_dw();
        flushed = true;
        tc_fence_before();
        mbar_arrive(d2_empty);
      }
      const int64_t q0 = (blockIdx.x + t * gridDim.x) * TM + wq * 32;          // first row of this warp
      const int acc = (int)(t % kAcc);
      const uint32_t tacc = tmem_base + lane_base + acc * C::kAccCols;
      mbar_wait(d1_full + acc, (uint32_t)(t / kAcc) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < D; c0 += 32) {
        {
          float ds[32], dp[32];
          tmem_ld32(tacc + c0, ds);
          tmem_ld32(tacc + D + c0, dp);
          if (c0 == D - 32) {
            tc_fence_before();
            mbar_arrive(d1_empty + acc);                    // this dS / dP accumulator stage is free again
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            *reinterpret_cast<float4*>(stg + lane * 256 + ((k ^ (lane & 15)) << 4)) =
                make_float4(ds[4 * k], ds[4 * k + 1], ds[4 * k + 2], ds[4 * k + 3]);
            *reinterpret_cast<float4*>(stg + lane * 256 + (((8 + k) ^ (lane & 15)) << 4)) =
                make_float4(dp[4 * k], dp[4 * k + 1], dp[4 * k + 2], dp[4 * k + 3]);
          }
        }
        __syncwarp();
        // row-contiguous pass: lane -> row 4 i + (lane >> 3), 16-byte chunk lane & 7; four rows' loads in flight at a time
        const int k = lane & 7;
#pragma unroll 1
        for (int i0 = 0; i0 < 8; i0 += 4) {
          float4 ev[4], lv[4], gv[4];
          int64_t g4[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = 4 * (i0 + i) + (lane >> 3);
            const int64_t q = q0 + rr;
            g4[i] = -1;
            if (q < n) {
              g4[i] = ((row_list ? (int64_t)__ldg(row_list + q) : q) * D + c0) / 4 + k;
              ev[i] = __ldg(reinterpret_cast<const float4*>(E) + g4[i]);
              lv[i] = __ldg(reinterpret_cast<const float4*>(LE) + g4[i]);
              gv[i] = *(reinterpret_cast<const float4*>(G) + g4[i]);
            }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (g4[i] < 0) continue;
            const int rr = 4 * (i0 + i) + (lane >> 3);
            const float4 ds = *reinterpret_cast<const float4*>(stg + rr * 256 + ((k ^ (rr & 15)) << 4));
            const float4 dp = *reinterpret_cast<const float4*>(stg + rr * 256 + (((8 + k) ^ (rr & 15)) << 4));
            float4 tt, gg = gv[i];
            tt.x = fmaf(dp.x, ev[i].x, ds.x); tt.y = fmaf(dp.y, ev[i].y, ds.y);
            tt.z = fmaf(dp.z, ev[i].z, ds.z); tt.w = fmaf(dp.w, ev[i].w, ds.w);
            gg.x += fmaf(dp.x, lv[i].x, ds.x); gg.y += fmaf(dp.y, lv[i].y, ds.y);
            gg.z += fmaf(dp.z, lv[i].z, ds.z); gg.w += fmaf(dp.w, lv[i].w, ds.w);
            reinterpret_cast<float4*>(T)[g4[i]] = tt;
            reinterpret_cast<float4*>(G)[g4[i]] = gg;
          }
        }
        __syncwarp();
      }
    }
    if (my_tiles > 0) {
      mbar_wait(d2_full, fph);
      tc_fence_after();
      drain_dw();
    } else {                                     // a CTA without a tile (short row list) contributes zeros
      for (int idx = r; idx < 2 * D * D; idx += 128) ws[(size_t)blockIdx.x * 2 * D * D + idx] = 0.f;
    }
  } else if (lane == 0) {
    // ================= MMA issuer =================
    uint32_t s = 0, ph = 0, first2 = 1, n_flush = 0;
    for (int64_t t = 0; t < my_tiles; ++t) {
      // ---- GEMM1: [dS | dP] = dZ . [W1 | W2]
      const int acc = (int)(t % kAcc);
      mbar_wait(d1_empty + acc, ((uint32_t)(t / kAcc) & 1u) ^ 1u);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < kSlabs; ++j) {
        mbar_wait(full + s, ph);
        tc_fence_after();
        const uint32_t st = s2u(ring + (size_t)s * kStage);
#pragma unroll
        for (int x = 0; x < kXG; ++x) {
          const uint32_t a[2] = {st, st + kBlk}, b[2] = {st + (1 + x) * kPair, st + (1 + x) * kPair + kBlk};
          const int pa[3] = {1, 0, 0}, pb[3] = {0, 1, 0};           // lo.hi, hi.lo, hi.hi
#pragma unroll
          for (int p = 0; p < 3; ++p) {
            const uint64_t ad = sw128_desc(a[pa[p]]);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t bd = mn_desc(b[pb[p]] + (uint32_t)(ks * 2) * C::kSboW, C::kLbo, C::kSboW);
              umma_tf32(tmem_base + acc * C::kAccCols + x * 128, ad + 2 * ks, bd, idesc1, (j == 0 && p == 0 && ks == 0) ? 0u : 1u);
            }
          }
        }
        umma_commit(empty + s);
        if (++s == (uint32_t)kNst) { s = 0; ph ^= 1; }
      }
      umma_commit(d1_full + acc);
      // ---- GEMM2: [dW1 | dW2]^T += [S | P]^T . dZ, four 32-row chunks
      if (t > 0 && t % kFlushTiles == 0) {               // the epilogue has taken the previous tiles' sums out of TMEM
        mbar_wait(d2_empty, n_flush & 1);
        ++n_flush;
        tc_fence_after();
        first2 = 1;
      }
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        mbar_wait(full + s, ph);
        tc_fence_after();
        const uint32_t st = s2u(ring + (size_t)s * kStage);
#pragma unroll
        for (int x = 0; x < kXG; ++x) {
          const uint32_t a[2] = {st + (1 + x) * kPair, st + (1 + x) * kPair + kBlk}, b[2] = {st, st + kBlk};
          const int pa[3] = {1, 0, 0}, pb[3] = {0, 1, 0};
#pragma unroll
          for (int p = 0; p < 3; ++p) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t ad = mn_desc(a[pa[p]] + (uint32_t)(ks * 2) * C::kSboSP, C::kLbo, C::kSboSP);
              const uint64_t bd = mn_desc(b[pb[p]] + (uint32_t)(ks * 2) * C::kSboZ, C::kLbo, C::kSboZ);
              umma_tf32(tmem_base + C::kColG2 + x * 128, ad, bd, idesc2, (first2 && p == 0 && ks == 0) ? 0u : 1u);
            }
          }
        }
        first2 = 0;
        umma_commit(empty + s);
        if (++s == (uint32_t)kNst) { s = 0; ph ^= 1; }
      }
      if ((t + 1) % kFlushTiles == 0 || t + 1 == my_tiles) umma_commit(d2_full);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) tmem_dealloc(tmem_base, C::kTmemCols);
}

}  // namespace yr

using namespace yr;

template <int D>
static int bwd_tc_launch(const float* E, const float* LE, const float* En, const float* Gn, const float* W1, const float* W2,
                         float slope, int64_t n, float* G, float* T, float* ws, int* n_parts, cudaStream_t s,
                         const int32_t* row_list, const int32_t* row_count, int64_t row_cap, int reserve_sms) {
  using C = BwdTc<D>;
  static yr::AttrOnce attr;
  { int rc_ = attr.set(ngcf_dense_bwd_tc_kernel<D>, (int)C::kSmem); if (rc_) return rc_; }
  const int64_t n_tiles = ((row_list ? row_cap : n) + kBwdTM - 1) / kBwdTM;
  const int64_t sms = yr_sm_count();
  int64_t grid = sms - reserve_sms;
  if (grid > n_tiles) grid = n_tiles;
  if (grid < 1) grid = 1;
  // the split weights live behind the per-CTA dW partials: yr_ngcf_layer_bwd_ws_bytes() sizes the workspace for two CTAs
  // per SM (the FP32-pipe kernel), this kernel runs one
  unsigned char* wsplit = reinterpret_cast<unsigned char*>(ws) + (size_t)sms * 2 * D * D * sizeof(float);
  ngcf_split_weights_bwd_kernel<D><<<16, 256, 0, s>>>(W1, W2, wsplit);
  YR_CHECK_LAUNCH();
  ngcf_dense_bwd_tc_kernel<D><<<(unsigned)grid, kBwdThreads, C::kSmem, s>>>(E, LE, En, Gn, wsplit, slope, n, G, T, ws, row_list,
                                                                            row_count);
  YR_CHECK_LAUNCH();
  *n_parts = (int)grid;
  return YR_OK;
}

// internal launcher used by yr_ngcf_dense_bwd / yr_ngcf_train_step (ngcf.cu); returns the number of CTAs (= dW partials) in
// *n_parts. ws: yr_ngcf_layer_bwd_ws_bytes(d) bytes.
int yr_ngcf_dense_bwd_tc_launch(int d, const float* E, const float* LE, const float* En, const float* Gn, const float* W1,
                                const float* W2, float slope, int64_t n, float* G, float* T, float* ws, int* n_parts,
                                cudaStream_t s, const int32_t* row_list, const int32_t* row_count, int64_t row_cap,
                                int reserve_sms) {
  if (d == 64) return bwd_tc_launch<64>(E, LE, En, Gn, W1, W2, slope, n, G, T, ws, n_parts, s, row_list, row_count, row_cap, reserve_sms);
  if (d == 128) return bwd_tc_launch<128>(E, LE, En, Gn, W1, W2, slope, n, G, T, ws, n_parts, s, row_list, row_count, row_cap, reserve_sms);
  return YR_ERR_BAD_DIM;
}
