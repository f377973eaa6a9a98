// NGCF dense transforms on the 5th-gen tensor cores (sm_100a), backward of one layer on its rows
// (autograd of reference models/ngcf.py:64-72):
//     dZ = G_next * leaky'(E_next)
//     [dS | dP] = dZ . [W1 | W2]                       GEMM1  (M = 128 rows, N = 128, K = 64)
//     T = dS + dP * E ;  G += dS + dP * LE             epilogue
//     [dW1 | dW2]^T = [S | P]^T . dZ                   GEMM2  (M = 128 features, N = 64, K = rows) accumulated in TMEM
// Both GEMMs are tcgen05.mma.kind::tf32 with the 3xTF32 split (lo.hi + hi.lo + hi.hi, fp32 accumulation in TMEM).
//
// Operand layouts (notes/README.md, verified by notes/mn_test): GEMM1 reads dZ K-major (128B swizzle, as the forward
// does) and the weights AS STORED ([out][in] row-major = N contiguous) through an MN-major descriptor; GEMM2 reduces
// over ROWS, so both its operands are the row-major tiles themselves read MN-major (SWIZZLE_128B_BASE32B atoms of
// 32 columns x 4 rows). dZ is therefore staged twice. Shared memory: W 64 KB + dZ(K) 64 KB + 64-row halves of dZ(MN)
// 32 KB and [S|P](MN) 64 KB = 224 KB, single-buffered; the halves alternate under the MMAs of the other operands.
//
// Per CTA (persistent over 128-row tiles, one CTA per SM), 416 threads:
//   warps 0-7 : loaders  — coalesced float4 reads (next tile's G_next / E_next prefetched in registers), dZ, S, P,
//                          hi/lo split, stores in the two layouts
//   warps 8-11: epilogue — thread = row = TMEM lane: tcgen05.ld of dS / dP, T and G rows; at the end the dW partial
//   warp  12  : TMEM allocator + MMA issuer (one elected lane): 24 + 2 x 24 MMAs per tile
// dW partials: one [2 x 64 x 64] block per CTA, reduced in CTA order by reduce_partials_kernel (deterministic).
//
// STATUS (round 1): correct (same parity tests as the FP32-pipe kernel) but NOT the default: 85 us per layer against
// 75 us for ngcf_dense_bwd_kernel. Measured by disabling parts: with no global traffic the kernel runs in ~35 us; the
// epilogue's row-per-thread reads of E / LE / G (TMEM lane = row forces that pattern, 128 threads, 12 loads in flight)
// cost ~25 us, the loaders' exposed latencies ~7 us, the 72 MMAs per tile ~9 us. 13 warps cap the kernel at 128
// registers per thread (4 warps share one 16 K-register scheduler partition), which is what forbids deeper
// software pipelining. Selected with yr_ngcf_set_dense_mode(2).
#include "tc_common.cuh"

namespace yr {

constexpr int kBwdThreads = 416;      // 8 loader warps + 4 epilogue warps + 1 MMA / TMEM warp
constexpr int kBwdTM = 128;
constexpr int kBwdD = 64;

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

__global__ void __launch_bounds__(kBwdThreads, 1)
ngcf_dense_bwd_tc_kernel(const float* __restrict__ E, const float* __restrict__ LE, const float* __restrict__ Enext,
                         const float* __restrict__ Gnext, const float* __restrict__ W1, const float* __restrict__ W2,
                         float slope, int64_t n, float* __restrict__ G, float* __restrict__ T, float* __restrict__ ws,
                         const int32_t* __restrict__ row_list, const int32_t* __restrict__ row_count) {
  constexpr int D = kBwdD, TM = kBwdTM;
  if (row_list) n = *row_count;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (s2u(smem_raw) & 1023u)) & 1023u);
  unsigned char* Bw[2]  = {smem, smem + 32768};                       // [W1|W2] hi, lo   MN-major (K = o 64, N = i' 128)
  unsigned char* dZk[2] = {smem + 65536, smem + 98304};               // dZ hi, lo        K-major  (M = r 128, K = o 64)
  unsigned char* dZm[2] = {smem + 131072, smem + 147456};             // dZ hi, lo        MN-major (K = r 64,  N = o 64)
  unsigned char* SPm[2] = {smem + 163840, smem + 196608};             // [S|P] hi, lo     MN-major (K = r 64,  M = i' 128)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 229376);
  uint64_t* g1_full = bars; uint64_t* g1_empty = bars + 1; uint64_t* g2_full = bars + 2; uint64_t* g2_empty = bars + 3;
  uint64_t* d1_full = bars + 4; uint64_t* d1_empty = bars + 6; uint64_t* d2_done = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  constexpr uint32_t kLbo = 512, kSboW = 2048, kSboZ = 1024, kSboSP = 2048;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n_tiles = (n + TM - 1) / TM;

  if (tid == 0) {
    mbar_init(g1_full, 1); mbar_init(g1_empty, 1); mbar_init(g2_full, 1); mbar_init(g2_empty, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(d1_full + a, 1); mbar_init(d1_empty + a, 128); }
    mbar_init(d2_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 12) tmem_alloc(tmem_slot, 512);          // D1: 2 stages x 128 columns, D2: 64 columns at 256
  // weights as stored ([o][i]): row o = K index, column i (W1) / 64 + i (W2) = N index
  for (int idx = tid; idx < D * (D / 4); idx += kBwdThreads) {
    const int o = idx / (D / 4), c4 = idx % (D / 4);
    float4 hi, lo;
    split4(__ldg(reinterpret_cast<const float4*>(W1 + o * D) + c4), hi, lo);
    *reinterpret_cast<float4*>(Bw[0] + mn_off(c4 * 4, o, kLbo, kSboW)) = hi;
    *reinterpret_cast<float4*>(Bw[1] + mn_off(c4 * 4, o, kLbo, kSboW)) = lo;
    split4(__ldg(reinterpret_cast<const float4*>(W2 + o * D) + c4), hi, lo);
    *reinterpret_cast<float4*>(Bw[0] + mn_off(D + c4 * 4, o, kLbo, kSboW)) = hi;
    *reinterpret_cast<float4*>(Bw[1] + mn_off(D + c4 * 4, o, kLbo, kSboW)) = lo;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // D = f32, A = B = tf32; GEMM1: A K-major, B MN-major, N = 128; GEMM2: both MN-major, N = 64; M = 128
  const uint32_t idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) |
                          ((uint32_t)(128 >> 4) << 24);

  if (warp < 8) {
    // ================= loaders (256 threads) =================
    // thread t owns float4 #(t + 256 i), i = 0..7, of every 128 x 64 tile: row (t >> 4) + 16 i, columns 4 (t & 15)..+3;
    // i < 4 is the first 64-row half, i >= 4 the second.
    const int c4 = tid & 15, rb = tid >> 4;
    uint32_t ph1 = 0, ph2 = 0;                  // parities of g1_empty / g2_empty waits
    float4 gn[8], en[8];
    auto row_of = [&](int64_t tile, int i) -> int64_t {
      const int64_t q = tile * TM + rb + 16 * i;
      if (q >= n) return -1;
      return row_list ? (int64_t)row_list[q] : q;
    };
    auto fetch = [&](int64_t tile) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t r = row_of(tile, i);
        gn[i] = make_float4(0.f, 0.f, 0.f, 0.f); en[i] = gn[i];
        if (r >= 0) {
          gn[i] = __ldg(reinterpret_cast<const float4*>(Gnext + r * D) + c4);
          en[i] = __ldg(reinterpret_cast<const float4*>(Enext + r * D) + c4);
        }
      }
    };
    if ((int64_t)blockIdx.x < n_tiles) fetch(blockIdx.x);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      float4 dz[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        dz[i].x = en[i].x > 0.f ? gn[i].x : gn[i].x * slope; dz[i].y = en[i].y > 0.f ? gn[i].y : gn[i].y * slope;
        dz[i].z = en[i].z > 0.f ? gn[i].z : gn[i].z * slope; dz[i].w = en[i].w > 0.f ? gn[i].w : gn[i].w * slope;
      }
      // E / LE of the whole tile: requested now, consumed after the dZ stores and the barrier waits
      float4 e[8], le[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t r = row_of(tile, i);
        e[i] = make_float4(0.f, 0.f, 0.f, 0.f); le[i] = e[i];
        if (r >= 0) {
          e[i] = __ldg(reinterpret_cast<const float4*>(E + r * D) + c4);
          le[i] = __ldg(reinterpret_cast<const float4*>(LE + r * D) + c4);
        }
      }
      auto store_half = [&](int h) {               // dZ (MN) and [S|P] (MN) of rows 64 h .. 64 h + 63
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int k = rb + 16 * i;                // row inside the half
          float4 hi, lo;
          split4(dz[4 * h + i], hi, lo);
          const uint32_t oz = mn_off(c4 * 4, k, kLbo, kSboZ);
          *reinterpret_cast<float4*>(dZm[0] + oz) = hi;
          *reinterpret_cast<float4*>(dZm[1] + oz) = lo;
          const float4 ee = e[4 * h + i], ll = le[4 * h + i];
          const float4 s = make_float4(ll.x + ee.x, ll.y + ee.y, ll.z + ee.z, ll.w + ee.w);
          const float4 p = make_float4(ee.x * ll.x, ee.y * ll.y, ee.z * ll.z, ee.w * ll.w);
          split4(s, hi, lo);
          const uint32_t os = mn_off(c4 * 4, k, kLbo, kSboSP);
          *reinterpret_cast<float4*>(SPm[0] + os) = hi;
          *reinterpret_cast<float4*>(SPm[1] + os) = lo;
          split4(p, hi, lo);
          const uint32_t op = mn_off(D + c4 * 4, k, kLbo, kSboSP);
          *reinterpret_cast<float4*>(SPm[0] + op) = hi;
          *reinterpret_cast<float4*>(SPm[1] + op) = lo;
        }
      };
      // ---- dZ, K-major, whole tile -> GEMM1
      mbar_wait(g1_empty, ph1 ^ 1);
      ph1 ^= 1;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 hi, lo;
        split4(dz[i], hi, lo);
        const uint32_t off = sw_off(TM, rb + 16 * i, c4);
        *reinterpret_cast<float4*>(dZk[0] + off) = hi;
        *reinterpret_cast<float4*>(dZk[1] + off) = lo;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (tid == 0) mbar_arrive(g1_full);
      // ---- first half of the MN-major operands -> GEMM2
      mbar_wait(g2_empty, ph2 ^ 1);
      ph2 ^= 1;
      store_half(0);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (tid == 0) mbar_arrive(g2_full);
      // ---- second half
      mbar_wait(g2_empty, ph2 ^ 1);
      ph2 ^= 1;
      store_half(1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (tid == 0) mbar_arrive(g2_full);
      if (tile + gridDim.x < n_tiles) fetch(tile + gridDim.x);
    }
  } else if (warp < 12) {
    // ================= epilogue: thread = row = TMEM lane, 16 columns of dS and of dP per step =================
    uint32_t acc = 0, aph = 0;
    const int r = (warp & 3) * 32 + lane;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t q = tile * TM + r;
      const int64_t row = (q < n) ? (row_list ? (int64_t)row_list[q] : q) : -1;
      mbar_wait(d1_full + acc, aph);
      tc_fence_after();
      const uint32_t trow = tmem_base + lane_base + acc * 128;
#pragma unroll 1
      for (int c0 = 0; c0 < D; c0 += 16) {
        float ds[16], dp[16];
        tmem_ld16(trow + c0, ds);
        tmem_ld16(trow + D + c0, dp);
        if (row >= 0) {
          const float4* e4 = reinterpret_cast<const float4*>(E + row * D + c0);
          const float4* le4 = reinterpret_cast<const float4*>(LE + row * D + c0);
          float4* g4 = reinterpret_cast<float4*>(G + row * D + c0);
          float4* t4 = reinterpret_cast<float4*>(T + row * D + c0);
          float4 ev[4], lv[4], gv[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) { ev[c] = __ldg(e4 + c); lv[c] = __ldg(le4 + c); gv[c] = g4[c]; }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float4 t, g = gv[c];
            t.x = fmaf(dp[4 * c + 0], ev[c].x, ds[4 * c + 0]); t.y = fmaf(dp[4 * c + 1], ev[c].y, ds[4 * c + 1]);
            t.z = fmaf(dp[4 * c + 2], ev[c].z, ds[4 * c + 2]); t.w = fmaf(dp[4 * c + 3], ev[c].w, ds[4 * c + 3]);
            g.x += fmaf(dp[4 * c + 0], lv[c].x, ds[4 * c + 0]); g.y += fmaf(dp[4 * c + 1], lv[c].y, ds[4 * c + 1]);
            g.z += fmaf(dp[4 * c + 2], lv[c].z, ds[4 * c + 2]); g.w += fmaf(dp[4 * c + 3], lv[c].w, ds[4 * c + 3]);
            t4[c] = t;
            g4[c] = g;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(d1_empty + acc);
      if (++acc == 2) { acc = 0; aph ^= 1; }
    }
    // ---- dW partial of this CTA: TMEM lane = feature i' (S columns 0..63 -> dW1, P columns -> dW2), column = o
    float* my = ws + (size_t)blockIdx.x * 2 * D * D + (size_t)(r >> 6) * D * D + (r & 63);
    if ((int64_t)blockIdx.x < n_tiles) {
      mbar_wait(d2_done, 0);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < D; c0 += 16) {
        float v[16];
        tmem_ld16(tmem_base + lane_base + 256 + c0, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) my[(size_t)(c0 + j) * D] = v[j];    // ws[..][o][i]: lanes write consecutive i
      }
    } else {                                     // a CTA without a tile (short row list) contributes zeros
      for (int o = 0; o < D; ++o) my[(size_t)o * D] = 0.f;
    }
  } else if (lane == 0) {
    // ================= MMA issuer =================
    uint32_t p1 = 0, p2 = 0, acc = 0, aph = 0, first2 = 1;
    // (A, B) hi/lo index per pass, small terms first
    const int pa[3] = {1, 0, 0};
    const int pb[3] = {0, 1, 0};
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      // ---- GEMM1: [dS | dP] = dZ . [W1 | W2]
      mbar_wait(d1_empty + acc, aph ^ 1);
      mbar_wait(g1_full, p1);
      p1 ^= 1;
      tc_fence_after();
      const uint32_t t1 = tmem_base + acc * 128;
      uint32_t first = 1;
#pragma unroll 1
      for (int p = 0; p < 3; ++p) {
        const uint32_t abase = s2u(dZk[pa[p]]), bbase = s2u(Bw[pb[p]]);
        for (int sl = 0; sl < 2; ++sl) {
          const uint64_t ad = sw128_desc(abase + sl * TM * 128);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t bd = mn_desc(bbase + (uint32_t)(sl * 8 + ks * 2) * kSboW, kLbo, kSboW);
            umma_tf32(t1, ad + 2 * ks, bd, idesc1, first ? 0u : 1u);
            first = 0;
          }
        }
      }
      umma_commit(g1_empty);
      umma_commit(d1_full + acc);
      if (++acc == 2) { acc = 0; aph ^= 1; }
      // ---- GEMM2: [dW1 | dW2]^T += [S | P]^T . dZ, two 64-row halves
      for (int h = 0; h < 2; ++h) {
        mbar_wait(g2_full, p2);
        p2 ^= 1;
        tc_fence_after();
#pragma unroll 1
        for (int p = 0; p < 3; ++p) {
          const uint32_t abase = s2u(SPm[pa[p]]), bbase = s2u(dZm[pb[p]]);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t ad = mn_desc(abase + (uint32_t)(ks * 2) * kSboSP, kLbo, kSboSP);
            const uint64_t bd = mn_desc(bbase + (uint32_t)(ks * 2) * kSboZ, kLbo, kSboZ);
            umma_tf32(tmem_base + 256, ad, bd, idesc2, first2 ? 0u : 1u);
            first2 = 0;
          }
        }
        umma_commit(g2_empty);
      }
    }
    umma_commit(d2_done);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) tmem_dealloc(tmem_base, 512);
}

}  // namespace yr

using namespace yr;

// internal launcher used by yr_ngcf_dense_bwd (ngcf.cu); returns the number of CTAs (= dW partials) in *n_parts
int yr_ngcf_dense_bwd_tc_launch(const float* E, const float* LE, const float* En, const float* Gn, const float* W1,
                                const float* W2, float slope, int64_t n, float* G, float* T, float* ws, int* n_parts,
                                cudaStream_t s, const int32_t* row_list, const int32_t* row_count, int64_t row_cap) {
  const size_t smem = 229376 + 128 + 1024;
  static yr::AttrOnce attr;
  { int rc_ = attr.set(ngcf_dense_bwd_tc_kernel, (int)smem); if (rc_) return rc_; }
  const int64_t n_tiles = ((row_list ? row_cap : n) + kBwdTM - 1) / kBwdTM;
  int64_t grid = yr_sm_count();
  if (grid > n_tiles) grid = n_tiles;
  if (grid < 1) grid = 1;
  ngcf_dense_bwd_tc_kernel<<<(unsigned)grid, kBwdThreads, smem, s>>>(E, LE, En, Gn, W1, W2, slope, n, G, T, ws, row_list,
                                                                     row_count);
  YR_CHECK_LAUNCH();
  *n_parts = (int)grid;
  return YR_OK;
}
