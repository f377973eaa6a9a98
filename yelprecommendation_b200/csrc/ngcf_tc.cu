// NGCF dense transforms on the 5th-gen tensor cores (sm_100a), forward:
//     E_next = leaky_relu( [LE+E | E*LE] . [W1 | W2]^T )          (reference models/ngcf.py:64-72)
// as tcgen05.mma.kind::tf32 with the 3xTF32 split (x = hi + lo, hi = top 11 significand bits):
//     A.B ~= A_lo.B_hi + A_hi.B_lo + A_hi.B_hi      (relative error ~2^-21 per product, fp32 accumulation in TMEM)
// which keeps the layer inside the 1e-5 parity bar while the 2*N*d*2d flops leave the FP32 pipe.
//
// Per CTA (persistent over 128-row tiles, one CTA per SM):
//   warps 0-7 : loaders   — coalesced float4 reads of E, LE (next tile prefetched into registers), S = LE+E, P = E*LE, split hi/lo, store the
//                           four [128 x 64] operands in the K-major 128B-swizzled layout the UMMA descriptor expects
//   warps 8-11: epilogue  — thread = row = TMEM lane: tcgen05.ld 64 columns, LeakyReLU, store E_next
//   warp  12  : TMEM allocator + MMA issuer (one elected lane): 48 MMAs (M=128, N=64, K=8) per tile
// W1/W2 are nn.Linear weights [out x in] = K-major B operands as stored; their hi/lo copies are built once per CTA.
// Two TMEM accumulator stages let the epilogue of tile t overlap the loads + MMAs of tile t+1.
#include <stdlib.h>
#include "tc_common.cuh"

namespace yr {

constexpr int kFwdThreads = 416;   // 8 loader warps + 4 epilogue warps + 1 MMA/TMEM warp
constexpr int kFwdTM = 128;

__global__ void __launch_bounds__(kFwdThreads, 1)
ngcf_dense_fwd_tc_kernel(const float* __restrict__ E, const float* __restrict__ LE, const float* __restrict__ W1,
                         const float* __restrict__ W2, float slope, int64_t n, float* __restrict__ Eout,
                         const int32_t* __restrict__ row_list, const int32_t* __restrict__ row_count) {
  // row_list != NULL: tile row q is graph row row_list[q] (only those rows are transformed: the batch rows of the
  // last layer in a BPR step)
  constexpr int D = 64;
  if (row_list) n = *row_count;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (s2u(smem_raw) & 1023u)) & 1023u);
  // A operands: S_hi, S_lo, P_hi, P_lo  [128 x 64] each = 32 KB;  B operands: W1_hi, W1_lo, W2_hi, W2_lo [64 x 64] = 16 KB
  unsigned char* A[4];
  unsigned char* B[4];
  for (int i = 0; i < 4; ++i) A[i] = smem + (size_t)i * 32768;
  for (int i = 0; i < 4; ++i) B[i] = smem + 4 * 32768 + (size_t)i * 16384;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 4 * 32768 + 4 * 16384);
  uint64_t* a_full = bars; uint64_t* a_empty = bars + 1; uint64_t* d_full = bars + 2; uint64_t* d_empty = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n_tiles = (n + kFwdTM - 1) / kFwdTM;

  if (tid == 0) {
    mbar_init(a_full, 1); mbar_init(a_empty, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(d_full + a, 1); mbar_init(d_empty + a, 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 12) tmem_alloc(tmem_slot, 128);         // 2 accumulator stages x 64 columns
  // weights -> hi/lo K-major swizzled B operands (row = output feature o, k contiguous)
  for (int idx = tid; idx < D * (D / 4); idx += kFwdThreads) {
    const int o = idx / (D / 4), c4 = idx % (D / 4);
    float4 hi, lo;
    split4(__ldg(reinterpret_cast<const float4*>(W1 + o * D) + c4), hi, lo);
    *reinterpret_cast<float4*>(B[0] + sw_off(D, o, c4)) = hi;
    *reinterpret_cast<float4*>(B[1] + sw_off(D, o, c4)) = lo;
    split4(__ldg(reinterpret_cast<const float4*>(W2 + o * D) + c4), hi, lo);
    *reinterpret_cast<float4*>(B[2] + sw_off(D, o, c4)) = hi;
    *reinterpret_cast<float4*>(B[3] + sw_off(D, o, c4)) = lo;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // D=f32, A=B=tf32, K-major both, N=64, M=128
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(D >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

  if (warp < 8) {
    // ================= loaders (256 threads) =================
    // The 128-row tile of E (and of LE) is one contiguous 32 KB block: thread t owns float4 #(t + 256 i), i = 0..7,
    // so every warp request is 512 contiguous bytes. The NEXT tile is prefetched into registers right after this
    // tile's operands are handed to the MMA warp, so the global loads overlap the MMAs and the epilogue.
    uint32_t ph = 0;
    float4 e[8], le[8];
    auto fetch = [&](int64_t tile) {
      const int64_t base4 = tile * kFwdTM * (D / 4);
      const int64_t lim4 = n * (D / 4);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int64_t g = base4 + tid + 256 * i;
        e[i] = make_float4(0.f, 0.f, 0.f, 0.f); le[i] = e[i];
        if (g < lim4) {
          if (row_list) g = (int64_t)row_list[g >> 4] * (D / 4) + (g & 15);
          e[i] = __ldg(reinterpret_cast<const float4*>(E) + g);
          le[i] = __ldg(reinterpret_cast<const float4*>(LE) + g);
        }
      }
    };
    if ((int64_t)blockIdx.x < n_tiles) fetch(blockIdx.x);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      mbar_wait(a_empty, ph ^ 1);                        // previous tile's MMAs have consumed the operands
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int g = tid + 256 * i;
        const int r = g >> 4, c4 = g & 15;
        const float4 s = make_float4(le[i].x + e[i].x, le[i].y + e[i].y, le[i].z + e[i].z, le[i].w + e[i].w);
        const float4 p = make_float4(e[i].x * le[i].x, e[i].y * le[i].y, e[i].z * le[i].z, e[i].w * le[i].w);
        float4 hi, lo;
        const uint32_t off = sw_off(kFwdTM, r, c4);
        split4(s, hi, lo);
        *reinterpret_cast<float4*>(A[0] + off) = hi;
        *reinterpret_cast<float4*>(A[1] + off) = lo;
        split4(p, hi, lo);
        *reinterpret_cast<float4*>(A[2] + off) = hi;
        *reinterpret_cast<float4*>(A[3] + off) = lo;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");     // all 128 rows written
      if (tid == 0) mbar_arrive(a_full);
      ph ^= 1;
      if (tile + gridDim.x < n_tiles) fetch(tile + gridDim.x);
    }
  } else if (warp < 12) {
    // ================= epilogue =================
    uint32_t acc = 0, aph = 0;
    const int r = (warp & 3) * 32 + lane;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t row = tile * kFwdTM + r;
      mbar_wait(d_full + acc, aph);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + acc * D;
      float v0[32], v1[32];
      tmem_ld32(trow, v0);
      tmem_ld32(trow + 32, v1);
      tc_fence_before();
      mbar_arrive(d_empty + acc);                        // accumulator stage free again
      if (row < n) {
        float4* o4 = reinterpret_cast<float4*>(Eout + (row_list ? (int64_t)row_list[row] : row) * D);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float4 o;
          o.x = v0[4 * c + 0] > 0.f ? v0[4 * c + 0] : v0[4 * c + 0] * slope;
          o.y = v0[4 * c + 1] > 0.f ? v0[4 * c + 1] : v0[4 * c + 1] * slope;
          o.z = v0[4 * c + 2] > 0.f ? v0[4 * c + 2] : v0[4 * c + 2] * slope;
          o.w = v0[4 * c + 3] > 0.f ? v0[4 * c + 3] : v0[4 * c + 3] * slope;
          o4[c] = o;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float4 o;
          o.x = v1[4 * c + 0] > 0.f ? v1[4 * c + 0] : v1[4 * c + 0] * slope;
          o.y = v1[4 * c + 1] > 0.f ? v1[4 * c + 1] : v1[4 * c + 1] * slope;
          o.z = v1[4 * c + 2] > 0.f ? v1[4 * c + 2] : v1[4 * c + 2] * slope;
          o.w = v1[4 * c + 3] > 0.f ? v1[4 * c + 3] : v1[4 * c + 3] * slope;
          o4[8 + c] = o;
        }
      }
      if (++acc == 2) { acc = 0; aph ^= 1; }
    }
  } else if (lane == 0) {
    // ================= MMA issuer =================
    uint32_t ph = 0, acc = 0, aph = 0;
    // (A operand, B operand) per pass, small terms first: S_lo.W1_hi, S_hi.W1_lo, P_lo.W2_hi, P_hi.W2_lo, S_hi.W1_hi, P_hi.W2_hi
    const int pa[6] = {1, 0, 3, 2, 0, 2};
    const int pb[6] = {0, 1, 2, 3, 0, 2};
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      mbar_wait(d_empty + acc, aph ^ 1);
      mbar_wait(a_full, ph);
      tc_fence_after();
      const uint32_t tacc = tmem_base + acc * D;
      uint32_t first = 1;
#pragma unroll 1
      for (int p = 0; p < 6; ++p) {
        const uint32_t abase = s2u(A[pa[p]]), bbase = s2u(B[pb[p]]);
        for (int sl = 0; sl < 2; ++sl) {
          const uint64_t ad = sw128_desc(abase + sl * kFwdTM * 128);
          const uint64_t bd = sw128_desc(bbase + sl * D * 128);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            umma_tf32(tacc, ad + 2 * ks, bd + 2 * ks, idesc, first ? 0u : 1u);
            first = 0;
          }
        }
      }
      umma_commit(a_empty);
      umma_commit(d_full + acc);
      ph ^= 1;
      if (++acc == 2) { acc = 0; aph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) tmem_dealloc(tmem_base, 128);
}

}  // namespace yr

using namespace yr;

// internal launcher used by yr_ngcf_layer_fwd (ngcf.cu)
int yr_ngcf_dense_fwd_tc_launch(const float* E, const float* LE, const float* W1, const float* W2, float slope,
                                int64_t n, float* Eout, cudaStream_t s, const int32_t* row_list,
                                const int32_t* row_count, int64_t row_cap) {
  const size_t smem = 4 * 32768 + 4 * 16384 + 64 + 1024;
  static yr::AttrOnce attr;
  { int rc_ = attr.set(ngcf_dense_fwd_tc_kernel, (int)smem); if (rc_) return rc_; }
  const int64_t n_tiles = ((row_list ? row_cap : n) + kFwdTM - 1) / kFwdTM;
  int64_t grid = yr_sm_count();
  if (grid > n_tiles) grid = n_tiles;
  if (grid < 1) grid = 1;
  ngcf_dense_fwd_tc_kernel<<<(unsigned)grid, kFwdThreads, smem, s>>>(E, LE, W1, W2, slope, n, Eout, row_list, row_count);
  YR_CHECK_LAUNCH();
  return YR_OK;
}
