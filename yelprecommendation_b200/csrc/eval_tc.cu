// Full-catalog evaluation on the 5th-gen tensor cores (sm_100a): same contract and BIT-IDENTICAL outputs as
// eval.cu (reference trainers/mf_trainer.py:134-178, metric.py:7-109), ~30x fewer FP32-pipe flops.
//
// Idea: TF32 tcgen05.mma scores are only a FILTER. For user u and item i let s be the canonical fp32 fma-chain
// score and s~ the tensor-core score; |s~ - s| <= eps_u := c * ||u||_2 * max_i ||v_i||_2 (c covers TF32 operand
// truncation 2*2^-10, the MMA accumulation and the chain's own rounding). If tau~ is the K-th largest s~ among the
// unmasked items seen so far, every true top-K item satisfies s~ >= tau~_final - 2 eps >= tau~_running - 2 eps.
// So one thread per user (= one TMEM lane) scans its row of the accumulator, keeps the K best s~ values and
// appends every unmasked item with s~ >= tau~ - 2 eps to a small candidate buffer (compacted in place when full).
// After the last item tile the few survivors (~K + a handful) are re-scored with the exact fp32 chain, ordered by
// (score desc, item id asc), and fed to the same metric code. Rows the filter cannot decide (candidate overflow, or
// fewer than K unmasked items) are appended to a fallback list that the exact kernel of eval.cu evaluates.
//
// Pipeline per CTA (one per SM, persistent over 128-user tiles), 384 threads:
//   warp 0   : TMA producer  — item tiles [128 x 32 floats] boxes, 128B swizzle, mbarrier expect_tx
//   warp 1   : MMA issuer    — tcgen05.mma.kind::tf32, M=128, N=128, K=8 per instruction, A/B from shared memory
//   warp 2   : TMEM allocator (4 accumulator stages of 128 columns = all 512 columns)
//   warps 4-11: epilogue     — two threads per user row (column halves), tcgen05.ld 32x32b.x32, filter, candidate
//                              buffers in global scratch, the K best s~ in registers; the two half-streams exchange
//                              their thresholds through shared memory; final exact re-score + metrics
// The user tile (gathered rows of eval_uid) is written to shared memory by all threads in the 128B-swizzled K-major
// layout the UMMA descriptor expects.
#include <float.h>
#include <stdlib.h>
#include "tc_common.cuh"

namespace yr {

constexpr int kTcTM = 128;
constexpr int kTcTN = 128;         // items per MMA tile (N)
constexpr int kTcAcc = 4;          // TMEM accumulator stages (4 x 128 columns)
constexpr int kTcThreads = 384;    // 4 control warps + 8 epilogue warps
constexpr int kTcMaxK = 16;
constexpr int kTcCap = 128;        // candidate buffer entries per half-stream (global scratch): compaction is rare
constexpr int kTcPend = 4;         // deferred candidates per epilogue thread (shared memory ring)
constexpr int kTcDrain = 8;        // chunks of 32 columns between two drains of the rings
constexpr float kTcErrCoef = 0.0025f;   // > 2*2^-10 (TF32 operands) + accumulation + fp32-chain rounding

struct TcParams {
  const float* Uemb; int64_t nU;
  const float* Vemb; int64_t nI; int d;
  const int64_t* eval_uid; int64_t n_eval;
  const int32_t* mask_ptr; const int32_t* mask_idx;
  const int32_t* act_ptr; const int32_t* act_idx; const int32_t* act_nuniq;
  const double* inv_log2; int K;
  int64_t* topk_out; float* topk_score; double* user_metrics;
  int32_t* err;
  const float* vmax;             // device scalar: max item row norm
  int32_t* fb_count; int32_t* fb_rows;
  int* cand;                     // global scratch for the candidate buffers: [grid][2 arrays][2][CAP][128]
  int SPS, NST, CAP;             // 32-float slabs per stage, pipeline stages, candidate buffer entries per half-stream
  int MCAP;                      // mask entries per row staged in shared memory (0 = none)
  // item-sliced evaluation (yr_eval_topk_metrics_tc_slice): this call covers one of n_slices disjoint item slices; gx
  // [n_slices x n_eval] carries every slice's current lower bound of a row's K-th best score between the concurrent calls
  float* gx; int slice, n_slices;
};

__device__ __forceinline__ float exact_score(const float* __restrict__ u, const float* __restrict__ v, int d) {
  float acc = 0.f;
  const float4* u4 = reinterpret_cast<const float4*>(u);
  const float4* v4 = reinterpret_cast<const float4*>(v);
  for (int k = 0; k < d / 4; ++k) {
    const float4 a = __ldg(u4 + k), c = __ldg(v4 + k);
    acc = fmaf(a.x, c.x, acc); acc = fmaf(a.y, c.y, acc);
    acc = fmaf(a.z, c.z, acc); acc = fmaf(a.w, c.w, acc);
  }
  return acc;
}

// Select v[j] for a per-lane j without local memory: 5-level mux over the 32 registers.
__device__ __forceinline__ float select32(const float (&v)[32], int j) {
  float a[16], b[8], c[4], e[2];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (j & 1) ? v[2 * i + 1] : v[2 * i];
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = (j & 2) ? a[2 * i + 1] : a[2 * i];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i] = (j & 4) ? b[2 * i + 1] : b[2 * i];
#pragma unroll
  for (int i = 0; i < 2; ++i) e[i] = (j & 8) ? c[2 * i + 1] : c[2 * i];
  return (j & 16) ? e[1] : e[0];
}

__device__ __forceinline__ float select16(const float (&v)[16], int j) {
  float a[8], b[4], c[2];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = (j & 1) ? v[2 * i + 1] : v[2 * i];
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = (j & 2) ? a[2 * i + 1] : a[2 * i];
#pragma unroll
  for (int i = 0; i < 2; ++i) c[i] = (j & 4) ? b[2 * i + 1] : b[2 * i];
  return (j & 8) ? c[1] : c[0];
}

__global__ void __launch_bounds__(kTcThreads, 1)
eval_tc_kernel(const __grid_constant__ CUtensorMap vmap, TcParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // the swizzled tiles need 1024-byte alignment in the shared window; the launch adds 1 KB of slack for this
  unsigned char* smem = smem_raw + ((1024u - (s2u(smem_raw) & 1023u)) & 1023u);
  const int d = P.d, K = P.K, SPS = P.SPS, NST = P.NST, CAP = P.CAP, MCAP = P.MCAP;
  constexpr int TN = kTcTN;
  const int n_slabs = d / 32;
  const int n_kc = n_slabs / SPS;                       // stages per item tile
  const uint32_t stage_bytes = (uint32_t)SPS * TN * 128;
  unsigned char* Us = smem;                             // [n_slabs][128 rows][128 B] swizzled
  unsigned char* Vs = Us + (size_t)n_slabs * 16384;     // [NST][SPS][TN rows][128 B] swizzled (TMA)
  float* xch = reinterpret_cast<float*>(Vs + (size_t)NST * stage_bytes);  // [2 halves][tau, mid][128] threshold exchange
  // candidate buffers live in global scratch (L2): touched ~100 times per user, and 64 KB of shared memory is
  // worth more as pipeline stages
  int* cid = P.cand + (size_t)blockIdx.x * 4 * CAP * kTcTM;               // [2][CAP][128] candidate item ids
  float* csc = reinterpret_cast<float*>(cid + 2 * CAP * kTcTM);           // [2][CAP][128] candidate s~
  float* tauB = xch + 4 * kTcTM;                                          // [128] tau of the second half-stream
  int* cntB = reinterpret_cast<int*>(tauB + kTcTM);                       // [128] its candidate count (-1 = overflow)
  uint64_t* bars = reinterpret_cast<uint64_t*>(cntB + kTcTM);             // full[4] empty[4] tfull[4] tempty[4]
  uint64_t* full = bars; uint64_t* empty = bars + 4; uint64_t* tfull = bars + 8; uint64_t* tempty = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  // The first MCAP entries of every row's sorted mask list, staged per user tile (row stride MCAP + 1: the 32 lanes of a warp
  // walk 32 different rows). The candidate path advances a cursor through this list; from global memory every step was a
  // dependent L2 round trip inside a divergent loop — the dominant cost of the epilogue in round 1's profile.
  int* msk = reinterpret_cast<int*>(tmem_slot + 4);
  // Deferred candidate handling: the scan only PUSHES (item, s~) into a small per-thread ring; the expensive part (mask
  // cursor, candidate buffer, sorted K-best list, threshold exchange) runs for the whole warp at once every kTcDrain chunks.
  // With 32 rows per warp nearly every 32-column chunk has a candidate in SOME lane, so handling candidates inline made
  // the whole warp walk the slow path once per chunk; batched, that cost is shared by all the lanes that have work.
  float* pend_s = reinterpret_cast<float*>(msk + kTcTM * (MCAP + 1));       // [kTcPend][256]
  int* pend_i = reinterpret_cast<int*>(pend_s + kTcPend * 256);             // [kTcPend][256]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n_utiles = (P.n_eval + kTcTM - 1) / kTcTM;
  const int n_itiles = (int)((P.nI + TN - 1) / TN);
  constexpr uint32_t tmem_cols = kTcAcc * TN;           // 4 accumulator stages x 128 columns = all of TMEM

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    for (int a = 0; a < kTcAcc; ++a) { mbar_init(tfull + a, 1); mbar_init(tempty + a, 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&vmap) : "memory");
  }
  if (warp == 2) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // instruction descriptor: D=f32, A=B=tf32, both K-major, N=TN, M=128
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

  uint32_t st_p = 0, ph_p = 0;      // producer ring position / phase
  uint32_t st_m = 0, ph_m = 0;      // MMA ring position / phase
  uint32_t acc_m = 0, aph_m = 0;    // MMA accumulator stage / phase
  uint32_t acc_e = 0, aph_e = 0;    // epilogue accumulator stage / phase

  for (int64_t ut = blockIdx.x; ut < n_utiles; ut += gridDim.x) {
    const int64_t e0 = ut * kTcTM;
    // ---- user tile -> shared memory, K-major SW128: byte = slab*16384 + (r/8)*1024 + (r%8)*128 + ((c ^ (r%8))*16)
    for (int idx = tid; idx < kTcTM * (d / 4); idx += kTcThreads) {
      const int r = idx / (d / 4), c4 = idx % (d / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int64_t e = e0 + r;
      if (e < P.n_eval) {
        const int64_t uid = P.eval_uid[e];
        if (uid >= 0 && uid < P.nU) v = __ldg(reinterpret_cast<const float4*>(P.Uemb + uid * d) + c4);
        else if (P.err) atomicExch(P.err, 1);
      }
      const int slab = c4 >> 3, c = c4 & 7;
      *reinterpret_cast<float4*>(Us + (size_t)slab * 16384 + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)) = v;
    }
    if (MCAP > 0) {
      for (int idx = tid; idx < kTcTM * MCAP; idx += kTcThreads) {
        const int r = idx / MCAP, j = idx - r * MCAP;
        const int64_t e = e0 + r;
        if (e < P.n_eval) {
          const int m0 = P.mask_ptr[e];
          if (j < P.mask_ptr[e + 1] - m0) msk[r * (MCAP + 1) + j] = P.mask_idx[m0 + j];
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to the MMA (async proxy)
    __syncthreads();

    if (warp == 0) {
      // ================= TMA producer =================
      if (lane == 0) {
        for (int it = 0; it < n_itiles; ++it) {
          for (int kc = 0; kc < n_kc; ++kc) {
            mbar_wait(empty + st_p, ph_p ^ 1);
            mbar_expect_tx(full + st_p, stage_bytes);
            unsigned char* dst = Vs + (size_t)st_p * stage_bytes;
            for (int sl = 0; sl < SPS; ++sl)
              tma_load_2d(dst + (size_t)sl * TN * 128, &vmap, (kc * SPS + sl) * 32, it * TN, full + st_p);
            if (++st_p == (uint32_t)NST) { st_p = 0; ph_p ^= 1; }
          }
        }
      }
    } else if (warp == 1) {
      // ================= MMA issuer =================
      if (lane == 0) {
        for (int it = 0; it < n_itiles; ++it) {
          mbar_wait(tempty + acc_m, aph_m ^ 1);
          tc_fence_after();
          const uint32_t tacc = tmem_base + acc_m * TN;
          for (int kc = 0; kc < n_kc; ++kc) {
            mbar_wait(full + st_m, ph_m);
            tc_fence_after();
            const uint32_t bbase = s2u(Vs + (size_t)st_m * stage_bytes);
            for (int sl = 0; sl < SPS; ++sl) {
              const uint64_t ad = sw128_desc(s2u(Us + (size_t)(kc * SPS + sl) * 16384));
              const uint64_t bd = sw128_desc(bbase + (uint32_t)sl * TN * 128);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)       // 4 x (K = 8 tf32 = 32 B) per 128-byte slab
                umma_tf32(tacc, ad + 2 * ks, bd + 2 * ks, idesc, (kc | sl | ks) != 0);
            }
            umma_commit(empty + st_m);             // smem stage free once these MMAs retire
            if (++st_m == (uint32_t)NST) { st_m = 0; ph_m ^= 1; }
          }
          umma_commit(tfull + acc_m);              // accumulator ready for the epilogue
          if (++acc_m == kTcAcc) { acc_m = 0; aph_m ^= 1; }
        }
      }
    } else if (warp >= 4) {
      // ================= epilogue: two threads per user row (column halves of every tile) =================
      const int half = (warp - 4) >> 2;            // 0: columns 0..63, 1: columns 64..127
      const int r = (warp & 3) * 32 + lane;        // TMEM lane == user row of the tile
      const int64_t e = e0 + r;
      const bool live = e < P.n_eval;
      int64_t uid = live ? P.eval_uid[e] : 0;
      if (uid < 0 || uid >= P.nU) uid = 0;
      const float* urow = P.Uemb + uid * d;
      float eps2 = 0.f;
      if (live) {
        float nn = 0.f;
        for (int k = 0; k < d; ++k) nn = fmaf(urow[k], urow[k], nn);
        eps2 = 2.f * kTcErrCoef * sqrtf(nn) * (*P.vmax) * 1.0001f + FLT_MIN;
      }
      int* my_cid = cid + (size_t)half * CAP * kTcTM + r;
      float* my_csc = csc + (size_t)half * CAP * kTcTM + r;
      // the K best s~ values of this half-stream live in REGISTERS, ascending: tkr[0] is the K-th best (tau). Slots
      // K..15 hold +inf so that one static 15-step bubble pass serves every K <= 16 without dynamic indexing.
      float tkr[kTcMaxK];
#pragma unroll
      for (int j = 0; j < kTcMaxK; ++j) tkr[j] = (j < K) ? -INFINITY : INFINITY;
      float tau = -INFINITY, theta = -INFINITY;
      // Threshold exchange between the two half-streams of a row (shared memory, monotone floats, stale reads are
      // still valid bounds): besides its tau each stream publishes `mid`, its ceil(K/2)-th best value. At least
      // 2*ceil(K/2) >= K items of the row score >= min(mid_a, mid_b), so that minimum is a lower bound of the K-th
      // best of the UNION — about the threshold a single stream over all columns would have, which halves the
      // number of candidates either stream lets through.
      const int midx = K - (K + 1) / 2;                   // ascending list: index of the ceil(K/2)-th best
      float mid = -INFINITY;
      float* xch_own = xch + (size_t)half * 2 * kTcTM + r;         // [half][tau, mid][128]
      float* xch_oth = xch + (size_t)(half ^ 1) * 2 * kTcTM + r;
      xch_own[0] = -INFINITY; xch_own[kTcTM] = -INFINITY;
      asm volatile("bar.sync 1, 256;" ::: "memory");               // both halves initialised before anyone reads
      int cnt = 0;
      bool overflow = false;
      const int mbeg = live ? P.mask_ptr[e] : 0;
      int mcur = mbeg;
      const int mend = live ? P.mask_ptr[e + 1] : 0;
      const int* mrow = msk + r * (MCAP + 1);
      auto mask_at = [&](int m) -> int {                     // staged prefix from shared memory, the tail from global
        if (m >= mend) return 0x7fffffff;
        return (m - mbeg < MCAP) ? mrow[m - mbeg] : P.mask_idx[m];
      };
      int mnext = mask_at(mcur);

      auto handle = [&](int item, float s) {          // candidate -> buffer (+ the K best s~ values)
        if (s < theta) return;                                      // theta may have risen since the scan
        while (mnext < item) { ++mcur; mnext = mask_at(mcur); }
        if (mnext == item) return;                                  // masked: never a candidate (see fallback rule)
        if (cnt == CAP) {                                           // compact: keep what is still inside the window
          int w = 0;
          for (int j = 0; j < CAP; ++j) {
            const float sj = my_csc[j * kTcTM];
            if (sj >= theta) { my_csc[w * kTcTM] = sj; my_cid[w * kTcTM] = my_cid[j * kTcTM]; ++w; }
          }
          cnt = w;
          if (cnt == CAP) { overflow = true; return; }
        }
        my_cid[cnt * kTcTM] = item; my_csc[cnt * kTcTM] = s; ++cnt;
        if (s > tau) {                                              // replace the K-th best, one bubble pass re-sorts
          tkr[0] = s;
#pragma unroll
          for (int j = 0; j + 1 < kTcMaxK; ++j) {
            const float lo = fminf(tkr[j], tkr[j + 1]), hi = fmaxf(tkr[j], tkr[j + 1]);
            tkr[j] = lo; tkr[j + 1] = hi;
          }
          tau = tkr[0];
          mid = select16(tkr, midx);
          xch_own[0] = tau; xch_own[kTcTM] = mid;
          theta = fmaxf(theta, tau - eps2);
        }
      };

      const int et = tid - 128;                          // epilogue thread 0..255
      int npend = 0, chunk_ctr = 0;
      auto drain = [&]() {                               // warp-collective
        int nmax = npend;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(kFull, nmax, o));
        for (int q = 0; q < nmax; ++q)
          if (q < npend) handle(pend_i[q * 256 + et], pend_s[q * 256 + et]);
        npend = 0;
      };

      for (int it = 0; it < n_itiles; ++it) {
        mbar_wait(tfull + acc_e, aph_e);
        tc_fence_after();
        {
          const float tau_o = xch_oth[0], mid_o = xch_oth[kTcTM];
          float bound = fmaxf(tau_o, fminf(mid, mid_o));
          if (P.n_slices > 1 && live) {
            // The K-th best of ANY item slice is a lower bound of the K-th best of the whole catalog, so the concurrent calls
            // on the other slices of this row lend their thresholds (global memory, monotone floats, stale reads are still
            // valid bounds). Without this every slice pays the warm-up of its own running threshold — K ln(n / K) candidate
            // events per slice, which is where the kernel's time goes — and slicing gains nothing (measured).
            for (int s2 = 0; s2 < P.n_slices; ++s2)
              if (s2 != P.slice) bound = fmaxf(bound, __ldcg(P.gx + (size_t)s2 * P.n_eval + e));
            if (half == 0) __stcg(P.gx + (size_t)P.slice * P.n_eval + e, fmaxf(tau, fmaxf(tau_o, fminf(mid, mid_o))));
          }
          theta = fmaxf(theta, bound - eps2);
        }
        const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + acc_e * TN + half * (TN / 2);
#pragma unroll 1
        for (int c0 = 0; c0 < TN / 2; c0 += 32) {
          float v[32];
          tmem_ld32(trow + c0, v);
          // tree max (depth 4) instead of a 31-deep chain: this warp is alone on its scheduler
          float m8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) m8[i] = fmaxf(fmaxf(v[4 * i], v[4 * i + 1]), fmaxf(v[4 * i + 2], v[4 * i + 3]));
          const float m = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
          const bool need = live && !overflow && m >= theta;
          if (__any_sync(kFull, need)) {
            unsigned mask = 0;
            if (need) {
#pragma unroll
              for (int j = 0; j < 32; ++j) mask |= (v[j] >= theta) ? (1u << j) : 0u;
            }
            const int item0 = it * TN + half * (TN / 2) + c0;
            // columns past the catalog are zero-filled TMA rows (score 0): never candidates
            if ((int64_t)item0 + 32 > P.nI) mask = (item0 < P.nI) ? (mask & ((1u << (int)(P.nI - item0)) - 1u)) : 0u;
            const bool single = __popc(mask) == 1;                  // the usual case: the only candidate is the max
            while (__any_sync(kFull, mask != 0)) {                    // few lanes, few bits: pushed, handled later
              if (mask && npend < kTcPend) {
                const int j = __ffs(mask) - 1;
                mask &= mask - 1;
                float sj;
                if (single) sj = m; else sj = select32(v, j);
                pend_i[npend * 256 + et] = item0 + j;
                pend_s[npend * 256 + et] = sj;
                ++npend;
              }
              if (__any_sync(kFull, npend == kTcPend)) drain();       // a full ring: everybody empties theirs
            }
          }
          if (((++chunk_ctr) & (kTcDrain - 1)) == 0 && __any_sync(kFull, npend > 0)) drain();
        }
        tc_fence_before();
        mbar_arrive(tempty + acc_e);
        if (++acc_e == kTcAcc) { acc_e = 0; aph_e ^= 1; }
      }

      if (__any_sync(kFull, npend > 0)) drain();
      // ---- hand the second half-stream's state to the first; first half finishes the row ----
      if (half == 1) { tauB[r] = tau; cntB[r] = overflow ? -1 : cnt; }
      xch_own[0] = tau; xch_own[kTcTM] = mid;
      asm volatile("bar.sync 1, 256;" ::: "memory");            // epilogue warps only
      if (half == 0 && live) {
        const float tau_b = tauB[r];
        const int cnt_b = cntB[r];
        // all three are lower bounds of the K-th best s~ of the row
        float tau_f = fmaxf(fmaxf(tau, tau_b), fminf(mid, xch_oth[kTcTM]));
        const bool sliced = P.n_slices > 1;
        if (sliced)
          for (int s2 = 0; s2 < P.n_slices; ++s2)
            if (s2 != P.slice) tau_f = fmaxf(tau_f, __ldcg(P.gx + (size_t)s2 * P.n_eval + e));
        const float theta_f = tau_f - eps2;
        bool fallback = overflow || cnt_b < 0 || tau_f == -INFINITY;
        int topi[kTcMaxK];
        float tops[kTcMaxK];
        int nt = 0;
        if (!fallback) {
          for (int h = 0; h < 2; ++h) {
            const int n_h = h == 0 ? cnt : cnt_b;
            const int* b_id = cid + (size_t)h * CAP * kTcTM + r;
            const float* b_sc = csc + (size_t)h * CAP * kTcTM + r;
            for (int j = 0; j < n_h; ++j) {
              if (b_sc[j * kTcTM] < theta_f) continue;
              const int item = b_id[j * kTcTM];
              const float s = exact_score(urow, P.Vemb + (int64_t)item * d, d);
              int pos = nt < K ? nt : K;                             // insertion by (score desc, id asc)
              while (pos > 0 && (tops[pos - 1] < s || (tops[pos - 1] == s && topi[pos - 1] > item))) --pos;
              if (pos < K) {
                const int last = nt < K ? nt : K - 1;
                for (int q = last; q > pos; --q) { tops[q] = tops[q - 1]; topi[q] = topi[q - 1]; }
                tops[pos] = s; topi[pos] = item;
                if (nt < K) ++nt;
              }
            }
          }
          // a slice of a sliced evaluation may keep fewer than K items (the other slices hold the rest of the top K): the
          // row's list is padded with id -1 / -inf, which yr_topk_merge skips
          if (nt < K && !sliced) fallback = true;
        }
        if (fallback) {
          P.fb_rows[atomicAdd(P.fb_count, 1)] = (int32_t)e;
        } else {
          for (int j = 0; j < K; ++j) {
            P.topk_out[e * K + j] = j < nt ? topi[j] : -1;
            if (P.topk_score) P.topk_score[e * K + j] = j < nt ? tops[j] : -INFINITY;
          }
          if (!sliced) {             // (a slice's list is not the row's recommendation: metrics follow the merge)
          // metric terms, metric.py:7-109 (quirks Q6-Q8), Python's summation order
          const int a0 = P.act_ptr[e], LA = P.act_ptr[e + 1] - a0;
          int hits = 0;
          double ap = 0.0, dcg = 0.0, idcg = 0.0;
          int firstpos[kTcMaxK];
          for (int j = 0; j < K; ++j) {
            firstpos[j] = 0x7fffffff;
            for (int a = 0; a < LA; ++a)
              if (P.act_idx[a0 + a] == topi[j]) { firstpos[j] = a; break; }
          }
          for (int i = 1; i <= K; ++i) {
            if (firstpos[i - 1] == 0x7fffffff) continue;
            ++hits;
            int c = 0;
            for (int j = 0; j < i; ++j) c += (firstpos[j] < i) ? 1 : 0;
            ap += (double)c / (double)i;
            if (i <= LA) dcg += P.inv_log2[i - 1];
          }
          for (int i = 1; i <= K && i <= LA; ++i) idcg += P.inv_log2[i - 1];
          const int nun = P.act_nuniq[e];
          double* um = P.user_metrics + e * 4;
          um[0] = (double)hits / (double)K;
          um[1] = (nun > 0) ? (double)hits / (double)nun : 0.0;
          um[2] = (LA > 0) ? ap / (double)LA : 0.0;
          um[3] = (nun > 0 && idcg > 0.0) ? dcg / idcg : 0.0;
          }
        }
      }
    }
    __syncthreads();     // user tile fully consumed before Us / the per-row state are overwritten
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, tmem_cols);
}

// max_i ||V_i||_2 (device scalar, non-negative floats order like their bit patterns).
__global__ void item_norm_max_kernel(const float* __restrict__ V, int64_t nI, int d, float* vmax) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float best = 0.f;
  for (int64_t i = w; i < nI; i += nw) {
    float acc = 0.f;
    for (int k = lane; k < d; k += 32) { const float x = V[i * d + k]; acc = fmaf(x, x, acc); }
    acc = warp_sum(acc);
    best = fmaxf(best, sqrtf(acc));
  }
  if (lane == 0) atomicMax(reinterpret_cast<int*>(vmax), __float_as_int(best * 1.0001f));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

struct TcConfig { int SPS, NST, CAP, MCAP; size_t smem; };

static bool tc_config(int d, int K, TcConfig* c) {
  if (d < 32 || d > 256 || (d % 32) != 0 || K < 1 || K > kTcMaxK) return false;
  c->SPS = (d % 64 == 0 && d <= 128) ? 2 : 1;
  c->CAP = kTcCap;
  const size_t fixed = (size_t)(d / 32) * 16384 + 4 * kTcTM * 4 + 2 * kTcTM * 4 + 16 * 8 + 32;   // Us, xch, tauB/cntB, barriers, tmem slot
  const size_t stage = (size_t)c->SPS * kTcTN * 128;
  int nst = (int)((224 * 1024 - fixed) / stage);
  if (nst > 4) nst = 4;
  if (nst < 2) return false;
  c->NST = nst;
  // what is left (after >= 3 pipeline stages where they fit) stages the first entries of the rows' mask lists
  const char* e = getenv("YR_EVAL_MCAP");
  int want = (e && *e) ? atoi(e) : 32;      // measured (profiles/README.md): 0 -> 1.46 ms, 32 -> 1.43 ms, 64 -> 1.58 ms on the MF workload
  if (want < 0) want = 0;
  if (want > 256) want = 256;
  auto need = [&](int n_st, int mcap) {
    return fixed + (size_t)n_st * stage + (size_t)kTcTM * (mcap + 1) * 4 + (size_t)2 * kTcPend * 256 * 4 + 1024;
  };
  while (want > 0 && need(c->NST, want) > 225 * 1024) {
    if (c->NST > 3) --c->NST; else want /= 2;
  }
  c->MCAP = want;
  c->smem = need(c->NST, c->MCAP);                  // + slack for the 1024-byte alignment of the swizzled tiles
  return true;
}

}  // namespace yr

using namespace yr;

static size_t tc_cand_bytes() { return (size_t)yr_sm_count() * 4 * kTcCap * kTcTM * sizeof(int); }

extern "C" size_t yr_eval_tc_ws_bytes(int64_t n_eval) {
  const size_t rows = ((size_t)(n_eval > 0 ? n_eval : 0) * sizeof(int32_t) + 255) / 256 * 256;
  return 256 + rows + tc_cand_bytes();
}

extern "C" int yr_eval_tc_supported(int d, int K) {
  TcConfig c;
  return tc_config(d, K, &c) && encode_tiled_fn() != nullptr;
}

static int eval_tc_impl(const float* Uemb, int64_t nU, const float* Vemb, const float* Vt,
                        int64_t ldt, int64_t nI, int d, const int64_t* eval_uid, int64_t n_eval,
                        const int32_t* mask_ptr, const int32_t* mask_idx, const int32_t* act_ptr,
                        const int32_t* act_idx, const int32_t* act_nuniq, const double* inv_log2,
                        int K, int64_t* topk_out, float* topk_score, double* user_metrics,
                        double* metric_sums, void* ws, size_t ws_bytes, int32_t* err,
                        int slice, int n_slices, float* xchg, yr_stream stream) {
  if (n_slices < 1 || slice < 0 || slice >= n_slices || (n_slices > 1 && (!xchg || !topk_score))) return YR_ERR_BAD_ARG;
  if (!Uemb || !Vemb || !Vt || !eval_uid || !mask_ptr || !mask_idx || !act_ptr || !act_idx || !act_nuniq ||
      !inv_log2 || !topk_out || !user_metrics || !metric_sums || !ws)
    return YR_ERR_BAD_ARG;
  if (n_eval < 0 || nI <= 0 || nI >= (1 << 24)) return YR_ERR_BAD_ARG;
  TcConfig cfg;
  if (!tc_config(d, K, &cfg)) return YR_ERR_BAD_DIM;
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return YR_ERR_BAD_DIM;
  if (ws_bytes < yr_eval_tc_ws_bytes(n_eval)) return YR_ERR_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(Vemb) & 15) != 0) return YR_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  float* vmax = reinterpret_cast<float*>(ws);
  int32_t* fb_count = reinterpret_cast<int32_t*>(ws) + 1;
  int32_t* fb_rows = reinterpret_cast<int32_t*>(ws) + 64;
  int* cand = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(ws) + yr_eval_tc_ws_bytes(n_eval) - tc_cand_bytes());
  YR_CUDA(cudaMemsetAsync(ws, 0, 64, s));
  if (n_eval > 0) {
    item_norm_max_kernel<<<yr_sm_count() * 4, 256, 0, s>>>(Vemb, nI, d, vmax);
    YR_CHECK_LAUNCH();
    CUtensorMap vmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)nI};
    const cuuint64_t gstride[1] = {(cuuint64_t)d * sizeof(float)};
    const cuuint32_t box[2] = {32u, (cuuint32_t)kTcTN};
    const cuuint32_t estr[2] = {1u, 1u};
    CUresult cr = enc(&vmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(Vemb), gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return YR_ERR_BAD_ARG;
    TcParams P;
    P.Uemb = Uemb; P.nU = nU; P.Vemb = Vemb; P.nI = nI; P.d = d; P.eval_uid = eval_uid; P.n_eval = n_eval;
    P.mask_ptr = mask_ptr; P.mask_idx = mask_idx; P.act_ptr = act_ptr; P.act_idx = act_idx; P.act_nuniq = act_nuniq;
    P.inv_log2 = inv_log2; P.K = K; P.topk_out = topk_out; P.topk_score = topk_score; P.user_metrics = user_metrics;
    P.err = err; P.vmax = vmax; P.fb_count = fb_count; P.fb_rows = fb_rows; P.cand = cand;
    P.SPS = cfg.SPS; P.NST = cfg.NST; P.CAP = cfg.CAP; P.MCAP = cfg.MCAP;
    P.gx = xchg; P.slice = slice; P.n_slices = n_slices;
    YR_CUDA(cudaFuncSetAttribute(eval_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem));
    const int64_t n_utiles = (n_eval + kTcTM - 1) / kTcTM;
    int64_t grid = yr_sm_count();
    if (grid > n_utiles) grid = n_utiles;
    eval_tc_kernel<<<(unsigned)grid, kTcThreads, cfg.smem, s>>>(vmap, P);
    YR_CHECK_LAUNCH();
    // rows the filter could not decide: exact FP32 kernel over the fallback list (usually empty)
    int rc = yr_eval_exact_launch(Uemb, nU, Vt, ldt, nI, d, eval_uid, n_eval, mask_ptr, mask_idx, act_ptr, act_idx,
                                  act_nuniq, inv_log2, K, topk_out, topk_score, user_metrics, err, fb_rows, fb_count,
                                  stream);
    if (rc) return rc;
  }
  if (n_slices > 1) return YR_OK;          // a slice's lists carry no metrics: yr_topk_merge + yr_topk_metrics follow
  return yr_eval_reduce_launch(user_metrics, act_ptr, act_nuniq, n_eval, metric_sums, stream);
}

extern "C" int yr_eval_topk_metrics_tc(const float* Uemb, int64_t nU, const float* Vemb, const float* Vt,
                                       int64_t ldt, int64_t nI, int d, const int64_t* eval_uid, int64_t n_eval,
                                       const int32_t* mask_ptr, const int32_t* mask_idx, const int32_t* act_ptr,
                                       const int32_t* act_idx, const int32_t* act_nuniq, const double* inv_log2,
                                       int K, int64_t* topk_out, float* topk_score, double* user_metrics,
                                       double* metric_sums, void* ws, size_t ws_bytes, int32_t* err,
                                       yr_stream stream) {
  return eval_tc_impl(Uemb, nU, Vemb, Vt, ldt, nI, d, eval_uid, n_eval, mask_ptr, mask_idx, act_ptr, act_idx, act_nuniq, inv_log2,
                      K, topk_out, topk_score, user_metrics, metric_sums, ws, ws_bytes, err, 0, 1, nullptr, stream);
}

extern "C" int yr_eval_topk_metrics_tc_slice(const float* Uemb, int64_t nU, const float* Vemb, const float* Vt,
                                             int64_t ldt, int64_t nI, int d, const int64_t* eval_uid, int64_t n_eval,
                                             const int32_t* mask_ptr, const int32_t* mask_idx, const int32_t* act_ptr,
                                             const int32_t* act_idx, const int32_t* act_nuniq, const double* inv_log2,
                                             int K, int64_t* topk_out, float* topk_score, double* user_metrics,
                                             double* metric_sums, void* ws, size_t ws_bytes, int32_t* err,
                                             int slice, int n_slices, float* xchg, yr_stream stream) {
  return eval_tc_impl(Uemb, nU, Vemb, Vt, ldt, nI, d, eval_uid, n_eval, mask_ptr, mask_idx, act_ptr, act_idx, act_nuniq, inv_log2,
                      K, topk_out, topk_score, user_metrics, metric_sums, ws, ws_bytes, err, slice, n_slices, xchg, stream);
}
