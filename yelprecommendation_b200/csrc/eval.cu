// Full-catalog evaluation (sm_100a): MFTrainer.evaluate / NGCFTrainer.evaluate
// (reference trainers/mf_trainer.py:134-178, trainers/ngcf_trainer.py:134-182, metric.py:7-109) fused into
// one persistent kernel: user-tile x item-tile score GEMM on the FP32 pipe (one fma chain per score, so
// scores are bit-identical to the oracle's canonical order), lazy train-item masking (-3.40282e+38),
// per-user top-K by (score desc, item id asc), and the reference's Precision/Recall/MAP/NDCG quirks.
// Item tiles stream through shared memory with cp.async.bulk (UBLKCP) + mbarrier; the U x I score matrix
// only ever exists as 8x8 register tiles.
#include <float.h>
#include "common.cuh"

namespace yr {

constexpr int kTU = 128;        // users per CTA tile
constexpr int kTI = 128;        // items per tile
constexpr int kKC = 32;         // k-chunk (rows of Vt per pipeline stage)
constexpr int kStages = 3;
constexpr int kEvalThreads = 256;
constexpr int kQCap = 2048;     // candidate queue entries
constexpr int kMCap = 1024;     // per-tile mask list entries
constexpr float kMaskValue = -3.40282e+38f;   // trainers/mf_trainer.py:167 (Q4)
constexpr int kMaxK = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct EvalSmem {
  // dynamic layout (floats unless noted):
  //   Us   [k_res][kTU]      k_res = d_pad when the whole user tile fits, else the largest multiple of kKC that does;
  //                          rows k >= k_res of the transposed user tile are then read from UT (global) instead
  //   Vs   [kStages][kKC][kTI]
  //   topS [kTU][K], topI [kTU][K] (int)
  //   qkey [kQCap] (u32), qval [kQCap]
  //   tmask[kMCap] (u32)
  //   thr  [kTU], mnext[kTU] (int), mcur[kTU] (int), mend[kTU] (int)
  //   bars [kStages] (u64), counters[4] (int)
};

__host__ __device__ inline size_t eval_smem_bytes(int d_pad, int K) {
  size_t f = (size_t)d_pad * kTU + (size_t)kStages * kKC * kTI + 2 * (size_t)kTU * K + 2 * (size_t)kQCap +
             kMCap + 4 * (size_t)kTU;
  return f * 4 + kStages * 8 + 16 + 16;
}

// largest multiple of kKC <= d_pad whose [k][kTU] user tile fits 227 KB next to everything else
inline int eval_resident_k(int d_pad, int K) {
  for (int k = d_pad; k >= kKC; k -= kKC)
    if (eval_smem_bytes(k, K) <= (size_t)227 * 1024) return k;
  return 0;
}
inline int64_t eval_ldu(int64_t n_eval) { return (n_eval + kTU - 1) / kTU * kTU; }

// UT[k][e] = Uemb[eval_uid[e]][k_res + k] for the rows of the user tile that do not fit shared memory (wide tables)
__global__ void __launch_bounds__(256)
gather_transpose_users_kernel(const float* __restrict__ Uemb, int64_t nU, int d, const int64_t* __restrict__ eval_uid,
                              int64_t n_eval, int k_res, int k_rows, int64_t ldu, float* __restrict__ UT) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ldu) return;
  int64_t uid = -1;
  if (e < n_eval) { uid = eval_uid[e]; if (uid < 0 || uid >= nU) uid = -1; }     // a bad id is flagged by the main kernel
  for (int k = blockIdx.y; k < k_rows; k += gridDim.y) {
    const int kk = k_res + k;
    UT[(size_t)k * ldu + e] = (uid >= 0 && kk < d) ? __ldg(Uemb + uid * d + kk) : 0.f;
  }
}

// acc[i][j] += sum_k us[k][user_i] * vs[k][item_j] over one k-chunk, k ascending
template <bool kGlobalU>
__device__ __forceinline__ void eval_chunk_fma(float2 (&acc)[8][4], const float* __restrict__ us, int64_t ustride,
                                               const float* __restrict__ vs, int tu, int ti) {
#pragma unroll 4
  for (int k = 0; k < kKC; ++k) {
    float4 a0, a1;
    if constexpr (kGlobalU) {
      a0 = __ldg(reinterpret_cast<const float4*>(us + k * ustride + tu * 4));
      a1 = __ldg(reinterpret_cast<const float4*>(us + k * ustride + 64 + tu * 4));
    } else {
      a0 = *reinterpret_cast<const float4*>(us + k * kTU + tu * 4);
      a1 = *reinterpret_cast<const float4*>(us + k * kTU + 64 + tu * 4);
    }
    const float4 b0 = *reinterpret_cast<const float4*>(vs + k * kTI + ti * 4);
    const float4 b1 = *reinterpret_cast<const float4*>(vs + k * kTI + 64 + ti * 4);
    const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float2 bv[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
    // packed FFMA2: each component is exactly fmaf(av[i], b, acc) — the canonical chain, two scores per instruction
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = __ffma2_rn(make_float2(av[i], av[i]), bv[j], acc[i][j]);
  }
}

__device__ __forceinline__ bool cand_better(float s_a, int i_a, float s_b, int i_b) {
  // does (s_a, i_a) rank strictly ahead of (s_b, i_b)?  score desc, item id asc
  return (s_a > s_b) || (s_a == s_b && i_a < i_b);
}

__global__ void __launch_bounds__(kEvalThreads, 1)
eval_topk_kernel(const float* __restrict__ Uemb, int64_t nU, const float* __restrict__ Vt, int64_t ldt,
                 int64_t nI, int d, int d_pad, const int64_t* __restrict__ eval_uid, int64_t n_eval,
                 const int32_t* __restrict__ mask_ptr, const int32_t* __restrict__ mask_idx,
                 const int32_t* __restrict__ act_ptr, const int32_t* __restrict__ act_idx,
                 const int32_t* __restrict__ act_nuniq, const double* __restrict__ inv_log2, int K,
                 int64_t* __restrict__ topk_out, float* __restrict__ topk_score,
                 double* __restrict__ user_metrics, int32_t* err, const int32_t* __restrict__ row_list,
                 const int32_t* __restrict__ n_rows_dev, int k_res, const float* __restrict__ UT, int64_t ldu) {
  // optional indirection: evaluate only rows row_list[0 .. *n_rows_dev) (the tensor-core path's undecided rows)
  if (row_list) n_eval = *n_rows_dev;
  auto ROW = [&](int64_t t) -> int64_t { return row_list ? (int64_t)row_list[t] : t; };
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* Us = reinterpret_cast<float*>(smem_raw);
  float* Vs = Us + (size_t)k_res * kTU;
  float* topS = Vs + (size_t)kStages * kKC * kTI;
  int* topI = reinterpret_cast<int*>(topS + (size_t)kTU * K);
  uint32_t* qkey = reinterpret_cast<uint32_t*>(topI + (size_t)kTU * K);
  float* qval = reinterpret_cast<float*>(qkey + kQCap);
  uint32_t* tmask = reinterpret_cast<uint32_t*>(qval + kQCap);
  float* thr = reinterpret_cast<float*>(tmask + kMCap);
  int* mnext = reinterpret_cast<int*>(thr + kTU);
  int* mcur = mnext + kTU;
  int* mend = mcur + kTU;
  uintptr_t bar_addr = (reinterpret_cast<uintptr_t>(mend + kTU) + 7) & ~(uintptr_t)7;
  uint64_t* bars = reinterpret_cast<uint64_t*>(bar_addr);
  int* ctr = reinterpret_cast<int*>(bars + kStages);   // [0] qcount, [1] tm_count

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tu = tid >> 4, ti = tid & 15;   // 16 x 16 thread grid, 8 x 8 scores per thread

  const int64_t n_utiles = (n_eval + kTU - 1) / kTU;
  const int64_t n_itiles = (nI + kTI - 1) / kTI;
  const int n_chunks = d_pad / kKC;
  const int64_t per_ut = n_itiles * n_chunks;
  const int64_t my_utiles = (n_utiles > blockIdx.x) ? (n_utiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t Q = my_utiles * per_ut;
  constexpr uint32_t kStageBytes = kKC * kTI * sizeof(float);

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(bars + s, 1);
    ctr[0] = 0; ctr[1] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto issue = [&](int64_t q) {   // one elected thread: fill stage q % kStages with V chunk q
    const int64_t r = q % per_ut;
    const int64_t it = r / n_chunks;
    const int ch = (int)(r % n_chunks);
    const int st = (int)(q % kStages);
    float* dst = Vs + (size_t)st * kKC * kTI;
    const float* src = Vt + (size_t)ch * kKC * ldt + it * kTI;
    mbar_expect_tx(bars + st, kStageBytes);
    for (int k = 0; k < kKC; ++k) bulk_g2s(dst + k * kTI, src + (size_t)k * ldt, kTI * sizeof(float), bars + st);
  };
  if (tid == 0)
    for (int64_t q = 0; q < kStages && q < Q; ++q) issue(q);

  int64_t q = 0;
  for (int64_t ut = blockIdx.x; ut < n_utiles; ut += gridDim.x) {
    const int64_t e0 = ut * kTU;
    // ---- user tile: gather + transpose into Us[k][u]; reset per-user state ----
    for (int idx = tid; idx < kTU * (k_res / 4); idx += kEvalThreads) {
      const int u = idx % kTU, c4 = idx / kTU;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int64_t e = e0 + u;
      if (e < n_eval && c4 * 4 < d) {
        const int64_t uid = eval_uid[ROW(e)];
        if (uid >= 0 && uid < nU) v = __ldg(reinterpret_cast<const float4*>(Uemb + uid * d) + c4);
        else if (err) atomicExch(err, 1);
      }
      Us[(c4 * 4 + 0) * kTU + u] = v.x; Us[(c4 * 4 + 1) * kTU + u] = v.y;
      Us[(c4 * 4 + 2) * kTU + u] = v.z; Us[(c4 * 4 + 3) * kTU + u] = v.w;
    }
    for (int idx = tid; idx < kTU * K; idx += kEvalThreads) { topS[idx] = -INFINITY; topI[idx] = 0x7fffffff; }
    if (tid < kTU) {
      const int64_t e = e0 + tid;
      thr[tid] = -INFINITY;
      int c = 0, en = 0;
      if (e < n_eval) { const int64_t er = ROW(e); c = mask_ptr[er]; en = mask_ptr[er + 1]; }
      mcur[tid] = c; mend[tid] = en;
      mnext[tid] = (c < en) ? mask_idx[c] : 0x7fffffff;
    }
    __syncthreads();

    for (int64_t it = 0; it < n_itiles; ++it) {
      const int i0 = (int)(it * kTI);
      // ---- masked (user,item) pairs that fall into this item tile ----
      if (tid < kTU) {
        const int tile_end = i0 + kTI;
        int nx = mnext[tid];
        if (nx < tile_end) {
          int c = mcur[tid];
          const int en = mend[tid];
          while (nx < tile_end) {
            const int p = atomicAdd(ctr + 1, 1);
            if (p < kMCap) tmask[p] = ((uint32_t)nx << 7) | (uint32_t)tid;
            ++c;
            nx = (c < en) ? mask_idx[c] : 0x7fffffff;
          }
          mcur[tid] = c; mnext[tid] = nx;
        }
      }
      // ---- scores: acc[i][j] = sum_k Us[k][user_i] * Vs[k][item_j], k ascending ----
      float2 acc[8][4];                  // [user][item pair]: score (i, j) = acc[i][j >> 1].x / .y
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
      for (int ch = 0; ch < n_chunks; ++ch, ++q) {
        const int st = (int)(q % kStages);
        mbar_wait(bars + st, (uint32_t)((q / kStages) & 1));
        const float* vs = Vs + (size_t)st * kKC * kTI;
        if (ch * kKC < k_res) eval_chunk_fma<false>(acc, Us + (size_t)ch * kKC * kTU, kTU, vs, tu, ti);
        else eval_chunk_fma<true>(acc, UT + (size_t)(ch * kKC - k_res) * ldu + e0, ldu, vs, tu, ti);
        __syncthreads();   // everyone is done with stage st
        if (tid == 0 && q + kStages < Q) issue(q + kStages);
      }

      // ---- candidate pass ----
      const int tmc = ctr[1];
      const bool mask_overflow = tmc > kMCap;
      auto masked = [&](int ul, int item) -> bool {
        if (!mask_overflow) {
          const uint32_t key = ((uint32_t)item << 7) | (uint32_t)ul;
          for (int m = 0; m < tmc; ++m)
            if (tmask[m] == key) return true;
          return false;
        }
        const int64_t e = ROW(e0 + ul);   // rare: more than kMCap masked pairs in one tile -> search the CSR
        int lo = mask_ptr[e], hi = mask_ptr[e + 1];
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          const int v = mask_idx[mid];
          if (v == item) return true;
          if (v < item) lo = mid + 1; else hi = mid;
        }
        return false;
      };
      auto try_push = [&](int i, int j) {
        const int ul = (i < 4) ? tu * 4 + i : 64 + tu * 4 + (i - 4);
        const int item = i0 + ((j < 4) ? ti * 4 + j : 64 + ti * 4 + (j - 4));
        float s = (j & 1) ? acc[i][j >> 1].y : acc[i][j >> 1].x;
        if (s >= thr[ul] && item < nI && e0 + ul < n_eval) {
          if (masked(ul, item)) s = kMaskValue;
          if (s >= thr[ul]) {
            const int p = atomicAdd(ctr + 0, 1);
            if (p < kQCap) { qkey[p] = ((uint32_t)item << 7) | (uint32_t)ul; qval[p] = s; }
          }
        }
      };
      auto drain = [&]() {   // warp `warp` owns users warp*16 .. warp*16+15
        const int n = min(ctr[0], kQCap);
        for (int base = 0; base < n; base += 32) {
          const int qi = base + lane;
          uint32_t key = 0; float val = 0.f;
          const bool valid = qi < n;
          if (valid) { key = qkey[qi]; val = qval[qi]; }
          unsigned mine = __ballot_sync(kFull, valid && (int)((key & 127u) >> 4) == warp);
          while (mine) {
            const int src = __ffs(mine) - 1;
            mine &= mine - 1;
            const uint32_t ck = __shfl_sync(kFull, key, src);
            const float cs = __shfl_sync(kFull, val, src);
            const int ul = (int)(ck & 127u), citem = (int)(ck >> 7);
            float s_j = -INFINITY; int i_j = 0x7fffffff;
            if (lane < K) { s_j = topS[ul * K + lane]; i_j = topI[ul * K + lane]; }
            const unsigned ahead = __ballot_sync(kFull, lane < K && cand_better(s_j, i_j, cs, citem));
            const int pos = __popc(ahead);
            const float s_up = __shfl_up_sync(kFull, s_j, 1);
            const int i_up = __shfl_up_sync(kFull, i_j, 1);
            if (pos < K && lane < K && lane >= pos) {
              const float ns = (lane == pos) ? cs : s_up;
              const int ni = (lane == pos) ? citem : i_up;
              topS[ul * K + lane] = ns; topI[ul * K + lane] = ni;
              if (lane == K - 1) thr[ul] = ns;
            }
            __syncwarp();
          }
        }
      };

#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) try_push(i, j);
      __syncthreads();
      if (ctr[0] <= kQCap) {
        drain();
        __syncthreads();
      } else {
        // queue overflow (first tiles of a user tile): redo in 8 bounded rounds
        for (int r = 0; r < 8; ++r) {
          __syncthreads();
          if (tid == 0) ctr[0] = 0;
          __syncthreads();
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (j == r) try_push(i, j);
          __syncthreads();
          drain();
        }
        __syncthreads();
      }
      if (tid == 0) { ctr[0] = 0; ctr[1] = 0; }
      __syncthreads();
    }

    // ---- finalize this user tile: top-K out + metric terms (metric.py quirks Q6-Q8) ----
    for (int uu = 0; uu < 16; ++uu) {
      const int ul = warp * 16 + uu;
      if (e0 + ul >= n_eval) break;
      const int64_t e = ROW(e0 + ul);
      int pi = -1; float ps = 0.f;
      if (lane < K) {
        pi = topI[ul * K + lane]; ps = topS[ul * K + lane];
        if (pi == 0x7fffffff) pi = -1;
        topk_out[e * K + lane] = (int64_t)pi;
        if (topk_score) topk_score[e * K + lane] = ps;
      }
      const int a0 = act_ptr[e], a1 = act_ptr[e + 1];
      const int LA = a1 - a0;
      int firstpos = 0x7fffffff;       // first index in `actual` holding my predicted item
      if (lane < K && pi >= 0)
        for (int a = 0; a < LA; ++a)
          if (act_idx[a0 + a] == pi) { firstpos = a; break; }
      const bool hit = firstpos != 0x7fffffff;
      const int hits = __popc(__ballot_sync(kFull, hit));
      // AP: rank i = lane+1 contributes |set(A[:i]) & set(P[:i])| / i when P[i-1] in A
      int c = 0;
      for (int j = 0; j < K; ++j) {
        const int fp = __shfl_sync(kFull, firstpos, j);
        if (j <= lane && fp < lane + 1) ++c;
      }
      const double ap_term = (lane < K && hit) ? (double)c / (double)(lane + 1) : 0.0;
      const double dcg_term = (lane < K && lane < LA && hit) ? inv_log2[lane] : 0.0;
      double ap = 0.0, dcg = 0.0, idcg = 0.0;
      for (int j = 0; j < K; ++j) {     // Python sum(): left to right, skipped terms are simply absent
        const double t1 = __shfl_sync(kFull, ap_term, j);
        const double t2 = __shfl_sync(kFull, dcg_term, j);
        const int h = __shfl_sync(kFull, (int)hit, j);
        if (h) ap += t1;
        if (h && j < LA) dcg += t2;
        if (j < LA) idcg += inv_log2[j];
      }
      if (lane == 0) {
        const int nun = act_nuniq[e];
        double* um = user_metrics + e * 4;
        um[0] = (double)hits / (double)K;
        um[1] = (nun > 0) ? (double)hits / (double)nun : 0.0;
        um[2] = (LA > 0) ? ap / (double)LA : 0.0;
        um[3] = (nun > 0 && idcg > 0.0) ? dcg / idcg : 0.0;
      }
    }
    __syncthreads();
  }
}

// Deterministic reduction of the per-row metric terms (row order within a thread, fixed tree across threads).
__global__ void __launch_bounds__(1024)
eval_reduce_kernel(const double* __restrict__ user_metrics, const int32_t* __restrict__ act_ptr,
                   const int32_t* __restrict__ act_nuniq, int64_t n_eval, double* __restrict__ sums) {
  __shared__ double sh[6][1024];
  double a[6] = {0, 0, 0, 0, 0, 0};
  for (int64_t e = threadIdx.x; e < n_eval; e += 1024) {
    a[0] += user_metrics[e * 4 + 0]; a[1] += user_metrics[e * 4 + 1];
    a[2] += user_metrics[e * 4 + 2]; a[3] += user_metrics[e * 4 + 3];
    a[4] += (act_nuniq[e] > 0) ? 1.0 : 0.0;
    a[5] += (act_ptr[e + 1] - act_ptr[e] > 0) ? 1.0 : 0.0;
  }
  for (int m = 0; m < 6; ++m) sh[m][threadIdx.x] = a[m];
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if (threadIdx.x < s)
      for (int m = 0; m < 6; ++m) sh[m][threadIdx.x] += sh[m][threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x < 6) sums[threadIdx.x] = sh[threadIdx.x][0];
}

__global__ void transpose_items_kernel(const float* __restrict__ V, int64_t nI, int d, float* __restrict__ Vt,
                                       int64_t ldt, int d_rows) {
  __shared__ float tile[32][33];
  const int64_t i0 = (int64_t)blockIdx.x * 32;
  const int k0 = blockIdx.y * 32;
  // read V[i0+ty.., k0+tx] coalesced along k
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int64_t i = i0 + r;
    const int k = k0 + threadIdx.x;
    tile[r][threadIdx.x] = (i < nI && k < d) ? V[i * d + k] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int k = k0 + r;
    const int64_t i = i0 + threadIdx.x;
    if (k < d_rows && i < ldt) Vt[(int64_t)k * ldt + i] = tile[threadIdx.x][r];
  }
}


}  // namespace yr

using namespace yr;

static inline int pad32(int d) { return (d + 31) / 32 * 32; }

// Vt has pad32(d) rows of ldt floats; rows >= d and columns >= nI are zero.
extern "C" int yr_transpose_items(const float* V, int64_t nI, int d, float* Vt, int64_t ldt,
                                  yr_stream stream) {
  if (!V || !Vt || nI <= 0 || d <= 0 || ldt < nI) return YR_ERR_BAD_ARG;
  const int d_rows = pad32(d);
  dim3 block(32, 8), grid((unsigned)((ldt + 31) / 32), (unsigned)(d_rows / 32));
  transpose_items_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(V, nI, d, Vt, ldt, d_rows);
  YR_CHECK_LAUNCH();
  return YR_OK;
}

extern "C" size_t yr_eval_ws_bytes(int64_t n_eval, int d, int K) {
  // tables up to ~288 floats wide keep all state in shared memory; wider ones spill the tail of the transposed user
  // tile, [pad32(d) - k_res][round_up(n_eval, 128)] floats, to the workspace
  if (n_eval <= 0 || d <= 0 || K <= 0 || K > kMaxK) return 256;
  const int d_pad = pad32(d), k_res = eval_resident_k(d_pad, K);
  return 256 + (size_t)(d_pad - k_res) * (size_t)eval_ldu(n_eval) * sizeof(float);
}

extern "C" int yr_eval_topk_metrics(const float* Uemb, int64_t nU, const float* Vt, int64_t ldt,
                                    int64_t nI, int d, const int64_t* eval_uid, int64_t n_eval,
                                    const int32_t* mask_ptr, const int32_t* mask_idx,
                                    const int32_t* act_ptr, const int32_t* act_idx,
                                    const int32_t* act_nuniq, const double* inv_log2, int K,
                                    int64_t* topk_out, float* topk_score, double* user_metrics,
                                    double* metric_sums, void* ws, size_t ws_bytes, int32_t* err,
                                    yr_stream stream) {
  int rc = yr_eval_exact_launch(Uemb, nU, Vt, ldt, nI, d, eval_uid, n_eval, mask_ptr, mask_idx, act_ptr, act_idx,
                                act_nuniq, inv_log2, K, topk_out, topk_score, user_metrics, err, nullptr, nullptr,
                                stream, ws, ws_bytes);
  if (rc) return rc;
  return yr_eval_reduce_launch(user_metrics, act_ptr, act_nuniq, n_eval, metric_sums, stream);
}

// Internal launchers shared with the tensor-core path (eval_tc.cu): the exact kernel over all rows, or over
// row_list[0 .. *n_rows_dev) when a list is given (n_eval then only bounds the grid).
int yr_eval_exact_launch(const float* Uemb, int64_t nU, const float* Vt, int64_t ldt, int64_t nI, int d,
                         const int64_t* eval_uid, int64_t n_eval, const int32_t* mask_ptr,
                         const int32_t* mask_idx, const int32_t* act_ptr, const int32_t* act_idx,
                         const int32_t* act_nuniq, const double* inv_log2, int K, int64_t* topk_out,
                         float* topk_score, double* user_metrics, int32_t* err, const int32_t* row_list,
                         const int32_t* n_rows_dev, yr_stream stream, void* ws, size_t ws_bytes) {
  if (!Uemb || !Vt || !eval_uid || !mask_ptr || !mask_idx || !act_ptr || !act_idx || !act_nuniq ||
      !inv_log2 || !topk_out || !user_metrics)
    return YR_ERR_BAD_ARG;
  if (n_eval < 0 || nI <= 0 || d <= 0 || K <= 0 || K > kMaxK) return YR_ERR_BAD_ARG;
  if ((d & 3) != 0 || nI >= (1 << 24)) return YR_ERR_BAD_DIM;
  if (ldt % kTI != 0 || ldt < nI || (ldt & 3) != 0) return YR_ERR_BAD_ARG;   // tiles must not run off Vt
  if (n_eval == 0) return YR_OK;
  const int d_pad = pad32(d);
  const int k_res = eval_resident_k(d_pad, K);
  if (k_res <= 0) return YR_ERR_BAD_DIM;
  const size_t smem = eval_smem_bytes(k_res, K);
  float* UT = nullptr;
  const int64_t ldu = eval_ldu(n_eval);
  if (k_res < d_pad) {            // wide table: rows k >= k_res of the transposed user tile live in the workspace
    if (row_list) return YR_ERR_BAD_DIM;                     // (the tensor-core path, d <= 256, never gets here)
    if (!ws || ws_bytes < yr_eval_ws_bytes(n_eval, d, K)) return YR_ERR_WORKSPACE;
    UT = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(ws) + 256);
    const int k_rows = d_pad - k_res;
    dim3 g((unsigned)((ldu + 255) / 256), (unsigned)(k_rows < 64 ? k_rows : 64));
    gather_transpose_users_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(Uemb, nU, d, eval_uid, n_eval, k_res, k_rows, ldu, UT);
    YR_CHECK_LAUNCH();
  }
  YR_CUDA(cudaFuncSetAttribute(eval_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t n_utiles = (n_eval + kTU - 1) / kTU;
  int64_t grid = yr_sm_count();
  if (grid > n_utiles) grid = n_utiles;
  eval_topk_kernel<<<(unsigned)grid, kEvalThreads, smem, (cudaStream_t)stream>>>(
      Uemb, nU, Vt, ldt, nI, d, d_pad, eval_uid, n_eval, mask_ptr, mask_idx, act_ptr, act_idx, act_nuniq,
      inv_log2, K, topk_out, topk_score, user_metrics, err, row_list, n_rows_dev, k_res, UT, ldu);
  YR_CHECK_LAUNCH();
  return YR_OK;
}

int yr_eval_reduce_launch(const double* user_metrics, const int32_t* act_ptr, const int32_t* act_nuniq,
                          int64_t n_eval, double* metric_sums, yr_stream stream) {
  if (!user_metrics || !act_ptr || !act_nuniq || !metric_sums) return YR_ERR_BAD_ARG;
  eval_reduce_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(user_metrics, act_ptr, act_nuniq, n_eval, metric_sums);
  YR_CHECK_LAUNCH();
  return YR_OK;
}

// ---------------------------------------------------------------------------------------------
// Stand-alone pieces of the same path: one-row masked top-K and metrics of given recommendations.
// ---------------------------------------------------------------------------------------------
namespace yr {

// one block; K rounds of block-wide arg-best over scores with the mask applied on the fly.
__global__ void __launch_bounds__(1024)
topk_masked_row_kernel(const float* __restrict__ pred, int64_t nI, const int64_t* __restrict__ mask_idx,
                       int64_t n_mask, int K, int64_t* __restrict__ topk_out, unsigned char* __restrict__ flags) {
  __shared__ float s_s[32];
  __shared__ int64_t s_i[32];
  __shared__ float best_s;
  __shared__ int64_t best_i;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int64_t i = tid; i < nI; i += blockDim.x) flags[i] = 0;
  __syncthreads();
  for (int64_t m = tid; m < n_mask; m += blockDim.x) {
    int64_t j = mask_idx[m];
    if (j < 0) j += nI;
    if (j >= 0 && j < nI) flags[j] = 1;
  }
  __syncthreads();
  float prev_s = INFINITY;
  int64_t prev_i = -1;
  for (int r = 0; r < K; ++r) {
    float bs = -INFINITY;
    int64_t bi = -1;
    for (int64_t i = tid; i < nI; i += blockDim.x) {
      const float s = flags[i] ? kMaskValue : pred[i];
      // strictly after the previous pick in (score desc, id asc) order
      const bool after = (s < prev_s) || (s == prev_s && i > prev_i);
      if (after && (bi < 0 || s > bs || (s == bs && i < bi))) { bs = s; bi = i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(kFull, bs, o);
      const int64_t oi = __shfl_xor_sync(kFull, bi, o);
      if (oi >= 0 && (bi < 0 || os > bs || (os == bs && oi < bi))) { bs = os; bi = oi; }
    }
    if (lane == 0) { s_s[warp] = bs; s_i[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
      bs = (lane < (int)(blockDim.x >> 5)) ? s_s[lane] : -INFINITY;
      bi = (lane < (int)(blockDim.x >> 5)) ? s_i[lane] : -1;
      for (int o = 16; o > 0; o >>= 1) {
        const float os = __shfl_xor_sync(kFull, bs, o);
        const int64_t oi = __shfl_xor_sync(kFull, bi, o);
        if (oi >= 0 && (bi < 0 || os > bs || (os == bs && oi < bi))) { bs = os; bi = oi; }
      }
      if (lane == 0) { best_s = bs; best_i = bi; topk_out[r] = bi; }
    }
    __syncthreads();
    prev_s = best_s;
    prev_i = best_i;
    __syncthreads();
  }
}

// Batched form for model-agnostic evaluators (DCN's chunked evaluator, trainers/dcn_trainer.py:145-203): one block per score
// row, the row's mask list comes from a CSR, masked positions take `mask_value` (DCN masks sigmoid outputs with 0, :191).
__global__ void __launch_bounds__(1024)
topk_masked_rows_kernel(const float* __restrict__ pred, int64_t ld, int64_t nI, const int32_t* __restrict__ mask_ptr,
                        const int32_t* __restrict__ mask_idx, float mask_value, int K, int64_t* __restrict__ topk_out,
                        unsigned char* __restrict__ flags_all) {
  __shared__ float s_s[32];
  __shared__ int64_t s_i[32];
  __shared__ float best_s;
  __shared__ int64_t best_i;
  const int64_t row = blockIdx.x;
  const float* p = pred + row * ld;
  unsigned char* flags = flags_all + row * nI;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int64_t i = tid; i < nI; i += blockDim.x) flags[i] = 0;
  __syncthreads();
  for (int m = mask_ptr[row] + tid; m < mask_ptr[row + 1]; m += blockDim.x) {
    const int j = mask_idx[m];
    if (j >= 0 && j < nI) flags[j] = 1;
  }
  __syncthreads();
  float prev_s = INFINITY;
  int64_t prev_i = -1;
  for (int r = 0; r < K; ++r) {
    float bs = -INFINITY;
    int64_t bi = -1;
    for (int64_t i = tid; i < nI; i += blockDim.x) {
      const float s = flags[i] ? mask_value : p[i];
      const bool after = (s < prev_s) || (s == prev_s && i > prev_i);
      if (after && (bi < 0 || s > bs || (s == bs && i < bi))) { bs = s; bi = i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(kFull, bs, o);
      const int64_t oi = __shfl_xor_sync(kFull, bi, o);
      if (oi >= 0 && (bi < 0 || os > bs || (os == bs && oi < bi))) { bs = os; bi = oi; }
    }
    if (lane == 0) { s_s[warp] = bs; s_i[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
      bs = (lane < (int)(blockDim.x >> 5)) ? s_s[lane] : -INFINITY;
      bi = (lane < (int)(blockDim.x >> 5)) ? s_i[lane] : -1;
      for (int o = 16; o > 0; o >>= 1) {
        const float os = __shfl_xor_sync(kFull, bs, o);
        const int64_t oi = __shfl_xor_sync(kFull, bi, o);
        if (oi >= 0 && (bi < 0 || os > bs || (os == bs && oi < bi))) { bs = os; bi = oi; }
      }
      if (lane == 0) { best_s = bs; best_i = bi; topk_out[row * K + r] = bi; }
    }
    __syncthreads();
    prev_s = best_s;
    prev_i = best_i;
    __syncthreads();
  }
}

// warp per row; tolerates duplicates inside `predicted` exactly like Python's sets do.
__global__ void __launch_bounds__(256)
topk_metrics_kernel(const int64_t* __restrict__ predicted, int64_t ldp, int64_t n,
                    const int32_t* __restrict__ act_ptr, const int32_t* __restrict__ act_idx,
                    const int32_t* __restrict__ act_nuniq, const double* __restrict__ inv_log2, int K,
                    double* __restrict__ user_metrics) {
  const int lane = threadIdx.x & 31;
  const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (e >= n) return;
  const int64_t pi = (lane < K) ? predicted[e * ldp + lane] : -1;
  const int a0 = act_ptr[e], LA = act_ptr[e + 1] - a0;
  int firstpos = 0x7fffffff;
  if (lane < K)
    for (int a = 0; a < LA; ++a)
      if ((int64_t)act_idx[a0 + a] == pi) { firstpos = a; break; }
  const bool hit = firstpos != 0x7fffffff;
  bool is_first = lane < K;
  for (int j = 0; j < K; ++j) {
    const int64_t pj = __shfl_sync(kFull, pi, j);
    if (j < lane && pj == pi) is_first = false;
  }
  const int hits = __popc(__ballot_sync(kFull, hit && is_first));
  int c = 0;
  for (int j = 0; j < K; ++j) {
    const int fp = __shfl_sync(kFull, firstpos, j);
    const int fj = __shfl_sync(kFull, (int)is_first, j);
    if (j <= lane && fj && fp < lane + 1) ++c;
  }
  const double ap_term = (lane < K && hit) ? (double)c / (double)(lane + 1) : 0.0;
  double ap = 0.0, dcg = 0.0, idcg = 0.0;
  for (int j = 0; j < K; ++j) {
    const double t1 = __shfl_sync(kFull, ap_term, j);
    const int h = __shfl_sync(kFull, (int)hit, j);
    if (h) ap += t1;
    if (h && j < LA) dcg += inv_log2[j];
    if (j < LA) idcg += inv_log2[j];
  }
  if (lane == 0) {
    const int nun = act_nuniq[e];
    double* um = user_metrics + e * 4;
    um[0] = (double)hits / (double)K;
    um[1] = (nun > 0) ? (double)hits / (double)nun : 0.0;
    um[2] = (LA > 0) ? ap / (double)LA : 0.0;
    um[3] = (nun > 0 && idcg > 0.0) ? dcg / idcg : 0.0;
  }
}
}  // namespace yr

extern "C" int yr_topk_masked_row(const float* pred, int64_t nI, const int64_t* mask_idx, int64_t n_mask, int K,
                                  int64_t* topk_out, yr_stream stream) {
  if (!pred || !topk_out || nI <= 0 || K <= 0 || K > nI || n_mask < 0 || (n_mask > 0 && !mask_idx))
    return YR_ERR_BAD_ARG;
  unsigned char* flags = nullptr;
  cudaStream_t s = (cudaStream_t)stream;
  YR_CUDA(cudaMallocAsync((void**)&flags, (size_t)nI, s));   // stream-ordered scratch, freed below
  topk_masked_row_kernel<<<1, 1024, 0, s>>>(pred, nI, mask_idx, n_mask, K, topk_out, flags);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(flags, s);
  return (int)e;
}

extern "C" int yr_topk_masked_rows(const float* pred, int64_t ld, int64_t n_rows, int64_t nI, const int32_t* mask_ptr,
                                   const int32_t* mask_idx, float mask_value, int K, int64_t* topk_out, void* ws,
                                   size_t ws_bytes, yr_stream stream) {
  if (!pred || !topk_out || !mask_ptr || !mask_idx || !ws || nI <= 0 || K <= 0 || K > nI || n_rows < 0 || ld < nI)
    return YR_ERR_BAD_ARG;
  if (ws_bytes < (size_t)n_rows * (size_t)nI) return YR_ERR_WORKSPACE;
  if (n_rows == 0) return YR_OK;
  topk_masked_rows_kernel<<<(unsigned)n_rows, 1024, 0, (cudaStream_t)stream>>>(pred, ld, nI, mask_ptr, mask_idx, mask_value, K,
                                                                              topk_out, (unsigned char*)ws);
  YR_CHECK_LAUNCH();
  return YR_OK;
}

// Merge of per-slice recommendations: the catalog is cut into S disjoint item slices, slice s delivered its K best items per
// row (local ids, exact scores); the K best of the S * K candidates by (score desc, GLOBAL item id asc) are the K best of the
// whole catalog, in the same order the unsliced evaluation gives. One thread per row, S * K <= 64.
namespace yr {
__global__ void __launch_bounds__(256)
topk_merge_kernel(const int64_t* __restrict__ ids, const float* __restrict__ sc, int S, int64_t n, int K,
                  const int64_t* __restrict__ id_offset, int64_t* __restrict__ out_ids, float* __restrict__ out_sc) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  unsigned long long taken = 0ull;
  for (int k = 0; k < K; ++k) {
    int best = -1;
    float bs = 0.f;
    int64_t bi = 0;
    for (int s = 0; s < S; ++s) {
      const int64_t off = id_offset[s];
      for (int j = 0; j < K; ++j) {
        const int c = s * K + j;
        if ((taken >> c) & 1ull) continue;
        const int64_t lid = ids[((int64_t)s * n + r) * K + j];
        if (lid < 0) continue;                        // padding: the slice kept fewer than K items for this row
        const float v = sc[((int64_t)s * n + r) * K + j];
        const int64_t id = lid + off;
        if (best < 0 || v > bs || (v == bs && id < bi)) { best = c; bs = v; bi = id; }
      }
    }
    if (best >= 0) taken |= 1ull << best; else { bi = -1; bs = -INFINITY; }
    out_ids[r * K + k] = bi;
    if (out_sc) out_sc[r * K + k] = bs;
  }
}
}  // namespace yr

extern "C" int yr_topk_merge(const int64_t* ids, const float* scores, int S, int64_t n, int K, const int64_t* id_offset,
                             int64_t* out_ids, float* out_scores, yr_stream stream) {
  if (!ids || !scores || !id_offset || !out_ids || S < 1 || K < 1 || S * K > 64 || n < 0) return YR_ERR_BAD_ARG;
  if (n == 0) return YR_OK;
  yr::topk_merge_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ids, scores, S, n, K, id_offset, out_ids,
                                                                                    out_scores);
  YR_CHECK_LAUNCH();
  return YR_OK;
}

extern "C" int yr_topk_metrics(const int64_t* predicted, int64_t ldp, int64_t n, const int32_t* act_ptr,
                               const int32_t* act_idx, const int32_t* act_nuniq, const double* inv_log2, int K,
                               double* user_metrics, double* metric_sums, yr_stream stream) {
  if (!predicted || !act_ptr || !act_idx || !act_nuniq || !inv_log2 || !user_metrics || !metric_sums || n < 0 ||
      K <= 0 || K > kMaxK || ldp < K)
    return YR_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (n > 0) {
    const int64_t blocks = (n * 32 + 255) / 256;
    topk_metrics_kernel<<<(unsigned)blocks, 256, 0, s>>>(predicted, ldp, n, act_ptr, act_idx, act_nuniq, inv_log2,
                                                         K, user_metrics);
    YR_CHECK_LAUNCH();
  }
  eval_reduce_kernel<<<1, 1024, 0, s>>>(user_metrics, act_ptr, act_nuniq, n, metric_sums);
  YR_CHECK_LAUNCH();
  return YR_OK;
}
