// BPR matrix factorisation kernels (sm_100a).
//   yr_mf_score / _bwd      — MatrixFactorization.forward (reference models/mf.py:20-23) and its gradient
//   yr_bpr_loss_fwd / _bwd  — BPRLoss (reference loss.py:19-27)
//   yr_bpr_mf_train         — MFTrainer.train hot loop (reference trainers/mf_trainer.py:100-116) as ONE
//                             persistent cooperative kernel: warp-per-triple gather + dot + log-sigmoid,
//                             sparse gradient accumulate (vector RED), one optimizer update per touched row.
//   yr_bpr_mf_validate      — MFTrainer.validate (reference trainers/mf_trainer.py:118-132)
#include <stdlib.h>
#include "common.cuh"

namespace yr {

// ---------------------------------------------------------------------------------------------
// score: one thread per (user,item) pair, one fp32 fma chain over k (canonical order).
// ---------------------------------------------------------------------------------------------
__global__ void mf_score_kernel(const float* __restrict__ U, const float* __restrict__ V, int64_t nU,
                                int64_t nI, int d, const int64_t* __restrict__ uid,
                                const int64_t* __restrict__ iid, int64_t B, float* __restrict__ out,
                                int32_t* err) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int64_t u = uid[b], i = iid[b];
  if (u < 0 || u >= nU || i < 0 || i >= nI) {
    if (err) atomicExch(err, 1);
    out[b] = 0.f;
    return;
  }
  const float* ur = U + u * d;
  const float* vr = V + i * d;
  float acc = 0.f;
  if ((d & 3) == 0) {
    const float4* u4 = reinterpret_cast<const float4*>(ur);
    const float4* v4 = reinterpret_cast<const float4*>(vr);
    for (int k = 0; k < d / 4; ++k) {
      const float4 a = __ldg(u4 + k), c = __ldg(v4 + k);
      acc = fmaf(a.x, c.x, acc); acc = fmaf(a.y, c.y, acc);
      acc = fmaf(a.z, c.z, acc); acc = fmaf(a.w, c.w, acc);
    }
  } else {
    for (int k = 0; k < d; ++k) acc = fmaf(ur[k], vr[k], acc);
  }
  out[b] = acc;
}

// gradient of score wrt the gathered rows, accumulated into dense gU/gV (autograd-compat path).
template <int VPL>
__global__ void mf_score_bwd_kernel(const float* __restrict__ U, const float* __restrict__ V,
                                    int64_t nU, int64_t nI, const int64_t* __restrict__ uid,
                                    const int64_t* __restrict__ iid, int64_t B,
                                    const float* __restrict__ gout, float* gU, float* gV) {
  constexpr int D = VPL * 32;
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= B) return;
  const int64_t u = uid[w], i = iid[w];
  if (u < 0 || u >= nU || i < 0 || i >= nI) return;
  const float g = gout[w];
  Row<VPL> ur = ld_row<VPL>(U + u * D, lane), vr = ld_row<VPL>(V + i * D, lane), a, c;
#pragma unroll
  for (int j = 0; j < VPL; ++j) { a.x[j] = g * vr.x[j]; c.x[j] = g * ur.x[j]; }
  red_row<VPL>(gU + u * D, lane, a);
  red_row<VPL>(gV + i * D, lane, c);
}

__global__ void mf_score_bwd_generic_kernel(const float* __restrict__ U, const float* __restrict__ V,
                                            int64_t nU, int64_t nI, int d,
                                            const int64_t* __restrict__ uid,
                                            const int64_t* __restrict__ iid, int64_t B,
                                            const float* __restrict__ gout, float* gU, float* gV) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= B) return;
  const int64_t u = uid[w], i = iid[w];
  if (u < 0 || u >= nU || i < 0 || i >= nI) return;
  const float g = gout[w];
  for (int k = lane; k < d; k += 32) {
    atomicAdd(gU + u * d + k, g * V[i * d + k]);
    atomicAdd(gV + i * d + k, g * U[u * d + k]);
  }
}

// ---------------------------------------------------------------------------------------------
// BPR loss, stand-alone (autograd-compat path). Single block: B is a batch (2048).
// ---------------------------------------------------------------------------------------------
__global__ void bpr_loss_fwd_kernel(const float* __restrict__ pos, const float* __restrict__ neg,
                                    int64_t B, float* __restrict__ loss) {
  __shared__ double part[32];
  double acc = 0.0;
  for (int64_t b = threadIdx.x; b < B; b += blockDim.x) acc += (double)neg_logsigmoid(pos[b] - neg[b]);
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = (threadIdx.x < (blockDim.x >> 5)) ? part[threadIdx.x] : 0.0;
    v = warp_sum_d(v);
    if (threadIdx.x == 0) *loss = (float)(v / (double)B);
  }
}

__global__ void bpr_loss_bwd_kernel(const float* __restrict__ pos, const float* __restrict__ neg,
                                    int64_t B, const float* __restrict__ gloss,
                                    float* __restrict__ gpos, float* __restrict__ gneg) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float go = gloss ? *gloss : 1.f;
  const float g = (go / (float)B) * neg_logsigmoid_grad(pos[b] - neg[b]);
  gpos[b] = g;
  gneg[b] = -g;
}

// ---------------------------------------------------------------------------------------------
// Fused persistent trainer.
//
// Two step flavours inside one cooperative kernel:
//  * register path — plain SGD (wd = 0), at most one triple per warp, yr_mf_state.deterministic == 0: the three
//    gradient rows stay in registers across the grid barrier and land as vector REDs of -lr * g straight in the tables.
//    Duplicate rows of a batch are then summed in arrival order (ulp-level run-to-run differences, inside 1e-5).
//  * ordered path — everything else (Adam / AdamW / weight decay = torch's dense semantics, large batches, wide rows,
//    or deterministic != 0): NO atomics on floats. A pre-pass (mf_sort_batches_kernel) has grouped every batch's ids by
//    table row; phase 1 stores the per-triple gradient rows g*p, g*n, g*u (plain stores); phase 2 forms each touched
//    row's gradient as torch's autograd does on the CPU — per embedding call a sequential sum in batch order
//    (embedding_dense_backward), then the two calls added (AccumulateGrad) — and applies ONE optimizer update.
//    Bit-identical from run to run, and bit-identical to oracle/yr_oracle.c (which restates the same order).
//
// counters layout (int32): [4] = parity of the next step; the rest is spare. Per-CTA loss partials (double, fixed
// summation order) live in the workspace.
// ---------------------------------------------------------------------------------------------
// 512-thread CTAs (half the CTAs at the grid barrier) up to d = 256; 256 threads beyond, where a thread holds up to
// four rows of 16 / 32 floats (255 registers available).
template <int VPL> struct TrainCfg {
  static constexpr int kThreads = VPL > 8 ? 256 : 512;
  static constexpr int kWarps = kThreads / 32;
};

constexpr int kMaxTrainCtas = 4096;            // loss partial slots per parity
constexpr int kSortSmemCap = 16384;            // composites a CTA sorts in shared memory (128 KB)
constexpr unsigned long long kBadKey = 0xFFFFFFFFull;

struct MfWs {
  unsigned long long* sortedU;   // [n_steps x strideU]  (uid << 32) | b, ascending
  unsigned long long* sortedV;   // [n_steps x strideV]  (item << 32) | (which << 31) | b, which 0 = positive call, 1 = negative
  float* GR;                     // [3 x B x d]: g*p, g*n, g*u of every triple of the current step
  double* cta_loss;              // [2 x kMaxTrainCtas]
  int64_t strideU, strideV;
};

__host__ __device__ inline int64_t pow2_ceil64(int64_t x) { int64_t p = 1; while (p < x) p <<= 1; return p; }

static size_t mf_ws_layout(int64_t n_triples, int64_t B, int d, void* base, MfWs* w) {
  const int64_t n_steps = (n_triples + B - 1) / B;
  const int64_t sU = (B <= kSortSmemCap) ? B : pow2_ceil64(B);
  const int64_t sV = (2 * B <= kSortSmemCap) ? 2 * B : pow2_ceil64(2 * B);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
  const size_t o_u = take((size_t)n_steps * sU * 8), o_v = take((size_t)n_steps * sV * 8);
  const size_t o_g = take((size_t)3 * B * d * 4), o_l = take((size_t)2 * kMaxTrainCtas * 8);
  if (base && w) {
    char* p = (char*)base;
    w->sortedU = (unsigned long long*)(p + o_u); w->sortedV = (unsigned long long*)(p + o_v);
    w->GR = (float*)(p + o_g); w->cta_loss = (double*)(p + o_l);
    w->strideU = sU; w->strideV = sV;
  }
  return off;
}

// ---- pre-pass: one CTA per (step, table) sorts that batch's composites -------------------------------------------
// A triple with an out-of-range id gets the key 0xFFFFFFFF in both arrays (sorted behind every real row, never read).
template <typename Ptr>
__device__ __forceinline__ void bitonic_sort_block(Ptr a, int64_t P) {
  for (int64_t k = 2; k <= P; k <<= 1)
    for (int64_t j = k >> 1; j > 0; j >>= 1) {
      for (int64_t i = threadIdx.x; i < P; i += blockDim.x) {
        const int64_t q = i ^ j;
        if (q > i) {
          const unsigned long long x = a[i], y = a[q];
          if ((x > y) == ((i & k) == 0)) { a[i] = y; a[q] = x; }
        }
      }
      __syncthreads();
    }
}

__global__ void __launch_bounds__(1024)
mf_sort_batches_kernel(const int64_t* __restrict__ uid, const int64_t* __restrict__ pos, const int64_t* __restrict__ neg,
                       int64_t n_triples, int B, int64_t nU, int64_t nI, MfWs ws, int64_t smem_cap) {
  extern __shared__ unsigned long long sk[];
  const int64_t s = blockIdx.x >> 1;
  const int table = blockIdx.x & 1;                       // 0 = user table, 1 = item table
  const int64_t base = s * (int64_t)B;
  const int nb = (int)((n_triples - base < B) ? (n_triples - base) : B);
  const int64_t n = table ? 2 * (int64_t)nb : nb;
  const int64_t P = pow2_ceil64(n);
  unsigned long long* out = table ? ws.sortedV + s * ws.strideV : ws.sortedU + s * ws.strideU;
  const bool in_smem = P <= smem_cap;                     // smem_cap = 0: every slice of this launch is sorted in place
  unsigned long long* a = in_smem ? sk : out;             // large batches: in place on the (power-of-two) output slice
  for (int64_t e = threadIdx.x; e < P; e += blockDim.x) {
    unsigned long long c = ~0ull;
    if (e < n) {
      const int which = (int)(e / nb);
      const int b = (int)(e - (int64_t)which * nb);
      const int64_t u = uid[base + b], p = pos[base + b], q = neg[base + b];
      const bool ok = u >= 0 && u < nU && p >= 0 && p < nI && q >= 0 && q < nI;
      const unsigned long long key = !ok ? kBadKey : (unsigned long long)(table ? (which ? q : p) : u);
      c = (key << 32) | ((unsigned long long)which << 31) | (unsigned long long)b;
    }
    a[e] = c;
  }
  __syncthreads();
  bitonic_sort_block(a, P);
  if (in_smem)
    for (int64_t e = threadIdx.x; e < n; e += blockDim.x) out[e] = sk[e];
}

// gradient of one table row from its segment of the sorted composites (torch's order: per call sequential in batch
// order, then the two calls added). USER: both chains run over the same entries (g*p and -(g*n)); ITEM: entries of the
// positive call (which = 0) come first and feed the first chain, those of the negative call the second.
template <int VPL, bool USER>
__device__ __forceinline__ Row<VPL> segment_grad(const unsigned long long* __restrict__ srt, int i, int n, unsigned key,
                                                 const float* __restrict__ GR, int B, int lane) {
  constexpr int D = VPL * 32;
  Row<VPL> A, Bm;
#pragma unroll
  for (int j = 0; j < VPL; ++j) { A.x[j] = 0.f; Bm.x[j] = 0.f; }
  for (; i < n; ++i) {
    const unsigned long long c = __ldg(srt + i);
    if ((unsigned)(c >> 32) != key) break;
    const int b = (int)(c & 0x7fffffffu);
    if (USER) {
      const Row<VPL> up = ld_row_cg<VPL>(GR + (int64_t)b * D, lane);
      const Row<VPL> un = ld_row_cg<VPL>(GR + ((int64_t)B + b) * D, lane);
#pragma unroll
      for (int j = 0; j < VPL; ++j) { A.x[j] = __fadd_rn(A.x[j], up.x[j]); Bm.x[j] = __fsub_rn(Bm.x[j], un.x[j]); }
    } else {
      const Row<VPL> gv = ld_row_cg<VPL>(GR + (2 * (int64_t)B + b) * D, lane);
      if (((unsigned)c >> 31) == 0u) {
#pragma unroll
        for (int j = 0; j < VPL; ++j) A.x[j] = __fadd_rn(A.x[j], gv.x[j]);
      } else {
#pragma unroll
        for (int j = 0; j < VPL; ++j) Bm.x[j] = __fsub_rn(Bm.x[j], gv.x[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < VPL; ++j) A.x[j] = __fadd_rn(A.x[j], Bm.x[j]);
  return A;
}

template <int VPL>
__global__ void __launch_bounds__(TrainCfg<VPL>::kThreads)
bpr_mf_train_kernel(yr_mf_state st, yr_opt opt, const int64_t* __restrict__ uid,
                    const int64_t* __restrict__ pos, const int64_t* __restrict__ neg,
                    int64_t n_triples, int B, double* loss_sum, float* step_loss, MfWs ws) {
  constexpr int D = VPL * 32;
  constexpr int kTrainWarps = TrainCfg<VPL>::kWarps;
  cg::grid_group grid = cg::this_grid();
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int gwarp = blockIdx.x * kTrainWarps + wib;
  const int nwarps = gridDim.x * kTrainWarps;
  const bool dense = (opt.kind != YR_OPT_SGD) || (opt.weight_decay != 0.0);
  const int64_t n_steps = (n_triples + B - 1) / B;
  int32_t* counters = st.counters;
  const int parity0 = __ldcg(counters + 4) & 1;
  __shared__ double s_part[kTrainWarps];

  // per-CTA loss partial of this step (fixed order inside the CTA), summed after the barrier by warp 0 of block 0 in a
  // fixed order over CTAs: no floating-point atomics anywhere in the loss
  auto publish_loss = [&](double wl, int par) {
    if (lane == 0) s_part[wib] = wl;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
#pragma unroll
      for (int i = 0; i < kTrainWarps; ++i) t += s_part[i];
      ws.cta_loss[par * kMaxTrainCtas + blockIdx.x] = t;
    }
  };
  auto finish_loss = [&](int par, int nb, int64_t s) {        // warp 0 of block 0, after the grid barrier
    if (blockIdx.x == 0 && wib == 0) {
      double t = 0.0;
      for (int i = lane; i < (int)gridDim.x; i += 32) t += __ldcg(ws.cta_loss + par * kMaxTrainCtas + i);
      t = warp_sum_d(t);
      if (lane == 0) {
        const float mean = (float)(t / (double)nb);           // batch mean, then .item()
        if (step_loss) step_loss[s] = mean;
        if (loss_sum) *loss_sum += (double)mean;              // Q1: sum of batch means
        if (s + 1 == n_steps) counters[4] = par ^ 1;
      }
    }
  };

  // ---- register path ----------------------------------------------------------------------------------------------
  if (VPL <= 8 && !dense && B <= nwarps && !st.deterministic) {
    const float neg_lr = -(float)opt.lr;
    int64_t u = 0, p = 0, n = 0;
    if (gwarp < ((n_triples < B) ? (int)n_triples : B)) { u = uid[gwarp]; p = pos[gwarp]; n = neg[gwarp]; }
    for (int64_t s = 0; s < n_steps; ++s) {
      const int par = (parity0 + (int)(s & 1)) & 1;
      const int64_t base = s * (int64_t)B;
      const int nb = (int)((n_triples - base < B) ? (n_triples - base) : B);
      const float inv_nb = 1.f / (float)nb;
      bool act = gwarp < nb;
      if (act && (u < 0 || u >= st.nU || p < 0 || p >= st.nI || n < 0 || n >= st.nI)) {
        if (lane == 0) atomicExch(st.err, 1);
        act = false;
      }
      Row<VPL> gu, gp, gn;
      double wl = 0.0;
      if (act) {
        const Row<VPL> ur = ld_row<VPL>(st.U + u * D, lane);
        const Row<VPL> pr = ld_row<VPL>(st.V + p * D, lane);
        const Row<VPL> nr = ld_row<VPL>(st.V + n * D, lane);
        const float x = __fsub_rn(warp_sum(dot_partial<VPL>(ur, pr)), warp_sum(dot_partial<VPL>(ur, nr)));
        const float g = __fmul_rn(neg_logsigmoid_grad_fast(x), inv_nb);
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
          const float a = __fsub_rn(__fmul_rn(g, pr.x[j]), __fmul_rn(g, nr.x[j]));   // Q2: two rounded products
          const float c = __fmul_rn(g, ur.x[j]);
          gu.x[j] = __fmul_rn(neg_lr, a); gp.x[j] = __fmul_rn(neg_lr, c); gn.x[j] = -gp.x[j];
        }
        wl = (double)neg_logsigmoid_fast(x);
      }
      const int64_t uu = u, pp = p, nn = n;
      // ids of the next step (read-only input): requested before the barrier, consumed after the second one
      const int64_t nbase = base + B;
      if (nbase + gwarp < n_triples && gwarp < B) { u = uid[nbase + gwarp]; p = pos[nbase + gwarp]; n = neg[nbase + gwarp]; }
      publish_loss(wl, par);
      grid_sync(grid);
      if (act) {
        red_row<VPL>(st.U + uu * D, lane, gu);
        red_row<VPL>(st.V + pp * D, lane, gp);
        red_row<VPL>(st.V + nn * D, lane, gn);
      }
      finish_loss(par, nb, s);
      if (s + 1 < n_steps) grid_sync(grid);
    }
    return;
  }

  // ---- ordered path -----------------------------------------------------------------------------------------------
  for (int64_t s = 0; s < n_steps; ++s) {
    const int par = (parity0 + (int)(s & 1)) & 1;
    const int64_t base = s * (int64_t)B;
    const int nb = (int)((n_triples - base < B) ? (n_triples - base) : B);
    const float inv_nb = 1.f / (float)nb;
    const unsigned long long* srtU = ws.sortedU + s * ws.strideU;
    const unsigned long long* srtV = ws.sortedV + s * ws.strideV;

    // ---- phase 1: gather, dots, loss, per-triple gradient rows (plain stores) ------------------------------------
    double wl = 0.0;
    for (int b = gwarp; b < nb; b += nwarps) {
      const int64_t u = uid[base + b], p = pos[base + b], n = neg[base + b];
      if (u < 0 || u >= st.nU || p < 0 || p >= st.nI || n < 0 || n >= st.nI) {
        if (lane == 0) atomicExch(st.err, 1);
        continue;
      }
      const Row<VPL> ur = ld_row<VPL>(st.U + u * D, lane);
      const Row<VPL> pr = ld_row<VPL>(st.V + p * D, lane);
      const Row<VPL> nr = ld_row<VPL>(st.V + n * D, lane);
      const float dp = warp_sum(dot_partial<VPL>(ur, pr));
      const float dn = warp_sum(dot_partial<VPL>(ur, nr));
      const float x = __fsub_rn(dp, dn);
      const float g = __fmul_rn(neg_logsigmoid_grad(x), inv_nb);
      Row<VPL> gup, gun, gv;
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        gup.x[j] = __fmul_rn(g, pr.x[j]);       // user row, positive call; the negative call contributes -(g*n) (Q2)
        gun.x[j] = __fmul_rn(g, nr.x[j]);
        gv.x[j] = __fmul_rn(g, ur.x[j]);        // item rows: +g*u (positive), -(g*u) (negative)
      }
      st_row<VPL>(ws.GR + (int64_t)b * D, lane, gup);
      st_row<VPL>(ws.GR + ((int64_t)B + b) * D, lane, gun);
      st_row<VPL>(ws.GR + (2 * (int64_t)B + b) * D, lane, gv);
      wl += (double)neg_logsigmoid(x);
    }
    if (dense) {
      // heads of the row segments leave (position + 1) in the row's flag: the sweep below finds its segment through it
      const int gthread = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
      for (int i = gthread; i < 3 * nb; i += nthreads) {
        const bool is_u = i < nb;
        const int k = is_u ? i : i - nb;
        const unsigned long long* srt = is_u ? srtU : srtV;
        const unsigned key = (unsigned)(__ldg(srt + k) >> 32);
        if (key == (unsigned)kBadKey) continue;
        if (k == 0 || (unsigned)(__ldg(srt + k - 1) >> 32) != key) (is_u ? st.flagU : st.flagV)[key] = k + 1;
      }
    }
    publish_loss(wl, par);
    grid_sync(grid);

    // ---- phase 2: one optimizer update per row ----------------------------------------------
    OptScalars os;
    opt_scalars_for_step(os, opt, opt.step + (int)s);
    if (!dense) {
      // plain SGD: only the rows of this batch move — one warp per segment head
      for (int i = gwarp; i < 3 * nb; i += nwarps) {
        const bool is_u = i < nb;
        const int k = is_u ? i : i - nb;
        const unsigned long long* srt = is_u ? srtU : srtV;
        const unsigned key = (unsigned)(__ldg(srt + k) >> 32);
        if (key == (unsigned)kBadKey) continue;
        if (k != 0 && (unsigned)(__ldg(srt + k - 1) >> 32) == key) continue;
        float* prow = (is_u ? st.U : st.V) + (int64_t)key * D;
        const Row<VPL> gv = is_u ? segment_grad<VPL, true>(srt, k, nb, key, ws.GR, B, lane)
                                 : segment_grad<VPL, false>(srt, k, 2 * nb, key, ws.GR, B, lane);
        Row<VPL> pv = ld_row<VPL>(prow, lane);
#pragma unroll
        for (int j = 0; j < VPL; ++j) { float m = 0.f, v = 0.f; opt_update(os, pv.x[j], gv.x[j], m, v); }
        st_row<VPL>(prow, lane, pv);
      }
    } else {
      const int64_t nrows = st.nU + st.nI;
      for (int64_t r0 = gwarp; r0 < nrows; r0 += nwarps) {
        const bool is_u = r0 < st.nU;
        const int64_t r = is_u ? r0 : r0 - st.nU;
        int32_t* flag = (is_u ? st.flagU : st.flagV) + r;
        const int f = __ldcg(flag);
        float* prow = (is_u ? st.U : st.V) + r * D;
        Row<VPL> pv = ld_row<VPL>(prow, lane), gv, mv, vv;
        if (f) {
          gv = is_u ? segment_grad<VPL, true>(srtU, f - 1, nb, (unsigned)r, ws.GR, B, lane)
                    : segment_grad<VPL, false>(srtV, f - 1, 2 * nb, (unsigned)r, ws.GR, B, lane);
        } else {
#pragma unroll
          for (int j = 0; j < VPL; ++j) gv.x[j] = 0.f;
        }
        if (opt.kind != YR_OPT_SGD) {
          float* mrow = (is_u ? st.mU : st.mV) + r * D;
          float* vrow = (is_u ? st.vU : st.vV) + r * D;
          mv = ld_row<VPL>(mrow, lane);
          vv = ld_row<VPL>(vrow, lane);
#pragma unroll
          for (int j = 0; j < VPL; ++j) opt_update(os, pv.x[j], gv.x[j], mv.x[j], vv.x[j]);
          st_row<VPL>(mrow, lane, mv);
          st_row<VPL>(vrow, lane, vv);
        } else {
#pragma unroll
          for (int j = 0; j < VPL; ++j) { float m = 0.f, v = 0.f; opt_update(os, pv.x[j], gv.x[j], m, v); }
        }
        st_row<VPL>(prow, lane, pv);
        if (f && lane == 0) *flag = 0;
      }
    }
    finish_loss(par, nb, s);
    if (s + 1 < n_steps) grid_sync(grid);
  }
}

// validate: one block per batch.
template <int VPL>
__global__ void __launch_bounds__(256)
bpr_mf_validate_kernel(const float* __restrict__ U, const float* __restrict__ V, int64_t nU,
                       int64_t nI, const int64_t* __restrict__ uid, const int64_t* __restrict__ pos,
                       const int64_t* __restrict__ neg, int64_t n_triples, int B, double* loss_sum,
                       float* step_loss, int32_t* err) {
  constexpr int D = VPL * 32;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t s = blockIdx.x;
  const int64_t base = s * (int64_t)B;
  const int nb = (int)((n_triples - base < B) ? (n_triples - base) : B);
  __shared__ double s_part[8];
  double wl = 0.0;
  for (int b = wib; b < nb; b += 8) {
    const int64_t u = uid[base + b], p = pos[base + b], n = neg[base + b];
    if (u < 0 || u >= nU || p < 0 || p >= nI || n < 0 || n >= nI) {
      if (lane == 0 && err) atomicExch(err, 1);
      continue;
    }
    const Row<VPL> ur = ld_row<VPL>(U + u * D, lane);
    const Row<VPL> pr = ld_row<VPL>(V + p * D, lane);
    const Row<VPL> nr = ld_row<VPL>(V + n * D, lane);
    const float x = warp_sum(dot_partial<VPL>(ur, pr)) - warp_sum(dot_partial<VPL>(ur, nr));
    wl += (double)neg_logsigmoid(x);
  }
  if (lane == 0) s_part[wib] = wl;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += s_part[i];
    const float mean = (float)(t / (double)nb);
    if (step_loss) step_loss[s] = mean;
    if (loss_sum) atomicAdd(loss_sum, (double)mean);
  }
}

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

}  // namespace yr

using namespace yr;

extern "C" int yr_version(void) { return 100; }

extern "C" int yr_device_sm_count(int* sm_count_h) {
  if (!sm_count_h) return YR_ERR_BAD_ARG;
  int dev = 0;
  YR_CUDA(cudaGetDevice(&dev));
  YR_CUDA(cudaDeviceGetAttribute(sm_count_h, cudaDevAttrMultiProcessorCount, dev));
  return YR_OK;
}

extern "C" int yr_mf_score(const float* U, const float* V, int64_t nU, int64_t nI, int d,
                           const int64_t* uid, const int64_t* iid, int64_t B, float* out,
                           int32_t* err, yr_stream stream) {
  if (!U || !V || !uid || !iid || !out || d <= 0 || B < 0) return YR_ERR_BAD_ARG;
  if (B == 0) return YR_OK;
  const int threads = 128;
  const int64_t blocks = (B + threads - 1) / threads;
  mf_score_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(U, V, nU, nI, d, uid, iid, B,
                                                                           out, err);
  YR_CHECK_LAUNCH();
  return YR_OK;
}

extern "C" int yr_mf_score_bwd(const float* U, const float* V, int64_t nU, int64_t nI, int d,
                               const int64_t* uid, const int64_t* iid, int64_t B, const float* gout,
                               float* gU, float* gV, yr_stream stream) {
  if (!U || !V || !uid || !iid || !gout || !gU || !gV || d <= 0 || B < 0) return YR_ERR_BAD_ARG;
  if (B == 0) return YR_OK;
  const int threads = 256;
  const int64_t blocks = (B * 32 + threads - 1) / threads;
  cudaStream_t s = (cudaStream_t)stream;
  switch (dim_vpl(d)) {
    case 1: mf_score_bwd_kernel<1><<<(unsigned)blocks, threads, 0, s>>>(U, V, nU, nI, uid, iid, B, gout, gU, gV); break;
    case 2: mf_score_bwd_kernel<2><<<(unsigned)blocks, threads, 0, s>>>(U, V, nU, nI, uid, iid, B, gout, gU, gV); break;
    case 4: mf_score_bwd_kernel<4><<<(unsigned)blocks, threads, 0, s>>>(U, V, nU, nI, uid, iid, B, gout, gU, gV); break;
    case 8: mf_score_bwd_kernel<8><<<(unsigned)blocks, threads, 0, s>>>(U, V, nU, nI, uid, iid, B, gout, gU, gV); break;
    default:
      mf_score_bwd_generic_kernel<<<(unsigned)blocks, threads, 0, s>>>(U, V, nU, nI, d, uid, iid, B, gout, gU, gV);
  }
  YR_CHECK_LAUNCH();
  return YR_OK;
}

extern "C" int yr_bpr_loss_fwd(const float* pos, const float* neg, int64_t B, float* loss,
                               yr_stream stream) {
  if (!pos || !neg || !loss || B <= 0) return YR_ERR_BAD_ARG;
  bpr_loss_fwd_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(pos, neg, B, loss);
  YR_CHECK_LAUNCH();
  return YR_OK;
}

extern "C" int yr_bpr_loss_bwd(const float* pos, const float* neg, int64_t B, const float* gloss,
                               float* gpos, float* gneg, yr_stream stream) {
  if (!pos || !neg || !gpos || !gneg || B <= 0) return YR_ERR_BAD_ARG;
  const int threads = 256;
  bpr_loss_bwd_kernel<<<(unsigned)((B + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
      pos, neg, B, gloss, gpos, gneg);
  YR_CHECK_LAUNCH();
  return YR_OK;
}

template <int VPL>
static int launch_train(const yr_mf_state* st, const yr_opt* opt, const int64_t* uid,
                        const int64_t* pos, const int64_t* neg, int64_t n_triples, int32_t B,
                        double* loss_sum, float* step_loss, cudaStream_t stream) {
  int dev = 0, sms = 0, occ = 0;
  YR_CUDA(cudaGetDevice(&dev));
  YR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  constexpr int kTrainThreads = TrainCfg<VPL>::kThreads, kTrainWarps = TrainCfg<VPL>::kWarps;
  YR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bpr_mf_train_kernel<VPL>, kTrainThreads, 0));
  if (occ < 1) return YR_ERR_COOP;
  const bool dense = (opt->kind != YR_OPT_SGD) || (opt->weight_decay != 0.0);
  // sparse steps only need ~B warps in flight (fewer CTAs = cheaper grid barrier);
  // dense-semantics steps stream all rows and want every warp slot.
  int per_sm = dense ? occ : (int)((((int64_t)B + kTrainWarps - 1) / kTrainWarps + sms - 1) / sms);
  per_sm = env_int(dense ? "YR_MF_DENSE_CTAS_PER_SM" : "YR_MF_SPARSE_CTAS_PER_SM", per_sm);
  if (per_sm < 1) per_sm = 1;
  if (per_sm > occ) per_sm = occ;
  while (sms * per_sm > kMaxTrainCtas && per_sm > 1) --per_sm;
  if (sms * per_sm > kMaxTrainCtas) return YR_ERR_COOP;
  dim3 grid((unsigned)(sms * per_sm)), block(kTrainThreads);
  if (!st->ws || st->ws_bytes < mf_ws_layout(n_triples, B, st->d, nullptr, nullptr)) return YR_ERR_WORKSPACE;
  MfWs w;
  mf_ws_layout(n_triples, B, st->d, st->ws, &w);
  const int nwarps = (int)grid.x * kTrainWarps;
  const bool reg_path = VPL <= 8 && !dense && B <= nwarps && !st->deterministic;
  if (!reg_path) {
    // group every batch's ids by table row (one CTA per step and table), ahead of the persistent kernel
    const int64_t n_steps = (n_triples + B - 1) / B;
    const int64_t Pmax = pow2_ceil64(2 * (int64_t)B);
    const int64_t smem_cap = Pmax <= kSortSmemCap ? Pmax : 0;
    const size_t smem = (size_t)(smem_cap ? smem_cap : 1) * 8;
    static AttrOnce attr;
    { int rc_ = attr.set(mf_sort_batches_kernel, kSortSmemCap * 8); if (rc_) return rc_; }
    mf_sort_batches_kernel<<<(unsigned)(2 * n_steps), 1024, smem, stream>>>(uid, pos, neg, n_triples, B, st->nU, st->nI, w,
                                                                              smem_cap);
    YR_CHECK_LAUNCH();
  }
  yr_mf_state st_v = *st;
  yr_opt opt_v = *opt;
  void* args[] = {&st_v, &opt_v, &uid, &pos, &neg, &n_triples, &B, &loss_sum, &step_loss, &w};
  YR_CUDA(cudaLaunchCooperativeKernel((const void*)bpr_mf_train_kernel<VPL>, grid, block, args, 0, stream));
  return YR_OK;
}

extern "C" size_t yr_bpr_mf_train_ws_bytes(int64_t n_triples, int32_t B, int d) {
  if (n_triples < 0 || B <= 0 || d <= 0) return 0;
  return mf_ws_layout(n_triples, B, d, nullptr, nullptr);
}

extern "C" int yr_bpr_mf_train(const yr_mf_state* st, const yr_opt* opt, const int64_t* uid,
                               const int64_t* pos, const int64_t* neg, int64_t n_triples, int32_t B,
                               double* loss_sum, float* step_loss, yr_stream stream) {
  if (!st || !opt || !uid || !pos || !neg || B <= 0 || n_triples < 0) return YR_ERR_BAD_ARG;
  if (!st->U || !st->V || !st->flagU || !st->flagV || !st->counters || !st->err) return YR_ERR_BAD_ARG;
  if (st->nU >= (int64_t)kBadKey || st->nI >= (int64_t)kBadKey || B >= (1 << 30)) return YR_ERR_BAD_DIM;
  if (opt->kind < YR_OPT_SGD || opt->kind > YR_OPT_ADAMW) return YR_ERR_BAD_OPT;
  if (opt->kind != YR_OPT_SGD && (!st->mU || !st->vU || !st->mV || !st->vV)) return YR_ERR_BAD_ARG;
  if (n_triples == 0) return YR_OK;
  cudaStream_t s = (cudaStream_t)stream;
  switch (dim_vpl(st->d)) {
    case 1: return launch_train<1>(st, opt, uid, pos, neg, n_triples, B, loss_sum, step_loss, s);
    case 2: return launch_train<2>(st, opt, uid, pos, neg, n_triples, B, loss_sum, step_loss, s);
    case 4: return launch_train<4>(st, opt, uid, pos, neg, n_triples, B, loss_sum, step_loss, s);
    case 8: return launch_train<8>(st, opt, uid, pos, neg, n_triples, B, loss_sum, step_loss, s);
    case 16: return launch_train<16>(st, opt, uid, pos, neg, n_triples, B, loss_sum, step_loss, s);
    case 32: return launch_train<32>(st, opt, uid, pos, neg, n_triples, B, loss_sum, step_loss, s);
    default: return YR_ERR_BAD_DIM;
  }
}

extern "C" int yr_bpr_mf_validate(const float* U, const float* V, int64_t nU, int64_t nI, int d,
                                  const int64_t* uid, const int64_t* pos, const int64_t* neg,
                                  int64_t n_triples, int32_t B, double* loss_sum, float* step_loss,
                                  int32_t* err, yr_stream stream) {
  if (!U || !V || !uid || !pos || !neg || B <= 0 || n_triples < 0) return YR_ERR_BAD_ARG;
  if (n_triples == 0) return YR_OK;
  const int64_t n_steps = (n_triples + B - 1) / B;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned g = (unsigned)n_steps;
  switch (dim_vpl(d)) {
    case 1: bpr_mf_validate_kernel<1><<<g, 256, 0, s>>>(U, V, nU, nI, uid, pos, neg, n_triples, B, loss_sum, step_loss, err); break;
    case 2: bpr_mf_validate_kernel<2><<<g, 256, 0, s>>>(U, V, nU, nI, uid, pos, neg, n_triples, B, loss_sum, step_loss, err); break;
    case 4: bpr_mf_validate_kernel<4><<<g, 256, 0, s>>>(U, V, nU, nI, uid, pos, neg, n_triples, B, loss_sum, step_loss, err); break;
    case 8: bpr_mf_validate_kernel<8><<<g, 256, 0, s>>>(U, V, nU, nI, uid, pos, neg, n_triples, B, loss_sum, step_loss, err); break;
    case 16: bpr_mf_validate_kernel<16><<<g, 256, 0, s>>>(U, V, nU, nI, uid, pos, neg, n_triples, B, loss_sum, step_loss, err); break;
    case 32: bpr_mf_validate_kernel<32><<<g, 256, 0, s>>>(U, V, nU, nI, uid, pos, neg, n_triples, B, loss_sum, step_loss, err); break;
    default: return YR_ERR_BAD_DIM;
  }
  YR_CHECK_LAUNCH();
  return YR_OK;
}
