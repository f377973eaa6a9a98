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
// counters layout (int32): [0..1] = #unique user rows / item rows of parity-0 steps, [2..3] = parity 1,
//                          [4] = parity of the next step, [5..7] spare.
// loss_acc (double[2]) sits right behind the counters (counters is 8 x int32 = 32 bytes; loss_acc at
// byte offset 32) — the state struct hands us one 64-byte block for both.
// ---------------------------------------------------------------------------------------------
// 512-thread CTAs (half the CTAs at the grid barrier) up to d = 256; 256 threads beyond, where a thread holds up to
// four rows of 16 / 32 floats (255 registers available).
template <int VPL> struct TrainCfg {
  static constexpr int kThreads = VPL > 8 ? 256 : 512;
  static constexpr int kWarps = kThreads / 32;
};

template <int VPL>
__global__ void __launch_bounds__(TrainCfg<VPL>::kThreads)
bpr_mf_train_kernel(yr_mf_state st, yr_opt opt, const int64_t* __restrict__ uid,
                    const int64_t* __restrict__ pos, const int64_t* __restrict__ neg,
                    int64_t n_triples, int B, double* loss_sum, float* step_loss) {
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
  double* loss_acc = reinterpret_cast<double*>(st.counters + 8);
  const int parity0 = __ldcg(counters + 4) & 1;
  int32_t* rowsU = st.rows;
  int32_t* rowsV = st.rows + B;
  __shared__ double s_part[kTrainWarps];

  // ---- plain SGD with at most one triple per warp: the three gradient rows stay in REGISTERS across the barrier and
  // are applied as vector REDs of -lr * g straight into the tables — no gradient scratch, no touched-row list, no
  // second pass over the rows. Two grid barriers per step remain (every read of step s precedes every update of
  // step s, every update precedes the reads of step s + 1): the reference's sequential step semantics.
  if (VPL <= 8 && !dense && B <= nwarps) {
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
        const float x = warp_sum(dot_partial<VPL>(ur, pr)) - warp_sum(dot_partial<VPL>(ur, nr));
        const float g = neg_logsigmoid_grad(x) * inv_nb;
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
          const float a = __fsub_rn(__fmul_rn(g, pr.x[j]), __fmul_rn(g, nr.x[j]));   // Q2: two rounded products
          const float c = g * ur.x[j];
          gu.x[j] = neg_lr * a; gp.x[j] = neg_lr * c; gn.x[j] = -gp.x[j];
        }
        wl = (double)neg_logsigmoid(x);
      }
      const int64_t uu = u, pp = p, nn = n;
      // ids of the next step (read-only input): requested before the barrier, consumed after the second one
      const int64_t nbase = base + B;
      if (nbase + gwarp < n_triples && gwarp < B) { u = uid[nbase + gwarp]; p = pos[nbase + gwarp]; n = neg[nbase + gwarp]; }
      if (lane == 0) s_part[wib] = wl;
      __syncthreads();
      if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < kTrainWarps; ++i) t += s_part[i];
        if (t != 0.0) atomicAdd(loss_acc + par, t);
      }
      grid_sync(grid);
      if (act) {
        red_row<VPL>(st.U + uu * D, lane, gu);
        red_row<VPL>(st.V + pp * D, lane, gp);
        red_row<VPL>(st.V + nn * D, lane, gn);
      }
      if (blockIdx.x == 0 && threadIdx.x == 0) {
        const float mean = (float)(__ldcg(loss_acc + par) / (double)nb);
        if (step_loss) step_loss[s] = mean;
        if (loss_sum) *loss_sum += (double)mean;
        loss_acc[par ^ 1] = 0.0;
        if (s + 1 == n_steps) counters[4] = par ^ 1;
      }
      if (s + 1 < n_steps) grid_sync(grid);
    }
    return;
  }

  for (int64_t s = 0; s < n_steps; ++s) {
    const int par = (parity0 + (int)(s & 1)) & 1;
    int32_t* cnt = counters + 2 * par;
    const int64_t base = s * (int64_t)B;
    const int nb = (int)((n_triples - base < B) ? (n_triples - base) : B);
    const float inv_nb = 1.f / (float)nb;

    // ---- phase 1: gather, dots, loss, gradient rows -> sparse accumulate --------------------
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
      const float x = dp - dn;
      const float g = neg_logsigmoid_grad(x) * inv_nb;
      Row<VPL> gu, gp, gn;
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        // user row is gathered twice by the reference (pos call + neg call, Q2): g*p + (-g)*n
        gu.x[j] = __fsub_rn(__fmul_rn(g, pr.x[j]), __fmul_rn(g, nr.x[j]));   // two rounded products, like autograd
        gp.x[j] = g * ur.x[j];
        gn.x[j] = -gp.x[j];
      }
      red_row<VPL>(st.gU + u * D, lane, gu);
      red_row<VPL>(st.gV + p * D, lane, gp);
      red_row<VPL>(st.gV + n * D, lane, gn);
      if (!dense) {
        if (lane == 0 && atomicExch(st.flagU + u, 1) == 0) rowsU[atomicAdd(cnt + 0, 1)] = (int32_t)u;
        if (lane == 1 && atomicExch(st.flagV + p, 1) == 0) rowsV[atomicAdd(cnt + 1, 1)] = (int32_t)p;
        __syncwarp();   // p == n cannot happen for a sampled negative, but stay correct if it does
        if (lane == 2 && atomicExch(st.flagV + n, 1) == 0) rowsV[atomicAdd(cnt + 1, 1)] = (int32_t)n;
      } else if (lane < 3) {
        if (lane == 0) st.flagU[u] = 1;
        if (lane == 1) st.flagV[p] = 1;
        if (lane == 2) st.flagV[n] = 1;
      }
      wl += (double)neg_logsigmoid(x);
    }
    if (lane == 0) s_part[wib] = wl;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
#pragma unroll
      for (int i = 0; i < kTrainWarps; ++i) t += s_part[i];
      if (t != 0.0) atomicAdd(loss_acc + par, t);
    }
    grid_sync(grid);

    // ---- phase 2: one optimizer update per row ----------------------------------------------
    OptScalars os;
    opt_scalars_for_step(os, opt, opt.step + (int)s);
    if (!dense) {
      const int nu = __ldcg(cnt + 0), nv = __ldcg(cnt + 1);
      for (int i = gwarp; i < nu + nv; i += nwarps) {
        const bool is_u = i < nu;
        const int64_t r = is_u ? __ldcg(rowsU + i) : __ldcg(rowsV + (i - nu));
        float* prow = (is_u ? st.U : st.V) + r * D;
        float* grow = (is_u ? st.gU : st.gV) + r * D;
        Row<VPL> pv = ld_row<VPL>(prow, lane);
        Row<VPL> gv = ld_row<VPL>(grow, lane);
        Row<VPL> z;
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
          float m = 0.f, v = 0.f;
          opt_update(os, pv.x[j], gv.x[j], m, v);
          z.x[j] = 0.f;
        }
        st_row<VPL>(prow, lane, pv);
        st_row<VPL>(grow, lane, z);
        if (lane == 0) (is_u ? st.flagU : st.flagV)[r] = 0;
      }
    } else {
      const int64_t nrows = st.nU + st.nI;
      for (int64_t r0 = gwarp; r0 < nrows; r0 += nwarps) {
        const bool is_u = r0 < st.nU;
        const int64_t r = is_u ? r0 : r0 - st.nU;
        int32_t* flag = (is_u ? st.flagU : st.flagV) + r;
        const bool touched = __ldcg(flag) != 0;
        float* prow = (is_u ? st.U : st.V) + r * D;
        float* grow = (is_u ? st.gU : st.gV) + r * D;
        Row<VPL> pv = ld_row<VPL>(prow, lane), gv, mv, vv;
        if (touched) {
          gv = ld_row<VPL>(grow, lane);
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
        if (touched) {
          Row<VPL> z;
#pragma unroll
          for (int j = 0; j < VPL; ++j) z.x[j] = 0.f;
          st_row<VPL>(grow, lane, z);
          if (lane == 0) *flag = 0;
        }
      }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      const float mean = (float)(__ldcg(loss_acc + par) / (double)nb);   // batch mean, then .item()
      if (step_loss) step_loss[s] = mean;
      if (loss_sum) *loss_sum += (double)mean;                          // Q1: sum of batch means
      // re-arm the other parity's accumulators for the next step
      counters[2 * (par ^ 1) + 0] = 0;
      counters[2 * (par ^ 1) + 1] = 0;
      loss_acc[par ^ 1] = 0.0;
      if (s + 1 == n_steps) counters[4] = par ^ 1;
    }
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
  dim3 grid((unsigned)(sms * per_sm)), block(kTrainThreads);
  yr_mf_state st_v = *st;
  yr_opt opt_v = *opt;
  void* args[] = {&st_v, &opt_v, &uid, &pos, &neg, &n_triples, &B, &loss_sum, &step_loss};
  YR_CUDA(cudaLaunchCooperativeKernel((const void*)bpr_mf_train_kernel<VPL>, grid, block, args, 0, stream));
  return YR_OK;
}

extern "C" int yr_bpr_mf_train(const yr_mf_state* st, const yr_opt* opt, const int64_t* uid,
                               const int64_t* pos, const int64_t* neg, int64_t n_triples, int32_t B,
                               double* loss_sum, float* step_loss, yr_stream stream) {
  if (!st || !opt || !uid || !pos || !neg || B <= 0 || n_triples < 0) return YR_ERR_BAD_ARG;
  if (!st->U || !st->V || !st->gU || !st->gV || !st->flagU || !st->flagV || !st->rows ||
      !st->counters || !st->err)
    return YR_ERR_BAD_ARG;
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
