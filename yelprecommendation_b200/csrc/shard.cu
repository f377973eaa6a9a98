// Row-sharded BPR-MF (BASELINE config 5): device-side pieces of one step; the NCCL collectives between them are
// issued by the caller (yelprecommendation_b200/trainers/sharded_mf_trainer.py) on the same stream.
//   yr_shard_gather_rows — owner fills the batch's rows it holds (zeros elsewhere) -> SUM all-reduce = exact gather
//   yr_bpr_rows_grad     — forward + BPR loss + per-triple gradient rows for this rank's slice of the batch
//   yr_shard_accumulate  — owner sums the gradient rows of its ids into scratch (duplicates summed, rows listed)
//   yr_shard_step        — owner steps its shard once: listed rows (plain SGD) or every row (dense-semantics Adam / L2)
// Same arithmetic as the single-GPU fused kernel (mf.cu): reference trainers/mf_trainer.py:104-114.
#include <math.h>
#include "common.cuh"

namespace yr {

template <int VPL>
__global__ void __launch_bounds__(256)
shard_gather_kernel(const float* __restrict__ T, int64_t row0, int64_t row1, int64_t n_rows_global,
                    const int64_t* __restrict__ ids, int64_t n, float* __restrict__ out, int64_t out_ld, int32_t* err) {
  constexpr int D = VPL * 32;
  const int lane = threadIdx.x & 31;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < n; j += nw) {
    const int64_t id = ids[j];
    Row<VPL> r;
#pragma unroll
    for (int q = 0; q < VPL; ++q) r.x[q] = 0.f;
    if (id < 0 || id >= n_rows_global) { if (lane == 0 && err) atomicExch(err, 1); }
    else if (id >= row0 && id < row1) r = ld_row<VPL>(T + (id - row0) * D, lane);
    st_row<VPL>(out + j * out_ld, lane, r);
  }
}

// packed gather for the all-to-all exchange: out[j] = (sel[j] ? T1 : T0)[row[j]] — an owner collects, in slot order, the
// rows of its user table (sel 0) and item table (sel 1) that a requester's slice of the batch needs
template <int VPL>
__global__ void __launch_bounds__(256)
shard_gather_local_kernel(const float* __restrict__ T0, const float* __restrict__ T1, const int32_t* __restrict__ sel,
                          const int32_t* __restrict__ row, int64_t n, float* __restrict__ out, int64_t out_ld) {
  constexpr int D = VPL * 32;
  const int lane = threadIdx.x & 31;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < n; j += nw) {
    const float* T = sel[j] ? T1 : T0;
    st_row<VPL>(out + j * out_ld, lane, ld_row<VPL>(T + (int64_t)row[j] * D, lane));
  }
}

template <int VPL>
__global__ void __launch_bounds__(256)
bpr_rows_grad_kernel(const float* __restrict__ R, int64_t B, int64_t b0, int64_t b1, float* __restrict__ G,
                     double* loss_acc) {
  constexpr int D = VPL * 32;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float inv_b = 1.f / (float)B;
  __shared__ double s_part[8];
  double wl = 0.0;
  for (int64_t b = b0 + (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); b < b1; b += nw) {
    const float* rb = R + b * 3 * D;
    const Row<VPL> ur = ld_row<VPL>(rb, lane), pr = ld_row<VPL>(rb + D, lane), nr = ld_row<VPL>(rb + 2 * D, lane);
    const float x = warp_sum(dot_partial<VPL>(ur, pr)) - warp_sum(dot_partial<VPL>(ur, nr));
    wl += (double)neg_logsigmoid(x);
    const float g = neg_logsigmoid_grad(x) * inv_b;
    Row<VPL> gu, gp, gn;
#pragma unroll
    for (int q = 0; q < VPL; ++q) {
      gu.x[q] = __fsub_rn(__fmul_rn(g, pr.x[q]), __fmul_rn(g, nr.x[q]));
      gp.x[q] = g * ur.x[q];
      gn.x[q] = -gp.x[q];
    }
    float* gb = G + b * 3 * D;
    st_row<VPL>(gb, lane, gu);
    st_row<VPL>(gb + D, lane, gp);
    st_row<VPL>(gb + 2 * D, lane, gn);
  }
  if (lane == 0) s_part[wib] = wl;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += s_part[i];
    if (t != 0.0) atomicAdd(loss_acc, t);
  }
}

// accumulate the gradient rows of owned ids into the shard's scratch; first toucher lists the row
template <int VPL>
__global__ void __launch_bounds__(256)
shard_scatter_kernel(yr_shard_state st, const int64_t* __restrict__ ids, int64_t n, const float* __restrict__ G,
                     int64_t g_ld, bool list_rows) {
  constexpr int D = VPL * 32;
  const int lane = threadIdx.x & 31;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < n; j += nw) {
    const int64_t id = ids[j];
    if (id < st.row0 || id >= st.row1) continue;
    const int64_t r = id - st.row0;
    const Row<VPL> g = ld_row<VPL>(G + j * g_ld, lane);
    red_row<VPL>(st.gscratch + r * D, lane, g);
    if (lane == 0) {
      if (list_rows) {
        if (atomicExch(st.flags + r, 1) == 0) st.rows_list[atomicAdd(st.counters, 1)] = (int32_t)r;
      } else {
        st.flags[r] = 1;
      }
    }
  }
}

template <int VPL>
__global__ void __launch_bounds__(256)
shard_update_kernel(yr_shard_state st, yr_opt opt, bool dense) {
  constexpr int D = VPL * 32;
  const int lane = threadIdx.x & 31;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  OptScalars os;
  opt_scalars_for_step(os, opt, opt.step);
  const int64_t n_items = dense ? (st.row1 - st.row0) : (int64_t)__ldcg(st.counters);
  for (int64_t i = w0; i < n_items; i += nw) {
    const int64_t r = dense ? i : (int64_t)st.rows_list[i];
    const bool touched = dense ? (__ldcg(st.flags + r) != 0) : true;
    float* prow = st.T + r * D;
    float* grow = st.gscratch + r * D;
    Row<VPL> pv = ld_row<VPL>(prow, lane), gv, mv, vv;
    if (touched) gv = ld_row<VPL>(grow, lane);
    else {
#pragma unroll
      for (int q = 0; q < VPL; ++q) gv.x[q] = 0.f;
    }
    if (opt.kind != YR_OPT_SGD) {
      mv = ld_row<VPL>(st.m + r * D, lane);
      vv = ld_row<VPL>(st.v + r * D, lane);
#pragma unroll
      for (int q = 0; q < VPL; ++q) opt_update(os, pv.x[q], gv.x[q], mv.x[q], vv.x[q]);
      st_row<VPL>(st.m + r * D, lane, mv);
      st_row<VPL>(st.v + r * D, lane, vv);
    } else {
#pragma unroll
      for (int q = 0; q < VPL; ++q) { float m = 0.f, v = 0.f; opt_update(os, pv.x[q], gv.x[q], m, v); }
    }
    st_row<VPL>(prow, lane, pv);
    if (touched) {
      Row<VPL> z;
#pragma unroll
      for (int q = 0; q < VPL; ++q) z.x[q] = 0.f;
      st_row<VPL>(grow, lane, z);
      if (lane == 0) st.flags[r] = 0;
    }
  }
}

__global__ void shard_reset_kernel(int32_t* counters) { counters[0] = 0; }

// ---- ordered (atomics-free) owner-side accumulate -----------------------------------------------------------------
// rows_sorted[j] = local row of the j-th gradient row in ROW order (a stable sort of the received list, so equal rows keep
// their arrival order: requester rank, then slot), src[j] = its position in G. One warp per segment head sums the
// segment left to right into the row's scratch (plain stores: every row has exactly one segment) and lists the row.
template <int VPL>
__global__ void __launch_bounds__(256)
shard_accumulate_sorted_kernel(yr_shard_state st, const int32_t* __restrict__ rows_sorted, const int32_t* __restrict__ src,
                               int64_t n, const float* __restrict__ G, int64_t g_ld, bool list_rows) {
  constexpr int D = VPL * 32;
  const int lane = threadIdx.x & 31;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < n; j += nw) {
    const int32_t r = rows_sorted[j];
    if (r < 0) continue;                                       // not a row of this shard (callers sort those to the front)
    if (j > 0 && rows_sorted[j - 1] == r) continue;            // not a segment head
    Row<VPL> acc = ld_row<VPL>(G + (int64_t)src[j] * g_ld, lane);
    for (int64_t q = j + 1; q < n && rows_sorted[q] == r; ++q) {
      const Row<VPL> g = ld_row<VPL>(G + (int64_t)src[q] * g_ld, lane);
#pragma unroll
      for (int v = 0; v < VPL; ++v) acc.x[v] = __fadd_rn(acc.x[v], g.x[v]);
    }
    st_row<VPL>(st.gscratch + (int64_t)r * D, lane, acc);
    if (lane == 0) {
      st.flags[r] = 1;
      if (list_rows) st.rows_list[atomicAdd(st.counters, 1)] = r;     // list order is irrelevant: one update per row
    }
  }
}

// ---- sparse-traffic ("catch-up") Adam / AdamW -------------------------------------------------------------------------
// torch's optimizer is dense: a row without a gradient still moves while its moments are non-zero. A row that is not
// touched evolves by a fixed recurrence (g = 0), so only the rows of the batch are visited: each replays the steps it
// missed with the SAME opt_update() calls and the same per-step scalars the dense sweep would have used, then takes the
// current step with its gradient — bit-identical to the sweep (tests/test_gpu_shard.py). last[r] = last step applied to
// row r; scal[2 * t], scal[2 * t + 1] = (float)(lr / (1 - beta1^t)), (float)sqrt(1 - beta2^t), filled by yr_adam_scalars
// (the same device function the dense sweep evaluates per step).
template <int VPL>
__device__ __forceinline__ void replay_row(OptScalars os, const float* __restrict__ scal, int from, int to, bool skip_if_idle,
                                           Row<VPL>& p, Row<VPL>& m, Row<VPL>& v) {
  if (skip_if_idle) {           // wd == 0: with m == v == 0 an update is exactly p + (-0) = p
    bool idle = true;
#pragma unroll
    for (int q = 0; q < VPL; ++q) idle = idle && (m.x[q] == 0.f) && (v.x[q] == 0.f);
    if (__all_sync(kFull, idle)) return;
  }
  for (int t = from; t <= to; ++t) {
    os.step_size = __ldg(scal + 2 * t);
    os.bc2_sqrt = __ldg(scal + 2 * t + 1);
#pragma unroll
    for (int q = 0; q < VPL; ++q) opt_update(os, p.x[q], 0.f, m.x[q], v.x[q]);
  }
}

// Rows that are about to be READ (gathered for a forward pass) must first be brought up to the previous step: one warp per
// segment head of the sorted row list replays last + 1 .. step - 1.
template <int VPL>
__global__ void __launch_bounds__(256)
shard_catch_up_kernel(yr_shard_state st, yr_opt opt, const float* __restrict__ scal, int32_t* __restrict__ last,
                      const int32_t* __restrict__ rows_sorted, int64_t n) {
  constexpr int D = VPL * 32;
  const int lane = threadIdx.x & 31;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  OptScalars os;
  opt_scalars_for_step(os, opt, opt.step);
  const bool skip_idle = (opt.weight_decay == 0.0);
  const int t_prev = opt.step - 1;
  for (int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < n; j += nw) {
    const int32_t r = rows_sorted[j];
    if (j > 0 && rows_sorted[j - 1] == r) continue;
    const int l = last[r];
    if (l >= t_prev) continue;
    float* prow = st.T + (int64_t)r * D;
    Row<VPL> pv = ld_row<VPL>(prow, lane), mv = ld_row<VPL>(st.m + (int64_t)r * D, lane), vv = ld_row<VPL>(st.v + (int64_t)r * D, lane);
    replay_row<VPL>(os, scal, l + 1, t_prev, skip_idle, pv, mv, vv);
    st_row<VPL>(prow, lane, pv);
    st_row<VPL>(st.m + (int64_t)r * D, lane, mv);
    st_row<VPL>(st.v + (int64_t)r * D, lane, vv);
    if (lane == 0) last[r] = t_prev;
  }
}

template <int VPL>
__global__ void __launch_bounds__(256)
shard_sparse_adam_kernel(yr_shard_state st, yr_opt opt, const float* __restrict__ scal, int32_t* __restrict__ last,
                         bool flush) {
  constexpr int D = VPL * 32;
  const int lane = threadIdx.x & 31;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  OptScalars os;
  opt_scalars_for_step(os, opt, opt.step);
  const bool skip_idle = (opt.weight_decay == 0.0);
  const int t_now = opt.step;                                   // flush: bring rows up to and including t_now
  const int64_t n_items = flush ? (st.row1 - st.row0) : (int64_t)__ldcg(st.counters);
  for (int64_t i = w0; i < n_items; i += nw) {
    const int64_t r = flush ? i : (int64_t)st.rows_list[i];
    const int l = last[r];
    if (flush && l >= t_now) continue;
    float* prow = st.T + r * D;
    Row<VPL> pv = ld_row<VPL>(prow, lane), mv = ld_row<VPL>(st.m + r * D, lane), vv = ld_row<VPL>(st.v + r * D, lane);
    replay_row<VPL>(os, scal, l + 1, flush ? t_now : t_now - 1, skip_idle, pv, mv, vv);
    if (!flush) {
      float* grow = st.gscratch + r * D;
      const Row<VPL> gv = ld_row<VPL>(grow, lane);
      os.step_size = __ldg(scal + 2 * t_now);
      os.bc2_sqrt = __ldg(scal + 2 * t_now + 1);
#pragma unroll
      for (int q = 0; q < VPL; ++q) opt_update(os, pv.x[q], gv.x[q], mv.x[q], vv.x[q]);
      Row<VPL> z;
#pragma unroll
      for (int q = 0; q < VPL; ++q) z.x[q] = 0.f;
      st_row<VPL>(grow, lane, z);
      if (lane == 0) st.flags[r] = 0;
    }
    st_row<VPL>(prow, lane, pv);
    st_row<VPL>(st.m + r * D, lane, mv);
    st_row<VPL>(st.v + r * D, lane, vv);
    if (lane == 0) last[r] = t_now;
  }
}

static unsigned warp_grid(int64_t n_warps, int per_block_warps = 8) {
  int64_t blocks = (n_warps + per_block_warps - 1) / per_block_warps;
  const int64_t cap = (int64_t)yr_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace yr

using namespace yr;

extern "C" int yr_shard_gather_rows(const float* T_local, int64_t row0, int64_t row1, int64_t n_rows_global, int d,
                                    const int64_t* ids, int64_t n, float* out, int64_t out_ld, int32_t* err,
                                    yr_stream stream) {
  if (!T_local || !ids || !out || n < 0 || row1 < row0 || out_ld < d) return YR_ERR_BAD_ARG;
  if (n == 0) return YR_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned g = warp_grid(n);
  switch (dim_vpl(d)) {
    case 1: shard_gather_kernel<1><<<g, 256, 0, s>>>(T_local, row0, row1, n_rows_global, ids, n, out, out_ld, err); break;
    case 2: shard_gather_kernel<2><<<g, 256, 0, s>>>(T_local, row0, row1, n_rows_global, ids, n, out, out_ld, err); break;
    case 4: shard_gather_kernel<4><<<g, 256, 0, s>>>(T_local, row0, row1, n_rows_global, ids, n, out, out_ld, err); break;
    case 8: shard_gather_kernel<8><<<g, 256, 0, s>>>(T_local, row0, row1, n_rows_global, ids, n, out, out_ld, err); break;
    default: return YR_ERR_BAD_DIM;
  }
  YR_CHECK_LAUNCH();
  return YR_OK;
}

extern "C" int yr_bpr_rows_grad(const float* R, int d, int64_t B, int64_t b0, int64_t b1, float* G, double* loss_acc,
                                yr_stream stream) {
  if (!R || !G || !loss_acc || B <= 0 || b0 < 0 || b1 > B || b1 < b0) return YR_ERR_BAD_ARG;
  if (b1 == b0) return YR_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned g = warp_grid(b1 - b0);
  switch (dim_vpl(d)) {
    case 1: bpr_rows_grad_kernel<1><<<g, 256, 0, s>>>(R, B, b0, b1, G, loss_acc); break;
    case 2: bpr_rows_grad_kernel<2><<<g, 256, 0, s>>>(R, B, b0, b1, G, loss_acc); break;
    case 4: bpr_rows_grad_kernel<4><<<g, 256, 0, s>>>(R, B, b0, b1, G, loss_acc); break;
    case 8: bpr_rows_grad_kernel<8><<<g, 256, 0, s>>>(R, B, b0, b1, G, loss_acc); break;
    case 16: bpr_rows_grad_kernel<16><<<g, 256, 0, s>>>(R, B, b0, b1, G, loss_acc); break;      // NGCF: d * (layers + 1) = 512
    case 32: bpr_rows_grad_kernel<32><<<g, 256, 0, s>>>(R, B, b0, b1, G, loss_acc); break;
    default: return YR_ERR_BAD_DIM;
  }
  YR_CHECK_LAUNCH();
  return YR_OK;
}

extern "C" int yr_shard_accumulate(const yr_shard_state* st, const yr_opt* opt, const int64_t* ids, int64_t n,
                                   const float* G, int64_t g_ld, yr_stream stream) {
  if (!st || !opt || !ids || !G || n < 0 || !st->T || !st->gscratch || !st->flags || !st->rows_list || !st->counters)
    return YR_ERR_BAD_ARG;
  if (g_ld < st->d) return YR_ERR_BAD_ARG;
  if (st->row1 <= st->row0 || n == 0) return YR_OK;
  const bool dense = (opt->kind != YR_OPT_SGD) || (opt->weight_decay != 0.0);
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned g = warp_grid(n);
  switch (dim_vpl(st->d)) {
    case 1: shard_scatter_kernel<1><<<g, 256, 0, s>>>(*st, ids, n, G, g_ld, !dense); break;
    case 2: shard_scatter_kernel<2><<<g, 256, 0, s>>>(*st, ids, n, G, g_ld, !dense); break;
    case 4: shard_scatter_kernel<4><<<g, 256, 0, s>>>(*st, ids, n, G, g_ld, !dense); break;
    case 8: shard_scatter_kernel<8><<<g, 256, 0, s>>>(*st, ids, n, G, g_ld, !dense); break;
    default: return YR_ERR_BAD_DIM;
  }
  YR_CHECK_LAUNCH();
  return YR_OK;
}

extern "C" int yr_shard_step(const yr_shard_state* st, const yr_opt* opt, int64_t max_rows, yr_stream stream) {
  if (!st || !opt || !st->T || !st->gscratch || !st->flags || !st->rows_list || !st->counters) return YR_ERR_BAD_ARG;
  if (opt->kind < YR_OPT_SGD || opt->kind > YR_OPT_ADAMW) return YR_ERR_BAD_OPT;
  if (opt->kind != YR_OPT_SGD && (!st->m || !st->v)) return YR_ERR_BAD_ARG;
  if (st->row1 <= st->row0) return YR_OK;
  const bool dense = (opt->kind != YR_OPT_SGD) || (opt->weight_decay != 0.0);
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned g = warp_grid(dense ? (st->row1 - st->row0) : (max_rows > 0 ? max_rows : 1));
  switch (dim_vpl(st->d)) {
    case 1: shard_update_kernel<1><<<g, 256, 0, s>>>(*st, *opt, dense); break;
    case 2: shard_update_kernel<2><<<g, 256, 0, s>>>(*st, *opt, dense); break;
    case 4: shard_update_kernel<4><<<g, 256, 0, s>>>(*st, *opt, dense); break;
    case 8: shard_update_kernel<8><<<g, 256, 0, s>>>(*st, *opt, dense); break;
    default: return YR_ERR_BAD_DIM;
  }
  YR_CHECK_LAUNCH();
  if (!dense) { shard_reset_kernel<<<1, 1, 0, s>>>(st->counters); YR_CHECK_LAUNCH(); }
  return YR_OK;
}

extern "C" int yr_shard_accumulate_sorted(const yr_shard_state* st, const yr_opt* opt, const int32_t* rows_sorted,
                                          const int32_t* src, int64_t n, const float* G, int64_t g_ld, int list_rows,
                                          yr_stream stream) {
  if (!st || !opt || !rows_sorted || !src || !G || n < 0 || !st->gscratch || !st->flags || !st->rows_list || !st->counters)
    return YR_ERR_BAD_ARG;
  if (g_ld < st->d) return YR_ERR_BAD_ARG;
  if (n == 0) return YR_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned g = warp_grid(n);
  switch (dim_vpl(st->d)) {
    case 1: shard_accumulate_sorted_kernel<1><<<g, 256, 0, s>>>(*st, rows_sorted, src, n, G, g_ld, list_rows != 0); break;
    case 2: shard_accumulate_sorted_kernel<2><<<g, 256, 0, s>>>(*st, rows_sorted, src, n, G, g_ld, list_rows != 0); break;
    case 4: shard_accumulate_sorted_kernel<4><<<g, 256, 0, s>>>(*st, rows_sorted, src, n, G, g_ld, list_rows != 0); break;
    case 8: shard_accumulate_sorted_kernel<8><<<g, 256, 0, s>>>(*st, rows_sorted, src, n, G, g_ld, list_rows != 0); break;
    default: return YR_ERR_BAD_DIM;
  }
  YR_CHECK_LAUNCH();
  return YR_OK;
}

// per-step Adam scalars for steps [0, n_steps) in the layout the sparse kernels read (entry 0 unused), computed ON THE
// DEVICE by the same function the dense sweep calls per step, so the two paths use bit-identical scalars
namespace yr {
__global__ void adam_scalars_kernel(yr_opt opt, int n_steps, float* __restrict__ scal) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_steps; t += gridDim.x * blockDim.x) {
    OptScalars os;
    opt_scalars_for_step(os, opt, t < 1 ? 1 : t);
    scal[2 * t] = os.step_size;
    scal[2 * t + 1] = os.bc2_sqrt;
  }
}
}  // namespace yr

extern "C" int yr_adam_scalars(const yr_opt* opt, int n_steps, float* scal, yr_stream stream) {
  if (!opt || !scal || n_steps < 1) return YR_ERR_BAD_ARG;
  yr::adam_scalars_kernel<<<(n_steps + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*opt, n_steps, scal);
  YR_CHECK_LAUNCH();
  return YR_OK;
}

extern "C" int yr_shard_step_sparse_adam(const yr_shard_state* st, const yr_opt* opt, const float* scal, int32_t n_scal_steps,
                                         int32_t* last, int64_t max_rows, int flush, yr_stream stream) {
  if (!st || !opt || !scal || !last || !st->T || !st->m || !st->v || !st->gscratch || !st->flags || !st->rows_list || !st->counters)
    return YR_ERR_BAD_ARG;
  if (opt->kind != YR_OPT_ADAM && opt->kind != YR_OPT_ADAMW) return YR_ERR_BAD_OPT;
  if (opt->step < (flush ? 0 : 1) || opt->step >= n_scal_steps) return YR_ERR_BAD_ARG;
  if (st->row1 <= st->row0) return YR_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned g = warp_grid(flush ? (st->row1 - st->row0) : (max_rows > 0 ? max_rows : 1));
  switch (dim_vpl(st->d)) {
    case 1: shard_sparse_adam_kernel<1><<<g, 256, 0, s>>>(*st, *opt, scal, last, flush != 0); break;
    case 2: shard_sparse_adam_kernel<2><<<g, 256, 0, s>>>(*st, *opt, scal, last, flush != 0); break;
    case 4: shard_sparse_adam_kernel<4><<<g, 256, 0, s>>>(*st, *opt, scal, last, flush != 0); break;
    case 8: shard_sparse_adam_kernel<8><<<g, 256, 0, s>>>(*st, *opt, scal, last, flush != 0); break;
    default: return YR_ERR_BAD_DIM;
  }
  YR_CHECK_LAUNCH();
  if (!flush) { shard_reset_kernel<<<1, 1, 0, s>>>(st->counters); YR_CHECK_LAUNCH(); }
  return YR_OK;
}

extern "C" int yr_shard_gather_local(const float* T0, const float* T1, int d, const int32_t* sel, const int32_t* row,
                                     int64_t n, float* out, int64_t out_ld, yr_stream stream) {
  if (!T0 || !T1 || !sel || !row || !out || n < 0 || out_ld < d) return YR_ERR_BAD_ARG;
  if (n == 0) return YR_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned g = warp_grid(n);
  switch (dim_vpl(d)) {
    case 1: shard_gather_local_kernel<1><<<g, 256, 0, s>>>(T0, T1, sel, row, n, out, out_ld); break;
    case 2: shard_gather_local_kernel<2><<<g, 256, 0, s>>>(T0, T1, sel, row, n, out, out_ld); break;
    case 4: shard_gather_local_kernel<4><<<g, 256, 0, s>>>(T0, T1, sel, row, n, out, out_ld); break;
    case 8: shard_gather_local_kernel<8><<<g, 256, 0, s>>>(T0, T1, sel, row, n, out, out_ld); break;
    default: return YR_ERR_BAD_DIM;
  }
  YR_CHECK_LAUNCH();
  return YR_OK;
}

extern "C" int yr_shard_catch_up(const yr_shard_state* st, const yr_opt* opt, const float* scal, int32_t n_scal_steps,
                                 int32_t* last, const int32_t* rows_sorted, int64_t n, yr_stream stream) {
  if (!st || !opt || !scal || !last || !rows_sorted || !st->T || !st->m || !st->v || n < 0) return YR_ERR_BAD_ARG;
  if (opt->kind != YR_OPT_ADAM && opt->kind != YR_OPT_ADAMW) return YR_ERR_BAD_OPT;
  if (opt->step < 1 || opt->step >= n_scal_steps) return YR_ERR_BAD_ARG;
  if (n == 0 || opt->step == 1) return YR_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned g = warp_grid(n);
  switch (dim_vpl(st->d)) {
    case 1: shard_catch_up_kernel<1><<<g, 256, 0, s>>>(*st, *opt, scal, last, rows_sorted, n); break;
    case 2: shard_catch_up_kernel<2><<<g, 256, 0, s>>>(*st, *opt, scal, last, rows_sorted, n); break;
    case 4: shard_catch_up_kernel<4><<<g, 256, 0, s>>>(*st, *opt, scal, last, rows_sorted, n); break;
    case 8: shard_catch_up_kernel<8><<<g, 256, 0, s>>>(*st, *opt, scal, last, rows_sorted, n); break;
    default: return YR_ERR_BAD_DIM;
  }
  YR_CHECK_LAUNCH();
  return YR_OK;
}
