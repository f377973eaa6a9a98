/*
 * yr_oracle.c — TEST INFRASTRUCTURE ONLY. Plain-C CPU restatement of the reference's BPR-MF / NGCF /
 * full-catalog-eval arithmetic. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it; the product (yelprecommendation_b200/) never does.
 *
 * Each function cites the reference lines (twndus/YelpRecommendation) it follows. Where the reference
 * leaves the fp32 summation order to PyTorch/NumPy, this file fixes ONE canonical order (a single fmaf
 * chain in index order) — the same order the CUDA kernels use, so integer outputs (top-K ids) and the
 * canonical-order floats can be compared bit-for-bit, while comparisons against the real reference's
 * outputs (tests/golden/) use the 1e-5 relative tolerance of BASELINE.json.
 *
 * Parity pin: tests/test_oracle_golden.py checks this file against fixtures produced by importing the
 * unmodified reference (tests/golden/make_golden.py) and against the reference's own known-answer
 * metric tests (test/test_metric.py:9-47).
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (oracle/Makefile). fmaf() is explicit, never implied.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MASK_VALUE (-3.40282e+38f) /* trainers/mf_trainer.py:167 */

/* ---- MatrixFactorization.forward, models/mf.py:20-23 -------------------------------------------- */
static float dot_chain(const float* a, const float* b, int d) {
  float acc = 0.f;
  for (int k = 0; k < d; ++k) acc = fmaf(a[k], b[k], acc);
  return acc;
}

void orc_mf_score(const float* U, const float* V, int d, const int64_t* uid, const int64_t* iid, int64_t B,
                  float* out) {
  for (int64_t b = 0; b < B; ++b) out[b] = dot_chain(U + uid[b] * d, V + iid[b] * d, d);
}

/* ---- BPRLoss.forward, loss.py:25-27; logsigmoid as ATen computes it ------------------------------
 * exp / log1p are taken CORRECTLY ROUNDED to fp32 (double evaluation, one rounding): the one definition the CUDA
 * kernels (common.cuh: exp_cr / log1p_cr) and this file can share bit for bit. glibc's expf, CUDA's expf and ATen's
 * Sleef kernels each differ from it (and from each other) by at most 1 ulp. */
static float exp_cr(float x) { return (float)exp((double)x); }
static float log1p_cr(float x) { return (float)log1p((double)x); }
static float neg_logsigmoid(float x) { return log1p_cr(exp_cr(-fabsf(x))) - fminf(x, 0.f); }
static float neg_logsigmoid_grad(float x) {
  const float z = exp_cr(-fabsf(x));
  const float q = z / (1.f + z);
  return -((x < 0.f) ? 1.f - q : q);
}

/* The trainers' row dot product as the kernels form it (common.cuh: dot_partial + warp_sum): lane l of a warp owns
 * d/32 elements of the row (d <= 64: the contiguous elements l*d/32 ..; d >= 128: the float4 number j*32 + l of every
 * 128-float4 group j), runs one fma chain over them, and the 32 partials are combined by the xor butterfly
 * 16, 8, 4, 2, 1 (every lane ends with the same value). d must be a multiple of 32. */
static float dot_warp(const float* a, const float* b, int d) {
  float s[32], t[32];
  const int vpl = d / 32;
  for (int l = 0; l < 32; ++l) {
    float acc = 0.f;
    if (vpl <= 2) {
      for (int j = 0; j < vpl; ++j) acc = fmaf(a[l * vpl + j], b[l * vpl + j], acc);
    } else {
      for (int j = 0; j < vpl / 4; ++j)
        for (int c = 0; c < 4; ++c) {
          const int k = 4 * (j * 32 + l) + c;
          acc = fmaf(a[k], b[k], acc);
        }
    }
    s[l] = acc;
  }
  for (int o = 16; o > 0; o >>= 1) {
    for (int l = 0; l < 32; ++l) t[l] = s[l] + s[l ^ o];
    memcpy(s, t, sizeof(s));
  }
  return s[0];
}

float orc_bpr_loss(const float* pos, const float* neg, int64_t B) {
  double acc = 0.0;
  for (int64_t b = 0; b < B; ++b) acc += (double)neg_logsigmoid(pos[b] - neg[b]);
  return (float)(acc / (double)B);
}

/* ---- torch.optim single-tensor updates, trainers/base_trainer.py:34-40 --------------------------- */
typedef struct {
  int32_t kind; /* 0 sgd 1 adam 2 adamw */
  int32_t step; /* 1-based */
  double lr, weight_decay, beta1, beta2, eps;
} orc_opt;

void orc_dense_opt_step(float* p, const float* g_in, float* m, float* v, int64_t n, const orc_opt* o) {
  const float lr = (float)o->lr, wd = (float)o->weight_decay;
  if (o->kind == 0) {
    for (int64_t i = 0; i < n; ++i) {
      float g = g_in[i];
      if (wd != 0.f) g = fmaf(p[i], wd, g);
      p[i] = fmaf(g, -lr, p[i]);
    }
    return;
  }
  const float w = (float)(1.0 - o->beta1), b2 = (float)o->beta2, omb2 = (float)(1.0 - o->beta2);
  const float eps = (float)o->eps;
  const double bc1 = 1.0 - pow(o->beta1, (double)o->step), bc2 = 1.0 - pow(o->beta2, (double)o->step);
  const float step_size = (float)(o->lr / bc1), bc2s = (float)sqrt(bc2);
  const float decay = (float)(1.0 - o->lr * o->weight_decay);
  for (int64_t i = 0; i < n; ++i) {
    float g = g_in[i], pi = p[i];
    if (o->kind == 2) pi = pi * decay;
    else if (wd != 0.f) g = fmaf(pi, wd, g);
    m[i] = fmaf(w, g - m[i], m[i]);
    v[i] = v[i] * b2;
    v[i] = fmaf(omb2 * g, g, v[i]);
    const float denom = sqrtf(v[i]) / bc2s + eps;
    p[i] = pi + (-step_size * m[i]) / denom;
  }
}

/* ---- MFTrainer.train, one batch (trainers/mf_trainer.py:104-114): dense grads like autograd --------
 * Gradient of a table row = what torch's CPU autograd produces: every embedding CALL (pos_pred = model(u, pos),
 * neg_pred = model(u, neg), trainers/mf_trainer.py:106-107) yields a dense gradient in which the rows of the batch are
 * summed sequentially in batch order (ATen embedding_dense_backward_cpu), and AccumulateGrad adds the two calls.
 * The row products are single rounded multiplications (g*p, -(g*n), g*u, -(g*u)).
 * gU/gV: caller-provided zeroed dense scratch [nU*d],[nI*d] (first call's part); the second call's part uses a
 * private scratch. Returns the batch-mean loss. */
float orc_mf_train_step(float* U, float* V, int64_t nU, int64_t nI, int d, float* mU, float* vU, float* mV,
                        float* vV, float* gU, float* gV, const int64_t* uid, const int64_t* pos,
                        const int64_t* neg, int64_t B, const orc_opt* o) {
  double lacc = 0.0;
  const float inv_b = 1.f / (float)B;
  float* gU2 = (float*)calloc((size_t)(nU * d), sizeof(float));
  float* gV2 = (float*)calloc((size_t)(nI * d), sizeof(float));
  for (int64_t b = 0; b < B; ++b) {
    const float* u = U + uid[b] * d;
    const float* p = V + pos[b] * d;
    const float* n = V + neg[b] * d;
    const float x = ((d % 32) == 0 ? dot_warp(u, p, d) : dot_chain(u, p, d)) -
                    ((d % 32) == 0 ? dot_warp(u, n, d) : dot_chain(u, n, d));
    lacc += (double)neg_logsigmoid(x);
    const float g = neg_logsigmoid_grad(x) * inv_b;
    float* gu = gU + uid[b] * d;  float* gu2 = gU2 + uid[b] * d;
    float* gp = gV + pos[b] * d;  float* gn2 = gV2 + neg[b] * d;
    for (int k = 0; k < d; ++k) {
      gu[k] = gu[k] + g * p[k];      /* positive call */
      gu2[k] = gu2[k] - g * n[k];    /* negative call: rows -(g*n), user row gathered twice (Q2) */
      gp[k] = gp[k] + g * u[k];
      gn2[k] = gn2[k] - g * u[k];
    }
  }
  for (int64_t i = 0; i < nU * d; ++i) gU[i] = gU[i] + gU2[i];
  for (int64_t i = 0; i < nI * d; ++i) gV[i] = gV[i] + gV2[i];
  free(gU2); free(gV2);
  orc_dense_opt_step(U, gU, mU, vU, nU * d, o);
  orc_dense_opt_step(V, gV, mV, vV, nI * d, o);
  memset(gU, 0, sizeof(float) * (size_t)(nU * d));
  memset(gV, 0, sizeof(float) * (size_t)(nI * d));
  return (float)(lacc / (double)B);
}

/* ---- torch.sparse.mm(L, E), models/ngcf.py:64,67 ------------------------------------------------------
 * Canonical order (include/yelprec_b200.h, yr_csr): rows with <= 128 non-zeros are one fma chain in column
 * order (seeded with Y's old value when accumulating); longer rows are the left-to-right sum of the partial
 * chains of their 128-non-zero chunks (plus Y_old when accumulating). */
#define ORC_SPMM_CHUNK 128
void orc_spmm_csr(const int32_t* rowptr, const int32_t* col, const float* val, int64_t n, int d, const float* X,
                  float* Y, int accumulate) {
  float* part = (float*)malloc(sizeof(float) * (size_t)d);
  float* tot = (float*)malloc(sizeof(float) * (size_t)d);
  for (int64_t r = 0; r < n; ++r) {
    float* y = Y + r * d;
    const int32_t s = rowptr[r], e = rowptr[r + 1];
    if (e - s <= ORC_SPMM_CHUNK) {
      if (!accumulate)
        for (int k = 0; k < d; ++k) y[k] = 0.f;
      for (int32_t j = s; j < e; ++j) {
        const float a = val[j];
        const float* x = X + (int64_t)col[j] * d;
        for (int k = 0; k < d; ++k) y[k] = fmaf(a, x[k], y[k]);
      }
      continue;
    }
    for (int32_t c = s; c < e; c += ORC_SPMM_CHUNK) {
      const int32_t ce = c + ORC_SPMM_CHUNK < e ? c + ORC_SPMM_CHUNK : e;
      for (int k = 0; k < d; ++k) part[k] = 0.f;
      for (int32_t j = c; j < ce; ++j) {
        const float a = val[j];
        const float* x = X + (int64_t)col[j] * d;
        for (int k = 0; k < d; ++k) part[k] = fmaf(a, x[k], part[k]);
      }
      for (int k = 0; k < d; ++k) tot[k] = (c == s) ? part[k] : tot[k] + part[k];
    }
    for (int k = 0; k < d; ++k) y[k] = accumulate ? y[k] + tot[k] : tot[k];
  }
  free(part); free(tot);
}

/* ---- NGCF.embedding_propagation, models/ngcf.py:60-72 (identity hoisted: (L+I)E = LE + E) ---------- */
void orc_ngcf_layer_fwd(const int32_t* rowptr, const int32_t* col, const float* val, int64_t n, int d,
                        const float* E, const float* W1, const float* W2, float slope, float* Enext, float* LE) {
  orc_spmm_csr(rowptr, col, val, n, d, E, LE, 0);
  for (int64_t r = 0; r < n; ++r) {
    const float* e = E + r * d;
    const float* le = LE + r * d;
    for (int o = 0; o < d; ++o) {
      float acc = 0.f;
      for (int k = 0; k < d; ++k) acc = fmaf(le[k] + e[k], W1[o * d + k], acc);
      for (int k = 0; k < d; ++k) acc = fmaf(e[k] * le[k], W2[o * d + k], acc);
      Enext[r * d + o] = acc > 0.f ? acc : acc * slope;
    }
  }
}

/* gradient of one layer; G accumulates, dW1/dW2 overwritten; T scratch [n*d] */
void orc_ngcf_layer_bwd(const int32_t* rowptrT, const int32_t* colT, const float* valT, int64_t n, int d,
                        const float* E, const float* LE, const float* Enext, const float* Gnext, const float* W1,
                        const float* W2, float slope, float* G, float* T, float* dW1, float* dW2) {
  double* a1 = (double*)calloc((size_t)d * d, sizeof(double));
  double* a2 = (double*)calloc((size_t)d * d, sizeof(double));
  float* dz = (float*)malloc(sizeof(float) * d);
  for (int64_t r = 0; r < n; ++r) {
    const float* e = E + r * d;
    const float* le = LE + r * d;
    for (int o = 0; o < d; ++o) {
      const float g = Gnext[r * d + o];
      dz[o] = Enext[r * d + o] > 0.f ? g : g * slope;
    }
    for (int i = 0; i < d; ++i) {
      float ds = 0.f, dp = 0.f;
      for (int o = 0; o < d; ++o) {
        ds = fmaf(dz[o], W1[o * d + i], ds);
        dp = fmaf(dz[o], W2[o * d + i], dp);
      }
      T[r * d + i] = fmaf(dp, e[i], ds);
      G[r * d + i] += fmaf(dp, le[i], ds);
    }
    for (int o = 0; o < d; ++o)
      for (int i = 0; i < d; ++i) {
        a1[o * d + i] += (double)dz[o] * (double)(le[i] + e[i]);
        a2[o * d + i] += (double)dz[o] * (double)(e[i] * le[i]);
      }
  }
  for (int i = 0; i < d * d; ++i) { dW1[i] = (float)a1[i]; dW2[i] = (float)a2[i]; }
  free(a1); free(a2); free(dz);
  orc_spmm_csr(rowptrT, colT, valT, n, d, T, G, 1);
}

/* ---- tail of NGCF.bpr_forward + BPRLoss, models/ngcf.py:37-45 --------------------------------------- */
float orc_ngcf_tail(const float* const* E_layers, float* const* G_layers, int n_layers, int64_t nU, int d,
                    const int64_t* uid, const int64_t* pos, const int64_t* neg, int64_t B, float* pos_out,
                    float* neg_out) {
  double lacc = 0.0;
  const float inv_b = 1.f / (float)B;
  for (int64_t b = 0; b < B; ++b) {
    float dp = 0.f, dn = 0.f;
    for (int l = 0; l <= n_layers; ++l) {
      const float* u = E_layers[l] + uid[b] * d;
      const float* p = E_layers[l] + (nU + pos[b]) * d;
      const float* n = E_layers[l] + (nU + neg[b]) * d;
      for (int k = 0; k < d; ++k) { dp = fmaf(u[k], p[k], dp); dn = fmaf(u[k], n[k], dn); }
    }
    if (pos_out) pos_out[b] = dp;
    if (neg_out) neg_out[b] = dn;
    const float x = dp - dn;
    lacc += (double)neg_logsigmoid(x);
    if (G_layers) {
      const float g = neg_logsigmoid_grad(x) * inv_b;
      for (int l = 0; l <= n_layers; ++l) {
        const float* u = E_layers[l] + uid[b] * d;
        const float* p = E_layers[l] + (nU + pos[b]) * d;
        const float* n = E_layers[l] + (nU + neg[b]) * d;
        float* gu = G_layers[l] + uid[b] * d;
        float* gp = G_layers[l] + (nU + pos[b]) * d;
        float* gn = G_layers[l] + (nU + neg[b]) * d;
        for (int k = 0; k < d; ++k) {
          gu[k] += g * p[k] - g * n[k];
          gp[k] += g * u[k];
          gn[k] -= g * u[k];
        }
      }
    }
  }
  return (float)(lacc / (double)B);
}

/* ---- MFTrainer.evaluate + _generate_top_k_recommendation + metric.py -------------------------------- */
/* trainers/mf_trainer.py:134-178; metric.py:7-109. Tie-break fixed to (score desc, item id asc) (Q5). */
static int ranks_ahead(float sa, int64_t ia, float sb, int64_t ib) { return sa > sb || (sa == sb && ia < ib); }

void orc_eval_topk_metrics(const float* Uemb, const float* Vemb, int64_t nI, int d, const int64_t* eval_uid,
                           int64_t n_eval, const int32_t* mask_ptr, const int32_t* mask_idx,
                           const int32_t* act_ptr, const int32_t* act_idx, const double* inv_log2, int K,
                           int64_t* topk_out, float* topk_score, double* user_metrics, double* sums) {
  float* sc = (float*)malloc(sizeof(float) * (size_t)nI);
  float* bs = (float*)malloc(sizeof(float) * (size_t)K);
  int64_t* bi = (int64_t*)malloc(sizeof(int64_t) * (size_t)K);
  for (int m = 0; m < 6; ++m) sums[m] = 0.0;
  for (int64_t e = 0; e < n_eval; ++e) {
    const float* u = Uemb + eval_uid[e] * d;
    for (int64_t i = 0; i < nI; ++i) sc[i] = dot_chain(u, Vemb + i * d, d);
    for (int32_t j = mask_ptr[e]; j < mask_ptr[e + 1]; ++j) sc[mask_idx[j]] = MASK_VALUE;
    int cnt = 0;
    for (int64_t i = 0; i < nI; ++i) {
      if (cnt == K && !ranks_ahead(sc[i], i, bs[K - 1], bi[K - 1])) continue;
      int pos = cnt < K ? cnt : K - 1;
      while (pos > 0 && ranks_ahead(sc[i], i, bs[pos - 1], bi[pos - 1])) {
        bs[pos] = bs[pos - 1]; bi[pos] = bi[pos - 1]; --pos;
      }
      bs[pos] = sc[i]; bi[pos] = i;
      if (cnt < K) ++cnt;
    }
    for (int j = 0; j < K; ++j) {
      topk_out[e * K + j] = j < cnt ? bi[j] : -1;
      if (topk_score) topk_score[e * K + j] = j < cnt ? bs[j] : -INFINITY;
    }
    /* metrics */
    const int32_t* A = act_idx + act_ptr[e];
    const int LA = act_ptr[e + 1] - act_ptr[e];
    int nuniq = 0;
    for (int a = 0; a < LA; ++a) {
      int seen = 0;
      for (int b = 0; b < a; ++b) seen |= (A[b] == A[a]);
      nuniq += !seen;
    }
    int hits = 0;
    double ap = 0.0, dcg = 0.0, idcg = 0.0;
    for (int i = 1; i <= K && i <= cnt; ++i) {
      const int64_t p = bi[i - 1];
      int in_a = 0;
      for (int a = 0; a < LA; ++a) in_a |= (A[a] == p);
      if (!in_a) continue;
      ++hits;
      /* |set(A[:i]) & set(P[:i])| / i   (metric.py:73-75, Q7) */
      int c = 0;
      for (int j = 0; j < i; ++j) {
        int f = 0;
        for (int a = 0; a < LA && a < i; ++a) f |= (A[a] == bi[j]);
        c += f;
      }
      ap += (double)c / (double)i;
      if (i <= LA) dcg += inv_log2[i - 1]; /* metric.py:107 (Q8) */
    }
    for (int i = 1; i <= K && i <= LA; ++i) idcg += inv_log2[i - 1];
    double* um = user_metrics + e * 4;
    um[0] = (double)hits / (double)K;
    um[1] = nuniq > 0 ? (double)hits / (double)nuniq : 0.0;
    um[2] = LA > 0 ? ap / (double)LA : 0.0;
    um[3] = (nuniq > 0 && idcg > 0.0) ? dcg / idcg : 0.0;
    sums[0] += um[0]; sums[1] += um[1]; sums[2] += um[2]; sums[3] += um[3];
    sums[4] += nuniq > 0 ? 1.0 : 0.0;
    sums[5] += LA > 0 ? 1.0 : 0.0;
  }
  free(sc); free(bs); free(bi);
}

/* ------------------------------------------------------------------------------------------------
 * Negative sampler (reference data/datasets/mf_dataset.py:18-22: draw uniformly over the items until the
 * draw is not one of the user's positives). The reference draws from NumPy's global Mersenne Twister, one
 * sample at a time inside DataLoader workers — a sequential stream a data-parallel sampler cannot reproduce,
 * so the stream is REDEFINED as counter-based Philox4x32-10 (Salmon et al., SC'11; Random123 known-answer
 * vectors pinned in tests/test_oracle_golden.py): triple t uses key = seed, counter = (t_lo, t_hi, block, 0),
 * words consumed in order; a word w maps to an item with Lemire's unbiased multiply-shift (reject w when
 * lo32(w * nI) < (2^32 - nI) mod nI), and the item is rejected when it is in the user's positive list.
 * Same distribution as the reference (uniform over the user's non-positives), different stream. */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* returns the number of triples that exhausted max_blocks Philox blocks (neg = -1 for those) */
int64_t orc_sample_negatives(const int64_t* uid, int64_t n, const int32_t* pos_ptr, const int32_t* pos_idx,
                             int64_t num_items, uint64_t seed, uint64_t offset, int max_blocks, int64_t* neg_out) {
  const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  const uint32_t nI = (uint32_t)num_items;
  const uint32_t thresh = (uint32_t)(0u - nI) % nI;
  int64_t failed = 0;
  for (int64_t t = 0; t < n; ++t) {
    const uint64_t idx = offset + (uint64_t)t;
    const int32_t lo = pos_ptr[uid[t]], hi = pos_ptr[uid[t] + 1];
    int64_t found = -1;
    for (int blk = 0; blk < max_blocks && found < 0; ++blk) {
      const uint32_t ctr[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)blk, 0u};
      uint32_t w[4];
      orc_philox4x32_10(ctr, key, w);
      for (int q = 0; q < 4 && found < 0; ++q) {
        const uint64_t m = (uint64_t)w[q] * nI;
        if ((uint32_t)m < thresh) continue;
        const int32_t cand = (int32_t)(m >> 32);
        int32_t a = lo, b = hi;                       /* binary search in the sorted positives */
        while (a < b) { const int32_t mid = (a + b) >> 1; if (pos_idx[mid] < cand) a = mid + 1; else b = mid; }
        if (a < hi && pos_idx[a] == cand) continue;
        found = cand;
      }
    }
    if (found < 0) ++failed;
    neg_out[t] = found;
  }
  return failed;
}
