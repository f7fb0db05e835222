// Noun-pronoun distillation (BASELINE config 5): the soft-KD loss between the teacher (noun) and student (pronoun)
// predictions with its own Hungarian matching of the unmatched queries (reference models/mdetr.py:520-599), a
// warp-parallel batched LSAP that keeps scipy's tie rules, and the k-means / memory-bank arithmetic of
// ClusterCriterion (models/mdetr.py:29-312, models/kmeans.py:21-133).
#include <math.h>

#include "boxmath.cuh"
#include "common.cuh"
#include "host_util.h"
#include "lsap_warp.cuh"

namespace toist {

// cost [P, ld_r, ld_c] fp32; problem p has n_rows[p] x n_cols[p] valid entries.  col_of_row [P, ld_r]: assigned
// column of every row (-1 = unassigned, only when n_rows > n_cols).  flags[0] |= 1 on NaN / -inf / infeasible.
__global__ void lsap_batched_kernel(const float* __restrict__ cost, const int* __restrict__ n_rows,
                                    const int* __restrict__ n_cols, int* __restrict__ col_of_row, int* __restrict__ flags,
                                    int ld_r, int ld_c) {
  pdl_prologue();
  __shared__ LsapShared w;
  const int p = blockIdx.x, lane = threadIdx.x;
  const int R = n_rows[p], C = n_cols[p];
  const float* c = cost + (size_t)p * ld_r * ld_c;
  int* out = col_of_row + (size_t)p * ld_r;
  for (int i = lane; i < ld_r; i += 32) out[i] = -1;
  if (R <= 0 || C <= 0) return;
  int bad = 0;
  for (int i = lane; i < R * C; i += 32) {
    const float v = c[(i / C) * ld_c + (i % C)];
    if (isnan(v) || v == -INFINITY) bad = 1;
  }
  bad = __any_sync(0xffffffffu, bad);
  if (bad) {
    if (lane == 0) atomicOr(flags, 1);
    return;
  }
  __syncwarp();
  int rc;
  if (R <= C) {
    rc = lsap_warp(c, ld_c, 1, R, C, w);
    if (rc == 0)
      for (int i = lane; i < R; i += 32) out[i] = w.col4row[i];
  } else {  // scipy transposes when there are more rows than columns
    rc = lsap_warp(c, 1, ld_c, C, R, w);
    if (rc == 0)
      for (int j = lane; j < C; j += 32) out[w.col4row[j]] = j;
  }
  if (rc != 0 && lane == 0) atomicOr(flags, 1);
}

// ------------------------------------------------------------------------------------------------ soft-KD
// bi[l, b, q] = (sum_{c < C-1} softmax(logits)[c], softmax(logits)[C-1])          (mdetr.py:552-556)
// one warp per (l, b, q) row
__global__ void biprob_kernel(const float* __restrict__ logits, float* __restrict__ bi, long long rows, int C) {
  pdl_prologue();
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* lg = logits + row * C;
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, lg[c]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int c = lane; c < C; c += 32) sum += expf(lg[c] - mx);
  sum = warp_sum(sum);
  float obj = 0.f;
  for (int c = lane; c < C - 1; c += 32) obj += __fdiv_rn(expf(lg[c] - mx), sum);
  obj = warp_sum(obj);
  if (lane == 0) {
    bi[row * 2 + 0] = obj;
    bi[row * 2 + 1] = __fdiv_rn(expf(lg[C - 1] - mx), sum);
  }
}

// Per (l, b): the unmatched ("false positive") queries of teacher and student in ascending order and their cost
// matrix  C[s, t] = L1(box_s, box_t) + KL(bi_t || bi_s) - GIoU(box_s, box_t)            (mdetr.py:520-541)
// rows = student (source), columns = teacher (target).  fp_* [L, B, Q] (first n_fp entries valid), cost [L, B, Q, Q].
__global__ void softkd_cost_kernel(const float* __restrict__ bi_n, const float* __restrict__ bi_s,
                                   const float* __restrict__ box_n, const float* __restrict__ box_s,
                                   const int* __restrict__ match_n, const int* __restrict__ match_s,
                                   const int* __restrict__ tgt_count, int* __restrict__ fp_n, int* __restrict__ fp_s,
                                   int* __restrict__ n_fp, float* __restrict__ cost, int B, int Q, int Tmax) {
  pdl_prologue();
  extern __shared__ int sk_smem[];
  int* tp_n = sk_smem;       // [Q] 1 = matched to a target
  int* tp_s = tp_n + Q;
  int* cnt = tp_s + Q;       // [2]
  const int lb = blockIdx.x, b = lb % B;
  const int T = min(tgt_count[b], Tmax);
  for (int q = threadIdx.x; q < Q; q += blockDim.x) tp_n[q] = tp_s[q] = 0;
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const int qn = match_n[(size_t)lb * Tmax + t], qs = match_s[(size_t)lb * Tmax + t];
    if (qn >= 0) tp_n[qn] = 1;
    if (qs >= 0) tp_s[qs] = 1;
  }
  __syncthreads();
  if (threadIdx.x < 2) {  // ascending compaction (boolean-mask indexing in the reference keeps query order)
    const int* tp = threadIdx.x == 0 ? tp_n : tp_s;
    int* dst = (threadIdx.x == 0 ? fp_n : fp_s) + (size_t)lb * Q;
    int n = 0;
    for (int q = 0; q < Q; ++q)
      if (!tp[q]) dst[n++] = q;
    for (int k = n; k < Q; ++k) dst[k] = -1;
    cnt[threadIdx.x] = n;
  }
  __syncthreads();
  const int ns = cnt[1], nt = cnt[0];
  if (threadIdx.x == 0) {  // n_fp [2, P]: plane 0 = teacher (columns), plane 1 = student (rows)
    n_fp[lb] = nt;
    n_fp[gridDim.x + lb] = ns;
  }
  const int* fn = fp_n + (size_t)lb * Q;
  const int* fs = fp_s + (size_t)lb * Q;
  float* cm = cost + (size_t)lb * Q * Q;
  for (int e = threadIdx.x; e < ns * nt; e += blockDim.x) {
    const int s = e / nt, t = e % nt;
    const int qs = fs[s], qt = fn[t];
    const float* ps = bi_s + ((size_t)lb * Q + qs) * 2;
    const float* pt = bi_n + ((size_t)lb * Q + qt) * 2;
    // (target * (log target - log source)).sum(-1), two classes
    const float k0 = __fmul_rn(pt[0], __fsub_rn(logf(pt[0]), logf(ps[0])));
    const float k1 = __fmul_rn(pt[1], __fsub_rn(logf(pt[1]), logf(ps[1])));
    const float c_class = __fadd_rn(k0, k1);
    const float* bs = box_s + ((size_t)lb * Q + qs) * 4;
    const float* bt = box_n + ((size_t)lb * Q + qt) * 4;
    const float c_bbox = box_l1(bs, bt);
    const float c_giou = -giou_xyxy(to_xyxy(bs[0], bs[1], bs[2], bs[3]), to_xyxy(bt[0], bt[1], bt[2], bt[3]));
    cm[s * Q + t] = __fadd_rn(__fadd_rn(c_bbox, c_class), c_giou);
  }
}

// Per (l, b): pair every student query with a teacher query (matched targets by target index, the rest by the LSAP
// result) and accumulate KL(teacher || student) over the Q pairs:  loss[l] += sum / Q / B   (kl_div batchmean, then the
// mean over images, mdetr.py:595-597).  pair_n[l, b, q_s] = teacher query of student query q_s (-1 = unpaired).
__device__ __forceinline__ float xlogy_diff(float t, float s) {  // F.kl_div pointwise: t * (log t - log s), 0 when t == 0
  return t > 0.f ? t * (logf(t) - logf(s)) : 0.f;
}
__global__ void softkd_loss_kernel(const float* __restrict__ bi_n, const float* __restrict__ bi_s,
                                   const int* __restrict__ match_n, const int* __restrict__ match_s,
                                   const int* __restrict__ tgt_count, const int* __restrict__ fp_n,
                                   const int* __restrict__ fp_s, const int* __restrict__ n_fp,
                                   const int* __restrict__ col_of_row, const int* __restrict__ flags, int* __restrict__ pair_n,
                                   float* __restrict__ loss, int B, int Q, int Tmax) {
  pdl_prologue();
  const int lb = blockIdx.x, b = lb % B, l = lb / B;
  const int T = min(tgt_count[b], Tmax);
  int* pn = pair_n + (size_t)lb * Q;
  for (int q = threadIdx.x; q < Q; q += blockDim.x) pn[q] = -1;
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const int qn = match_n[(size_t)lb * Tmax + t], qs = match_s[(size_t)lb * Tmax + t];
    if (qn >= 0 && qs >= 0) pn[qs] = qn;
  }
  const int ns = n_fp[gridDim.x + lb], nt = n_fp[lb];
  const int npair = min(ns, nt);
  for (int s = threadIdx.x; s < ns; s += blockDim.x) {
    const int t = col_of_row[(size_t)lb * Q + s];
    if (t >= 0) pn[fp_s[(size_t)lb * Q + s]] = fp_n[(size_t)lb * Q + t];
  }
  __syncthreads();
  float acc = 0.f;
  for (int q = threadIdx.x; q < Q; q += blockDim.x) {
    const int qn = pn[q];
    if (qn < 0) continue;
    const float* ps = bi_s + ((size_t)lb * Q + q) * 2;
    const float* pt = bi_n + ((size_t)lb * Q + qn) * 2;
    acc += xlogy_diff(pt[0], ps[0]) + xlogy_diff(pt[1], ps[1]);
  }
  acc = warp_sum(acc);
  __shared__ float part[32];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += part[i];
    const int rows = T + npair;  // batchmean divides by the number of (tp + fp) rows
    // NaN / infeasible costs: the reference's scipy call raises; poison the loss instead of synchronising
    atomicAdd(loss + l, flags[0] ? __int_as_float(0x7fc00000) : s / (float)max(rows, 1) / (float)B);
  }
}

// d loss[l] / d student logits.  With P = softmax(z), s0 = sum_{c<last} P_c, s1 = P_last and teacher pair (t0, t1):
//   d/dz_c [-(t0 log s0 + t1 log s1)] = (t1 / s1 - t0 / s0) * ds0/dz_c,  ds0/dz_c = c < last ? P_c * s1 : -s0 * s1
// and, because t0 + t1 = s0 + s1 = 1,  t1 / s1 - t0 / s0 = (t1 - s1) / (s0 s1): the gradient is
//   c < last: P_c (t1 - s1) / s0,   c = last: s1 - t1      (one subtraction of nearly equal numbers instead of two)
// one warp per (l, b, q); scale = gout[l] / rows / B
__global__ void softkd_bwd_kernel(const float* __restrict__ logits_s, const float* __restrict__ bi_n,
                                  const float* __restrict__ bi_s, const int* __restrict__ pair_n,
                                  const int* __restrict__ tgt_count, const int* __restrict__ n_fp,
                                  const float* __restrict__ gout, float* __restrict__ dlogits, long long rows, int B,
                                  int Q, int C, int Tmax) {
  pdl_prologue();
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const long long lb = row / Q;
  const int b = (int)(lb % B), l = (int)(lb / B);
  float* dz = dlogits + row * C;
  const int qn = pair_n[row];
  if (qn < 0) {
    for (int c = lane; c < C; c += 32) dz[c] = 0.f;
    return;
  }
  const int T = min(tgt_count[b], Tmax);
  const long long P = rows / Q;
  const int n_rows = max(T + min(n_fp[lb], n_fp[P + lb]), 1);
  const float scale = gout[l] / (float)n_rows / (float)B;
  const float* lg = logits_s + row * C;
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, lg[c]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int c = lane; c < C; c += 32) sum += expf(lg[c] - mx);
  sum = warp_sum(sum);
  const float s0 = bi_s[row * 2], s1 = bi_s[row * 2 + 1];
  const float t0 = bi_n[(lb * Q + qn) * 2], t1 = bi_n[(lb * Q + qn) * 2 + 1];
  const float diff = (t1 - s1) * scale;
  (void)t0;
  for (int c = lane; c < C; c += 32) {
    const float p = expf(lg[c] - mx) / sum;
    dz[c] = c < C - 1 ? p * diff / s0 : -diff;
  }
}

// ------------------------------------------------------------------------------------------------ k-means
// Lloyd iterations of models/kmeans.py:21-96 for ONE task, entirely on the device (the reference synchronises with
// the host once per iteration to test convergence).  X [N, D] memory bank, centers [K, D] updated in place, choice [N].
// One CTA; stops when (sum_k ||c_k - c_k_prev||)^2 < tol or after max_iter iterations.
//
// Memory access: the assignment step gives every thread one point and walks its D features in order (the summation
// order of the distance is part of the result: near-equidistant points must fall on the same side as before), so
// the kernel first writes the transposed bank XT [D, N] into `xt` (caller workspace, 1 MB for the 1024 x 256 bank:
// L2 resident): thread n then reads XT[d * N + n], coalesced across the warp, instead of striding through X by rows
// (32 sectors per warp load; 10 ms per call at N = 1024, D = 256).  The update step reads X by columns (coalesced) and
// adds `member ? x : 0` in point order, which leaves every partial sum bit-identical to the branchy form.
constexpr int kKmMaxK = 8;

template <int KK>
__device__ __forceinline__ void kmeans_assign(const float* __restrict__ XT, const float* __restrict__ cen,
                                              int* __restrict__ choice, int N, int D) {
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float acc[KK];
#pragma unroll
    for (int k = 0; k < KK; ++k) acc[k] = 0.f;
#pragma unroll 4
    for (int d = 0; d < D; ++d) {
      const float x = XT[(size_t)d * N + n];
#pragma unroll
      for (int k = 0; k < KK; ++k) {
        const float df = x - cen[k * D + d];
        acc[k] += df * df;
      }
    }
    float best = INFINITY;
    int bk = 0;
#pragma unroll
    for (int k = 0; k < KK; ++k)
      if (acc[k] < best) {  // first minimum on ties (torch.argmin)
        best = acc[k];
        bk = k;
      }
    choice[n] = bk;
  }
}

template <int KK>
__device__ __forceinline__ void kmeans_update(const float* __restrict__ X, const int* __restrict__ choice,
                                              const int* __restrict__ cnt, float* __restrict__ cen,
                                              float* __restrict__ shift2, int N, int D) {
  // thread d owns column d of every centre: ONE pass over X, `member ? x : 0` added to each cluster's running sum in
  // point order (bit-identical to summing the members only)
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc[KK];
#pragma unroll
    for (int k = 0; k < KK; ++k) acc[k] = 0.f;
#pragma unroll 8
    for (int n = 0; n < N; ++n) {
      const float x = X[(size_t)n * D + d];
      const int c = choice[n];
#pragma unroll
      for (int k = 0; k < KK; ++k) acc[k] += (c == k) ? x : 0.f;
    }
#pragma unroll
    for (int k = 0; k < KK; ++k) {
      if (cnt[k] == 0) continue;  // an empty cluster keeps its centre (kmeans.py:70-72)
      const float nc = acc[k] / (float)cnt[k];
      const float df = nc - cen[k * D + d];
      atomicAdd(&shift2[k], df * df);
      cen[k * D + d] = nc;
    }
  }
}

// grid = number of problems.  Problem p clusters bank `task_of[p]` of X_base [n_tasks, N, D] (task_of null: X_base is
// the single bank) from centers [p, K, D] (in place); choice [p, N]; iters [p]; xt [p, D, N] workspace (may be null);
// query [p, D] (may be null) -> qchoice[p] = nearest final centre of the query (kmeans_predict arithmetic).
__global__ void kmeans_kernel(const float* __restrict__ X_base, const int* __restrict__ task_of,
                              float* __restrict__ centers_all, int* __restrict__ choice_all, int* __restrict__ iters,
                              float* __restrict__ xt_all, const float* __restrict__ query, int* __restrict__ qchoice,
                              int N, int D, int K, float tol, int max_iter) {
  pdl_prologue();
  extern __shared__ float km_smem[];
  float* cen = km_smem;                 // [K, D]
  float* shift2 = cen + K * D;          // [K]
  int* cnt = reinterpret_cast<int*>(shift2 + K);  // [K]
  __shared__ int done;
  __shared__ float tile[32][33];
  const int prob = blockIdx.x;
  const float* X = X_base + (task_of != nullptr ? (size_t)task_of[prob] * N * D : 0);
  float* centers = centers_all + (size_t)prob * K * D;
  int* choice = choice_all + (size_t)prob * N;
  float* xt = xt_all != nullptr ? xt_all + (size_t)prob * D * N : nullptr;
  for (int i = threadIdx.x; i < K * D; i += blockDim.x) cen[i] = centers[i];
  if (threadIdx.x == 0) done = 0;
  const bool fast = xt != nullptr && K <= kKmMaxK && blockDim.x == 1024;
  if (fast) {  // XT = X^T through 32 x 32 shared-memory tiles (coalesced both ways)
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int n0 = 0; n0 < N; n0 += 32)
      for (int d0 = 0; d0 < D; d0 += 32) {
        tile[ty][tx] = (n0 + ty < N && d0 + tx < D) ? X[(size_t)(n0 + ty) * D + d0 + tx] : 0.f;
        __syncthreads();
        if (d0 + ty < D && n0 + tx < N) xt[(size_t)(d0 + ty) * N + n0 + tx] = tile[tx][ty];
        __syncthreads();
      }
    __threadfence_block();
  }
  __syncthreads();
  int it = 0;
  for (; it < max_iter; ++it) {
    // assignment: argmin_k sum_d (x - c)^2, first minimum on ties (torch.argmin)
    if (fast) {
      switch (K) {
        case 1: kmeans_assign<1>(xt, cen, choice, N, D); break;
        case 2: kmeans_assign<2>(xt, cen, choice, N, D); break;
        case 3: kmeans_assign<3>(xt, cen, choice, N, D); break;
        case 4: kmeans_assign<4>(xt, cen, choice, N, D); break;
        case 5: kmeans_assign<5>(xt, cen, choice, N, D); break;
        case 6: kmeans_assign<6>(xt, cen, choice, N, D); break;
        case 7: kmeans_assign<7>(xt, cen, choice, N, D); break;
        default: kmeans_assign<8>(xt, cen, choice, N, D); break;
      }
    } else {
      for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const float* x = X + (size_t)n * D;
        float best = INFINITY;
        int bk = 0;
        for (int k = 0; k < K; ++k) {
          float acc = 0.f;
          for (int d = 0; d < D; ++d) {
            const float df = x[d] - cen[k * D + d];
            acc += df * df;
          }
          if (acc < best) {
            best = acc;
            bk = k;
          }
        }
        choice[n] = bk;
      }
    }
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      cnt[k] = 0;
      shift2[k] = 0.f;
    }
    __threadfence_block();
    __syncthreads();
    for (int n = threadIdx.x; n < N; n += blockDim.x) atomicAdd(&cnt[choice[n]], 1);
    __syncthreads();
    // update: mean of the members
    if (K <= kKmMaxK) {
      switch (K) {
        case 1: kmeans_update<1>(X, choice, cnt, cen, shift2, N, D); break;
        case 2: kmeans_update<2>(X, choice, cnt, cen, shift2, N, D); break;
        case 3: kmeans_update<3>(X, choice, cnt, cen, shift2, N, D); break;
        case 4: kmeans_update<4>(X, choice, cnt, cen, shift2, N, D); break;
        case 5: kmeans_update<5>(X, choice, cnt, cen, shift2, N, D); break;
        case 6: kmeans_update<6>(X, choice, cnt, cen, shift2, N, D); break;
        case 7: kmeans_update<7>(X, choice, cnt, cen, shift2, N, D); break;
        default: kmeans_update<8>(X, choice, cnt, cen, shift2, N, D); break;
      }
    } else {
      for (int e = threadIdx.x; e < K * D; e += blockDim.x) {
        const int k = e / D, d = e % D;
        if (cnt[k] == 0) continue;
        float acc = 0.f;
        for (int n = 0; n < N; ++n)
          if (choice[n] == k) acc += X[(size_t)n * D + d];
        const float nc = acc / (float)cnt[k];
        const float df = nc - cen[e];
        atomicAdd(&shift2[k], df * df);
        cen[e] = nc;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int k = 0; k < K; ++k) s += sqrtf(shift2[k]);
      done = (s * s < tol) ? 1 : 0;
    }
    __syncthreads();
    if (done) {
      ++it;
      break;
    }
  }
  for (int i = threadIdx.x; i < K * D; i += blockDim.x) centers[i] = cen[i];
  if (threadIdx.x == 0 && iters != nullptr) iters[prob] = it;
  if (query != nullptr && threadIdx.x < 32) {  // nearest final centre of the query: kmeans_predict_kernel's arithmetic
    const int lane = threadIdx.x;
    const float* q = query + (size_t)prob * D;
    float best = INFINITY;
    int bk = 0;
    for (int k = 0; k < K; ++k) {
      float acc = 0.f;
      for (int d = lane; d < D; d += 32) {
        const float df = q[d] - cen[(size_t)k * D + d];
        acc += df * df;
      }
      acc = warp_sum(acc);
      if (acc < best) {
        best = acc;
        bk = k;
      }
    }
    if (lane == 0) qchoice[prob] = bk;
  }
}

// choice[m] = argmin_k ||x_m - c_k||^2 (kmeans_predict, kmeans.py:99-133); one warp per row
__global__ void kmeans_predict_kernel(const float* __restrict__ X, const float* __restrict__ centers,
                                      int* __restrict__ choice, int M, int D, int K) {
  pdl_prologue();
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (m >= M) return;
  float best = INFINITY;
  int bk = 0;
  for (int k = 0; k < K; ++k) {
    float acc = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float df = X[(size_t)m * D + d] - centers[(size_t)k * D + d];
      acc += df * df;
    }
    acc = warp_sum(acc);
    if (acc < best) {
      best = acc;
      bk = k;
    }
  }
  if (lane == 0) choice[m] = bk;
}

// out[m, d] = sum_t w[m, t] * x[t, m, d]   (x fp32 [T, M, D] sequence layout; w carries the 1/n of the token means)
__global__ void token_wsum_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ out,
                                  int T, int M, int D) {
  pdl_prologue();
  const int m = blockIdx.x;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = 0.f;
    for (int t = 0; t < T; ++t) {
      const float wt = w[m * T + t];
      if (wt != 0.f) acc += wt * x[((size_t)t * M + m) * D + d];  // NaN weights (empty selection) propagate
    }
    out[(size_t)m * D + d] = acc;
  }
}

// dx[t, m, d] = w[m, t] * dout[m, d]
__global__ void token_wsum_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ w, float* __restrict__ dx,
                                      int T, int M, int D) {
  pdl_prologue();
  const int m = blockIdx.x % M, t = blockIdx.x / M;
  const float wt = w[m * T + t];
  for (int d = threadIdx.x; d < D; d += blockDim.x) dx[((size_t)t * M + m) * D + d] = wt * dout[(size_t)m * D + d];
}

// x[t, m, :] = feat[m, :] for every selected token (in-place replacement of the caption tokens by the cluster centre)
__global__ void token_fill_kernel(float* __restrict__ x, const uint8_t* __restrict__ sel, const float* __restrict__ feat,
                                  int T, int M, int D) {
  pdl_prologue();
  const int m = blockIdx.x;
  for (int t = 0; t < T; ++t) {
    if (!sel[m * T + t]) continue;
    for (int d = threadIdx.x; d < D; d += blockDim.x)
      x[((size_t)t * M + m) * D + d] = feat != nullptr ? feat[(size_t)m * D + d] : 0.f;
  }
}

// F.mse_loss(a[m], b[m]) averaged over the rows with use[m] != 0 (loss_cluster_feature, mdetr.py:270-278) and its
// gradient w.r.t. a.  One CTA.
__global__ void mse_rows_kernel(const float* __restrict__ a, const float* __restrict__ b, const uint8_t* __restrict__ use,
                                float* __restrict__ loss, float* __restrict__ da, int M, int D) {
  pdl_prologue();
  int cnt = 0;
  for (int m = 0; m < M; ++m) cnt += use[m] ? 1 : 0;
  float acc = 0.f;
  for (int e = threadIdx.x; e < M * D; e += blockDim.x) {
    const int m = e / D;
    float g = 0.f;
    if (use[m]) {
      const float df = a[e] - b[e];
      acc += df * df;
      g = 2.f * df / (float)D / (float)cnt;
    }
    if (da != nullptr) da[e] = g;
  }
  acc = warp_sum(acc);
  __shared__ float part[32];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += part[i];
    loss[0] = cnt ? s / (float)D / (float)cnt : 0.f;
  }
}

// pairwise L1 distance (torch.cdist p=1, mdetr.py:98): a [n, D], b [m, D] -> out [n, m]; one warp per entry
__global__ void cdist_l1_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int n,
                                int m, int D) {
  pdl_prologue();
  const long long e = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (e >= (long long)n * m) return;
  const int i = (int)(e / m), j = (int)(e % m);
  float acc = 0.f;
  for (int d = lane; d < D; d += 32) acc += fabsf(a[(size_t)i * D + d] - b[(size_t)j * D + d]);
  acc = warp_sum(acc);
  if (lane == 0) out[e] = acc;
}

}  // namespace toist

using namespace toist;

extern "C" {

int toist_lsap_batched(const float* cost, const int32_t* n_rows, const int32_t* n_cols, int32_t* col_of_row,
                       int32_t* flags, int32_t n_problems, int32_t ld_rows, int32_t ld_cols, void* stream) {
  TOIST_REQUIRE(cost && n_rows && n_cols && col_of_row && flags, "toist_lsap_batched: null pointer");
  TOIST_REQUIRE(ld_rows >= 1 && ld_cols >= 1 && ld_rows <= kLsapMax && ld_cols <= kLsapMax,
                "toist_lsap_batched: problems up to %d x %d are supported (got %d x %d)", kLsapMax, kLsapMax, ld_rows,
                ld_cols);
  if (n_problems == 0) return TOIST_OK;
  launch_pdl(lsap_batched_kernel, dim3(n_problems), dim3(32), 0, (cudaStream_t)stream, cost, n_rows, n_cols, col_of_row, flags, ld_rows,
                                                                   ld_cols);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_softkd_fwd(const float* logits_noun, const float* logits_sth, const float* boxes_noun, const float* boxes_sth,
                     const int32_t* match_noun, const int32_t* match_sth, const int32_t* tgt_count, float* bi_noun,
                     float* bi_sth, int32_t* fp_noun, int32_t* fp_sth, int32_t* n_fp, float* cost, int32_t* col_of_row,
                     int32_t* pair_noun, int32_t* flags, float* loss, int32_t n_layers, int32_t batch, int32_t n_queries,
                     int32_t n_classes, int32_t t_max, void* stream) {
  TOIST_REQUIRE(logits_noun && logits_sth && boxes_noun && boxes_sth && match_noun && match_sth && tgt_count,
                "toist_softkd_fwd: null input");
  TOIST_REQUIRE(bi_noun && bi_sth && fp_noun && fp_sth && n_fp && cost && col_of_row && pair_noun && flags && loss,
                "toist_softkd_fwd: null workspace / output");
  TOIST_REQUIRE(n_queries >= 1 && n_queries <= kLsapMax, "toist_softkd_fwd: at most %d queries (got %d)", kLsapMax,
                n_queries);
  TOIST_REQUIRE(n_classes >= 2 && n_layers >= 1 && batch >= 1 && t_max >= 1, "toist_softkd_fwd: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  const long long rows = (long long)n_layers * batch * n_queries;
  const unsigned gb = (unsigned)((rows + 7) / 8);
  launch_pdl(biprob_kernel, dim3(gb), dim3(256), 0, st, logits_noun, bi_noun, rows, n_classes);
  launch_pdl(biprob_kernel, dim3(gb), dim3(256), 0, st, logits_sth, bi_sth, rows, n_classes);
  const int P = n_layers * batch;
  launch_pdl(softkd_cost_kernel, dim3(P), dim3(256), (2 * n_queries + 2) * sizeof(int), st, bi_noun, bi_sth, boxes_noun, boxes_sth,
                                                                       match_noun, match_sth, tgt_count, fp_noun, fp_sth,
                                                                       n_fp, cost, batch, n_queries, t_max);
  TOIST_CHECK_CUDA(cudaMemsetAsync(loss, 0, sizeof(float) * n_layers, st));
  // rows = student false positives (n_fp plane 1), columns = teacher false positives (plane 0)
  launch_pdl(lsap_batched_kernel, dim3(P), dim3(32), 0, st, cost, n_fp + P, n_fp, col_of_row, flags, n_queries, n_queries);
  launch_pdl(softkd_loss_kernel, dim3(P), dim3(128), 0, st, bi_noun, bi_sth, match_noun, match_sth, tgt_count, fp_noun, fp_sth, n_fp,
                                        col_of_row, flags, pair_noun, loss, batch, n_queries, t_max);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_softkd_bwd(const float* logits_sth, const float* bi_noun, const float* bi_sth, const int32_t* pair_noun,
                     const int32_t* tgt_count, const int32_t* n_fp, const float* gout, float* dlogits_sth,
                     int32_t n_layers, int32_t batch, int32_t n_queries, int32_t n_classes, int32_t t_max, void* stream) {
  TOIST_REQUIRE(logits_sth && bi_noun && bi_sth && pair_noun && tgt_count && n_fp && gout && dlogits_sth,
                "toist_softkd_bwd: null pointer");
  const long long rows = (long long)n_layers * batch * n_queries;
  if (rows == 0) return TOIST_OK;
  launch_pdl(softkd_bwd_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, (cudaStream_t)stream, 
      logits_sth, bi_noun, bi_sth, pair_noun, tgt_count, n_fp, gout, dlogits_sth, rows, batch, n_queries, n_classes,
      t_max);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_kmeans(const float* x, float* centers, int32_t* choice, int32_t* iters, float* xt_workspace, int32_t n,
                 int32_t dim, int32_t k, float tol, int32_t max_iter, void* stream) {
  TOIST_REQUIRE(x && centers && choice, "toist_kmeans: null pointer");
  TOIST_REQUIRE(n >= 1 && dim >= 1 && k >= 1 && max_iter >= 1, "toist_kmeans: bad sizes");
  const size_t smem = ((size_t)k * dim + 2 * k) * sizeof(float);
  TOIST_REQUIRE(smem <= 48 * 1024, "toist_kmeans: %d x %d centres do not fit shared memory", k, dim);
  launch_pdl(kmeans_kernel, dim3(1), dim3(1024), smem, (cudaStream_t)stream, x, (const int*)nullptr, centers, (int*)choice,
             (int*)iters, xt_workspace, (const float*)nullptr, (int*)nullptr, (int)n, (int)dim, (int)k, tol, (int)max_iter);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_kmeans_batched(const float* banks, const int32_t* task_of, float* centers, int32_t* choice, int32_t* iters,
                         float* xt_workspace, const float* query, int32_t* query_choice, int32_t n_problems, int32_t n,
                         int32_t dim, int32_t k, float tol, int32_t max_iter, void* stream) {
  TOIST_REQUIRE(banks && task_of && centers && choice, "toist_kmeans_batched: null pointer");
  TOIST_REQUIRE((query == nullptr) == (query_choice == nullptr), "toist_kmeans_batched: query needs query_choice");
  TOIST_REQUIRE(n >= 1 && dim >= 1 && k >= 1 && max_iter >= 1 && n_problems >= 0, "toist_kmeans_batched: bad sizes");
  if (n_problems == 0) return TOIST_OK;
  const size_t smem = ((size_t)k * dim + 2 * k) * sizeof(float);
  TOIST_REQUIRE(smem <= 48 * 1024, "toist_kmeans_batched: %d x %d centres do not fit shared memory", k, dim);
  launch_pdl(kmeans_kernel, dim3((unsigned)n_problems), dim3(1024), smem, (cudaStream_t)stream, banks, (const int*)task_of,
             centers, (int*)choice, (int*)iters, xt_workspace, query, (int*)query_choice, (int)n, (int)dim, (int)k, tol,
             (int)max_iter);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_kmeans_predict(const float* x, const float* centers, int32_t* choice, int32_t m, int32_t dim, int32_t k,
                         void* stream) {
  TOIST_REQUIRE(x && centers && choice, "toist_kmeans_predict: null pointer");
  if (m == 0) return TOIST_OK;
  launch_pdl(kmeans_predict_kernel, dim3((m + 3) / 4), dim3(128), 0, (cudaStream_t)stream, x, centers, choice, m, dim, k);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_token_wsum(const float* x, const float* w, float* out, int32_t n_tokens, int32_t batch, int32_t dim,
                     void* stream) {
  TOIST_REQUIRE(x && w && out, "toist_token_wsum: null pointer");
  if (batch == 0) return TOIST_OK;
  launch_pdl(token_wsum_kernel, dim3(batch), dim3(256), 0, (cudaStream_t)stream, x, w, out, n_tokens, batch, dim);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_token_wsum_bwd(const float* dout, const float* w, float* dx, int32_t n_tokens, int32_t batch, int32_t dim,
                         void* stream) {
  TOIST_REQUIRE(dout && w && dx, "toist_token_wsum_bwd: null pointer");
  if (batch * n_tokens == 0) return TOIST_OK;
  launch_pdl(token_wsum_bwd_kernel, dim3(batch * n_tokens), dim3(256), 0, (cudaStream_t)stream, dout, w, dx, n_tokens, batch, dim);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_token_fill(float* x, const uint8_t* sel, const float* feat, int32_t n_tokens, int32_t batch, int32_t dim,
                     void* stream) {
  TOIST_REQUIRE(x && sel, "toist_token_fill: null pointer");
  if (batch == 0) return TOIST_OK;
  launch_pdl(token_fill_kernel, dim3(batch), dim3(256), 0, (cudaStream_t)stream, x, sel, feat, n_tokens, batch, dim);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_mse_rows(const float* a, const float* b, const uint8_t* use, float* loss, float* da, int32_t rows, int32_t dim,
                   void* stream) {
  TOIST_REQUIRE(a && b && use && loss, "toist_mse_rows: null pointer");
  launch_pdl(mse_rows_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, a, b, use, loss, da, rows, dim);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_cdist_l1(const float* a, const float* b, float* out, int32_t n, int32_t m, int32_t dim, void* stream) {
  TOIST_REQUIRE(a && b && out, "toist_cdist_l1: null pointer");
  const long long e = (long long)n * m;
  if (e == 0) return TOIST_OK;
  launch_pdl(cdist_l1_kernel, dim3((unsigned)((e + 7) / 8)), dim3(256), 0, (cudaStream_t)stream, a, b, out, n, m, dim);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

}  // extern "C"
