// HungarianMatcher on the device: block-diagonal matching cost for every decoder layer in one launch, and a
// shortest-augmenting-path LSAP solved per (layer, image) problem without leaving the GPU.
// Reference: models/matcher.py:39-87 (cost + scipy LSAP per image), util/box_ops.py:11-61 (GIoU).
#include <math.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "host_util.h"
#include "boxmath.cuh"
#include "lsap.h"
#include "lsap_warp.cuh"

namespace toist {

// grid = (L*B, ceil(Q / 4)), block = 128 (4 warps): one query per warp (48 x 25 CTAs instead of 48 long-running ones)
// cost[l][b][q][t] for t < count[b]; padding columns are written as 0.
__global__ void match_cost_kernel(const float* __restrict__ logits, const float* __restrict__ boxes,
                                  const float* __restrict__ tgt_boxes, const int* __restrict__ tgt_count,
                                  const float* __restrict__ posmap, float* __restrict__ cost, int B, int Q, int C,
                                  int Tmax, float w_class, float w_bbox, float w_giou) {
  pdl_prologue();
  const int lb = blockIdx.x;
  const int b = lb % B;
  const int T = min(tgt_count[b], Tmax);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  extern __shared__ float prob[];  // [nwarps][C]
  float* myprob = prob + warp * C;
  for (int q = blockIdx.y * nwarps + warp; q < Q; q += nwarps * gridDim.y) {
    const float* lg = logits + ((size_t)lb * Q + q) * C;
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, lg[c]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float e = expf(lg[c] - mx);
      myprob[c] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    const float* bq = boxes + ((size_t)lb * Q + q) * 4;
    const float qcx = bq[0], qcy = bq[1], qw = bq[2], qh = bq[3];
    const BoxXYXY qb = to_xyxy(qcx, qcy, qw, qh);
    float* crow = cost + ((size_t)lb * Q + q) * Tmax;
    for (int t = 0; t < T; ++t) {
      const float* pm = posmap + ((size_t)b * Tmax + t) * C;
      float dot = 0.f;
      for (int c = lane; c < C; c += 32) dot += __fdiv_rn(myprob[c], sum) * pm[c];
      dot = warp_sum(dot);
      if (lane == 0) {
        const float* tb = tgt_boxes + ((size_t)b * Tmax + t) * 4;
        const float l1 = __fadd_rn(__fadd_rn(__fadd_rn(fabsf(__fsub_rn(qcx, tb[0])), fabsf(__fsub_rn(qcy, tb[1]))),
                                             fabsf(__fsub_rn(qw, tb[2]))),
                                   fabsf(__fsub_rn(qh, tb[3])));
        const float g = giou_xyxy(qb, to_xyxy(tb[0], tb[1], tb[2], tb[3]));
        const float c_class = -dot;
        const float c_giou = -g;
        crow[t] = __fadd_rn(__fadd_rn(__fmul_rn(w_bbox, l1), __fmul_rn(w_class, c_class)), __fmul_rn(w_giou, c_giou));
      }
    }
    for (int t = T + lane; t < Tmax; t += 32) crow[t] = 0.f;
    __syncwarp();
  }
}

// One block per problem, thread 0 runs the (inherently sequential) augmenting-path search in float64.
// cost: [P][Q][Tmax] fp32.  match_q[p][t] = query assigned to target t (or -1), flags[0] |= 1 on NaN / infeasible.
__global__ void lsap_kernel(const float* __restrict__ cost, const int* __restrict__ tgt_count, int* __restrict__ match_q,
                            int* __restrict__ flags, int B, int Q, int Tmax) {
  pdl_prologue();
  const int p = blockIdx.x;
  const int b = p % B;
  const int T = min(tgt_count[b], Tmax);
  const float* c = cost + (size_t)p * Q * Tmax;
  int* mq = match_q + (size_t)p * Tmax;
  for (int t = threadIdx.x; t < Tmax; t += blockDim.x) mq[t] = -1;
  if (T == 0) return;
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < Q * T; i += blockDim.x) {
    const float v = c[(i / T) * Tmax + (i % T)];
    if (isnan(v) || v == -INFINITY) bad = 1;
  }
  __syncthreads();
  if (bad) {
    if (threadIdx.x == 0) atomicOr(flags, 1);
    return;
  }
  if (threadIdx.x != 0) return;
  extern __shared__ unsigned char lsap_smem[];
  const int nr = min(Q, T), nc = max(Q, T);
  double* u = reinterpret_cast<double*>(lsap_smem);
  double* v = u + nr;
  double* shortest = v + nc;
  int* path = reinterpret_cast<int*>(shortest + nc);
  int* col4row = path + nc;
  int* row4col = col4row + nr;
  int* remaining = row4col + nc;
  uint8_t* SR = reinterpret_cast<uint8_t*>(remaining + nc);
  uint8_t* SC = SR + nr;
  LsapWork w{u, v, shortest, path, col4row, row4col, remaining, SR, SC};
  int rc;
  if (T <= Q) {
    // scipy transposes when rows (queries) > cols (targets): rows of the solved problem are targets
    rc = lsap_solve(nr, nc, [c, Tmax](int t, int q) { return (double)c[q * Tmax + t]; }, w);
    if (rc == 0)
      for (int t = 0; t < T; ++t) mq[t] = col4row[t];
  } else {
    rc = lsap_solve(nr, nc, [c, Tmax](int q, int t) { return (double)c[q * Tmax + t]; }, w);
    if (rc == 0)
      for (int q = 0; q < Q; ++q) mq[col4row[q]] = q;
  }
  if (rc != 0) atomicOr(flags, 1);
}

// One warp per problem (lsap_warp.cuh: the column scan of every augmentation step runs on 32 lanes, same scan order,
// tie rule and float64 arithmetic as the sequential solver above).  Used when max(Q, T) <= kLsapMax.
__global__ void lsap_warp_kernel(const float* __restrict__ cost, const int* __restrict__ tgt_count,
                                 int* __restrict__ match_q, int* __restrict__ flags, int B, int Q, int Tmax) {
  pdl_prologue();
  __shared__ LsapShared w;
  const int p = blockIdx.x, lane = threadIdx.x;
  const int b = p % B;
  const int T = min(tgt_count[b], Tmax);
  const float* c = cost + (size_t)p * Q * Tmax;
  int* mq = match_q + (size_t)p * Tmax;
  for (int t = lane; t < Tmax; t += 32) mq[t] = -1;
  if (T == 0) return;
  int bad = 0;
  for (int i = lane; i < Q * T; i += 32) {
    const float v = c[(i / T) * Tmax + (i % T)];
    if (isnan(v) || v == -INFINITY) bad = 1;
  }
  if (__any_sync(0xffffffffu, bad)) {
    if (lane == 0) atomicOr(flags, 1);
    return;
  }
  __syncwarp();
  int rc;
  if (T <= Q) {  // scipy transposes when rows (queries) > cols (targets): rows of the solved problem are targets
    rc = lsap_warp(c, 1, Tmax, T, Q, w);
    if (rc == 0)
      for (int t = lane; t < T; t += 32) mq[t] = w.col4row[t];
  } else {
    rc = lsap_warp(c, Tmax, 1, Q, T, w);
    if (rc == 0)
      for (int q = lane; q < Q; q += 32) mq[w.col4row[q]] = q;
  }
  if (rc != 0 && lane == 0) atomicOr(flags, 1);
}

}  // namespace toist

using namespace toist;

extern "C" {

int toist_match_cost(const float* logits, const float* boxes, const float* tgt_boxes, const int32_t* tgt_count,
                     const float* posmap, float* cost, int32_t n_layers, int32_t batch, int32_t n_queries,
                     int32_t n_classes, int32_t t_max, float w_class, float w_bbox, float w_giou, void* stream) {
  TOIST_REQUIRE(logits && boxes && tgt_boxes && tgt_count && posmap && cost, "toist_match_cost: null pointer");
  TOIST_REQUIRE(n_layers > 0 && batch > 0 && n_queries > 0 && n_classes > 0 && t_max > 0, "toist_match_cost: bad sizes");
  const int threads = 128;
  const size_t smem = (size_t)(threads / 32) * n_classes * sizeof(float);
  launch_pdl(match_cost_kernel, dim3(n_layers * batch, (n_queries + threads / 32 - 1) / (threads / 32)), dim3(threads), smem, (cudaStream_t)stream, 
      logits, boxes, tgt_boxes, tgt_count, posmap, cost, batch, n_queries, n_classes, t_max, w_class, w_bbox, w_giou);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_lsap_device(const float* cost, const int32_t* tgt_count, int32_t* match_q, int32_t* flags,
                      int32_t n_problems, int32_t batch, int32_t n_queries, int32_t t_max, void* stream) {
  TOIST_REQUIRE(cost && tgt_count && match_q && flags, "toist_lsap_device: null pointer");
  TOIST_REQUIRE(n_problems > 0 && batch > 0 && n_queries > 0 && t_max > 0, "toist_lsap_device: bad sizes");
  const int nmax = n_queries > t_max ? n_queries : t_max;
  if (nmax <= kLsapMax) {
    launch_pdl(lsap_warp_kernel, dim3(n_problems), dim3(32), 0, (cudaStream_t)stream, cost, tgt_count, match_q, flags,
               batch, n_queries, t_max);
    TOIST_CHECK_CUDA(cudaGetLastError());
    return TOIST_OK;
  }
  const size_t smem = (size_t)nmax * (3 * sizeof(double) + 4 * sizeof(int) + 2) + 64;
  TOIST_REQUIRE(smem <= 200 * 1024, "toist_lsap_device: problem too large for shared memory (%zu bytes)", smem);
  static bool configured = false;
  if (!configured && smem > 48 * 1024) {
    TOIST_CHECK_CUDA(cudaFuncSetAttribute(lsap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  launch_pdl(lsap_kernel, dim3(n_problems), dim3(64), smem, (cudaStream_t)stream, cost, tgt_count, match_q, flags, batch, n_queries, t_max);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

// Host LSAP with scipy.optimize.linear_sum_assignment semantics (models/matcher.py:85).  cost is row-major
// [n_rows][n_cols] float64; writes min(n_rows, n_cols) pairs sorted by row.  Returns the pair count or a negative
// toist_status (TOIST_ERR_NUMERIC for NaN / -inf entries or an infeasible matrix, like scipy's ValueError).
int toist_lsap_f64(const double* cost, int32_t n_rows, int32_t n_cols, int64_t* row_ind, int64_t* col_ind) {
  TOIST_REQUIRE(n_rows >= 0 && n_cols >= 0, "toist_lsap_f64: negative shape");
  if (n_rows == 0 || n_cols == 0) return 0;
  TOIST_REQUIRE(cost && row_ind && col_ind, "toist_lsap_f64: null pointer");
  for (int64_t i = 0; i < (int64_t)n_rows * n_cols; ++i)
    if (cost[i] != cost[i] || cost[i] == -INFINITY)
      return set_error(TOIST_ERR_NUMERIC, "matrix contains invalid numeric entries");
  const bool transposed = n_rows > n_cols;
  const int nr = transposed ? n_cols : n_rows, nc = transposed ? n_rows : n_cols;
  std::vector<double> u(nr), v(nc), shortest(nc);
  std::vector<int> path(nc), col4row(nr), row4col(nc), remaining(nc);
  std::vector<uint8_t> SR(nr), SC(nc);
  LsapWork w{u.data(), v.data(), shortest.data(), path.data(), col4row.data(), row4col.data(), remaining.data(),
             SR.data(), SC.data()};
  int rc;
  if (transposed)
    rc = lsap_solve(nr, nc, [cost, n_cols](int i, int j) { return cost[(int64_t)j * n_cols + i]; }, w);
  else
    rc = lsap_solve(nr, nc, [cost, n_cols](int i, int j) { return cost[(int64_t)i * n_cols + j]; }, w);
  if (rc != 0) return set_error(TOIST_ERR_NUMERIC, "cost matrix is infeasible");
  if (!transposed) {
    for (int i = 0; i < nr; ++i) {
      row_ind[i] = i;
      col_ind[i] = col4row[i];
    }
  } else {
    // solved rows are original columns; emit pairs sorted by original row (= assigned column of the solved problem)
    std::vector<int> order(nr);
    for (int i = 0; i < nr; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return col4row[a] < col4row[b]; });
    for (int k = 0; k < nr; ++k) {
      row_ind[k] = col4row[order[k]];
      col_ind[k] = order[k];
    }
  }
  return nr;
}

}  // extern "C"
