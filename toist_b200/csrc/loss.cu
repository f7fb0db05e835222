// SetCriterion on the device, all decoder layers per launch, forward value + "unit" gradients (d loss_l / d input
// for an upstream gradient of 1; the autograd wrapper scales them by the real upstream gradient per layer).
// Reference: models/mdetr.py:488-518 (loss_labels), :805-825 (loss_boxes), :783-803 (loss_cardinality),
// :601-666 (loss_contrastive_align); util/box_ops.py:11-61.
#include <math.h>

#include "common.cuh"
#include "host_util.h"

namespace toist {

__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < nw; ++i) t += red[i];  // fixed order: deterministic
  return t;
}

// ---------------------------------------------------------------------------------------------- soft-token CE
// grid = L*B*Q rows / 4 (one warp per row).  row_loss[l,b,q] = w * -(logp . target); unit grad into dlogits.
__global__ void token_ce_kernel(const float* __restrict__ logits, const int* __restrict__ match_q,
                                const int* __restrict__ tgt_count, const float* __restrict__ posmap,
                                const float* __restrict__ num_boxes, float* __restrict__ row_loss,
                                float* __restrict__ dlogits, int rows, int B, int Q, int C, int Tmax, float eos_coef) {
  pdl_prologue();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int q = row % Q;
  const int lb = row / Q;
  const int b = lb % B;
  const int T = min(tgt_count[b], Tmax);
  int t_match = -1;
  const int* mq = match_q + (size_t)lb * Tmax;
  for (int t = 0; t < T; ++t)
    if (mq[t] == q) t_match = t;
  const float* lg = logits + (size_t)row * C;
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, lg[c]);
  mx = warp_max(mx);
  float se = 0.f;
  for (int c = lane; c < C; c += 32) se += expf(lg[c] - mx);
  se = warp_sum(se);
  const float lse = mx + logf(se);
  const float* tgt = t_match >= 0 ? posmap + ((size_t)b * Tmax + t_match) * C : nullptr;
  float dot = 0.f, tsum = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float tv = tgt ? tgt[c] : (c == C - 1 ? 1.f : 0.f);
    dot += (lg[c] - lse) * tv;
    tsum += tv;
  }
  dot = warp_sum(dot);
  tsum = warp_sum(tsum);
  const float w = t_match >= 0 ? 1.f : eos_coef;
  if (lane == 0) row_loss[row] = -dot * w;
  if (dlogits != nullptr) {
    const float s = w / num_boxes[0];
    float* dl = dlogits + (size_t)row * C;
    for (int c = lane; c < C; c += 32) {
      const float tv = tgt ? tgt[c] : (c == C - 1 ? 1.f : 0.f);
      dl[c] = s * (expf(lg[c] - lse) * tsum - tv);
    }
  }
}

// ---------------------------------------------------------------------------------------------- box losses
struct GiouGrad {
  float loss, dcx, dcy, dw, dh;
};
__device__ GiouGrad giou_loss_grad(float cx, float cy, float w, float h, float tcx, float tcy, float tw, float th) {
  const float x0 = cx - 0.5f * w, x1 = cx + 0.5f * w, y0 = cy - 0.5f * h, y1 = cy + 0.5f * h;
  const float X0 = tcx - 0.5f * tw, X1 = tcx + 0.5f * tw, Y0 = tcy - 0.5f * th, Y1 = tcy + 0.5f * th;
  const float area_p = (x1 - x0) * (y1 - y0), area_t = (X1 - X0) * (Y1 - Y0);
  const float iw_raw = fminf(x1, X1) - fmaxf(x0, X0), ih_raw = fminf(y1, Y1) - fmaxf(y0, Y0);
  const float iw = fmaxf(iw_raw, 0.f), ih = fmaxf(ih_raw, 0.f);
  const float inter = iw * ih;
  const float uni = area_p + area_t - inter;
  const float iou = inter / uni;
  const float hw_raw = fmaxf(x1, X1) - fminf(x0, X0), hh_raw = fmaxf(y1, Y1) - fminf(y0, Y0);
  const float hw = fmaxf(hw_raw, 0.f), hh = fmaxf(hh_raw, 0.f);
  const float hull = hw * hh;
  const float giou = iou - (hull - uni) / hull;
  GiouGrad g;
  g.loss = 1.f - giou;
  // L = 1 - inter/uni + (hull - uni)/hull
  const float gU = inter / (uni * uni) - 1.f / hull;            // dL/d uni (total)
  const float gI = -1.f / uni - gU;                              // dL/d inter (uni = ap + at - inter)
  const float gH = uni / (hull * hull);                          // dL/d hull
  const float gAp = gU;                                          // dL/d area_p
  const float g_iw = (iw_raw >= 0.f) ? gI * ih : 0.f, g_ih = (ih_raw >= 0.f) ? gI * iw : 0.f;
  const float g_hw = (hw_raw >= 0.f) ? gH * hh : 0.f, g_hh = (hh_raw >= 0.f) ? gH * hw : 0.f;
  // binary min/max split the gradient evenly on exact ties, like torch.min/torch.max
  auto sel_le = [](float a, float b) { return a < b ? 1.f : (a == b ? 0.5f : 0.f); };
  const float dx1 = g_iw * sel_le(x1, X1) + g_hw * sel_le(X1, x1) + gAp * (y1 - y0);
  const float dx0 = -g_iw * sel_le(X0, x0) - g_hw * sel_le(x0, X0) - gAp * (y1 - y0);
  const float dy1 = g_ih * sel_le(y1, Y1) + g_hh * sel_le(Y1, y1) + gAp * (x1 - x0);
  const float dy0 = -g_ih * sel_le(Y0, y0) - g_hh * sel_le(y0, Y0) - gAp * (x1 - x0);
  g.dcx = dx0 + dx1;
  g.dcy = dy0 + dy1;
  g.dw = 0.5f * (dx1 - dx0);
  g.dh = 0.5f * (dy1 - dy0);
  return g;
}

// grid = L*B, block = 32*k.  pair_loss[l,b,t] = {l1, giou}; dboxes (pre-zeroed by this kernel for its rows).
__global__ void box_loss_kernel(const float* __restrict__ boxes, const int* __restrict__ match_q,
                                const int* __restrict__ tgt_count, const float* __restrict__ tgt_boxes,
                                const float* __restrict__ num_boxes, float* __restrict__ pair_l1,
                                float* __restrict__ pair_giou, float* __restrict__ dboxes_l1,
                                float* __restrict__ dboxes_giou, int B, int Q, int Tmax) {
  pdl_prologue();
  const int lb = blockIdx.x;
  const int b = lb % B;
  const int T = min(tgt_count[b], Tmax);
  if (dboxes_l1 != nullptr)
    for (int i = threadIdx.x; i < Q * 4; i += blockDim.x) {
      dboxes_l1[(size_t)lb * Q * 4 + i] = 0.f;
      dboxes_giou[(size_t)lb * Q * 4 + i] = 0.f;
    }
  __syncthreads();
  const float inv_nb = 1.f / num_boxes[0];
  for (int t = threadIdx.x; t < Tmax; t += blockDim.x) {
    float l1 = 0.f, gl = 0.f;
    const int q = t < T ? match_q[(size_t)lb * Tmax + t] : -1;
    if (q >= 0) {
      const float* p = boxes + ((size_t)lb * Q + q) * 4;
      const float* tb = tgt_boxes + ((size_t)b * Tmax + t) * 4;
      l1 = fabsf(p[0] - tb[0]) + fabsf(p[1] - tb[1]) + fabsf(p[2] - tb[2]) + fabsf(p[3] - tb[3]);
      const GiouGrad g = giou_loss_grad(p[0], p[1], p[2], p[3], tb[0], tb[1], tb[2], tb[3]);
      gl = g.loss;
      if (dboxes_l1 != nullptr) {
        float* d1 = dboxes_l1 + ((size_t)lb * Q + q) * 4;
        float* d2 = dboxes_giou + ((size_t)lb * Q + q) * 4;
        for (int k = 0; k < 4; ++k) {
          const float df = p[k] - tb[k];
          d1[k] = (df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f)) * inv_nb;
        }
        d2[0] = g.dcx * inv_nb;
        d2[1] = g.dcy * inv_nb;
        d2[2] = g.dw * inv_nb;
        d2[3] = g.dh * inv_nb;
      }
    }
    pair_l1[(size_t)lb * Tmax + t] = l1;
    pair_giou[(size_t)lb * Tmax + t] = gl;
  }
}

// ---------------------------------------------------------------------------------------------- cardinality
// grid = L*B, one warp per query row; card[l,b] = #(argmax != C-1)
__global__ void cardinality_kernel(const float* __restrict__ logits, int* __restrict__ card, int Q, int C) {
  pdl_prologue();
  const int lb = blockIdx.x;
  __shared__ int cnt;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int q = warp; q < Q; q += nw) {
    const float* lg = logits + ((size_t)lb * Q + q) * C;
    float best = -INFINITY;
    int bi = 0;
    for (int c = lane; c < C; c += 32) {
      const float v = lg[c];
      if (v > best) {
        best = v;
        bi = c;
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) {  // first maximum wins, like torch.argmax
        best = ov;
        bi = oi;
      }
    }
    if (lane == 0 && bi != C - 1) atomicAdd(&cnt, 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) card[lb] = cnt;
}

// ---------------------------------------------------------------------------------------------- contrastive align
// grid = L*B, block = 128.  logits[q,k] = pq[q,:] . pt[k,:] / temperature  (Q x K, K = text tokens <= 64).
// row_loss2[l,b] = (box->token + token->box) / 2 for this image; unit grads dpq [L,B,Q,D], dpt [L,B,K,D].
__global__ void contrastive_kernel(const float* __restrict__ pq, const float* __restrict__ pt,
                                   const int* __restrict__ match_q, const int* __restrict__ tgt_count,
                                   const uint8_t* __restrict__ tok_pos, const float* __restrict__ num_boxes,
                                   float* __restrict__ img_loss, float* __restrict__ dpq, float* __restrict__ dpt,
                                   int B, int Q, int K, int D, int Tmax, float inv_temp) {
  pdl_prologue();
  extern __shared__ float sm[];
  float* lg = sm;                                        // [Q][K]
  float* dl = lg + Q * K;                                // [Q][K]
  int* q2t = reinterpret_cast<int*>(dl + Q * K);         // [Q]
  float* red = reinterpret_cast<float*>(q2t + Q);        // [32]
  float* col_lse = red + 32;                             // [K]
  float* col_np = col_lse + K;                           // [K]
  const int lb = blockIdx.x;
  const int b = lb % B;
  const int T = min(tgt_count[b], Tmax);
  const float* q_ = pq + (size_t)lb * Q * D;
  const float* t_ = pt + (size_t)b * K * D;
  for (int q = threadIdx.x; q < Q; q += blockDim.x) q2t[q] = -1;
  __syncthreads();
  if (threadIdx.x == 0)
    for (int t = 0; t < T; ++t) {
      const int q = match_q[(size_t)lb * Tmax + t];
      if (q >= 0) q2t[q] = t;
    }
  for (int i = threadIdx.x; i < Q * K; i += blockDim.x) {
    const int q = i / K, k = i % K;
    float acc = 0.f;
    for (int d = 0; d < D; ++d) acc += q_[q * D + d] * t_[k * D + d];
    lg[i] = acc * inv_temp;
    dl[i] = 0.f;
  }
  __syncthreads();
  const uint8_t* tp = tok_pos + (size_t)b * Tmax * K;
  float part = 0.f;
  // box -> token (rows)
  for (int q = threadIdx.x; q < Q; q += blockDim.x) {
    const int t = q2t[q];
    if (t < 0) continue;
    int np = 0;
    float pos = 0.f, mx = -INFINITY;
    for (int k = 0; k < K; ++k) {
      mx = fmaxf(mx, lg[q * K + k]);
      if (tp[t * K + k]) {
        ++np;
        pos -= lg[q * K + k];
      }
    }
    if (np == 0) continue;
    float se = 0.f;
    for (int k = 0; k < K; ++k) se += expf(lg[q * K + k] - mx);
    const float lse = mx + logf(se);
    const float inv_np = 1.f / ((float)np + 1e-6f);
    part += pos * inv_np + lse;
    for (int k = 0; k < K; ++k)
      dl[q * K + k] += expf(lg[q * K + k] - lse) - (tp[t * K + k] ? inv_np : 0.f);
  }
  __syncthreads();
  // token -> box (columns)
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    int np = 0;
    float pos = 0.f, mx = -INFINITY;
    for (int q = 0; q < Q; ++q) {
      mx = fmaxf(mx, lg[q * K + k]);
      const int t = q2t[q];
      if (t >= 0 && tp[t * K + k]) {
        ++np;
        pos -= lg[q * K + k];
      }
    }
    col_np[k] = (float)np;
    if (np == 0) continue;
    float se = 0.f;
    for (int q = 0; q < Q; ++q) se += expf(lg[q * K + k] - mx);
    const float lse = mx + logf(se);
    col_lse[k] = lse;
    part += pos / ((float)np + 1e-6f) + lse;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Q * K; i += blockDim.x) {
    const int q = i / K, k = i % K;
    if (col_np[k] > 0.f) {
      const int t = q2t[q];
      const bool pm = t >= 0 && tp[t * K + k];
      dl[i] += expf(lg[i] - col_lse[k]) - (pm ? 1.f / (col_np[k] + 1e-6f) : 0.f);
    }
  }
  const float tot = block_sum(part, red);
  if (threadIdx.x == 0) img_loss[lb] = 0.5f * tot;
  __syncthreads();
  if (dpq == nullptr) return;
  const float s = 0.5f * inv_temp / num_boxes[0];
  float* dq_ = dpq + (size_t)lb * Q * D;
  float* dt_ = dpt + (size_t)lb * K * D;
  for (int i = threadIdx.x; i < Q * D; i += blockDim.x) {
    const int q = i / D, d = i % D;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc += dl[q * K + k] * t_[k * D + d];
    dq_[i] = acc * s;
  }
  for (int i = threadIdx.x; i < K * D; i += blockDim.x) {
    const int k = i / D, d = i % D;
    float acc = 0.f;
    for (int q = 0; q < Q; ++q) acc += dl[q * K + k] * q_[q * D + d];
    dt_[i] = acc * s;
  }
}

// ---------------------------------------------------------------------------------------------- final reduction
// grid = L, block = 256.  out[term][l]: 0 ce, 1 bbox, 2 giou, 3 cardinality error, 4 contrastive align.
__global__ void criterion_reduce_kernel(const float* __restrict__ row_loss, const float* __restrict__ pair_l1,
                                        const float* __restrict__ pair_giou, const int* __restrict__ card,
                                        const float* __restrict__ img_loss, const int* __restrict__ tgt_count,
                                        const float* __restrict__ num_boxes, const int* __restrict__ flags,
                                        float* __restrict__ out, int L, int B, int Q, int Tmax) {
  pdl_prologue();
  __shared__ float red[32];
  const int l = blockIdx.x;
  const float inv_nb = 1.f / num_boxes[0];
  float a = 0.f;
  for (int i = threadIdx.x; i < B * Q; i += blockDim.x) a += row_loss[(size_t)l * B * Q + i];
  a = block_sum(a, red);
  float l1 = 0.f, gi = 0.f;
  for (int i = threadIdx.x; i < B * Tmax; i += blockDim.x) {
    l1 += pair_l1[(size_t)l * B * Tmax + i];
    gi += pair_giou[(size_t)l * B * Tmax + i];
  }
  l1 = block_sum(l1, red);
  gi = block_sum(gi, red);
  float ce = 0.f, ca = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    ce += fabsf((float)card[l * B + b] - (float)min(tgt_count[b], Tmax));
    if (img_loss != nullptr) ca += img_loss[l * B + b];
  }
  ce = block_sum(ce, red);
  ca = block_sum(ca, red);
  if (threadIdx.x == 0) {
    // an invalid matching cost (scipy's ValueError in the reference) poisons every term: the caller's
    // non-finite-loss guard (engine.py:82-85) then aborts without a device synchronisation here
    if (flags != nullptr && flags[0] != 0) a = __int_as_float(0x7fc00000);
    out[0 * L + l] = a * inv_nb;
    out[1 * L + l] = l1 * inv_nb;
    out[2 * L + l] = gi * inv_nb;
    out[3 * L + l] = ce / (float)B;
    out[4 * L + l] = ca * inv_nb;
  }
}

// y[l, i] = x[l, i] * g[l]            (reduce == 0)
// y[i]    = sum_l x[l, i] * g[l]      (reduce == 1)
__global__ void scale_layers_kernel(const float* __restrict__ x, const float* __restrict__ g, float* __restrict__ y,
                                    int L, long long n, int reduce) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (reduce) {
    float acc = 0.f;
    for (int l = 0; l < L; ++l) acc += x[(size_t)l * n + i] * g[l];
    y[i] = acc;
  } else {
    for (int l = 0; l < L; ++l) y[(size_t)l * n + i] = x[(size_t)l * n + i] * g[l];
  }
}

// y[l, i] = x1[l, i] * g1[l] + x2[l, i] * g2[l]
__global__ void scale_layers2_kernel(const float* __restrict__ x1, const float* __restrict__ g1,
                                     const float* __restrict__ x2, const float* __restrict__ g2, float* __restrict__ y,
                                     int L, long long n) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int l = 0; l < L; ++l) y[(size_t)l * n + i] = x1[(size_t)l * n + i] * g1[l] + x2[(size_t)l * n + i] * g2[l];
}

}  // namespace toist

using namespace toist;

extern "C" {

int toist_token_ce(const float* logits, const int32_t* match_q, const int32_t* tgt_count, const float* posmap,
                   const float* num_boxes, float* row_loss, float* dlogits, int32_t n_layers, int32_t batch,
                   int32_t n_queries, int32_t n_classes, int32_t t_max, float eos_coef, void* stream) {
  TOIST_REQUIRE(logits && match_q && tgt_count && posmap && num_boxes && row_loss, "toist_token_ce: null pointer");
  const int rows = n_layers * batch * n_queries;
  launch_pdl(token_ce_kernel, dim3((rows + 3) / 4), dim3(128), 0, (cudaStream_t)stream, logits, match_q, tgt_count, posmap, num_boxes,
                                                                     row_loss, dlogits, rows, batch, n_queries,
                                                                     n_classes, t_max, eos_coef);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_box_loss(const float* boxes, const int32_t* match_q, const int32_t* tgt_count, const float* tgt_boxes,
                   const float* num_boxes, float* pair_l1, float* pair_giou, float* dboxes_l1, float* dboxes_giou,
                   int32_t n_layers, int32_t batch, int32_t n_queries, int32_t t_max, void* stream) {
  TOIST_REQUIRE(boxes && match_q && tgt_count && tgt_boxes && num_boxes && pair_l1 && pair_giou,
                "toist_box_loss: null pointer");
  TOIST_REQUIRE((dboxes_l1 == nullptr) == (dboxes_giou == nullptr), "toist_box_loss: pass both gradient buffers");
  launch_pdl(box_loss_kernel, dim3(n_layers * batch), dim3(64), 0, (cudaStream_t)stream, boxes, match_q, tgt_count, tgt_boxes, num_boxes,
                                                                     pair_l1, pair_giou, dboxes_l1, dboxes_giou, batch,
                                                                     n_queries, t_max);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_cardinality(const float* logits, int32_t* card, int32_t n_layers, int32_t batch, int32_t n_queries,
                      int32_t n_classes, void* stream) {
  TOIST_REQUIRE(logits && card, "toist_cardinality: null pointer");
  launch_pdl(cardinality_kernel, dim3(n_layers * batch), dim3(1024), 0, (cudaStream_t)stream, logits, card, n_queries, n_classes);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_contrastive_align(const float* proj_queries, const float* proj_tokens, const int32_t* match_q,
                            const int32_t* tgt_count, const uint8_t* tok_pos, const float* num_boxes, float* img_loss,
                            float* dpq, float* dpt, int32_t n_layers, int32_t batch, int32_t n_queries,
                            int32_t n_tokens, int32_t dim, int32_t t_max, float temperature, void* stream) {
  TOIST_REQUIRE(proj_queries && proj_tokens && match_q && tgt_count && tok_pos && num_boxes && img_loss,
                "toist_contrastive_align: null pointer");
  TOIST_REQUIRE((dpq == nullptr) == (dpt == nullptr), "toist_contrastive_align: pass both gradient buffers");
  const size_t smem = (size_t)(2 * n_queries * n_tokens + n_queries + 32 + 2 * n_tokens) * 4;
  TOIST_REQUIRE(smem <= 200 * 1024, "toist_contrastive_align: Q*K too large (%zu bytes of shared memory)", smem);
  static bool configured = false;
  if (!configured) {
    TOIST_CHECK_CUDA(cudaFuncSetAttribute(contrastive_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  launch_pdl(contrastive_kernel, dim3(n_layers * batch), dim3(512), smem, (cudaStream_t)stream, 
      proj_queries, proj_tokens, match_q, tgt_count, tok_pos, num_boxes, img_loss, dpq, dpt, batch, n_queries,
      n_tokens, dim, t_max, 1.f / temperature);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_criterion_reduce(const float* row_loss, const float* pair_l1, const float* pair_giou, const int32_t* card,
                           const float* img_loss, const int32_t* tgt_count, const float* num_boxes,
                           const int32_t* flags, float* out, int32_t n_layers, int32_t batch, int32_t n_queries,
                           int32_t t_max, void* stream) {
  TOIST_REQUIRE(row_loss && pair_l1 && pair_giou && card && tgt_count && num_boxes && out,
                "toist_criterion_reduce: null pointer");
  launch_pdl(criterion_reduce_kernel, dim3(n_layers), dim3(256), 0, (cudaStream_t)stream, row_loss, pair_l1, pair_giou, card, img_loss,
                                                                      tgt_count, num_boxes, flags, out, n_layers,
                                                                      batch, n_queries, t_max);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_scale_layers2(const float* x1, const float* g1, const float* x2, const float* g2, float* y, int32_t n_layers,
                        int64_t n, void* stream) {
  TOIST_REQUIRE(x1 && g1 && x2 && g2 && y, "toist_scale_layers2: null pointer");
  if (n == 0) return TOIST_OK;
  launch_pdl(scale_layers2_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, x1, g1, x2, g2, y, n_layers, n);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

int toist_scale_layers(const float* x, const float* g, float* y, int32_t n_layers, int64_t n, int32_t reduce,
                       void* stream) {
  TOIST_REQUIRE(x && g && y, "toist_scale_layers: null pointer");
  if (n == 0) return TOIST_OK;
  launch_pdl(scale_layers_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, x, g, y, n_layers, n, reduce);
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

}  // extern "C"
