/*
 * toist_b200 — C ABI of the B200-native (sm_100a) kernels behind the TOIST / MDETR training hot path.
 *
 * The reference (AIR-DISCOVER/TOIST) is pure Python/PyTorch and has no FFI of its own; every entry point below
 * replaces a *library call site* inside the reference's hot path (cited per function as reference file:line).
 * The reference-side binding a maintainer would add is the ctypes stub in INTEGRATION.md (toist_b200/_lib.py is
 * the shipped copy of it).
 *
 * Conventions
 *   - plain C, POD structs of raw device pointers + sizes; no torch / C++ types cross this boundary
 *   - the caller allocates every output and workspace; kernels never allocate and never synchronise
 *   - every launch is asynchronous on the cudaStream_t passed as `stream` (a void* here)
 *   - return value: 0 on success, negative toist_status otherwise; toist_last_error() gives the thread-local text
 *   - bf16 tensors are raw uint16 storage (torch.bfloat16); "f32" is IEEE binary32
 */
#ifndef TOIST_B200_H_
#define TOIST_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TOIST_ABI_VERSION 1

typedef enum toist_status {
  TOIST_OK = 0,
  TOIST_ERR_INVALID = -1, /* bad argument (shape, alignment, null pointer) */
  TOIST_ERR_CUDA = -2,    /* a CUDA runtime / driver call failed */
  TOIST_ERR_UNSUPPORTED = -3,
  TOIST_ERR_NUMERIC = -4 /* NaN / infeasible cost matrix (mirrors scipy's ValueError) */
} toist_status;

int toist_abi_version(void);
const char* toist_last_error(void);
/* sizeof(toist_gemm_desc) as compiled, so a foreign-language binding can verify its struct layout */
size_t toist_sizeof_gemm_desc(void);
/* 1 when the current device is compute capability 10.x (tcgen05 / TMEM present) */
int toist_device_ok(void);
/* Measurement hook used by bench.py's roofline pass only: while `on` is non-zero toist_gemm returns immediately
 * without launching, so (step time) - (step time without GEMMs) is the in-situ time of the tensor-core kernels.
 * Returns the previous setting.  Results computed while it is set are garbage by construction. */
int toist_debug_skip_gemm(int on);
/* Measurement hook (tools/gemm_trace.py): while `buf` is non-null every toist_gemm CTA writes 8 clock64() stamps
 * (entry, setup done, dependency wait done, first operands landed, last MMA issued, accumulator complete, epilogue
 * tile loop done, store done) and 2 %globaltimer stamps (entry, exit; nanoseconds) to buf[16 * cta_linear_index ..];
 * pass NULL to switch it off.  The buffer is the caller's (device memory, >= 128 bytes per CTA of the largest grid). */
int toist_debug_gemm_trace(void* buf);

/* ------------------------------------------------------------------------------------------------------------
 * Implicit-GEMM engine (tcgen05.mma + TMEM accumulators, operands staged by TMA, SWIZZLE_128B).
 *
 * One kernel family serves: nn.Linear fwd/dgrad/wgrad (reference models/transformer.py:275-277,301,342-344,405,
 * 483; models/mdetr.py:347-357,1031; transformers RobertaModel), nn.MultiheadAttention projections and the
 * batched QK^T / PV products (transformer.py:273,298,337-338,378,394), the backbone / input_proj / mask-head
 * convolutions with FrozenBatchNorm + ReLU + residual folded into the epilogue (models/backbone.py:48-58,75;
 * models/mdetr.py:351; models/segmentation.py:178-195), and their data / weight gradients.
 *
 * Both operands are bf16 tensors described as 4-D views (dim[0] innermost, stride[0] == 1, all other strides
 * multiples of 8 elements, base 16-byte aligned).  The M side of the product is a *pixel tile*
 * (tile_x * tile_y * tile_n == 128 rows) so that a 3x3 tap is just a shifted TMA box with hardware zero fill.
 *
 *   mode FWD   : D[pix, n] = sum_tap sum_c A[c, x*sx+dx, y*sy+dy, n_img+dn] * B[tap.col + c, n, (y), (n_img)]
 *                A K-major (channels innermost), B K-major.
 *   mode DGRAD : same A traversal, B is MN-major: B[tap.col + n, c, (y), (n_img)]   (reduction index c on dim 1)
 *   mode WGRAD : D[m, tap.col + n] (+)= sum_pix A[m, pix] * B[n, pix*s + tap]  (both operands MN-major,
 *                reduction over pixel tiles of 64, optional split over grid.z with fp32 atomics)
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct toist_tensor4 {
  const void* ptr;
  int64_t dim[4];    /* extents in elements, dim[0] innermost */
  int64_t stride[4]; /* strides in elements, stride[0] must be 1 */
} toist_tensor4;

typedef struct toist_tap {
  int16_t dx, dy, dn; /* coordinate offsets added to the pixel-tiled operand */
  int16_t pad_;
  int32_t col; /* FWD/DGRAD: column offset into B dim0; WGRAD: column offset into the output */
} toist_tap;

enum { TOIST_GEMM_FWD = 0, TOIST_GEMM_DGRAD = 1, TOIST_GEMM_WGRAD = 2 };
enum { TOIST_ACT_NONE = 0, TOIST_ACT_RELU = 1, TOIST_ACT_GELU = 2, TOIST_ACT_SIGMOID = 3 };
enum { TOIST_BF16 = 0, TOIST_F32 = 1 };
#define TOIST_MAX_TAPS 49

typedef struct toist_gemm_desc {
  int32_t mode;
  toist_tensor4 a, b;
  /* pixel space of the M side (FWD/DGRAD) or of the reduction (WGRAD) */
  int32_t ext_x, ext_y, ext_n;    /* logical extents iterated by tiles */
  int32_t tile_x, tile_y, tile_n; /* product 128 (FWD/DGRAD) or 64 (WGRAD) */
  int32_t stride_x, stride_y;     /* element stride of the pixel-tiled operand (conv stride), 1 or 2 */
  int32_t n_cols;                 /* N extent of the product (FWD/DGRAD: output columns; WGRAD: B rows used) */
  int32_t m_rows;                 /* WGRAD only: M extent (rows of D) */
  int32_t k_per_tap;              /* FWD/DGRAD: reduction length per tap (channels) */
  int32_t n_taps;
  toist_tap taps[TOIST_MAX_TAPS];
  int32_t b_batched; /* FWD/DGRAD: B dims 2,3 indexed by the tile's (y, n_img); WGRAD: unused */
  int32_t batch_y, batch_n; /* WGRAD: extra batch grid; A/B dims 2,3 and output offset by (by, bn) */
  int32_t splits;    /* WGRAD: split the pixel reduction over this many CTAs (atomic fp32 accumulate) */
  /* epilogue: v = acc * alpha; v = v * col_scale[n] + col_shift[n]; v *= row_scale[m]; v += res; v *= (mask > 0);
   *           aux = v (pre-activation, bf16); v = act(v); out (=|+=) v                                          */
  void* out;
  int32_t out_dtype; /* TOIST_BF16 | TOIST_F32 */
  int64_t out_sx, out_sy, out_sn; /* element strides of the output along pixel x / y / n (WGRAD: sx = row stride,
                                     sy / sn = batch strides); columns are contiguous */
  float alpha;
  const float* col_scale;
  const float* col_shift;
  const float* row_scale;
  const void* res; /* same addressing as out */
  int32_t res_dtype;
  const void* mask; /* bf16, same addressing as out */
  void* aux;        /* bf16, same addressing as out */
  int32_t act;
  int32_t accumulate; /* 1: out += v using fp32 atomics (out must be f32) */
} toist_gemm_desc;

/* Launches the product described by `d` on `stream`. */
int toist_gemm(const toist_gemm_desc* d, void* stream);
/* Scheduler workspace of the persistent variant of the engine (FWD / DGRAD with bf16 output): a ZEROED device buffer
 * of the current device, owned by the caller for as long as launches may run (8 bytes per launch slot, handed out
 * round robin; 512 KB = 65536 launches in flight / captured in graphs at once).  Every launch re-arms its slot before
 * it ends.  Without a workspace (or with NULL) every launch uses the one-tile-per-CTA kernel.  The library never
 * allocates device memory itself (safe under CUDA-graph capture). */
int toist_gemm_set_workspace(void* zeroed, int64_t bytes);
/* Tile-shape hint: the caller issues `chains` independent launch sequences side by side (e.g. the two half-batch
 * chains of the trunk, runtime.backbone_fwd), so a launch should aim at 1 / chains of the SMs when it chooses its
 * column-tile width.  Host-side state read at launch time; returns the previous value.  1 = default. */
int toist_gemm_concurrency(int chains);


/* ------------------------------------------------------------------------------------------------------------
 * HungarianMatcher (reference models/matcher.py:39-87, util/box_ops.py:11-61)
 *   targets are passed padded: tgt_boxes [B, t_max, 4] cxcywh, tgt_count [B], posmap [B, t_max, n_classes]
 *   logits [L, B, Q, n_classes], boxes [L, B, Q, 4]  ->  cost [L, B, Q, t_max]  (fp32, reference operation order)
 * ------------------------------------------------------------------------------------------------------------ */
int toist_match_cost(const float* logits, const float* boxes, const float* tgt_boxes, const int32_t* tgt_count,
                     const float* posmap, float* cost, int32_t n_layers, int32_t batch, int32_t n_queries,
                     int32_t n_classes, int32_t t_max, float w_class, float w_bbox, float w_giou, void* stream);
/* One shortest-augmenting-path solve per (layer, image) problem on the device (float64, scipy tie rules);
 * match_q [n_problems, t_max] = query assigned to each target (-1 = none); flags[0] |= 1 on NaN / infeasible. */
int toist_lsap_device(const float* cost, const int32_t* tgt_count, int32_t* match_q, int32_t* flags,
                      int32_t n_problems, int32_t batch, int32_t n_queries, int32_t t_max, void* stream);
/* Host LSAP replacing scipy.optimize.linear_sum_assignment (matcher.py:85, mdetr.py:100,539): row-major float64
 * cost [n_rows, n_cols]; writes min(n_rows, n_cols) pairs sorted by row; returns the pair count or a negative
 * status (TOIST_ERR_NUMERIC for NaN / -inf entries or an infeasible matrix, scipy's ValueError). */
int toist_lsap_f64(const double* cost, int32_t n_rows, int32_t n_cols, int64_t* row_ind, int64_t* col_ind);

/* ------------------------------------------------------------------------------------------------------------
 * SetCriterion terms for all decoder layers at once (reference models/mdetr.py:488-518, 601-666, 783-825).
 * Every kernel also emits the gradient for an upstream gradient of one ("unit" gradients, already divided by
 * num_boxes[0]); pass null gradient pointers for inference.
 * ------------------------------------------------------------------------------------------------------------ */
int toist_token_ce(const float* logits, const int32_t* match_q, const int32_t* tgt_count, const float* posmap,
                   const float* num_boxes, float* row_loss, float* dlogits, int32_t n_layers, int32_t batch,
                   int32_t n_queries, int32_t n_classes, int32_t t_max, float eos_coef, void* stream);
int toist_box_loss(const float* boxes, const int32_t* match_q, const int32_t* tgt_count, const float* tgt_boxes,
                   const float* num_boxes, float* pair_l1, float* pair_giou, float* dboxes_l1, float* dboxes_giou,
                   int32_t n_layers, int32_t batch, int32_t n_queries, int32_t t_max, void* stream);
int toist_cardinality(const float* logits, int32_t* card, int32_t n_layers, int32_t batch, int32_t n_queries,
                      int32_t n_classes, void* stream);
int toist_contrastive_align(const float* proj_queries, const float* proj_tokens, const int32_t* match_q,
                            const int32_t* tgt_count, const uint8_t* tok_pos, const float* num_boxes, float* img_loss,
                            float* dpq, float* dpt, int32_t n_layers, int32_t batch, int32_t n_queries,
                            int32_t n_tokens, int32_t dim, int32_t t_max, float temperature, void* stream);
/* out [5, L]: loss_ce, loss_bbox, loss_giou, cardinality_error, loss_contrastive_align (img_loss may be null).
 * flags (may be null): the word written by toist_lsap_device; when non-zero loss_ce becomes NaN so that the caller's
 * non-finite-loss guard (engine.py:82-85) fires, mirroring scipy's ValueError without a device synchronisation. */
int toist_criterion_reduce(const float* row_loss, const float* pair_l1, const float* pair_giou, const int32_t* card,
                           const float* img_loss, const int32_t* tgt_count, const float* num_boxes,
                           const int32_t* flags, float* out, int32_t n_layers, int32_t batch, int32_t n_queries,
                           int32_t t_max, void* stream);
/* y[l, i] = x1[l, i] * g1[l] + x2[l, i] * g2[l] */
int toist_scale_layers2(const float* x1, const float* g1, const float* x2, const float* g2, float* y, int32_t n_layers,
                        int64_t n, void* stream);
/* reduce == 0: y[l, i] = x[l, i] * g[l];  reduce == 1: y[i] = sum_l x[l, i] * g[l] */
int toist_scale_layers(const float* x, const float* g, float* y, int32_t n_layers, int64_t n, int32_t reduce,
                       void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * HBM-bound helpers
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct toist_prep_item {
  const float* src;       /* fp32 master weight, [rows, cols] row-major */
  void* dst;              /* bf16 shadow, leading dimension ldd */
  const float* row_scale; /* optional FrozenBatchNorm scale folded per output channel (backbone.py:48-58) */
  int32_t rows, cols, ldd;
  int32_t first_block; /* prefix sum of ceil(rows * cols / 2048) over the preceding items */
  int32_t taps;        /* > 1: src is a conv weight [rows][cols/taps][taps] (OIHW), dst gets [rows][taps][cols/taps] */
  int32_t pad_;
} toist_prep_item;
int toist_weight_prep(const void* items_dev, int32_t n_items, int32_t total_blocks, void* stream);
int toist_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream);
int toist_cast_bf16_f32(const void* src, float* dst, int64_t n, void* stream);
/* nn.Dropout: out[i] = keep(i) ? x[i] / (1 - p) : 0 (+ res[i]); keep(i) is a counter-based hash of (seed[0] on the
 * device, site, i), so the backward pass re-derives the mask by calling this again on the upstream gradient.
 * dtype TOIST_BF16 | TOIST_F32 for x, res and out; out may alias x. */
int toist_dropout(const void* x, const void* res, void* out, int64_t n, int32_t dtype, float p, const uint64_t* seed,
                  uint32_t site, void* stream);
/* dst[r, 0:ld] = bf16(src[r, 0:n]), zero padded (rows narrower than 16 bytes cannot be described to TMA) */
int toist_cast_pad_f32_bf16(const float* src, void* dst, int64_t rows, int32_t n, int32_t ld, void* stream);
int toist_add_bf16(const void* a, const void* b, const void* c /* may be null */, void* out, int64_t n, void* stream);
/* 7x7 stride-2 pad-3 stem (torchvision resnet conv1 via backbone.py:75): fp32 NCHW -> bf16 patches [N*Ho*Wo, ldk] */
int toist_stem_im2col(const float* images, void* patches, int32_t n, int32_t h, int32_t w, int32_t ldk, void* stream);
/* Stem input of the implicit-GEMM path (no im2col matrix): images fp32 NCHW [n, 3, h, w] -> out bf16 [n, hp, wp, 8]
 * (hp >= h + 6, wp >= w + 6; 3 zero pixels on top / left, channels 3..7 zero).  The 7 x 7 stride-2 convolution of
 * torchvision's ResNet stem (models/backbone.py:75) is then toist_gemm with one tap per kernel row over a tensor map whose
 * rows overlap (row pitch = 2 padded pixels, 64 elements per row). */
int toist_stem_pad_nhwc8(const float* images, void* out, int32_t n, int32_t h, int32_t w, int32_t hp, int32_t wp,
                         void* stream);
int toist_maxpool3x3s2(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c, void* stream);
int toist_colsum(const void* x, int32_t dtype, float* out, int64_t rows, int32_t cols, int64_t ld, void* stream);
int toist_gelu_bwd(const void* dy, const void* pre, void* dx, int64_t n, void* stream);
/* out = (dy + dy2) * (y > 0), bf16; dy2 may be null (ReLU derivative at a residual join, torchvision Bottleneck) */
int toist_relu_bwd(const void* dy, const void* dy2, const void* y, void* out, int64_t n, void* stream);
int toist_sigmoid_bwd(const float* dy, const float* y, float* dx, int64_t n, void* stream);
int toist_sum_mid(const void* x, int32_t dtype, float* out, int32_t a, int32_t r, int32_t c, int32_t accumulate,
                  void* stream);
int toist_bcast_mid(const float* x, void* out, int32_t a, int32_t r, int32_t c, void* stream);
int toist_nchw_to_nhwc(const float* x, void* y, int32_t n, int32_t c, int32_t hw, void* stream);
int toist_nhwc_to_nchw(const void* x, float* y, int32_t n, int32_t c, int32_t hw, void* stream);
/* dst[a, c, b] = src[a, b, c] (fp32): conv weight gradients [Cout][taps][Cin] -> the parameter's [Cout][Cin][taps] */
int toist_permute_021(const float* src, float* dst, int32_t a, int32_t b, int32_t c, void* stream);
/* nearest resize of the padding mask [B,in_h,in_w] -> small_mask [B,out_h,out_w] (models/backbone.py:78) and the
 * encoder key-padding mask key_mask [B, out_h*out_w + n_text] (transformer.py:102,134,146); text_attention is the
 * tokenizer's int64 attention_mask [B, n_text] (1 = keep).  Either output may be null. */
int toist_key_mask(const uint8_t* pad_mask, const int64_t* text_attention, uint8_t* small_mask, uint8_t* key_mask,
                   int32_t batch, int32_t in_h, int32_t in_w, int32_t out_h, int32_t out_w, int32_t n_text,
                   void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Row-wise normalisation / softmax / embeddings (transformer.py:279-280,346-349,484; mdetr.py:430-433;
 * nn.MultiheadAttention softmax with key padding; position_encoding.py:30-49; RobertaEmbeddings)
 * ------------------------------------------------------------------------------------------------------------ */
int toist_layernorm_fwd(const void* x, int32_t x_dtype, const float* gamma, const float* beta, void* y_bf16,
                        float* y_f32, float* mean, float* rstd, int64_t rows, int32_t n, float eps, void* stream);
int toist_layernorm_bwd(const void* dy, const void* dy2, int32_t dy_dtype, const void* x, int32_t x_dtype,
                        const float* mean, const float* rstd, const float* gamma, void* dx, int32_t dx_dtype,
                        float* dgamma, float* dbeta, int64_t rows, int32_t n, void* stream);
/* Fused residual branch of a transformer layer (transformer.py:298-303,378-407: dropoutN + residual + normN, and the
 * `+ pos` / `+ query_pos` in front of the next attention), bf16 rows of 256 / 512 / 768 elements:
 *   s = res + dropout(x)  (res nullable; p_drop = 0: no dropout), sum_out (nullable) = s, y = LayerNorm(s),
 *   y_add = y + add (both nullable).  Every value is rounded to bf16 where the separate kernels round it.
 * toist_layernorm_bwd_drop additionally writes dx_drop = dropout-backward(dx) with the decisions of `site`.
 * Return TOIST_ERR_UNSUPPORTED for other widths / alignments (the caller falls back to the separate kernels). */
int toist_layernorm_fused_fwd(const void* x, const void* res, const void* add, const float* gamma, const float* beta,
                              void* sum_out, void* y, void* y_add, float* mean, float* rstd, int64_t rows, int32_t n,
                              float eps, float p_drop, const uint64_t* seed, uint32_t site, void* stream);
int toist_layernorm_bwd_drop(const void* dy, const void* dy2, const void* x, const float* mean, const float* rstd,
                             const float* gamma, void* dx, void* dx_drop, float* dgamma, float* dbeta, int64_t rows,
                             int32_t n, float p_drop, const uint64_t* seed, uint32_t site, void* stream);
int toist_l2norm_fwd(const float* x, float* y, float* nrm, int64_t rows, int32_t n, float eps, void* stream);
int toist_l2norm_bwd(const float* dy, const float* y, const float* nrm, float* dx, int64_t rows, int32_t n,
                     void* stream);
/* probs_dropped (may be null): dropout(probs) with probability p_drop, the operand of the PV product in training
 * (nn.MultiheadAttention dropout on the attention weights); the mask is a counter-based hash of (seed[0], site, index)
 * and is regenerated, never stored, by toist_attn_softmax_bwd when it is given the same seed / site. */
int toist_attn_softmax_fwd(const float* scores, const uint8_t* key_mask, void* probs, void* probs_dropped,
                           int64_t rows, int32_t sk, int32_t ld_s, int32_t ld_p, int32_t rows_per_batch, float p_drop,
                           const uint64_t* seed, uint32_t site, void* stream);
int toist_attn_softmax_bwd(const float* dprobs, const void* probs, void* dscores, int64_t rows, int32_t sk,
                           int32_t ld_s, int32_t ld_p, float scale, float p_drop, const uint64_t* seed, uint32_t site,
                           void* stream);
/* ------------------------------------------------------------------------------------------------------------
 * Fused attention core (csrc/attention.cu): O = dropout(softmax(Q K^T / sqrt(d) + key mask)) V in one launch with the
 * scores in TMEM, and its backward (dQ, dK, dV) in one launch + a deterministic reduction.  Replaces the
 * bmm / softmax / dropout / bmm sequence inside nn.MultiheadAttention (reference models/transformer.py:273,337-338
 * -> F.multi_head_attention_forward) and inside RobertaSelfAttention (transformer.py:130).
 *
 * q, k, v, out, dout, dq, dk, dv are bf16 with element (s, b, head * d + i) at ptr[s * ss + b * sb + head * d + i]
 * (ss, sb in elements, multiples of 8; bases 16-byte aligned): the reference's [S, B, E] layout, or column slices of
 * a packed projection.  key_mask: uint8 [b, sk], non-zero = key ignored (key_padding_mask), may be null.
 * lse: f32 [b, h, sq, 2] row statistics (m2, l): the row maximum of the scaled scores in the log2 domain and the softmax
 * denominator sum_k exp2(s_k * scale * log2(e) - m2); written by the forward when non-null, required by the backward
 * (natural log-sum-exp = m2 * ln 2 + ln l; kept apart because their sum loses the softmax for very large logits).
 * Dropout (p_drop > 0): decisions are a hash of (seed[0], site, row, key pair) regenerated in the backward;
 * p is quantised to round(p * 65536) / 65536.  Supported: d in {32, 64}, sk <= 448 (toist_attention_supported). */
typedef struct toist_attn_desc {
  const void* q;
  const void* k;
  const void* v;
  void* out;
  float* lse;
  const uint8_t* key_mask;
  const uint64_t* seed;
  int64_t q_ss, q_sb, k_ss, k_sb, v_ss, v_sb, o_ss, o_sb;
  int32_t sq, sk, b, h, d;
  float p_drop;
  uint32_t site;
  int32_t reserved;
} toist_attn_desc;

typedef struct toist_attn_bwd_desc {
  toist_attn_desc fwd; /* the forward call's descriptor (out and lse as the forward wrote them) */
  const void* dout;
  void* dq;
  void* dk;
  void* dv;
  void* workspace; /* toist_attention_bwd_workspace() bytes, f32 partial dQ per 128-key tile */
  int64_t do_ss, do_sb, dq_ss, dq_sb, dk_ss, dk_sb, dv_ss, dv_sb;
} toist_attn_bwd_desc;

size_t toist_sizeof_attn_desc(void);
size_t toist_sizeof_attn_bwd_desc(void);
int toist_attention_supported(int32_t sq, int32_t sk, int32_t d);
int toist_attention_fwd(const toist_attn_desc* a, void* stream);
int64_t toist_attention_bwd_workspace(int32_t sq, int32_t sk, int32_t b, int32_t h, int32_t d);
int toist_attention_bwd(const toist_attn_bwd_desc* a, void* stream);
/* keep[b, h, sq, sk] in {0, 1}: the dropout decisions of the fused kernels for (seed, site) - test support */
int toist_attention_dropout_mask(uint8_t* keep, int32_t b, int32_t h, int32_t sq, int32_t sk, float p_drop,
                                 const uint64_t* seed, uint32_t site, void* stream);

int toist_pos_sine(const uint8_t* mask, float* pos_f32, void* pos_bf16, int32_t batch, int32_t h, int32_t w,
                   int32_t num_pos_feats, float temperature, void* stream);
/* ids [batch, len]; rows of out / pos_ids / dx are b*len + l, or l*batch + b when seq_first != 0 */
int toist_embed_gather(const int64_t* ids, const float* word, const float* pos, const float* type0, float* out,
                       int32_t* pos_ids, int32_t batch, int32_t len, int32_t dim, int32_t pad_id, int32_t seq_first,
                       void* stream);
/* backward of the three embedding lookups (atomic adds into zeroed tables).  pad_id: nn.Embedding(padding_idx) of
 * transformers' RobertaEmbeddings -- rows `pad_id` of the word and position tables receive no gradient (-1: none). */
int toist_embed_scatter(const void* dx, int32_t dx_dtype, const int64_t* ids, const int32_t* pos_ids, float* dword,
                        float* dpos, float* dtype0, int32_t batch, int32_t len, int32_t dim, int32_t seq_first,
                        int32_t pad_id, void* stream);
/* Sparse data-parallel exchange of the word-embedding gradient (reference main.py:336 all-reduces the dense table):
 * ids [n_rows] / rows [n_rows, dim] = the (token id, gradient row) pairs of ALL ranks in rank order (all-gathered by the
 * caller); table[id, :] = scale * sum of the rows carrying that id, summed in list order and written (deterministic:
 * bit-identical on every rank); `pad_id` rows are skipped. */
int toist_embed_rows_merge(float* table, const int64_t* ids, const float* rows, int32_t n_rows, int32_t dim,
                           int64_t pad_id, float scale, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Mask branch of DETRsegm (reference models/segmentation.py:157-241, 244-273; models/mdetr.py:827-853).
 * Maps are NHWC bf16 [n_maps = B*Q, H, W, C]; the 3x3 / 1x1 convolutions run on toist_gemm.
 * ------------------------------------------------------------------------------------------------------------ */
/* x0[(b*Q+q), pix, :] = concat(src_proj[pix, b, :dim], attn[b, :, q, pix])   (MaskHeadSmallConv input,
 * segmentation.py:205-207); src_proj bf16 [hw, B, dim], attn bf16 [B, n_heads, Q, ld_attn] */
int toist_mask_input(const void* src_proj, const void* attn, void* x0, int32_t batch, int32_t n_queries, int32_t hw,
                     int32_t dim, int32_t n_heads, int32_t ld_attn, void* stream);
/* dsrc_proj (bf16, may be null) = sum over queries; dattn fp32 [B, n_heads, Q, ld_attn] */
int toist_mask_input_bwd(const void* dx0, void* dsrc_proj, float* dattn, int32_t batch, int32_t n_queries, int32_t hw,
                         int32_t dim, int32_t n_heads, int32_t ld_attn, void* stream);
/* a = relu(GroupNorm(z)) with `groups` groups (torch.nn.GroupNorm(8, C) + F.relu, segmentation.py:209-240);
 * mean / rstd [n_maps, groups] are outputs kept for the backward pass */
int toist_groupnorm_relu_fwd(const void* z, const float* gamma, const float* beta, void* a, float* mean, float* rstd,
                             int32_t n_maps, int32_t hw, int32_t channels, int32_t groups, float eps, void* stream);
/* dz from da; dgamma / dbeta accumulate (may be null together); scratch: 2 * n_maps * groups floats */
int toist_groupnorm_relu_bwd(const void* da, const void* z, const float* mean, const float* rstd, const float* gamma,
                             const float* beta, void* dz, float* dgamma, float* dbeta, float* scratch, int32_t n_maps,
                             int32_t hw, int32_t channels, int32_t groups, void* stream);
/* out = fpn[map / n_queries] + nearest_upsample(xs)   (cur_fpn + F.interpolate(x, size), segmentation.py:217-220) */
int toist_upsample_add(const void* xs, const void* fpn, void* out, int32_t n_maps, int32_t n_queries, int32_t out_h,
                       int32_t out_w, int32_t in_h, int32_t in_w, int32_t channels, void* stream);
/* dxs = window sums of dout; dfpn (may be null) = sum over the queries of an image */
int toist_upsample_add_bwd(const void* dout, void* dxs, void* dfpn, int32_t n_maps, int32_t n_queries, int32_t out_h,
                           int32_t out_w, int32_t in_h, int32_t in_w, int32_t channels, void* stream);
/* loss_masks (mdetr.py:827-853): bilinear upsample (align_corners = false) of the matched predictions to the target
 * size fused with sigmoid focal (alpha .25, gamma 2) and dice.  pred_masks fp32 [B, Q, mh, mw]; tgt_masks uint8
 * [B, t_max, th, tw]; match_q int32 [B, t_max]; sums [B, t_max, 4] scratch kept for the backward pass;
 * out[0] = loss_mask, out[1] = loss_dice (both already divided by num_boxes[0]) */
int toist_mask_loss_fwd(const float* pred_masks, const uint8_t* tgt_masks, const int32_t* match_q,
                        const int32_t* tgt_count, const float* num_boxes, float* sums, float* out, int32_t batch,
                        int32_t n_queries, int32_t t_max, int32_t mask_h, int32_t mask_w, int32_t tgt_h, int32_t tgt_w,
                        void* stream);
/* dpred [B, Q, mh, mw] = gout[0] * d loss_mask + gout[1] * d loss_dice (gout on the device) */
int toist_mask_loss_bwd(const float* pred_masks, const uint8_t* tgt_masks, const int32_t* match_q,
                        const int32_t* tgt_count, const float* sums, const float* num_boxes, const float* gout,
                        float* dpred, int32_t batch, int32_t n_queries, int32_t t_max, int32_t mask_h, int32_t mask_w,
                        int32_t tgt_h, int32_t tgt_w, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Noun-pronoun distillation (BASELINE config 5).
 * Soft-KD (reference models/mdetr.py:520-599, called per decoder layer at mdetr.py:969-988): teacher (noun) and student
 * (pronoun) predictions of all layers at once.  Per (layer, image): two-class probabilities
 * (sum of the first C-1 softmax entries, last entry); matched queries are paired by target index, the unmatched ones by
 * an LSAP on  L1 + KL(teacher || student) - GIoU  (rows = student, scipy tie rules, float64);
 * loss[l] = mean over images of kl_div(log student, teacher, "batchmean").  Gradient: student logits only.
 *   logits_* [L, B, Q, C], boxes_* [L, B, Q, 4], match_* [L, B, t_max] (query of each target, toist_lsap_device)
 *   workspace (caller allocated): bi_* [L, B, Q, 2] f32, fp_* [L, B, Q] i32, n_fp [2, L*B] i32, cost [L, B, Q, Q] f32,
 *   col_of_row [L, B, Q] i32, pair_noun [L, B, Q] i32 (teacher query paired with each student query);
 *   flags[0] |= 1 on NaN / infeasible costs (scipy's ValueError).
 * ------------------------------------------------------------------------------------------------------------ */
int toist_softkd_fwd(const float* logits_noun, const float* logits_sth, const float* boxes_noun, const float* boxes_sth,
                     const int32_t* match_noun, const int32_t* match_sth, const int32_t* tgt_count, float* bi_noun,
                     float* bi_sth, int32_t* fp_noun, int32_t* fp_sth, int32_t* n_fp, float* cost, int32_t* col_of_row,
                     int32_t* pair_noun, int32_t* flags, float* loss, int32_t n_layers, int32_t batch, int32_t n_queries,
                     int32_t n_classes, int32_t t_max, void* stream);
/* dlogits_sth [L, B, Q, C] = sum_l gout[l] * d loss[l] / d logits_sth */
int toist_softkd_bwd(const float* logits_sth, const float* bi_noun, const float* bi_sth, const int32_t* pair_noun,
                     const int32_t* tgt_count, const int32_t* n_fp, const float* gout, float* dlogits_sth,
                     int32_t n_layers, int32_t batch, int32_t n_queries, int32_t n_classes, int32_t t_max, void* stream);
/* Batched LSAP, one warp per problem (scipy.optimize.linear_sum_assignment semantics incl. ties; mdetr.py:100,539):
 * cost [P, ld_rows, ld_cols] f32 with n_rows[p] x n_cols[p] valid entries (each <= 128);
 * col_of_row [P, ld_rows] = assigned column per row or -1. */
int toist_lsap_batched(const float* cost, const int32_t* n_rows, const int32_t* n_cols, int32_t* col_of_row,
                       int32_t* flags, int32_t n_problems, int32_t ld_rows, int32_t ld_cols, void* stream);
/* ClusterCriterion arithmetic (models/mdetr.py:29-312, models/kmeans.py:21-133).
 * toist_kmeans: Lloyd iterations on x [n, dim] from `centers` [k, dim] (updated in place) until
 * (sum_k ||shift_k||)^2 < tol, without host round trips; choice [n]; iters (may be null) = iterations run;
 * xt_workspace: f32 [dim, n] scratch for the transposed bank (may be null: slower row-strided access, same results). */
int toist_kmeans(const float* x, float* centers, int32_t* choice, int32_t* iters, float* xt_workspace, int32_t n,
                 int32_t dim, int32_t k, float tol, int32_t max_iter, void* stream);
/* One launch for several INDEPENDENT problems (one CTA each): problem p clusters bank task_of[p] of banks [n_tasks, n, dim]
 * from centers [p, k, dim] (in place); choice [p, n]; iters [p] (may be null); xt_workspace [p, dim, n] (may be null);
 * query [p, dim] -> query_choice[p] = nearest final centre (toist_kmeans_predict arithmetic), both may be null. */
int toist_kmeans_batched(const float* banks, const int32_t* task_of, float* centers, int32_t* choice, int32_t* iters,
                         float* xt_workspace, const float* query, int32_t* query_choice, int32_t n_problems, int32_t n,
                         int32_t dim, int32_t k, float tol, int32_t max_iter, void* stream);
int toist_kmeans_predict(const float* x, const float* centers, int32_t* choice, int32_t m, int32_t dim, int32_t k,
                         void* stream);
/* out[b, :] = sum_t w[b, t] * x[t, b, :]   (x f32 [n_tokens, batch, dim]; w = selection / count gives the token means of
 * mdetr.py:141,256) and its backward dx[t, b, :] = w[b, t] * dout[b, :] */
int toist_token_wsum(const float* x, const float* w, float* out, int32_t n_tokens, int32_t batch, int32_t dim,
                     void* stream);
int toist_token_wsum_bwd(const float* dout, const float* w, float* dx, int32_t n_tokens, int32_t batch, int32_t dim,
                         void* stream);
/* x[t, b, :] = feat[b, :] (0 when feat is null) for every selected token (mdetr.py:208,266) */
int toist_token_fill(float* x, const uint8_t* sel, const float* feat, int32_t n_tokens, int32_t batch, int32_t dim,
                     void* stream);
/* loss[0] = mean over the rows with use[r] != 0 of mse(a[r], b[r]); da (may be null) = its gradient w.r.t. a */
int toist_mse_rows(const float* a, const float* b, const uint8_t* use, float* loss, float* da, int32_t rows, int32_t dim,
                   void* stream);
/* out [n, m] = torch.cdist(a [n, dim], b [m, dim], p=1)   (mdetr.py:98) */
int toist_cdist_l1(const float* a, const float* b, float* out, int32_t n, int32_t m, int32_t dim, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Optimizer side of the step (csrc/optim.cu), multi-tensor: `items` is a DEVICE array of 48-byte records
 *   { float* a; float* b; float* c; float* d; int64_t n; int32_t first_block; int32_t group; }
 * (toist_sizeof_opt_item()), first_block = running sum of ceil(n / 2048) over the preceding records, total_blocks the
 * sum over all of them.  All tensors are contiguous f32.
 *   toist_grad_sqnorm      a = grad.  partial: f32 [total_blocks] scratch.  norm_and_coef[0] = L2 norm over all items,
 *                          [1] = min(1, max_norm / (norm + 1e-6))   -- torch.nn.utils.clip_grad_norm_ (engine.py:89-90)
 *   toist_grad_clip_scale  a = grad, scaled in place by norm_and_coef[1]
 *   toist_adamw_step       a = param, b = grad, c = exp_avg, d = exp_avg_sq, group = parameter group; hyper (HOST) holds 8
 *                          floats per group {1 - lr * wd, 1 - beta1, beta2, 1 - beta2, eps, lr / (1 - beta1^t), sqrt(1 - beta2^t), 0},
 *                          each formed in double precision as torch's Python code does and rounded once
 *                          -- torch.optim.AdamW.step (main.py:351-392, engine.py:91)
 *   toist_ema_update       a = ema tensor, b = model tensor: a = a * decay + one_minus_decay * b  (util/optim.py:9-26) */
size_t toist_sizeof_opt_item(void);
int toist_grad_sqnorm(const void* items_dev, int32_t n_items, int32_t total_blocks, float* partial, float max_norm,
                      float* norm_and_coef, void* stream);
int toist_grad_clip_scale(const void* items_dev, int32_t n_items, int32_t total_blocks, const float* norm_and_coef,
                          void* stream);
int toist_adamw_step(const void* items_dev, int32_t n_items, int32_t total_blocks, const float* hyper_host,
                     int32_t n_groups, void* stream);
int toist_ema_update(const void* items_dev, int32_t n_items, int32_t total_blocks, float decay, float one_minus_decay,
                     void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * The two ends of the path (csrc/io.cu).
 *
 * Batch assembly -- replaces NestedTensor.from_tensor_list's per-image copy + mask fill (util/misc.py:185-209) and,
 * for the u8 entry, ToTensor + Normalize of the CPU data workers (datasets/transforms.py:257-272):
 *   image_ptrs [B] DEVICE array of device pointers, image_hw [B, 2] DEVICE int32 (height, width);
 *   u8 images are HWC (3 channels), f32 images CHW; out [B, C, H, W] f32 zero padded, mask [B, H, W] u8 (1 = padding);
 *   mean3_host / std3_host: 3 HOST floats.  Arithmetic per pixel: ((x / 255) - mean) / std, each step rounded once.
 * PostProcess (models/postprocessors.py:15-58): scores = 1 - softmax(logits)[..., -1], labels = 1, boxes cxcywh ->
 *   xyxy * (w, h, w, h) with sizes [B, 2] = (height, width) as f32 OR i64 (pass the other as null); optional
 *   is_final [B, Q] -> scores_refexp = scores * sigmoid(is_final).
 * PostProcessSegm (models/postprocessors.py:61-109), one image per call: pred_masks [Q, mask_h, mask_w] f32 ->
 *   bilinear to (stage1_h, stage1_w) [the padded batch size], crop to (crop_h, crop_w), bilinear to (out_h, out_w),
 *   sigmoid > threshold -> out [Q, out_h, out_w] u8.  Both interpolations align_corners = false, fused into one pass.
 * ------------------------------------------------------------------------------------------------------------ */
int toist_pad_normalize_u8(const uint64_t* image_ptrs, const int32_t* image_hw, float* out, uint8_t* mask, int32_t batch,
                           int32_t height, int32_t width, const float* mean3_host, const float* std3_host, void* stream);
int toist_pad_batch_f32(const uint64_t* image_ptrs, const int32_t* image_hw, float* out, uint8_t* mask, int32_t batch,
                        int32_t channels, int32_t height, int32_t width, void* stream);
int toist_postprocess_boxes(const float* logits, const float* boxes, const float* sizes_f32, const int64_t* sizes_i64,
                            const float* is_final, float* scores, int64_t* labels, float* out_boxes, float* scores_refexp,
                            int32_t batch, int32_t n_queries, int32_t n_classes, void* stream);
int toist_postprocess_masks(const float* pred_masks, uint8_t* out, int32_t n_queries, int32_t mask_h, int32_t mask_w,
                            int32_t stage1_h, int32_t stage1_w, int32_t crop_h, int32_t crop_w, int32_t out_h,
                            int32_t out_w, float threshold, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TOIST_B200_H_ */
