// Fused multi-head attention core for sm_100a: softmax(Q K^T / sqrt(d) + key mask) (+ dropout) V in ONE launch, and its
// backward in one launch (+ a small deterministic dQ reduction).  Replaces the three (forward) / five (backward)
// launches of the unfused path (QK^T GEMM -> softmax -> PV GEMM through gemm.cu + norm.cu) and their fp32 score /
// bf16 probability round trips through HBM: scores live in TMEM, probabilities go TMEM -> registers -> shared memory
// -> tensor core, only O (bf16) and the row log-sum-exp (fp32) are written.
//
// Replaces, on the reference path, the attention core inside nn.MultiheadAttention (transformer.py:273,337-338:
// F.multi_head_attention_forward -> bmm / softmax / dropout / bmm) for the encoder self-attention (S = 416), the
// decoder self-attention (Q = 100), the decoder cross-attention (100 x 416) and RoBERTa (16 x 16, d = 64).
//
// Forward, one CTA per (128-query tile, head, batch), 192 threads:
//   warp 0      TMA producer: Q tile, all K rows, all V rows (SWIZZLE_128B, head_dim padded to 64 by TMA zero fill)
//   warp 1      TMEM owner + tcgen05.mma issuer: S = Q K^T (N up to 256 per instruction), then O += P_j V_j per
//               64-key block j as the softmax warps hand the blocks over through a shared-memory ring
//   warps 2..5  softmax: thread r owns query row r = TMEM lane r.  Pass 1 reads S once for the row maximum, pass 2
//               reads it again, exponentiates (ex2.approx), applies the dropout decision, sums, packs bf16 and writes
//               the K-major operand tile for the PV product; the epilogue scales O by keep_scale / l.
// Backward, one CTA per (128-key tile, head, batch), 320 threads, loop over 128-query chunks:
//   S^T = K Q^T and dP^T = V dO^T (TMEM) -> 8 compute warps: p = exp2((s - mx) * scale) / l, dropout decision regenerated,
//   Pd^T and dS^T (bf16, K-major in shared memory) -> dV += Pd^T dO, dK += dS^T Q (accumulated in TMEM over the
//   chunks), dQ_chunk = dS K (TMEM -> fp32 partial per key tile; attn_dq_reduce_kernel sums the key tiles in a fixed
//   order and writes bf16).  Nothing is atomic: results are deterministic run to run.
//
// Dropout: one 32-bit hash per PAIR of adjacent keys of a row (two 16-bit decisions), keyed by (seed, site); the
// probability is therefore quantised to thr16 / 65536 (p = 0.1 -> 0.100006) and keep_scale is 1 / (1 - thr16 / 65536).
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "host_util.h"

namespace toist {

constexpr int kAttNP = 4;  // depth of the P ring (forward)
constexpr int kFwdThreads = 320;
constexpr int kAttMaxKB = 7;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnFwdParams {
  int sq, sk, b, h, d, kb;
  float scale, scale_log2;
  __nv_bfloat16* out;
  long long o_ss, o_sb;
  float* lse;
  const uint8_t* kmask;
  const unsigned long long* seed;
  uint32_t site, thr16;
  float keep_scale;
  int tmem_cols;
};

// ------------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(kFwdThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ AttnFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int kb = p.kb;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + 16384;
  uint8_t* sV = sK + kb * 8192;
  uint8_t* sP = sV + kb * 8192;
  float* sRed = reinterpret_cast<float*>(sP + kAttNP * 16384);  // [2][128] row maxima, [2][128] row sums
  uint32_t* sBits = reinterpret_cast<uint32_t*>(sRed + 512);    // one mask word per 32 keys
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBits + 16);
  uint64_t* bar_qk = bars + 0;
  uint64_t* bar_v = bars + 1;
  uint64_t* bar_s = bars + 2;  // [2]
  uint64_t* bar_o = bars + 4;
  uint64_t* p_full = bars + 5;            // [kAttNP]
  uint64_t* p_empty = p_full + kAttNP;    // [kAttNP]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_empty + kAttNP);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int ncols = kb * 64;
  const int nchunks = (ncols + 255) / 256;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(&bar_s[0], 1);
    mbar_init(&bar_s[1], 1);
    mbar_init(bar_o, 1);
    for (int s = 0; s < kAttNP; ++s) {
      mbar_init(&p_full[s], 128);
      mbar_init(&p_empty[s], 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + (uint32_t)ncols;
  pdl_wait();

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(bar_qk, 16384u + (uint32_t)kb * 8192u);
      tma_load_4d(sQ, &tm_q, bar_qk, 0, q0, h, b);
      for (int j = 0; j < kb; ++j) tma_load_4d(sK + j * 8192, &tm_k, bar_qk, 0, j * 64, h, b);
      mbar_expect_tx(bar_v, (uint32_t)kb * 8192u);
      for (int j = 0; j < kb; ++j) tma_load_4d(sV + j * 8192, &tm_v, bar_v, 0, j * 64, h, b);
    }
  } else if (warp == 1) {
    const int kd = p.d >> 4;
    mbar_wait(bar_qk, 0);
    tc_fence_after();
    if (elect_one()) {
      for (int c = 0; c < nchunks; ++c) {
        const int n = min(256, ncols - c * 256);
        const uint32_t idesc = umma_idesc_bf16(n, false, false);
        for (int ks = 0; ks < kd; ++ks) {
          const uint64_t da = umma_smem_desc(smem_u32(sQ) + ks * 32, 16, 1024);
          const uint64_t db = umma_smem_desc(smem_u32(sK) + c * 256 * 128 + ks * 32, 16, 1024);
          umma_f16(tmem_base + (uint32_t)(c * 256), da, db, idesc, ks > 0 ? 1u : 0u);
        }
        umma_commit(&bar_s[c]);
      }
    }
    __syncwarp();
    mbar_wait(bar_v, 0);
    const uint32_t idesc_o = umma_idesc_bf16(p.d, false, true);
    for (int j = 0; j < kb; ++j) {
      const int slot = j % kAttNP;
      mbar_wait(&p_full[slot], (uint32_t)((j / kAttNP) & 1));
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t da = umma_smem_desc(smem_u32(sP) + slot * 16384 + ks * 32, 16, 1024);
          const uint64_t db = umma_smem_desc(smem_u32(sV) + j * 8192 + ks * 2048, 8192, 1024);
          umma_f16(tmem_o, da, db, idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(&p_empty[slot]);
        if (j == kb - 1) umma_commit(bar_o);
      }
      __syncwarp();
    }
  } else {
    // 8 softmax warps: thread (r, hf) owns query row r = TMEM lane r and the 64-key blocks j with j % 2 == hf
    const int quad = warp & 3;
    const int hf = (warp - 2) >> 2;
    const int r = quad * 32 + lane;
    const int qrow = q0 + r;
    const bool warp_valid = q0 + quad * 32 < p.sq;  // warp-uniform: rows past the last query do no arithmetic
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    // key mask as one bit per key (1 = masked or past the last key)
    for (int w = warp - 2; w < kb * 2; w += 8) {
      const int key = w * 32 + lane;
      const bool m = key >= p.sk || (p.kmask != nullptr && p.kmask[(long long)b * p.sk + key] != 0);
      const uint32_t bits = __ballot_sync(0xffffffffu, m);
      if (lane == 0) sBits[w] = bits;
    }
    named_barrier_sync(1, 256);
    // ---- pass 1: row maximum of the raw scores over the unmasked keys
    float mx = -INFINITY;
    int s_ready = 0;
    if (warp_valid) {
      for (int j = hf; j < kb; j += 2) {
        while (s_ready <= (j >> 2)) {
          mbar_wait(&bar_s[s_ready], 0);
          ++s_ready;
        }
        tc_fence_after();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int c0 = j * 64 + half * 32;
          uint32_t raw[32];
          tmem_ld_32x32(lane_addr + (uint32_t)c0, raw);
          tmem_ld_wait();
          const uint32_t bits = sBits[c0 >> 5];
          if (bits == 0u) {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(raw[i]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (!((bits >> i) & 1u)) mx = fmaxf(mx, __uint_as_float(raw[i]));
          }
        }
      }
    }
    sRed[hf * 128 + r] = mx;
    named_barrier_sync(1, 256);
    mx = fmaxf(sRed[r], sRed[128 + r]);
    // ---- pass 2: e = exp2((s - max) * scale * log2 e), dropout, bf16 operand tiles for the PV product.
    // The subtraction comes first: s - mx is exactly 0 for the row maximum and <= 0 for every other key, so e <= 1 and
    // 1 <= l <= Sk whatever the magnitude of the logits (random-init ResNet-101 features give logits of 1e10, where
    // fma(s, c, -mx * c) is off by up to half an ulp of mx * c, i.e. by dozens of octaves).
    const bool drop = p.seed != nullptr;
    uint32_t k0 = 0, k1 = 0;
    if (drop) {
      const uint64_t key = dropout_key(p.seed, p.site);
      k0 = (uint32_t)key;
      k1 = (uint32_t)(key >> 32);
    }
    const uint32_t pair_base = (uint32_t)(((long long)(b * p.h + h) * p.sq + qrow) * (kb * 32)) + k0;
    const uint32_t rx = (uint32_t)(r & 7);
    float l = 0.f;
    for (int j = hf; j < kb; j += 2) {
      const int slot = j % kAttNP;
      mbar_wait(&p_empty[slot], (uint32_t)(((j / kAttNP) & 1) ^ 1));
      if (warp_valid) {
        while (s_ready <= (j >> 2)) {
          mbar_wait(&bar_s[s_ready], 0);
          ++s_ready;
        }
        tc_fence_after();
        uint8_t* prow = sP + slot * 16384 + r * 128;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int c0 = j * 64 + half * 32;
          uint32_t raw[32];
          tmem_ld_32x32(lane_addr + (uint32_t)c0, raw);
          tmem_ld_wait();
          const uint32_t bits = sBits[c0 >> 5];
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            // exactly this expression is re-evaluated by the backward pass (same operands, one rounding)
            float e0 = ex2_approx(__fmul_rn(__fsub_rn(__uint_as_float(raw[i]), mx), p.scale_log2));
            float e1 = ex2_approx(__fmul_rn(__fsub_rn(__uint_as_float(raw[i + 1]), mx), p.scale_log2));
            if (bits != 0u) {  // warp-uniform
              if ((bits >> i) & 1u) e0 = 0.f;
              if ((bits >> (i + 1)) & 1u) e1 = 0.f;
            }
            l += e0 + e1;
            if (drop) {
              const uint32_t w = hash32(pair_base + (uint32_t)((c0 + i) >> 1)) ^ k1;
              if ((w & 0xffffu) < p.thr16) e0 = 0.f;
              if ((w >> 16) < p.thr16) e1 = 0.f;
            }
            pk[i >> 1] = pack_bf16(e0, e1);
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint32_t chunk = (uint32_t)(half * 4 + g) ^ rx;
            *reinterpret_cast<uint4*>(prow + (chunk << 4)) =
                make_uint4(pk[g * 4], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]);
          }
        }
        tc_fence_before();
        fence_proxy_async();
      }
      mbar_arrive(&p_full[slot]);
    }
    pdl_trigger();
    sRed[256 + hf * 128 + r] = l;
    named_barrier_sync(1, 256);
    l = sRed[256 + r] + sRed[384 + r];
    // ---- epilogue: O * keep_scale / l -> bf16 (this thread: d/2 columns), log-sum-exp for the backward pass
    mbar_wait(bar_o, 0);
    tc_fence_after();
    const float inv = p.keep_scale / l;
    const int hc = p.d >> 1;
    __nv_bfloat16* op = p.out + (long long)qrow * p.o_ss + (long long)b * p.o_sb + h * p.d + hf * hc;
    if (hc == 16) {
      uint32_t raw[16];
      tmem_ld_32x16(lane_addr + (uint32_t)(ncols + hf * 16), raw);
      tmem_ld_wait();
      if (qrow < p.sq) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint4 u;
          u.x = pack_bf16(__uint_as_float(raw[g * 8 + 0]) * inv, __uint_as_float(raw[g * 8 + 1]) * inv);
          u.y = pack_bf16(__uint_as_float(raw[g * 8 + 2]) * inv, __uint_as_float(raw[g * 8 + 3]) * inv);
          u.z = pack_bf16(__uint_as_float(raw[g * 8 + 4]) * inv, __uint_as_float(raw[g * 8 + 5]) * inv);
          u.w = pack_bf16(__uint_as_float(raw[g * 8 + 6]) * inv, __uint_as_float(raw[g * 8 + 7]) * inv);
          reinterpret_cast<uint4*>(op)[g] = u;
        }
      }
    } else {
      uint32_t raw[32];
      tmem_ld_32x32(lane_addr + (uint32_t)(ncols + hf * 32), raw);
      tmem_ld_wait();
      if (qrow < p.sq) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          u.x = pack_bf16(__uint_as_float(raw[g * 8 + 0]) * inv, __uint_as_float(raw[g * 8 + 1]) * inv);
          u.y = pack_bf16(__uint_as_float(raw[g * 8 + 2]) * inv, __uint_as_float(raw[g * 8 + 3]) * inv);
          u.z = pack_bf16(__uint_as_float(raw[g * 8 + 4]) * inv, __uint_as_float(raw[g * 8 + 5]) * inv);
          u.w = pack_bf16(__uint_as_float(raw[g * 8 + 6]) * inv, __uint_as_float(raw[g * 8 + 7]) * inv);
          reinterpret_cast<uint4*>(op)[g] = u;
        }
      }
    }
    // Row statistics for the backward pass: (mx, l) = (raw row maximum of q.k, softmax denominator), kept as two
    // numbers.  Folding them into one log-sum-exp loses the softmax when |logit| is large: at random initialisation
    // the first encoder layer sees logits of ~1e10 (ulp 1e3), the backward recomputed exp(s - lse) from two
    // differently rounded large numbers and produced inf -> NaN gradients for the whole backbone.
    if (hf == 0 && p.lse != nullptr && qrow < p.sq)
      reinterpret_cast<float2*>(p.lse)[(long long)(b * p.h + h) * p.sq + qrow] = make_float2(mx, l);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------ backward
struct AttnBwdParams {
  int sq, sk, b, h, d, kb, nq;
  float scale, scale_log2;
  const __nv_bfloat16* o;
  long long o_ss, o_sb;
  const __nv_bfloat16* dout;
  long long do_ss, do_sb;
  const float* lse;
  const uint8_t* kmask;
  const unsigned long long* seed;
  uint32_t site, thr16;
  float keep_scale;
  float* dq_part;  // [key tiles][sq][b][h*d]
  __nv_bfloat16* dk;
  long long dk_ss, dk_sb;
  __nv_bfloat16* dv;
  long long dv_ss, dv_sb;
};

constexpr int kBwdThreads = 320;
constexpr int kBwdMaxNQ = 8;  // query chunks of 128 (sq <= 1024)

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                const __grid_constant__ AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + 16384;
  uint8_t* sRing = sV + 16384;      // 2 x (Q chunk 16 KB | dO chunk 16 KB)
  uint8_t* sPd = sRing + 65536;     // Pd^T: 2 blocks of [128 keys][64 queries]
  uint8_t* sdS = sPd + 32768;       // dS^T: same layout
  float* sLse = reinterpret_cast<float*>(sdS + 32768);  // [nq <= 8][128] raw row maximum of q.k (inf past the last query)
  float* sDelta = sLse + kBwdMaxNQ * 128;                // [nq][128] rowsum(dO * O)
  float* sInvL = sDelta + kBwdMaxNQ * 128;               // [nq][128] 1 / softmax denominator (0 past the last query)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sInvL + kBwdMaxNQ * 128);
  uint64_t* bar_kv = bars + 0;
  uint64_t* full = bars + 1;    // [2]
  uint64_t* empty = bars + 3;   // [2]
  uint64_t* bar_st = bars + 5;
  uint64_t* bar_pds = bars + 6;
  uint64_t* bar_dq = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int key0 = kt * 128;
  const int kd = p.d >> 4;
  constexpr uint32_t kColST = 0, kColDP = 128, kColDQ = 256, kColDV = 320, kColDK = 384;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_do);
    mbar_init(bar_kv, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(bar_st, 1);
    mbar_init(bar_pds, 256);
    mbar_init(bar_dq, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(bar_kv, 32768u);
      tma_load_4d(sK, &tm_k, bar_kv, 0, key0, h, b);
      tma_load_4d(sV, &tm_v, bar_kv, 0, key0, h, b);
      for (int c = 0; c < p.nq; ++c) {
        const int st = c & 1;
        mbar_wait(&empty[st], (uint32_t)(((c >> 1) & 1) ^ 1));
        mbar_expect_tx(&full[st], 32768u);
        tma_load_4d(sRing + st * 32768, &tm_q, &full[st], 0, c * 128, h, b);
        tma_load_4d(sRing + st * 32768 + 16384, &tm_do, &full[st], 0, c * 128, h, b);
      }
    }
  } else if (warp == 1) {
    const uint32_t id_s = umma_idesc_bf16(128, false, false);
    const uint32_t id_kv = umma_idesc_bf16(p.d, false, true);
    const uint32_t id_q = umma_idesc_bf16(p.d, true, true);
    const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aPd = smem_u32(sPd), adS = smem_u32(sdS);
    mbar_wait(bar_kv, 0);
    for (int c = 0; c < p.nq; ++c) {
      const int st = c & 1;
      const uint32_t aQ = smem_u32(sRing + st * 32768), aDO = aQ + 16384;
      mbar_wait(&full[st], (uint32_t)((c >> 1) & 1));
      tc_fence_after();
      if (elect_one()) {
        for (int ks = 0; ks < kd; ++ks)  // S^T = K Q^T
          umma_f16(tmem_base + kColST, umma_smem_desc(aK + ks * 32, 16, 1024), umma_smem_desc(aQ + ks * 32, 16, 1024),
                   id_s, ks > 0 ? 1u : 0u);
        for (int ks = 0; ks < kd; ++ks)  // dP^T = V dO^T
          umma_f16(tmem_base + kColDP, umma_smem_desc(aV + ks * 32, 16, 1024), umma_smem_desc(aDO + ks * 32, 16, 1024),
                   id_s, ks > 0 ? 1u : 0u);
        umma_commit(bar_st);
      }
      __syncwarp();
      mbar_wait(bar_pds, (uint32_t)(c & 1));
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {  // dV += Pd^T dO ; dK += dS^T Q   (reduction over the chunk's queries)
          const uint32_t aoff = (uint32_t)((ks >> 2) * 16384 + (ks & 3) * 32);
          const uint32_t acc = (c > 0 || ks > 0) ? 1u : 0u;
          umma_f16(tmem_base + kColDV, umma_smem_desc(aPd + aoff, 16, 1024), umma_smem_desc(aDO + ks * 2048, 8192, 1024),
                   id_kv, acc);
          umma_f16(tmem_base + kColDK, umma_smem_desc(adS + aoff, 16, 1024), umma_smem_desc(aQ + ks * 2048, 8192, 1024),
                   id_kv, acc);
        }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)  // dQ_chunk = dS K   (reduction over this CTA's 128 keys)
          umma_f16(tmem_base + kColDQ, umma_smem_desc(adS + ks * 2048, 16384, 1024),
                   umma_smem_desc(aK + ks * 2048, 8192, 1024), id_q, ks > 0 ? 1u : 0u);
        umma_commit(&empty[st]);
        umma_commit(bar_dq);
      }
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = quad * 32 + lane;  // key row (element phase) / query row (delta and dQ phases)
    const int key = key0 + r;
    const bool key_ok = key < p.sk && !(p.kmask != nullptr && p.kmask[(long long)b * p.sk + key] != 0);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    const bool drop = p.seed != nullptr;
    uint32_t k0 = 0, k1 = 0;
    if (drop) {
      const uint64_t kk = dropout_key(p.seed, p.site);
      k0 = (uint32_t)kk;
      k1 = (uint32_t)(kk >> 32);
    }
    const uint32_t pairs_per_row = (uint32_t)(p.kb * 32);
    const uint32_t rx = (uint32_t)(r & 7);
    const bool warp_keys = key0 + quad * 32 < p.sk;  // warp-uniform: at least one of this warp's 32 keys exists
    // delta[q] = sum_d dO[q, d] * O[q, d] and the softmax statistics of every query of this (b, h), once, straight from
    // global memory: all loads are in flight together and overlap the K / V / Q tile loads and the first products
    for (int qrow = (int)threadIdx.x - 64; qrow < p.nq * 128; qrow += 256) {
      float dl = 0.f, ls = INFINITY, il = 0.f;
      if (qrow < p.sq) {
        const __nv_bfloat16* orow = p.o + (long long)qrow * p.o_ss + (long long)b * p.o_sb + h * p.d;
        const __nv_bfloat16* drow = p.dout + (long long)qrow * p.do_ss + (long long)b * p.do_sb + h * p.d;
        for (int g = 0; g < (p.d >> 3); ++g) {
          const uint4 ud = __ldg(reinterpret_cast<const uint4*>(drow) + g);
          const uint4 uo = __ldg(reinterpret_cast<const uint4*>(orow) + g);
          const uint32_t* pd_ = &ud.x;
          const uint32_t* po_ = &uo.x;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 fd = unpack_bf16(pd_[k]), fo = unpack_bf16(po_[k]);
            dl += fd.x * fo.x + fd.y * fo.y;
          }
        }
        const float2 st2 = reinterpret_cast<const float2*>(p.lse)[(long long)(b * p.h + h) * p.sq + qrow];
        ls = st2.x;
        il = 1.f / st2.y;
      }
      sDelta[qrow] = dl;
      sLse[qrow] = ls;
      sInvL[qrow] = il;
    }
    named_barrier_sync(1, 256);
    for (int c = 0; c < p.nq; ++c) {
      const int st = c;
      const int qc0 = c * 128;
      mbar_wait(bar_st, (uint32_t)(c & 1));
      tc_fence_after();
      uint8_t* pd_row = sPd + half * 16384 + r * 128;
      uint8_t* ds_row = sdS + half * 16384 + r * 128;
#pragma unroll 1
      for (int g = 0; g < 2; ++g) {
        const int qi0 = half * 64 + g * 32;
        if (!warp_keys || qc0 + qi0 >= p.sq) {
          // warp-uniform: no valid key in this warp's 32 rows, or no valid query in these 32 columns -> exact zeros
          // (rows of keys past the end are written once, in the first chunk that reaches here, and stay zero)
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const uint32_t chunk = (uint32_t)(g * 4 + q4) ^ rx;
            *reinterpret_cast<uint4*>(pd_row + (chunk << 4)) = make_uint4(0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(ds_row + (chunk << 4)) = make_uint4(0u, 0u, 0u, 0u);
          }
          continue;
        }
        uint32_t st_raw[32], dp_raw[32];
        tmem_ld_32x32(lane_addr + kColST + (uint32_t)qi0, st_raw);
        tmem_ld_32x32(lane_addr + kColDP + (uint32_t)qi0, dp_raw);
        tmem_ld_wait();
        uint32_t ppk[16], dpk[16];
        // pair index of (query row0 + i, this key) = (row0 + i) * pairs_per_row + key / 2; lanes 2m and 2m+1 hold the
        // two keys of one pair, so each computes the hash of one of two consecutive queries and they swap
        const uint32_t hbase = (uint32_t)((b * p.h + h) * p.sq + qc0 + qi0) * pairs_per_row + (uint32_t)(key >> 1) + k0;
        const float4* lse4 = reinterpret_cast<const float4*>(sLse + st * 128 + qi0);
        const float4* del4 = reinterpret_cast<const float4*>(sDelta + st * 128 + qi0);
        const float4* inv4 = reinterpret_cast<const float4*>(sInvL + st * 128 + qi0);
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 ls4 = lse4[i4], dl4 = del4[i4], il4 = inv4[i4];
          const float lsv[4] = {ls4.x, ls4.y, ls4.z, ls4.w}, dlv[4] = {dl4.x, dl4.y, dl4.z, dl4.w};
          const float ilv[4] = {il4.x, il4.y, il4.z, il4.w};
#pragma unroll
          for (int i2 = 0; i2 < 2; ++i2) {
            const int i = i4 * 4 + i2 * 2;
            bool keep0 = true, keep1 = true;
            if (drop) {
              const uint32_t mine = hash32(hbase + (uint32_t)(i + (lane & 1)) * pairs_per_row) ^ k1;
              const uint32_t other = __shfl_xor_sync(0xffffffffu, mine, 1);
              const uint32_t w0 = (lane & 1) ? other : mine, w1 = (lane & 1) ? mine : other;
              keep0 = ((lane & 1) ? (w0 >> 16) : (w0 & 0xffffu)) >= p.thr16;
              keep1 = ((lane & 1) ? (w1 >> 16) : (w1 & 0xffffu)) >= p.thr16;
            }
            // p = e / l with e exactly as the forward formed it; a probability never exceeds 1 (the clamp only acts
            // if S^T and S differ in the last bit at logit magnitudes where one ulp is worth many octaves)
            float pr0 = fminf(ex2_approx(__fmul_rn(fminf(__fsub_rn(__uint_as_float(st_raw[i]), lsv[i2 * 2]), 0.f),
                                                   p.scale_log2)) * ilv[i2 * 2], 1.f);
            float pr1 = fminf(ex2_approx(__fmul_rn(fminf(__fsub_rn(__uint_as_float(st_raw[i + 1]), lsv[i2 * 2 + 1]), 0.f),
                                                   p.scale_log2)) * ilv[i2 * 2 + 1], 1.f);
            if (!key_ok) pr0 = pr1 = 0.f;
            const float dp0 = keep0 ? __uint_as_float(dp_raw[i]) * p.keep_scale : 0.f;
            const float dp1 = keep1 ? __uint_as_float(dp_raw[i + 1]) * p.keep_scale : 0.f;
            ppk[i >> 1] = pack_bf16(keep0 ? pr0 * p.keep_scale : 0.f, keep1 ? pr1 * p.keep_scale : 0.f);
            dpk[i >> 1] = pack_bf16(pr0 * p.scale * (dp0 - dlv[i2 * 2]), pr1 * p.scale * (dp1 - dlv[i2 * 2 + 1]));
          }
        }
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const uint32_t chunk = (uint32_t)(g * 4 + q4) ^ rx;
          *reinterpret_cast<uint4*>(pd_row + (chunk << 4)) =
              make_uint4(ppk[q4 * 4], ppk[q4 * 4 + 1], ppk[q4 * 4 + 2], ppk[q4 * 4 + 3]);
          *reinterpret_cast<uint4*>(ds_row + (chunk << 4)) =
              make_uint4(dpk[q4 * 4], dpk[q4 * 4 + 1], dpk[q4 * 4 + 2], dpk[q4 * 4 + 3]);
        }
      }
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(bar_pds);
      // ---- dQ of this chunk: TMEM lane = query row; this warp owns d/2 of its columns
      mbar_wait(bar_dq, (uint32_t)(c & 1));
      tc_fence_after();
      {
        const int qrow = qc0 + r;
        const int hc = p.d >> 1;  // columns per half: 16 (d = 32) or 32 (d = 64)
        float* dst = p.dq_part + (((long long)kt * p.sq + qrow) * p.b + b) * (long long)(p.h * p.d) + h * p.d + half * hc;
        if (hc == 16) {
          uint32_t raw[16];
          tmem_ld_32x16(lane_addr + kColDQ + (uint32_t)(half * 16), raw);
          tmem_ld_wait();
          if (qrow < p.sq) {
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4)
              reinterpret_cast<float4*>(dst)[g4] =
                  make_float4(__uint_as_float(raw[g4 * 4]), __uint_as_float(raw[g4 * 4 + 1]),
                              __uint_as_float(raw[g4 * 4 + 2]), __uint_as_float(raw[g4 * 4 + 3]));
          }
        } else {
          uint32_t raw[32];
          tmem_ld_32x32(lane_addr + kColDQ + (uint32_t)(half * 32), raw);
          tmem_ld_wait();
          if (qrow < p.sq) {
#pragma unroll
            for (int g4 = 0; g4 < 8; ++g4)
              reinterpret_cast<float4*>(dst)[g4] =
                  make_float4(__uint_as_float(raw[g4 * 4]), __uint_as_float(raw[g4 * 4 + 1]),
                              __uint_as_float(raw[g4 * 4 + 2]), __uint_as_float(raw[g4 * 4 + 3]));
          }
        }
      }
      tc_fence_before();
    }
    pdl_trigger();
    // ---- dK / dV of this key tile (accumulated over all chunks): TMEM lane = key row
    {
      const int hc = p.d >> 1;
      for (int which = 0; which < 2; ++which) {
        const uint32_t col = (which == 0 ? kColDK : kColDV) + (uint32_t)(half * hc);
        __nv_bfloat16* base = which == 0 ? p.dk : p.dv;
        const long long ss = which == 0 ? p.dk_ss : p.dv_ss, sb = which == 0 ? p.dk_sb : p.dv_sb;
        __nv_bfloat16* dst = base + (long long)key * ss + (long long)b * sb + h * p.d + half * hc;
        if (hc == 16) {
          uint32_t raw[16];
          tmem_ld_32x16(lane_addr + col, raw);
          tmem_ld_wait();
          if (key < p.sk) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              uint4 u;
              u.x = pack_bf16(__uint_as_float(raw[g * 8 + 0]), __uint_as_float(raw[g * 8 + 1]));
              u.y = pack_bf16(__uint_as_float(raw[g * 8 + 2]), __uint_as_float(raw[g * 8 + 3]));
              u.z = pack_bf16(__uint_as_float(raw[g * 8 + 4]), __uint_as_float(raw[g * 8 + 5]));
              u.w = pack_bf16(__uint_as_float(raw[g * 8 + 6]), __uint_as_float(raw[g * 8 + 7]));
              reinterpret_cast<uint4*>(dst)[g] = u;
            }
          }
        } else {
          uint32_t raw[32];
          tmem_ld_32x32(lane_addr + col, raw);
          tmem_ld_wait();
          if (key < p.sk) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 u;
              u.x = pack_bf16(__uint_as_float(raw[g * 8 + 0]), __uint_as_float(raw[g * 8 + 1]));
              u.y = pack_bf16(__uint_as_float(raw[g * 8 + 2]), __uint_as_float(raw[g * 8 + 3]));
              u.z = pack_bf16(__uint_as_float(raw[g * 8 + 4]), __uint_as_float(raw[g * 8 + 5]));
              u.w = pack_bf16(__uint_as_float(raw[g * 8 + 6]), __uint_as_float(raw[g * 8 + 7]));
              reinterpret_cast<uint4*>(dst)[g] = u;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// dq[q, b, e] (bf16, strided) = sum over key tiles of dq_part[kt][q][b][e], 8 elements per thread, fixed order
__global__ void attn_dq_reduce_kernel(const float* __restrict__ part, __nv_bfloat16* __restrict__ dq, long long dq_ss,
                                      long long dq_sb, int sq, int b, int e, int kt) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int e8 = e >> 3;
  const long long total = (long long)sq * b * e8;
  if (i >= total) return;
  const int c = (int)(i % e8) * 8;
  const long long qb = i / e8;
  const int bb = (int)(qb % b);
  const long long q = qb / b;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  const long long plane = (long long)sq * b * e;
  for (int t = 0; t < kt; ++t) {
    const float4* src = reinterpret_cast<const float4*>(part + t * plane + qb * e + c);
    const float4 a = __ldg(src), b4 = __ldg(src + 1);
    acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
    acc[4] += b4.x; acc[5] += b4.y; acc[6] += b4.z; acc[7] += b4.w;
  }
  uint4 u;
  u.x = pack_bf16(acc[0], acc[1]);
  u.y = pack_bf16(acc[2], acc[3]);
  u.z = pack_bf16(acc[4], acc[5]);
  u.w = pack_bf16(acc[6], acc[7]);
  *reinterpret_cast<uint4*>(dq + q * dq_ss + (long long)bb * dq_sb + c) = u;
}

// keep[b, h, q, k] in {0, 1}: the dropout decisions of the fused kernels, for tests
__global__ void attn_dropout_mask_kernel(uint8_t* __restrict__ keep, int bh, int sq, int sk, int kb,
                                         const unsigned long long* __restrict__ seed, uint32_t site, uint32_t thr16) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)bh * sq * sk;
  if (i >= total) return;
  const int k = (int)(i % sk);
  const long long row = i / sk;
  const uint64_t key = dropout_key(seed, site);
  const uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
  const uint32_t w = hash32((uint32_t)row * (uint32_t)(kb * 32) + (uint32_t)(k >> 1) + k0) ^ k1;
  const uint32_t v16 = (k & 1) ? (w >> 16) : (w & 0xffffu);
  keep[i] = v16 >= thr16 ? 1 : 0;
}

static int pow2_cols(int need) {
  int c = 32;
  while (c < need) c <<= 1;
  return c;
}

static int make_qkv_map(CUtensorMap* m, const void* ptr, int d, int s, int h, int b, int64_t ss, int64_t sb, int rows) {
  const int64_t dim[4] = {d, s, h, b};
  const int64_t str[4] = {1, ss, d, sb};
  const uint32_t box[4] = {64, (uint32_t)rows, 1, 1};
  const uint32_t ones[4] = {1, 1, 1, 1};
  return encode_tmap_bf16_4d(m, ptr, dim, str, box, ones);
}

static uint32_t thr16_of(float p) {
  long v = lroundf(p * 65536.f);
  if (v < 1) v = 1;
  if (v > 65535) v = 65535;
  return (uint32_t)v;
}

}  // namespace toist

using namespace toist;

extern "C" size_t toist_sizeof_attn_desc(void) { return sizeof(toist_attn_desc); }
extern "C" size_t toist_sizeof_attn_bwd_desc(void) { return sizeof(toist_attn_bwd_desc); }

extern "C" int toist_attention_supported(int32_t sq, int32_t sk, int32_t d) {
  return (d == 32 || d == 64) && sk >= 1 && sk <= kAttMaxKB * 64 && sq >= 1 && sq <= kBwdMaxNQ * 128 ? 1 : 0;
}

extern "C" int toist_attention_fwd(const toist_attn_desc* a, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  TOIST_REQUIRE(a != nullptr && a->q && a->k && a->v && a->out, "toist_attention_fwd: null pointer");
  TOIST_REQUIRE(toist_attention_supported(a->sq, a->sk, a->d), "toist_attention_fwd: unsupported shape sq=%d sk=%d d=%d",
                a->sq, a->sk, a->d);
  TOIST_REQUIRE(a->p_drop == 0.f || (a->seed != nullptr && a->p_drop > 0.f && a->p_drop < 1.f),
                "toist_attention_fwd: dropout needs a seed and 0 < p < 1");
  TOIST_REQUIRE(a->o_ss % 8 == 0 && a->o_sb % 8 == 0 && (reinterpret_cast<uintptr_t>(a->out) & 15) == 0,
                "toist_attention_fwd: output must be 16-byte aligned");
  AttnFwdParams p;
  memset(&p, 0, sizeof(p));
  p.sq = a->sq; p.sk = a->sk; p.b = a->b; p.h = a->h; p.d = a->d;
  p.kb = (a->sk + 63) / 64;
  p.scale = 1.f / sqrtf((float)a->d);
  p.scale_log2 = p.scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__nv_bfloat16*>(a->out);
  p.o_ss = a->o_ss; p.o_sb = a->o_sb;
  p.lse = a->lse;
  p.kmask = a->key_mask;
  p.keep_scale = 1.f;
  if (a->p_drop > 0.f) {
    p.seed = reinterpret_cast<const unsigned long long*>(a->seed);
    p.site = a->site;
    p.thr16 = thr16_of(a->p_drop);
    p.keep_scale = 1.f / (1.f - (float)p.thr16 / 65536.f);
  }
  p.tmem_cols = pow2_cols(p.kb * 64 + a->d);
  CUtensorMap mq, mk, mv;
  int rc;
  if ((rc = make_qkv_map(&mq, a->q, a->d, a->sq, a->h, a->b, a->q_ss, a->q_sb, 128)) != TOIST_OK) return rc;
  if ((rc = make_qkv_map(&mk, a->k, a->d, a->sk, a->h, a->b, a->k_ss, a->k_sb, 64)) != TOIST_OK) return rc;
  if ((rc = make_qkv_map(&mv, a->v, a->d, a->sk, a->h, a->b, a->v_ss, a->v_sb, 64)) != TOIST_OK) return rc;
  const int smem = 16384 + 2 * p.kb * 8192 + kAttNP * 16384 + 2048 + 64 + 16 * 8 + 16 + 1024;
  static bool configured = false;
  if (!configured) {
    TOIST_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          16384 + 2 * kAttMaxKB * 8192 + kAttNP * 16384 + 2048 + 64 + 16 * 8 + 16 + 1024));
    configured = true;
  }
  dim3 grid((unsigned)((a->sq + 127) / 128), (unsigned)a->h, (unsigned)a->b);
  TOIST_CHECK_CUDA(launch_pdl(attn_fwd_kernel, grid, dim3(kFwdThreads), (size_t)smem, stream, mq, mk, mv, p));
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

extern "C" int64_t toist_attention_bwd_workspace(int32_t sq, int32_t sk, int32_t b, int32_t h, int32_t d) {
  return (int64_t)((sk + 127) / 128) * sq * b * h * d * (int64_t)sizeof(float);
}

extern "C" int toist_attention_bwd(const toist_attn_bwd_desc* a, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  TOIST_REQUIRE(a != nullptr, "toist_attention_bwd: null descriptor");
  const toist_attn_desc* f = &a->fwd;
  TOIST_REQUIRE(f->q && f->k && f->v && f->out && f->lse && a->dout && a->dq && a->dk && a->dv && a->workspace,
                "toist_attention_bwd: null pointer");
  TOIST_REQUIRE(toist_attention_supported(f->sq, f->sk, f->d), "toist_attention_bwd: unsupported shape sq=%d sk=%d d=%d",
                f->sq, f->sk, f->d);
  TOIST_REQUIRE(f->p_drop == 0.f || (f->seed != nullptr && f->p_drop > 0.f && f->p_drop < 1.f),
                "toist_attention_bwd: dropout needs a seed and 0 < p < 1");
  const int e = f->h * f->d;
  TOIST_REQUIRE(a->dq_ss % 8 == 0 && a->dq_sb % 8 == 0 && a->dk_ss % 8 == 0 && a->dk_sb % 8 == 0 && a->dv_ss % 8 == 0 &&
                    a->dv_sb % 8 == 0 && f->o_ss % 8 == 0 && f->o_sb % 8 == 0 && a->do_ss % 8 == 0 && a->do_sb % 8 == 0,
                "toist_attention_bwd: strides must be multiples of 8 elements");
  AttnBwdParams p;
  memset(&p, 0, sizeof(p));
  p.sq = f->sq; p.sk = f->sk; p.b = f->b; p.h = f->h; p.d = f->d;
  p.kb = (f->sk + 63) / 64;
  p.nq = (f->sq + 127) / 128;
  p.scale = 1.f / sqrtf((float)f->d);
  p.scale_log2 = p.scale * 1.4426950408889634f;
  p.o = reinterpret_cast<const __nv_bfloat16*>(f->out);
  p.o_ss = f->o_ss; p.o_sb = f->o_sb;
  p.dout = reinterpret_cast<const __nv_bfloat16*>(a->dout);
  p.do_ss = a->do_ss; p.do_sb = a->do_sb;
  p.lse = f->lse;
  p.kmask = f->key_mask;
  p.keep_scale = 1.f;
  if (f->p_drop > 0.f) {
    p.seed = reinterpret_cast<const unsigned long long*>(f->seed);
    p.site = f->site;
    p.thr16 = thr16_of(f->p_drop);
    p.keep_scale = 1.f / (1.f - (float)p.thr16 / 65536.f);
  }
  p.dq_part = reinterpret_cast<float*>(a->workspace);
  p.dk = reinterpret_cast<__nv_bfloat16*>(a->dk); p.dk_ss = a->dk_ss; p.dk_sb = a->dk_sb;
  p.dv = reinterpret_cast<__nv_bfloat16*>(a->dv); p.dv_ss = a->dv_ss; p.dv_sb = a->dv_sb;
  CUtensorMap mq, mk, mv, mdo;
  int rc;
  if ((rc = make_qkv_map(&mq, f->q, f->d, f->sq, f->h, f->b, f->q_ss, f->q_sb, 128)) != TOIST_OK) return rc;
  if ((rc = make_qkv_map(&mk, f->k, f->d, f->sk, f->h, f->b, f->k_ss, f->k_sb, 128)) != TOIST_OK) return rc;
  if ((rc = make_qkv_map(&mv, f->v, f->d, f->sk, f->h, f->b, f->v_ss, f->v_sb, 128)) != TOIST_OK) return rc;
  if ((rc = make_qkv_map(&mdo, a->dout, f->d, f->sq, f->h, f->b, a->do_ss, a->do_sb, 128)) != TOIST_OK) return rc;
  const int smem = 32768 + 65536 + 65536 + 3 * kBwdMaxNQ * 512 + 8 * 8 + 16 + 1024;
  static bool configured = false;
  if (!configured) {
    TOIST_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int kt = (f->sk + 127) / 128;
  dim3 grid((unsigned)kt, (unsigned)f->h, (unsigned)f->b);
  TOIST_CHECK_CUDA(launch_pdl(attn_bwd_kernel, grid, dim3(kBwdThreads), (size_t)smem, stream, mq, mk, mv, mdo, p));
  const long long total = (long long)f->sq * f->b * (e / 8);
  TOIST_CHECK_CUDA(launch_pdl(attn_dq_reduce_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream,
                              (const float*)p.dq_part, reinterpret_cast<__nv_bfloat16*>(a->dq), (long long)a->dq_ss,
                              (long long)a->dq_sb, f->sq, f->b, e, kt));
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

extern "C" int toist_attention_dropout_mask(uint8_t* keep, int32_t b, int32_t h, int32_t sq, int32_t sk, float p_drop,
                                            const uint64_t* seed, uint32_t site, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  TOIST_REQUIRE(keep && seed && p_drop > 0.f && p_drop < 1.f, "toist_attention_dropout_mask: bad arguments");
  const long long total = (long long)b * h * sq * sk;
  TOIST_CHECK_CUDA(launch_pdl(attn_dropout_mask_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, keep,
                              b * h, sq, sk, (sk + 63) / 64, (const unsigned long long*)seed, site, thr16_of(p_drop)));
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}
