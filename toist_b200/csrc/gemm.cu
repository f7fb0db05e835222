// Implicit-GEMM engine for sm_100a: TMA (SWIZZLE_128B, 4-D shifted boxes with hardware zero fill) -> shared memory
// -> tcgen05.mma (M = 128, fp32 accumulators in TMEM) -> tcgen05.ld -> fused epilogue -> global.
//
// Warp roles per CTA (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..5 = epilogue (warp w reads TMEM lane quadrant w % 4).  One CTA computes one 128 x BN output tile; two
// CTAs are co-resident per SM for BN <= 128 so one tile's epilogue overlaps the other's main loop.
//
// See include/toist_b200.h for the three traversal modes (FWD / DGRAD / WGRAD) and the epilogue contract.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "host_util.h"

namespace toist {

constexpr int kBM = 128;          // rows per tile == TMEM lanes
constexpr int kBK = 64;           // reduction elements per pipeline stage (64 bf16 = one 128-byte swizzle row)
constexpr int kABytes = kBM * 128;
constexpr int kThreads = 192;

struct GemmKParams {
  int ext_x, ext_y, ext_n;
  int tile_x, tile_y, tile_n;
  int tiles_x, tiles_y, tiles_n;
  int stride_x, stride_y;
  int n_cols, m_rows;
  int kblocks;  // FWD/DGRAD: 64-wide reduction blocks per tap
  int n_taps;
  int b_batched;
  int batch_y, batch_n, splits;
  int n_tiles;  // WGRAD: number of BN-wide column tiles (grid.y = n_taps * n_tiles)
  void* out;
  int out_dtype;
  long long out_sx, out_sy, out_sn;
  float alpha;
  const float* col_scale;
  const float* col_shift;
  const float* row_scale;
  const void* res;
  int res_dtype;
  const void* mask;
  void* aux;
  int act;
  int accumulate;
  int vec_ok;   // 1: every row base and column tile is 16-byte aligned for all epilogue pointers
  int slabs;    // persistent kernel: 16 KB slabs of the epilogue ring
  int stages;   // depth of the TMA -> MMA shared-memory ring
  int epi;      // epilogue flavour (template parameter EPI of the kernel)
  int cluster;  // 2: CTA pairs share the B tile by TMA multicast (PAIR 1); 3: cta_group::2 MMA over the pair (PAIR 2)
  long long* trace;  // measurement hook (toist_debug_gemm_trace): 8 clock stamps per CTA, nullptr in production
  toist_tap taps[TOIST_MAX_TAPS];
};

// Phase stamps of one CTA (tools/gemm_trace.py): 0 entry, 1 setup done (barriers, TMEM, cluster syncs), 2 dependency wait
// done, 3 first operand stage landed, 4 last MMA issued, 5 accumulator complete, 6 epilogue tile loop done, 7 store done
__device__ __forceinline__ void trace_stamp(const GemmKParams& p, int slot) {
  if (p.trace != nullptr) {
    const long long cta = (long long)blockIdx.x + (long long)gridDim.x * (blockIdx.y + (long long)gridDim.y * blockIdx.z);
    long long t;
    if (slot >= 8) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));  // slots 8 / 9: entry / exit in ns, comparable across SMs
    else t = clock64();
    p.trace[cta * 16 + slot] = t;
  }
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == TOIST_ACT_RELU) return fmaxf(v, 0.f);
  if (act == TOIST_ACT_GELU) return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
  if (act == TOIST_ACT_SIGMOID) return 1.f / (1.f + __expf(-v));
  return v;
}

// EPI: 0 = bf16 tile through shared memory + TMA (res / mask by TMA load; act none / relu)
//      1 = fp32 direct stores (alpha, column / row scaling, activation, optional atomic accumulate)
//      2 = generic direct path (any dtype, aux output, every activation, unaligned shapes)
// One epilogue per instantiation keeps the SASS small: the tiles are short-lived and instruction fetch of a
// 100+ KB kernel showed up as the second largest stall reason (profiles/r01_ncu_gemm_notes.md).
// CL = 2: the two CTAs of a cluster compute neighbouring M tiles of the same N tile; each loads its own A tile and
// HALF of the shared B tile, multicast to both (the B = weight traffic from L2 per SM halves: the layer3 convolutions
// are bound by the ~50 B/clk an SM can pull from L2, two thirds of which was B).
// PAIR = 2 (EXPERIMENTAL, TOIST_GEMM_2SM=1, written at the end of round 1 and not yet run on a GPU): the CTA pair issues
// ONE tcgen05.mma.cta_group::2 per k-step over a 256 x BN tile; each CTA stages its own 128 A rows and only BN / 2 rows of
// B, so an SM ingests 32 KB instead of 48 KB per 512 MMA clocks (BN = 256) - the measured bound of the layer3 convolutions.
template <int BN, int MODE, int EPI, int PAIR>
__global__ void __launch_bounds__(kThreads, BN == 256 ? 1 : 3)  // BN <= 128: up to three CTAs per SM (<= 112 registers)
gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
            const __grid_constant__ CUtensorMap tma_out, const __grid_constant__ CUtensorMap tma_res,
            const __grid_constant__ CUtensorMap tma_mask, const __grid_constant__ GemmKParams p) {
  constexpr int CL = PAIR ? 2 : 1;          // CTAs per cluster
  constexpr bool kTwoSm = (PAIR == 2);      // cta_group::2 MMA (PAIR == 1: independent MMAs, multicast B)
  const int STAGES = p.stages;
  constexpr int kBBytes = (kTwoSm ? BN / 2 : BN) * 128;
  constexpr int kStageBytes = kABytes + kBBytes;
  constexpr bool kAMN = (MODE == TOIST_GEMM_WGRAD);
  constexpr bool kBMN = (MODE != TOIST_GEMM_FWD);

  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment; dynamic smem only guarantees 16.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint64_t* epi_bar = accum_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(epi_bar + 1);
  float* s_col = reinterpret_cast<float*>(smem + STAGES * kStageBytes + 256);  // [2][BN] per-column scale | shift (EPI 0)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    trace_stamp(p, 0);
    trace_stamp(p, 8);
  }

  // ---------------- tile decode
  int x0 = 0, y0 = 0, i0 = 0;  // FWD/DGRAD: pixel-tile origin
  int n0 = 0, m0 = 0;          // column / row origin of the output tile
  int tap_w = 0;               // WGRAD: tap handled by this CTA
  int by = 0, bn = 0;          // WGRAD batch coordinates
  int it_begin = 0, it_end = 0;
  if constexpr (MODE != TOIST_GEMM_WGRAD) {
    int t = blockIdx.x;
    const int tx = t % p.tiles_x;
    t /= p.tiles_x;
    const int ty = t % p.tiles_y;
    const int tn = t / p.tiles_y;
    x0 = tx * p.tile_x;
    y0 = ty * p.tile_y;
    i0 = tn * p.tile_n;
    n0 = blockIdx.y * BN;
    it_begin = 0;
    it_end = p.n_taps * p.kblocks;
  } else {
    m0 = blockIdx.x * kBM;
    tap_w = blockIdx.y / p.n_tiles;
    n0 = (blockIdx.y % p.n_tiles) * BN;
    int z = blockIdx.z;
    const int split = z % p.splits;
    z /= p.splits;
    by = z % p.batch_y;
    bn = z / p.batch_y;
    const int total = p.tiles_x * p.tiles_y * p.tiles_n;
    const int per = (total + p.splits - 1) / p.splits;
    it_begin = split * per;
    it_end = min(total, it_begin + per);
  }
  const int n_iters = it_end - it_begin;

  // ---------------- one-time setup
  if constexpr (kTwoSm) cluster_sync();  // both CTAs of the pair are resident before the paired TMEM allocation
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int s = 0; s < STAGES; ++s) {
      // PAIR 2: the leader's full barrier collects its own expect_tx arrival and the peer's "loads issued" arrival
      mbar_init(&full_bar[s], kTwoSm ? 2 : 1);
      // PAIR 1: every CTA that received the stage's multicast B releases it; PAIR 2: one multicast commit of the leader
      mbar_init(&empty_bar[s], PAIR == 1 ? 2 : 1);
    }
    mbar_init(accum_bar, 1);
    mbar_init(epi_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    if constexpr (kTwoSm) {
      tmem_alloc_2sm(tmem_slot, BN);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_slot, BN);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync();  // the peer's barriers exist before anything is multicast to or signalled on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t cl_rank = CL > 1 ? cluster_ctarank() : 0u;
  constexpr uint16_t kClMask = (uint16_t)((1u << CL) - 1u);
  if (threadIdx.x == 0) trace_stamp(p, 1);
  pdl_wait();  // barrier init, descriptor prefetch and the TMEM allocation above overlap the previous kernel's tail
  if (threadIdx.x == 0) trace_stamp(p, 2);

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = it_begin; it < it_end; ++it) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * kStageBytes;
        uint8_t* sb = sa + kABytes;
        if constexpr (kTwoSm) {
          if (cl_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * kStageBytes);  // both CTAs' pieces land on this barrier
        } else {
          mbar_expect_tx(&full_bar[stage], kStageBytes);
        }
        if constexpr (kTwoSm) {
          // every load completes on the LEADER's full barrier; this CTA stages its A rows and its half of B
          const int t = it / p.kblocks;
          const int kb = it - t * p.kblocks;
          const toist_tap tp = p.taps[t];
          tma_load_4d_2sm(sa, &tma_a, &full_bar[stage], kb * kBK, x0 * p.stride_x + tp.dx, y0 * p.stride_y + tp.dy,
                          i0 + tp.dn);
          if constexpr (MODE == TOIST_GEMM_FWD) {
            tma_load_4d_2sm(sb, &tma_b, &full_bar[stage], tp.col + kb * kBK, n0 + (int)cl_rank * (BN / 2), 0, 0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 128; ++j)
              tma_load_4d_2sm(sb + j * 8192, &tma_b, &full_bar[stage], tp.col + n0 + ((int)cl_rank * (BN / 128) + j) * 64,
                              kb * kBK, 0, 0);
          }
          if (cl_rank != 0) mbar_arrive_leader(&full_bar[stage]);
        } else if constexpr (MODE != TOIST_GEMM_WGRAD) {
          const int t = it / p.kblocks;
          const int kb = it - t * p.kblocks;
          const toist_tap tp = p.taps[t];
          tma_load_4d(sa, &tma_a, &full_bar[stage], kb * kBK, x0 * p.stride_x + tp.dx, y0 * p.stride_y + tp.dy,
                      i0 + tp.dn);
          const int cy = p.b_batched ? y0 : 0;
          const int cn = p.b_batched ? i0 : 0;
          if constexpr (CL > 1) {  // this CTA fetches 1/CL of the B tile for the whole cluster
            if constexpr (MODE == TOIST_GEMM_FWD) {
              constexpr int kRows = BN / CL;  // tma_b's box holds BN / CL rows
              tma_load_4d_mc(sb + cl_rank * kRows * 128, &tma_b, &full_bar[stage], tp.col + kb * kBK,
                             n0 + (int)cl_rank * kRows, cy, cn, kClMask);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64 / CL; ++j) {
                const int jj = (int)cl_rank * (BN / 64 / CL) + j;
                tma_load_4d_mc(sb + jj * 8192, &tma_b, &full_bar[stage], tp.col + n0 + jj * 64, kb * kBK, cy, cn, kClMask);
              }
            }
          } else if constexpr (MODE == TOIST_GEMM_FWD) {
            tma_load_4d(sb, &tma_b, &full_bar[stage], tp.col + kb * kBK, n0, cy, cn);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_4d(sb + j * 8192, &tma_b, &full_bar[stage], tp.col + n0 + j * 64, kb * kBK, cy, cn);
          }
        } else {
          int t = it;
          const int px = t % p.tiles_x;
          t /= p.tiles_x;
          const int py = t % p.tiles_y;
          const int pn = t / p.tiles_y;
          const toist_tap tp = p.taps[tap_w];
          const int ax = px * p.tile_x, ay = py * p.tile_y, an = pn * p.tile_n;
#pragma unroll
          for (int j = 0; j < 2; ++j)
            tma_load_4d(sa + j * 8192, &tma_a, &full_bar[stage], m0 + j * 64, ax, ay + by, an + bn);
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            tma_load_4d(sb + j * 8192, &tma_b, &full_bar[stage], n0 + j * 64, ax * p.stride_x + tp.dx,
                        ay * p.stride_y + tp.dy + by, an + tp.dn + bn);
        }
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (PAIR 2: the leader CTA issues for both) =======================
    constexpr uint32_t idesc = umma_idesc_bf16(BN, kAMN, kBMN, kTwoSm ? 256 : 128);
    int stage = 0;
    uint32_t phase = 0;
    const int mma_iters = (kTwoSm && cl_rank != 0) ? 0 : n_iters;
    for (int it = 0; it < mma_iters; ++it) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (it == 0 && lane == 0) trace_stamp(p, 3);
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + stage * kStageBytes);
        const uint32_t sb = sa + kABytes;
#pragma unroll
        for (int ks = 0; ks < kBK / 16; ++ks) {
          const uint64_t da = kAMN ? umma_smem_desc(sa + ks * 2048, 8192, 1024) : umma_smem_desc(sa + ks * 32, 16, 1024);
          const uint64_t db = kBMN ? umma_smem_desc(sb + ks * 2048, 8192, 1024) : umma_smem_desc(sb + ks * 32, 16, 1024);
          if constexpr (kTwoSm) umma_f16_2sm(tmem_base, da, db, idesc, (it > 0 || ks > 0) ? 1u : 0u);
          else umma_f16(tmem_base, da, db, idesc, (it > 0 || ks > 0) ? 1u : 0u);
        }
        if constexpr (kTwoSm) {
          umma_commit_2sm(&empty_bar[stage], kClMask);                       // frees the stage in both CTAs
          if (it == n_iters - 1) umma_commit_2sm(accum_bar, kClMask);        // both CTAs' epilogues may start
        } else {
          if constexpr (CL > 1) umma_commit_mc(&empty_bar[stage], kClMask);  // ... in every CTA that fills it
          else umma_commit(&empty_bar[stage]);            // frees the smem stage once these MMAs retire
          if (it == n_iters - 1) umma_commit(accum_bar);  // accumulator complete
        }
      }
      __syncwarp();
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
    if (lane == 0 && mma_iters > 0) trace_stamp(p, 4);
  } else {
    // ======================= epilogue (4 warps, one TMEM lane quadrant each) =======================
    const int q = warp & 3;
    const int r = q * 32 + lane;  // row of the tile owned by this thread
    bool row_ok;
    long long row_off;
    int row_global;  // index for row_scale
    if constexpr (MODE != TOIST_GEMM_WGRAD) {
      const int dx = r % p.tile_x;
      const int t2 = r / p.tile_x;
      const int dy = t2 % p.tile_y;
      const int dn = t2 / p.tile_y;
      const int x = x0 + dx, y = y0 + dy, n = i0 + dn;
      row_ok = (x < p.ext_x) && (y < p.ext_y) && (n < p.ext_n);
      row_off = (long long)x * p.out_sx + (long long)y * p.out_sy + (long long)n * p.out_sn;
      row_global = 0;
    } else {
      const int m = m0 + r;
      row_ok = m < p.m_rows;
      row_off = (long long)m * p.out_sx + (long long)by * p.out_sy + (long long)bn * p.out_sn + p.taps[tap_w].col;
      row_global = m;
    }
    const float rscale = (p.row_scale != nullptr && row_ok) ? __ldg(p.row_scale + row_global) : 1.f;

    // EPI 0: per-column constants of this tile -> shared memory while the main loop runs (alpha folded into the scale)
    float* s_scale = s_col;
    float* s_shift = s_col + BN;
    const bool has_scale = p.col_scale != nullptr || p.alpha != 1.f;
    if constexpr (EPI == 0) {
      const int last = p.n_cols - 1;  // columns past n_cols are clipped by the TMA store: clamp, do not branch
      for (int c = (int)threadIdx.x - 64; c < BN; c += 128) {
        const int col = min(n0 + c, last);
        s_scale[c] = p.alpha * (p.col_scale != nullptr ? __ldg(p.col_scale + col) : 1.f);
        s_shift[c] = p.col_shift != nullptr ? __ldg(p.col_shift + col) : 0.f;
      }
      named_barrier_sync(1, 128);
    }

    if (n_iters > 0) {
      mbar_wait(accum_bar, 0);
      tc_fence_after();
    }
    if (warp == 2 && lane == 0) trace_stamp(p, 5);
    // Main loop done: let the next kernel's CTAs be scheduled into the SM slots this grid frees from here on (its
    // prologue then overlaps our epilogue; it still waits for this grid's completion before touching memory).
    // Triggering at kernel entry instead made many-wave grids slower (39 -> 51 us on the 1600-CTA layer1 conv).
    pdl_trigger();
    if constexpr (EPI == 0) {
      {
        // ---- bf16 epilogue staged through shared memory.  The accumulator barrier implies that every MMA has
        // retired, so the operand ring is idle and is reused: [out | res | mask], each BN/64 slabs of 128 rows x 128 B
        // in the SWIZZLE_128B layout (16-byte chunk index XOR row % 8), which makes the per-row accesses of the 128
        // epilogue threads bank-conflict free and lets TMA move whole tiles with full-line transactions.
        const bool has_res = p.res != nullptr, has_mask = p.mask != nullptr;
        // The output tile is written IN PLACE over the residual tile (or over the mask tile when there is no
        // residual): every thread reads its 64-byte pieces of res / mask before it writes the same addresses of out,
        // and nobody else touches them.  Staging is therefore max(1, res + mask) tiles, which lets short reductions
        // run with a 2-deep ring (dispatch_bn) and three CTAs per SM.
        uint8_t* st_out = smem;
        uint8_t* st_res = smem;
        uint8_t* st_mask = smem + (has_res ? BN * 256 : 0);
        const bool leader = (warp == 2 && lane == 0);
        if (has_res || has_mask) {
          if (leader) {
            mbar_expect_tx(epi_bar, (uint32_t)((has_res ? 1 : 0) + (has_mask ? 1 : 0)) * BN * 256);
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) {
              if (has_res) tma_load_4d(st_res + j * 16384, &tma_res, epi_bar, n0 + j * 64, x0, y0, i0);
              if (has_mask) tma_load_4d(st_mask + j * 16384, &tma_mask, epi_bar, n0 + j * 64, x0, y0, i0);
            }
          }
          mbar_wait(epi_bar, 0);
        }
        // One epilogue warp per scheduler: nothing hides a dependent instruction's latency except the warp's own
        // independent work (measured with toist_debug_gemm_trace: ~1000 clocks per 32-column group when every group
        // waited for its TMEM load and fetched 64 per-column constants through L1; the 128 x 256 tile's epilogue was
        // 30 % of the CTA's lifetime).  Hence: the per-column constants come from shared memory (staged during the
        // main loop), the TMEM load of group c+1 is in flight while group c is processed, and each 64-column slab is
        // handed to the TMA store as soon as it is complete.
        const uint32_t rx = (uint32_t)(r & 7);
        const int nch = min(BN / 16, (p.n_cols - n0 + 15) >> 4);  // warp-uniform: 16-column groups with valid columns
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        const bool relu = p.act == TOIST_ACT_RELU;
        auto process = [&](uint32_t (&raw)[16], int c0) {  // 16 columns of this thread's row: raw -> bf16 in st_out
          const float4* sc4 = reinterpret_cast<const float4*>(s_scale + c0);
          const float4* sh4 = reinterpret_cast<const float4*>(s_shift + c0);
          float v[16];
          if (has_scale) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const float4 a = sc4[g], b = sh4[g];
              v[g * 4 + 0] = fmaf(__uint_as_float(raw[g * 4 + 0]), a.x, b.x);
              v[g * 4 + 1] = fmaf(__uint_as_float(raw[g * 4 + 1]), a.y, b.y);
              v[g * 4 + 2] = fmaf(__uint_as_float(raw[g * 4 + 2]), a.z, b.z);
              v[g * 4 + 3] = fmaf(__uint_as_float(raw[g * 4 + 3]), a.w, b.w);
            }
          } else {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const float4 b = sh4[g];
              v[g * 4 + 0] = __uint_as_float(raw[g * 4 + 0]) + b.x;
              v[g * 4 + 1] = __uint_as_float(raw[g * 4 + 1]) + b.y;
              v[g * 4 + 2] = __uint_as_float(raw[g * 4 + 2]) + b.z;
              v[g * 4 + 3] = __uint_as_float(raw[g * 4 + 3]) + b.w;
            }
          }
          const uint32_t row_base = (uint32_t)(c0 >> 6) * 16384u + (uint32_t)r * 128u;
          const uint32_t cb = (uint32_t)(c0 & 63) >> 3;  // first 16-byte chunk of this 16-column group (0, 2, 4, 6)
          if (has_res) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              const uint4 u = *reinterpret_cast<const uint4*>(st_res + row_base + (((cb + g) ^ rx) << 4));
              const float2 f0 = unpack_bf16(u.x), f1 = unpack_bf16(u.y), f2 = unpack_bf16(u.z), f3 = unpack_bf16(u.w);
              v[g * 8 + 0] += f0.x; v[g * 8 + 1] += f0.y; v[g * 8 + 2] += f1.x; v[g * 8 + 3] += f1.y;
              v[g * 8 + 4] += f2.x; v[g * 8 + 5] += f2.y; v[g * 8 + 6] += f3.x; v[g * 8 + 7] += f3.y;
            }
          }
          if (has_mask) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              const uint4 u = *reinterpret_cast<const uint4*>(st_mask + row_base + (((cb + g) ^ rx) << 4));
              const float2 f0 = unpack_bf16(u.x), f1 = unpack_bf16(u.y), f2 = unpack_bf16(u.z), f3 = unpack_bf16(u.w);
              if (!(f0.x > 0.f)) v[g * 8 + 0] = 0.f;
              if (!(f0.y > 0.f)) v[g * 8 + 1] = 0.f;
              if (!(f1.x > 0.f)) v[g * 8 + 2] = 0.f;
              if (!(f1.y > 0.f)) v[g * 8 + 3] = 0.f;
              if (!(f2.x > 0.f)) v[g * 8 + 4] = 0.f;
              if (!(f2.y > 0.f)) v[g * 8 + 5] = 0.f;
              if (!(f3.x > 0.f)) v[g * 8 + 6] = 0.f;
              if (!(f3.y > 0.f)) v[g * 8 + 7] = 0.f;
            }
          }
          if (relu) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
          }
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            uint4 u;
            u.x = pack_bf16(v[g * 8 + 0], v[g * 8 + 1]);
            u.y = pack_bf16(v[g * 8 + 2], v[g * 8 + 3]);
            u.z = pack_bf16(v[g * 8 + 4], v[g * 8 + 5]);
            u.w = pack_bf16(v[g * 8 + 6], v[g * 8 + 7]);
            *reinterpret_cast<uint4*>(st_out + row_base + (((cb + g) ^ rx) << 4)) = u;
          }
        };
        uint32_t raw_a[16], raw_b[16];
        if (n_iters > 0) {
          tmem_ld_32x16(t_addr, raw_a);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) raw_a[i] = raw_b[i] = 0u;
        }
#pragma unroll 1
        for (int c = 0; c < nch; c += 2) {
          const bool more1 = c + 1 < nch, more2 = c + 2 < nch;
          if (more1 && n_iters > 0) tmem_ld_32x16(t_addr + (uint32_t)((c + 1) * 16), raw_b);
          process(raw_a, c * 16);
          if (more1) {
            if (n_iters > 0) {
              tmem_ld_wait();
              if (more2) tmem_ld_32x16(t_addr + (uint32_t)((c + 2) * 16), raw_a);
            }
            process(raw_b, (c + 1) * 16);
            if (more2 && n_iters > 0) tmem_ld_wait();
          }
          if ((c & 2) != 0 || !more2) {  // the 64-column slab c / 4 is complete: hand it to the TMA store
            fence_proxy_async();         // generic-proxy writes above -> visible to the TMA (async proxy) reads below
            named_barrier_sync(1, 128);  // the four epilogue warps
            if (leader) tma_store_4d(&tma_out, st_out + (c >> 2) * 16384, n0 + (c >> 2) * 64, x0, y0, i0);
          }
        }
        if (leader) {
          trace_stamp(p, 6);
          tma_store_commit();
          tma_store_wait_read();
          trace_stamp(p, 7);
        }
      }
    }
    if constexpr (EPI == 1) {
      // ---- fp32 outputs (weight gradients, attention scores, prediction heads): direct stores / atomics
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        if (n0 + c0 >= p.n_cols) break;  // warp-uniform
        uint32_t raw[32];
        if (n_iters > 0) {
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, raw);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) raw[i] = 0u;
        }
        if (!row_ok) continue;
        const int ncol = n0 + c0;
        const int last = p.n_cols - 1;
        const int nvalid = min(32, p.n_cols - ncol);
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]) * p.alpha;
        if (p.col_scale != nullptr) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] *= __ldg(p.col_scale + min(ncol + i, last));
        }
        if (p.col_shift != nullptr) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += __ldg(p.col_shift + min(ncol + i, last));
        }
        if (p.row_scale != nullptr) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] *= rscale;
        }
        if (p.act == TOIST_ACT_RELU) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
        } else if (p.act == TOIST_ACT_SIGMOID) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 1.f / (1.f + __expf(-v[i]));
        }
        float* op = reinterpret_cast<float*>(p.out) + row_off + ncol;
        if (p.accumulate) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < nvalid) atomicAdd(op + i, v[i]);
        } else if (p.vec_ok && nvalid == 32) {
#pragma unroll
          for (int g = 0; g < 8; ++g)
            reinterpret_cast<float4*>(op)[g] = make_float4(v[g * 4 + 0], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < nvalid) op[i] = v[i];
        }
      }
    }
    if constexpr (EPI == 3) {
      // ---- fp32 weight gradients accumulated across split-K CTAs: the 128 x BN tile is staged in shared memory
      // (BN/32 slabs of 128 rows x 128 B, SWIZZLE_128B) and added to global memory by TMA reduce, i.e. one full-line
      // L2 reduction per row and 32 columns instead of 32 scattered fp32 atomics per thread and column.
      // Columns past n_cols hold exact zeros (their B operand was TMA zero fill) and rows past m_rows are clipped
      // by the tensor map, so the whole tile is written unconditionally.
      const bool leader = (warp == 2 && lane == 0);
      const uint32_t rx = (uint32_t)(r & 7);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        if (n0 + c0 >= p.n_cols) break;  // warp-uniform
        uint32_t raw[32];
        if (n_iters > 0) {
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, raw);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) raw[i] = 0u;
        }
        const float sc = p.alpha * rscale;
        uint8_t* slab = smem + (c0 >> 5) * 16384 + r * 128;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          float4 f;
          f.x = __uint_as_float(raw[g * 4 + 0]) * sc;
          f.y = __uint_as_float(raw[g * 4 + 1]) * sc;
          f.z = __uint_as_float(raw[g * 4 + 2]) * sc;
          f.w = __uint_as_float(raw[g * 4 + 3]) * sc;
          *reinterpret_cast<float4*>(slab + ((((uint32_t)g) ^ rx) << 4)) = f;
        }
      }
      fence_proxy_async();
      named_barrier_sync(1, 128);
      if (leader) {
        const int col0 = p.taps[tap_w].col + n0;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32)
          if (n0 + c0 < p.n_cols) tma_reduce_add_4d(&tma_out, smem + (c0 >> 5) * 16384, col0 + c0, m0, by, bn);
        tma_store_commit();
        tma_store_wait_read();
      }
    }
    if constexpr (EPI == 2) {
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (n0 + c0 >= p.n_cols) break;  // warp-uniform
      uint32_t raw[32];
      if (n_iters > 0) {
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, raw);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) raw[i] = 0u;
      }
      if (!row_ok) continue;
      const int ncol = n0 + c0;
      const long long off = row_off + ncol;
      const int nvalid = min(32, p.n_cols - ncol);
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]) * p.alpha;
      if (p.col_scale != nullptr) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < nvalid) v[i] *= __ldg(p.col_scale + ncol + i);
      }
      if (p.col_shift != nullptr) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < nvalid) v[i] += __ldg(p.col_shift + ncol + i);
      }
      if (p.row_scale != nullptr) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= rscale;
      }
      const bool vec = p.vec_ok && nvalid == 32;
      if (p.res != nullptr) {
        if (p.res_dtype == TOIST_BF16) {
          const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.res) + off;
          if (vec) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint4 u = __ldg(reinterpret_cast<const uint4*>(rp) + g);
              const float2 f0 = unpack_bf16(u.x), f1 = unpack_bf16(u.y), f2 = unpack_bf16(u.z), f3 = unpack_bf16(u.w);
              v[g * 8 + 0] += f0.x; v[g * 8 + 1] += f0.y; v[g * 8 + 2] += f1.x; v[g * 8 + 3] += f1.y;
              v[g * 8 + 4] += f2.x; v[g * 8 + 5] += f2.y; v[g * 8 + 6] += f3.x; v[g * 8 + 7] += f3.y;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < nvalid) v[i] += __bfloat162float(rp[i]);
          }
        } else {
          const float* rp = reinterpret_cast<const float*>(p.res) + off;
          if (vec) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const float4 f = __ldg(reinterpret_cast<const float4*>(rp) + g);
              v[g * 4 + 0] += f.x; v[g * 4 + 1] += f.y; v[g * 4 + 2] += f.z; v[g * 4 + 3] += f.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < nvalid) v[i] += rp[i];
          }
        }
      }
      if (p.mask != nullptr) {
        const __nv_bfloat16* mp = reinterpret_cast<const __nv_bfloat16*>(p.mask) + off;
        if (vec) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(mp) + g);
            const float2 f0 = unpack_bf16(u.x), f1 = unpack_bf16(u.y), f2 = unpack_bf16(u.z), f3 = unpack_bf16(u.w);
            if (!(f0.x > 0.f)) v[g * 8 + 0] = 0.f;
            if (!(f0.y > 0.f)) v[g * 8 + 1] = 0.f;
            if (!(f1.x > 0.f)) v[g * 8 + 2] = 0.f;
            if (!(f1.y > 0.f)) v[g * 8 + 3] = 0.f;
            if (!(f2.x > 0.f)) v[g * 8 + 4] = 0.f;
            if (!(f2.y > 0.f)) v[g * 8 + 5] = 0.f;
            if (!(f3.x > 0.f)) v[g * 8 + 6] = 0.f;
            if (!(f3.y > 0.f)) v[g * 8 + 7] = 0.f;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < nvalid && !(__bfloat162float(mp[i]) > 0.f)) v[i] = 0.f;
        }
      }
      if (p.aux != nullptr) {
        __nv_bfloat16* ap = reinterpret_cast<__nv_bfloat16*>(p.aux) + off;
        if (vec) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 u;
            u.x = pack_bf16(v[g * 8 + 0], v[g * 8 + 1]);
            u.y = pack_bf16(v[g * 8 + 2], v[g * 8 + 3]);
            u.z = pack_bf16(v[g * 8 + 4], v[g * 8 + 5]);
            u.w = pack_bf16(v[g * 8 + 6], v[g * 8 + 7]);
            reinterpret_cast<uint4*>(ap)[g] = u;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < nvalid) ap[i] = __float2bfloat16_rn(v[i]);
        }
      }
      if (p.act != TOIST_ACT_NONE) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = apply_act(v[i], p.act);
      }
      if (p.out_dtype == TOIST_BF16) {
        __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + off;
        if (vec) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 u;
            u.x = pack_bf16(v[g * 8 + 0], v[g * 8 + 1]);
            u.y = pack_bf16(v[g * 8 + 2], v[g * 8 + 3]);
            u.z = pack_bf16(v[g * 8 + 4], v[g * 8 + 5]);
            u.w = pack_bf16(v[g * 8 + 6], v[g * 8 + 7]);
            reinterpret_cast<uint4*>(op)[g] = u;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < nvalid) op[i] = __float2bfloat16_rn(v[i]);
        }
      } else {
        float* op = reinterpret_cast<float*>(p.out) + off;
        if (p.accumulate) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < nvalid) atomicAdd(op + i, v[i]);
        } else if (vec) {
#pragma unroll
          for (int g = 0; g < 8; ++g)
            reinterpret_cast<float4*>(op)[g] = make_float4(v[g * 4 + 0], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < nvalid) op[i] = v[i];
        }
      }
    }
    }
  }

  // ---------------- teardown
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) trace_stamp(p, 9);
  if constexpr (kTwoSm) {
    cluster_sync();  // both CTAs are done with the paired TMEM (and with each other's barriers) before it is freed
    if (warp == 1) {
      tc_fence_after();
      tmem_dealloc_2sm(tmem_base, BN);
    }
  } else {
    if (warp == 1) {
      tc_fence_after();
      tmem_dealloc(tmem_base, BN);
    }
    if constexpr (CL > 1) cluster_sync();  // the peer may still signal this CTA's empty barriers: keep the smem alive
  }
}

}  // namespace toist
#include "gemm_persist.cuh"
namespace toist {

// counters of the persistent kernel's tile scheduler: {next tile, finished CTAs} per launch, handed out round robin
// from a caller-provided zeroed device buffer (toist_gemm_set_workspace); each launch re-arms its pair before it ends
// Launches captured into CUDA graphs keep their slot for the life of the graph and may replay next to anything:
// they take slots from the upper half, never reused (when it is used up, captured launches fall back to the
// one-tile kernel); eager launches cycle through the lower half (a slot comes around again after 32768 launches).
static unsigned int* g_persist_ws[16] = {};
static int g_persist_slots[16] = {};
static unsigned int g_persist_next[16] = {};
static unsigned int g_persist_next_graph[16] = {};

static unsigned int* persist_slot(int dev, cudaStream_t stream) {
  const unsigned int half = (unsigned)g_persist_slots[dev] / 2;
  if (half == 0) return nullptr;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &st) != cudaSuccess) return nullptr;
  if (st == cudaStreamCaptureStatusNone) return g_persist_ws[dev] + 2 * (g_persist_next[dev]++ % half);
  if (g_persist_next_graph[dev] >= half) return nullptr;
  return g_persist_ws[dev] + 2 * (half + g_persist_next_graph[dev]++);
}

// TOIST_GEMM_PERSIST=1 sends eligible launches to the persistent kernel.  OFF by default: measured on B200 in the same
// process (tools/profile_kernels.py --graph --inner 10) it loses to the one-tile kernel on every bench shape
// (3x3 256->256 @40^2: 28.0 vs 20.8 us, 1x1 256->1024 + residual: 20.9 vs 18.3 us, FFN1 16.8 vs 11.1 us): a lone CTA
// per SM issues dependent tcgen05.mma into ONE accumulator and those retire no faster than one per ~135 clocks whatever
// N is (k-block of 4 MMAs: 567 / 538 / 605 clocks for N = 64 / 128 / 256, independent of the ring depth), so N < 256
// tiles need two or more co-resident CTAs (independent accumulation chains) to fill the tensor pipe - which the
// one-tile kernel gets for free.  Kept as the starting point for a version that interleaves tiles per CTA.
static bool persist_enabled() {
  static const bool on = []() {
    const char* e = getenv("TOIST_GEMM_PERSIST");
    return e ? atoi(e) != 0 : false;
  }();
  return on;
}

template <int BN, int MODE>
static int launch_persist(const CUtensorMap* maps, GemmKParams& kp, int total_tiles, unsigned int* counter,
                          cudaStream_t stream, int k_iters) {
  using Cfg = PersistCfg<BN>;
  // split of the 224 KB tile area: three rounds of slabs (residual + mask slabs of an output slab are prefetched while
  // two stores are in flight), four when the reduction is short (the epilogue is then the bottleneck and the ring
  // cannot use more than k_iters + 1 stages anyway); the operand ring takes the rest
  const int per = std::max(1, (kp.res != nullptr ? 1 : 0) + (kp.mask != nullptr ? 1 : 0));
  int slabs = std::min(kPMaxSlabs, (k_iters <= 4 ? 4 : 3) * per);
  int stages = std::min(kPMaxStages, (kPTileBytes - slabs * 16384) / Cfg::kStageBytes);
  stages = std::min(stages, std::max(2, k_iters + 1));
  static const int force_stages = []() {  // TOIST_GEMM_PERSIST_STAGES: ring depth override (experiments)
    const char* e = getenv("TOIST_GEMM_PERSIST_STAGES");
    return e ? atoi(e) : 0;
  }();
  if (force_stages > 0) stages = std::max(2, std::min(stages, force_stages));
  slabs = std::min(kPMaxSlabs, (kPTileBytes - stages * Cfg::kStageBytes) / 16384);
  kp.stages = stages;
  kp.slabs = slabs;
  static bool configured = false;
  auto kfn = gemm_persist_kernel<BN, MODE>;
  if (!configured) {
    TOIST_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
    configured = true;
  }
  static const int n_sm = []() {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    const char* e = getenv("TOIST_GEMM_PERSIST_CTAS");
    return e ? atoi(e) : n;
  }();
  const int grid = std::min(total_tiles, n_sm);
  TOIST_CHECK_CUDA(launch_pdl(kfn, dim3((unsigned)grid), dim3(kPThreads), (size_t)Cfg::kSmem, stream, maps[0], maps[1],
                              maps[2], maps[3], maps[4], kp, counter, total_tiles));
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

template <int BN, int MODE, int EPI, int PAIR = 0>
static int launch_gemm(const CUtensorMap* maps, const GemmKParams& kp, dim3 grid, cudaStream_t stream) {
  constexpr int CL = PAIR ? 2 : 1;
  constexpr int kStage = kABytes + (PAIR == 2 ? BN / 2 : BN) * 128;
  constexpr int kMaxStages = PAIR == 2 ? 6 : ((BN == 128) ? 3 : 4);
  constexpr int max_smem = kMaxStages * kStage + 1024 /*align slack*/ + 256 /*barriers*/ + 2048 /*column constants*/;
  const int smem = kp.stages * kStage + 1024 + 256 + 2048;
  static bool configured = false;
  auto kfn = gemm_kernel<BN, MODE, EPI, PAIR>;
  if (!configured) {
    TOIST_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    configured = true;
  }
  if constexpr (CL > 1) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = CL;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    TOIST_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kfn, maps[0], maps[1], maps[2], maps[3], maps[4], kp));
  } else {
    launch_pdl(kfn, dim3(grid), dim3(kThreads), smem, stream, maps[0], maps[1], maps[2], maps[3], maps[4], kp);
  }
  TOIST_CHECK_CUDA(cudaGetLastError());
  return TOIST_OK;
}

template <int BN, int MODE>
static int dispatch_epi(const CUtensorMap* maps, const GemmKParams& kp, dim3 grid, cudaStream_t stream) {
  if constexpr (MODE == TOIST_GEMM_WGRAD) {
    // weight gradients are fp32; the batched dV / dK products of attention write bf16 through the generic path
    if (kp.epi == 3) return launch_gemm<BN, MODE, 3>(maps, kp, grid, stream);
    if (kp.epi == 1) return launch_gemm<BN, MODE, 1>(maps, kp, grid, stream);
    return launch_gemm<BN, MODE, 2>(maps, kp, grid, stream);
  } else {
    if constexpr (BN >= 128) {
      if (kp.epi == 0 && kp.cluster == 2) return launch_gemm<BN, MODE, 0, 1>(maps, kp, grid, stream);
    }
    if constexpr (BN == 256) {
      if (kp.epi == 0 && kp.cluster == 3) return launch_gemm<BN, MODE, 0, 2>(maps, kp, grid, stream);
    }
    switch (kp.epi) {
      case 0: return launch_gemm<BN, MODE, 0>(maps, kp, grid, stream);
      case 1: return launch_gemm<BN, MODE, 1>(maps, kp, grid, stream);
      default: return launch_gemm<BN, MODE, 2>(maps, kp, grid, stream);
    }
  }
}

static int short_k_stages() {  // TOIST_GEMM_SHORTK_STAGES: ring depth for reductions of <= 4 k-blocks (A/B knob, 0 = off)
  static const int v = []() {
    const char* e = getenv("TOIST_GEMM_SHORTK_STAGES");
    return e ? atoi(e) : 2;
  }();
  return v;
}

template <int MODE>
static int dispatch_bn(int bn, const CUtensorMap* maps, GemmKParams& kp, dim3 grid, cudaStream_t stream, int k_iters) {
  // ring depth: 2 CTAs per SM for BN <= 128 (one tile's epilogue overlaps the other's main loop).  A reduction of at
  // most 4 k-blocks (the 1x1 convolutions with Cin <= 256, the attention projections) never has more than 4 stages in
  // flight anyway and is bound by the latency chain of one short-lived CTA (prologue -> loads -> 4 MMAs -> residual
  // tile -> epilogue -> store): a 2-deep ring halves the shared memory per CTA so that twice as many CTAs are resident
  // and overlap each other's chains.
  int sk = short_k_stages();
  static const int short_max = []() {  // TOIST_GEMM_SHORTK_MAX: longest reduction (k-blocks) given the shallow ring
    const char* e = getenv("TOIST_GEMM_SHORTK_MAX");
    return e ? atoi(e) : 8;  // bench step with the two trunk chains: 4 -> 8.775 ms, 8 -> 8.728 ms (same box)
  }();
  const bool short_k = sk > 0 && k_iters <= short_max;
  if (short_k && kp.epi == 0) {  // the bf16 epilogue stages [res = out | mask] tiles of BN * 256 bytes in the idle ring
    const int stage_bytes = kABytes + bn * 128;
    const int staging = bn * 256 * std::max(1, (kp.res != nullptr ? 1 : 0) + (kp.mask != nullptr ? 1 : 0));
    sk = std::max(sk, (staging + stage_bytes - 1) / stage_bytes);
  }
  if (kp.cluster == 3 && bn == 256) {  // 32 KB stages: six of them, the epilogue needs at most four (res + mask tiles)
    static const int st2 = []() {  // TOIST_GEMM_2SM_STAGES: ring depth of the cta_group::2 variant (experiments)
      const char* e = getenv("TOIST_GEMM_2SM_STAGES");
      return e ? std::max(4, std::min(6, atoi(e))) : 6;
    }();
    kp.stages = st2;
    return dispatch_epi<256, MODE>(maps, kp, grid, stream);
  }
  switch (bn) {
    case 256: kp.stages = short_k ? std::min(sk, 4) : 4; return dispatch_epi<256, MODE>(maps, kp, grid, stream);
    case 128: kp.stages = short_k ? std::min(sk, 3) : 3; return dispatch_epi<128, MODE>(maps, kp, grid, stream);
    default:  kp.stages = short_k ? std::min(sk, 4) : 4; return dispatch_epi<64, MODE>(maps, kp, grid, stream);
  }
}

static int g_concurrent = 1;  // toist_gemm_concurrency

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static bool cluster_enabled() {  // TOIST_GEMM_CLUSTER=0 disables the CTA-pair / multicast variant (A/B runs)
  static const bool on = []() {
    const char* e = getenv("TOIST_GEMM_CLUSTER");
    return e ? atoi(e) != 0 : false;
  }();
  return on;
}

static bool two_sm_enabled() {  // TOIST_GEMM_2SM=1: cta_group::2 MMA over CTA pairs (experimental, see gemm_kernel)
  static const bool on = []() {
    const char* e = getenv("TOIST_GEMM_2SM");
    return e ? atoi(e) != 0 : false;
  }();
  return on;
}

static bool reduce_epilogue_enabled() {  // TOIST_GEMM_ATOMIC_WGRAD=1 selects the scattered-atomics epilogue (A/B runs)
  static const bool on = []() {
    const char* e = getenv("TOIST_GEMM_ATOMIC_WGRAD");
    return !(e && atoi(e) != 0);
  }();
  return on;
}

}  // namespace toist

using namespace toist;

extern "C" int toist_gemm_concurrency(int chains) {
  const int prev = g_concurrent;
  g_concurrent = chains < 1 ? 1 : (chains > 8 ? 8 : chains);
  return prev;
}

extern "C" int toist_gemm_set_workspace(void* zeroed, int64_t bytes) {
  int dev = 0;
  TOIST_CHECK_CUDA(cudaGetDevice(&dev));
  TOIST_REQUIRE(dev >= 0 && dev < 16, "toist_gemm_set_workspace: device index %d out of range", dev);
  TOIST_REQUIRE(zeroed == nullptr || bytes >= 8, "toist_gemm_set_workspace: need at least 8 bytes");
  g_persist_ws[dev] = reinterpret_cast<unsigned int*>(zeroed);
  g_persist_slots[dev] = zeroed ? (int)std::min<int64_t>(bytes / 8, 1 << 20) : 0;
  g_persist_next[dev] = 0;
  g_persist_next_graph[dev] = 0;
  return TOIST_OK;
}

extern "C" int toist_gemm(const toist_gemm_desc* d, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  if (skip_gemm()) return TOIST_OK;
  TOIST_REQUIRE(d != nullptr, "toist_gemm: null descriptor");
  TOIST_REQUIRE(d->mode >= 0 && d->mode <= 2, "toist_gemm: bad mode %d", d->mode);
  TOIST_REQUIRE(d->out != nullptr && d->a.ptr != nullptr && d->b.ptr != nullptr, "toist_gemm: null tensor");
  TOIST_REQUIRE(d->n_taps >= (d->mode == TOIST_GEMM_WGRAD ? 1 : 0) && d->n_taps <= TOIST_MAX_TAPS,
                "toist_gemm: n_taps %d out of range", d->n_taps);
  TOIST_REQUIRE(d->tile_x >= 1 && d->tile_y >= 1 && d->tile_n >= 1, "toist_gemm: bad tile");
  TOIST_REQUIRE(d->stride_x >= 1 && d->stride_x <= 8 && d->stride_y >= 1 && d->stride_y <= 8, "toist_gemm: bad stride");
  TOIST_REQUIRE(d->n_cols >= 1, "toist_gemm: n_cols must be positive");
  TOIST_REQUIRE(d->ext_x >= 1 && d->ext_y >= 1 && d->ext_n >= 1, "toist_gemm: empty pixel space");
  TOIST_REQUIRE(!d->accumulate || d->out_dtype == TOIST_F32, "toist_gemm: accumulate needs an f32 output");
  const int tile_rows = d->tile_x * d->tile_y * d->tile_n;

  GemmKParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.ext_x = d->ext_x; kp.ext_y = d->ext_y; kp.ext_n = d->ext_n;
  kp.tile_x = d->tile_x; kp.tile_y = d->tile_y; kp.tile_n = d->tile_n;
  kp.tiles_x = (int)ceil_div(d->ext_x, d->tile_x);
  kp.tiles_y = (int)ceil_div(d->ext_y, d->tile_y);
  kp.tiles_n = (int)ceil_div(d->ext_n, d->tile_n);
  kp.stride_x = d->stride_x; kp.stride_y = d->stride_y;
  kp.n_cols = d->n_cols; kp.m_rows = d->m_rows;
  kp.n_taps = d->n_taps;
  kp.b_batched = d->b_batched;
  kp.batch_y = d->batch_y > 0 ? d->batch_y : 1;
  kp.batch_n = d->batch_n > 0 ? d->batch_n : 1;
  kp.splits = d->splits > 0 ? d->splits : 1;
  kp.out = d->out; kp.out_dtype = d->out_dtype;
  kp.out_sx = d->out_sx; kp.out_sy = d->out_sy; kp.out_sn = d->out_sn;
  kp.alpha = d->alpha;
  kp.col_scale = d->col_scale; kp.col_shift = d->col_shift; kp.row_scale = d->row_scale;
  kp.res = d->res; kp.res_dtype = d->res_dtype; kp.mask = d->mask; kp.aux = d->aux;
  kp.act = d->act; kp.accumulate = d->accumulate;
  for (int i = 0; i < d->n_taps; ++i) kp.taps[i] = d->taps[i];
  kp.trace = gemm_trace_buffer();

  // 16-byte vector path: every row start and every 32-column group must be 16-byte aligned in all streams.
  bool vec = (d->out_sx % 8 == 0) && (d->out_sy % 8 == 0) && (d->out_sn % 8 == 0) && aligned16(d->out) &&
             (d->res == nullptr || aligned16(d->res)) && (d->mask == nullptr || aligned16(d->mask)) &&
             (d->aux == nullptr || aligned16(d->aux));
  if (d->mode == TOIST_GEMM_WGRAD)
    for (int i = 0; i < d->n_taps; ++i) vec = vec && (d->taps[i].col % 8 == 0);
  kp.vec_ok = vec ? 1 : 0;

  // pick the column-tile width: the widest tile that still yields enough CTAs to cover the 148 SMs
  const int64_t m_tiles = (d->mode == TOIST_GEMM_WGRAD) ? ceil_div(d->m_rows, kBM)
                                                         : (int64_t)kp.tiles_x * kp.tiles_y * kp.tiles_n;
  const int64_t z_mult = (d->mode == TOIST_GEMM_WGRAD) ? (int64_t)kp.batch_y * kp.batch_n * kp.splits * d->n_taps : 1;
  int bn = 64;
  const int cands[3] = {256, 128, 64};
  static const int min_ctas_env = []() {  // tuning knob: smallest grid for which a wider column tile is preferred
    const char* e = getenv("TOIST_GEMM_MIN_CTAS");
    return e ? atoi(e) : 160;  // measured on the bench step: 96 -> 9.55, 130 -> 9.47, 160 -> 9.32, 210 -> 9.83 ms
  }();
  // the caller runs `g_concurrent` independent launch chains side by side (toist_gemm_concurrency): each launch needs
  // only its share of the SMs
  const int min_ctas = std::max(32, min_ctas_env / std::max(1, g_concurrent));
  // short reductions are epilogue bound: keep two CTAs per SM (BN <= 128) so epilogues overlap main loops
  const int64_t k_iters = (d->mode == TOIST_GEMM_WGRAD) ? 1 << 20 : (int64_t)ceil_div(d->k_per_tap, kBK) * d->n_taps;
  static const int bn256_min_k = []() {  // TOIST_GEMM_BN256_MINK: shortest reduction (k-blocks) that may use 256-wide tiles
    const char* e = getenv("TOIST_GEMM_BN256_MINK");
    return e ? atoi(e) : 8;
  }();
  for (int c = 0; c < 3; ++c) {
    const int cand = cands[c];
    if (cand > 64 && d->n_cols <= cand / 2) continue;  // more than half the tile would be padding
    if (cand == 256 && k_iters < bn256_min_k) continue;
    const int64_t ctas = m_tiles * ceil_div(d->n_cols, cand) * z_mult;
    if (ctas >= min_ctas || cand == 64) {
      bn = cand;
      break;
    }
  }
  const int n_tiles = (int)ceil_div(d->n_cols, bn);
  kp.n_tiles = n_tiles;

  // epilogue flavour: fp32 outputs without residual / mask / aux take the compact direct path, everything unusual the
  // generic one; aligned bf16 outputs are switched to the TMA path below
  kp.epi = (d->out_dtype == TOIST_F32 && d->res == nullptr && d->mask == nullptr && d->aux == nullptr &&
            d->act != TOIST_ACT_GELU) ? 1 : 2;
  CUtensorMap maps[5];
  memset(maps, 0, sizeof(maps));
  CUtensorMap& ma = maps[0];
  CUtensorMap& mb = maps[1];
  uint32_t ones[4] = {1, 1, 1, 1};
  int rc;
  if (d->mode != TOIST_GEMM_WGRAD) {
    // bf16 outputs go through the shared-memory / TMA epilogue (full-line stores, res / mask tiles by TMA load)
    if (vec && d->out_dtype == TOIST_BF16 && (d->res == nullptr || d->res_dtype == TOIST_BF16) && d->aux == nullptr &&
        !d->accumulate && d->row_scale == nullptr && (d->act == TOIST_ACT_NONE || d->act == TOIST_ACT_RELU)) {
      const int64_t odim[4] = {d->n_cols, d->ext_x, d->ext_y, d->ext_n};
      const int64_t ostr[4] = {1, d->out_sx, d->out_sy, d->out_sn};
      const uint32_t obox[4] = {64, (uint32_t)d->tile_x, (uint32_t)d->tile_y, (uint32_t)d->tile_n};
      if ((rc = encode_tmap_bf16_4d(&maps[2], d->out, odim, ostr, obox, ones)) != TOIST_OK) return rc;
      if (d->res && (rc = encode_tmap_bf16_4d(&maps[3], d->res, odim, ostr, obox, ones)) != TOIST_OK) return rc;
      if (d->mask && (rc = encode_tmap_bf16_4d(&maps[4], d->mask, odim, ostr, obox, ones)) != TOIST_OK) return rc;
      kp.epi = 0;
    }
    TOIST_REQUIRE(tile_rows == kBM, "toist_gemm: FWD/DGRAD pixel tile must hold 128 rows (got %d)", tile_rows);
    TOIST_REQUIRE(d->k_per_tap >= 1, "toist_gemm: k_per_tap must be positive");
    kp.kblocks = (int)ceil_div(d->k_per_tap, kBK);
    uint32_t abox[4] = {64, (uint32_t)(d->tile_x * d->stride_x), (uint32_t)(d->tile_y * d->stride_y),
                        (uint32_t)d->tile_n};
    uint32_t aes[4] = {1, (uint32_t)d->stride_x, (uint32_t)d->stride_y, 1};
    if ((rc = encode_tmap_bf16_4d(&ma, d->a.ptr, d->a.dim, d->a.stride, abox, aes)) != TOIST_OK) return rc;
    // CTA pairs with a multicast B tile for the long, L2-feed-bound reductions (see gemm_kernel)
    kp.cluster = (cluster_enabled() && kp.epi == 0 && !d->b_batched && bn >= 128 && m_tiles % 2 == 0 && k_iters >= 8) ? 2 : 1;
    if (two_sm_enabled() && kp.epi == 0 && !d->b_batched && d->n_cols >= 256 && m_tiles % 2 == 0 && k_iters >= 8) {
      kp.cluster = 3;  // cta_group::2 pairs, always with 256-wide tiles
      bn = 256;
      kp.n_tiles = (int)ceil_div(d->n_cols, bn);
    }
    uint32_t bbox_fwd[4] = {64, (uint32_t)(kp.cluster > 1 ? bn / 2 : bn), 1, 1};
    uint32_t bbox_dg[4] = {64, 64, 1, 1};
    if ((rc = encode_tmap_bf16_4d(&mb, d->b.ptr, d->b.dim, d->b.stride,
                                  d->mode == TOIST_GEMM_FWD ? bbox_fwd : bbox_dg, ones)) != TOIST_OK)
      return rc;
    dim3 grid((unsigned)m_tiles, (unsigned)kp.n_tiles, 1);
    // persistent kernel: bf16 / TMA epilogue, plain (non-batched, non-clustered) launches whose per-column constants
    // can be read as aligned float4 groups; needs the scheduler's counter workspace
    int dev = 0;
    if (persist_enabled() && kp.epi == 0 && kp.cluster <= 1 && !d->b_batched && k_iters >= 1 && d->n_cols % 16 == 0 &&
        aligned16(d->col_scale) && aligned16(d->col_shift) && cudaGetDevice(&dev) == cudaSuccess && dev < 16 &&
        g_persist_ws[dev] != nullptr) {
      const int64_t total = m_tiles * kp.n_tiles;
      static const int min_tiles = []() {  // TOIST_GEMM_PERSIST_MIN: smallest tile count sent to the persistent kernel
        const char* e = getenv("TOIST_GEMM_PERSIST_MIN");
        return e ? atoi(e) : 1;
      }();
      unsigned int* ctr = (total >= min_tiles && total < (1 << 30)) ? persist_slot(dev, stream) : nullptr;
      if (ctr != nullptr) {
        const bool fwd = d->mode == TOIST_GEMM_FWD;
        switch (bn) {
          case 256: return fwd ? launch_persist<256, TOIST_GEMM_FWD>(maps, kp, (int)total, ctr, stream, (int)k_iters)
                               : launch_persist<256, TOIST_GEMM_DGRAD>(maps, kp, (int)total, ctr, stream, (int)k_iters);
          case 128: return fwd ? launch_persist<128, TOIST_GEMM_FWD>(maps, kp, (int)total, ctr, stream, (int)k_iters)
                               : launch_persist<128, TOIST_GEMM_DGRAD>(maps, kp, (int)total, ctr, stream, (int)k_iters);
          default:  return fwd ? launch_persist<64, TOIST_GEMM_FWD>(maps, kp, (int)total, ctr, stream, (int)k_iters)
                               : launch_persist<64, TOIST_GEMM_DGRAD>(maps, kp, (int)total, ctr, stream, (int)k_iters);
        }
      }
    }
    if (d->mode == TOIST_GEMM_FWD) return dispatch_bn<TOIST_GEMM_FWD>(bn, maps, kp, grid, stream, (int)k_iters);
    return dispatch_bn<TOIST_GEMM_DGRAD>(bn, maps, kp, grid, stream, (int)k_iters);
  }
  TOIST_REQUIRE(tile_rows == kBK, "toist_gemm: WGRAD pixel tile must hold 64 rows (got %d)", tile_rows);
  TOIST_REQUIRE(d->m_rows >= 1, "toist_gemm: WGRAD needs m_rows");
  const int64_t total_tiles = (int64_t)kp.tiles_x * kp.tiles_y * kp.tiles_n;
  if (kp.splits > total_tiles) kp.splits = (int)total_tiles;
  // every split must own at least one pixel tile
  while (kp.splits > 1 && (int64_t)(kp.splits - 1) * ceil_div(total_tiles, kp.splits) >= total_tiles) --kp.splits;
  TOIST_REQUIRE(kp.splits == 1 || d->accumulate, "toist_gemm: WGRAD splits > 1 requires accumulate");
  uint32_t abox[4] = {64, (uint32_t)d->tile_x, (uint32_t)d->tile_y, (uint32_t)d->tile_n};
  if ((rc = encode_tmap_bf16_4d(&ma, d->a.ptr, d->a.dim, d->a.stride, abox, ones)) != TOIST_OK) return rc;
  uint32_t bbox[4] = {64, (uint32_t)(d->tile_x * d->stride_x), (uint32_t)(d->tile_y * d->stride_y),
                      (uint32_t)d->tile_n};
  uint32_t bes[4] = {1, (uint32_t)d->stride_x, (uint32_t)d->stride_y, 1};
  if ((rc = encode_tmap_bf16_4d(&mb, d->b.ptr, d->b.dim, d->b.stride, bbox, bes)) != TOIST_OK) return rc;
  // accumulating fp32 weight gradients: TMA reduce-add epilogue when every row / tap offset is 16-byte aligned
  if (kp.epi == 1 && d->accumulate && vec && d->out_sx % 4 == 0 && d->act == TOIST_ACT_NONE &&
      d->col_scale == nullptr && d->col_shift == nullptr && reduce_epilogue_enabled()) {
    bool taps_ok = true;
    for (int i = 0; i < d->n_taps; ++i) taps_ok = taps_ok && (d->taps[i].col % 4 == 0);
    if (taps_ok) {
      const int64_t odim[4] = {d->out_sx, d->m_rows, kp.batch_y, kp.batch_n};
      const int64_t ostr[4] = {1, d->out_sx, d->out_sy, d->out_sn};
      const uint32_t obox[4] = {32, 128, 1, 1};
      if ((rc = encode_tmap_f32_4d(&maps[2], d->out, odim, ostr, obox, ones)) != TOIST_OK) return rc;
      kp.epi = 3;
    }
  }
  dim3 grid((unsigned)m_tiles, (unsigned)(n_tiles * d->n_taps), (unsigned)(kp.batch_y * kp.batch_n * kp.splits));
  return dispatch_bn<TOIST_GEMM_WGRAD>(bn, maps, kp, grid, stream, 1 << 20);
}
