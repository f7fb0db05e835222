// Persistent, warp-specialised variant of the implicit-GEMM engine (FWD / DGRAD, bf16 output through TMA):
// one CTA per SM walks output tiles handed out by a global counter; operand loads, tcgen05.mma and the epilogue of
// consecutive tiles overlap (two TMEM accumulators), eight epilogue warps share a tile, residual / ReLU-mask slabs
// are prefetched by their own warp into a ring of 16 KB slabs that doubles as the staging of the TMA stores.
//
// Why (measured with toist_debug_gemm_trace on the one-tile-per-CTA kernel, B200): a 128 x 128 tile of the layer3
// 256 -> 1024 convolution lives 13.7 k clocks of which the main loop is 3.0 k and the epilogue 6.3 k (one epilogue warp
// per scheduler, every dependent instruction exposed); setup + dependency wait + first operand latency are another
// 2.6 k per tile.  The short-reduction shapes (every 1 x 1 convolution with Cin <= 512, the stem, the transformer's
// linear layers: ~40 % of the step's GEMM launches) are therefore epilogue- and latency-bound, not tensor-bound.
// Here the per-tile cost in steady state is max(main loop, epilogue / 2) and the fixed costs are paid once per CTA.
//
// Warp roles (384 threads): 0 = tile scheduler + TMA producer of the A operand, 11 = TMA producer of the B operand,
// 1 = TMEM owner + MMA issuer, 2 = slab loader (residual / mask tiles by TMA), 3..10 = epilogue (warp w reads TMEM lane
// quadrant w % 4; warps 3-6 take the first 32 columns of every 64-column slab, warps 7-10 the second).
#pragma once

namespace toist {

constexpr int kPThreads = 384;
constexpr int kPQueue = 16;  // tile-id slots; the producer is never more than ring depth + 3 tiles ahead of the epilogue

constexpr int kPMaxStages = 8, kPMaxSlabs = 8;
constexpr int kPTileBytes = 220 * 1024;  // operand ring + slab ring, split per launch (GemmKParams.stages / .slabs)

template <int BN>
struct PersistCfg {
  static constexpr int kStageBytes = kABytes + BN * 128;
  static constexpr int kSmem = kPTileBytes + 512 /*barriers, queue*/ + 4 * BN * 4 /*column constants*/ + 1024 /*align slack*/;
};

template <int BN, int MODE>
__global__ void __launch_bounds__(kPThreads, 1)
gemm_persist_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    const __grid_constant__ CUtensorMap tma_out, const __grid_constant__ CUtensorMap tma_res,
                    const __grid_constant__ CUtensorMap tma_mask, const __grid_constant__ GemmKParams p,
                    unsigned int* __restrict__ counter, int total_tiles) {
  using Cfg = PersistCfg<BN>;
  constexpr int kStageBytes = Cfg::kStageBytes;
  constexpr bool kBMN = (MODE != TOIST_GEMM_FWD);
  const int S = p.stages;              // operand ring depth (host: as deep as the slab ring allows)
  const uint32_t NS = (uint32_t)p.slabs;  // slab ring: >= 3 x (residual + mask slabs per output slab, at least 1)

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* slabs = smem + kPTileBytes - NS * 16384;  // (1024-byte aligned: the slab ring sits at the end of the tile area)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kPTileBytes);
  uint64_t* op_full = bars;                           // [S]
  uint64_t* op_empty = op_full + kPMaxStages;         // [S]
  uint64_t* acc_full = op_empty + kPMaxStages;        // [2]
  uint64_t* acc_empty = acc_full + 2;                 // [2]
  uint64_t* slab_full = acc_empty + 2;                // [NS]
  uint64_t* slab_empty = slab_full + kPMaxSlabs;      // [NS]
  uint64_t* q_full = slab_empty + kPMaxSlabs;         // [kPQueue]
  int* tile_q = reinterpret_cast<int*>(q_full + kPQueue);  // [kPQueue]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tile_q + kPQueue);
  float* s_col = reinterpret_cast<float*>(smem + kPTileBytes + 512);  // [2 accumulators][scale | shift][BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_iters = p.n_taps * p.kblocks;
  const bool has_res = p.res != nullptr, has_mask = p.mask != nullptr;
  // trace slots (toist_debug_gemm_trace): 0 entry, 1 setup done, 2 dependency wait done, 3 first operands landed,
  // 4 last MMA issued, 5 first accumulator complete, 6 last store read, 7 = number of tiles this CTA processed
  if (threadIdx.x == 0) {
    trace_stamp(p, 0);
    trace_stamp(p, 8);
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    tma_prefetch_desc(&tma_out);
    for (int s = 0; s < S; ++s) {
      mbar_init(&op_full[s], 2);  // the A producer and the B producer each announce their own bytes
      mbar_init(&op_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 8);  // one arrival per epilogue warp
    }
    for (uint32_t b = 0; b < NS; ++b) {
      mbar_init(&slab_full[b], 1);
      mbar_init(&slab_empty[b], 1);
    }
    for (int b = 0; b < kPQueue; ++b) mbar_init(&q_full[b], 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 2 * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) trace_stamp(p, 1);
  pdl_wait();  // everything above overlaps the previous kernel's tail; nothing below may precede its completion
  if (threadIdx.x == 0) trace_stamp(p, 2);

  // tile t -> (pixel tile, column tile); column tiles fastest: CTAs running side by side share the A rows in L2
  auto decode = [&](int t, int& x0, int& y0, int& i0, int& n0) {
    const int tn = t % p.n_tiles;
    int tm = t / p.n_tiles;
    const int tx = tm % p.tiles_x;
    tm /= p.tiles_x;
    const int ty = tm % p.tiles_y;
    const int tz = tm / p.tiles_y;
    x0 = tx * p.tile_x;
    y0 = ty * p.tile_y;
    i0 = tz * p.tile_n;
    n0 = tn * BN;
  };
  auto next_tile = [&](int i) -> int {  // consumers: the i-th tile of this CTA (-1: no more)
    mbar_wait(&q_full[i % kPQueue], (uint32_t)((i / kPQueue) & 1));
    return *reinterpret_cast<volatile int*>(&tile_q[i % kPQueue]);
  };

  if (warp == 0) {
    // ======================= scheduler + operand producer (one thread) =======================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0;; ++i) {
        int t = (i == 0) ? (int)blockIdx.x : (int)(atomicAdd(counter, 1u) + gridDim.x);
        if (t >= total_tiles) t = -1;
        *reinterpret_cast<volatile int*>(&tile_q[i % kPQueue]) = t;
        mbar_arrive(&q_full[i % kPQueue]);  // (release at CTA scope: the slot is visible before the phase completes)
        if (t < 0) break;
        int x0, y0, i0, n0;
        decode(t, x0, y0, i0, n0);
        for (int it = 0; it < n_iters; ++it) {
          mbar_wait(&op_empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStageBytes;
          mbar_expect_tx(&op_full[stage], kABytes);
          const int tp_i = it / p.kblocks;
          const int kb = it - tp_i * p.kblocks;
          const toist_tap tp = p.taps[tp_i];
          tma_load_4d(sa, &tma_a, &op_full[stage], kb * kBK, x0 * p.stride_x + tp.dx, y0 * p.stride_y + tp.dy, i0 + tp.dn);
          if (++stage == S) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 11) {
    // ======================= B (weight) producer: a second thread issuing TMA loads, see the A producer =======================
    // One thread issuing every box of a stage sustained ~59 B/clk (32 KB stages landing every ~540 clocks, measured
    // with toist_debug_gemm_trace), two co-resident CTAs of the one-tile kernel together twice that: the issue rate
    // of a single thread limits the ingest, not the SM.
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0;; ++i) {
        const int t = next_tile(i);
        if (t < 0) break;
        int x0, y0, i0, n0;
        decode(t, x0, y0, i0, n0);
        for (int it = 0; it < n_iters; ++it) {
          mbar_wait(&op_empty[stage], phase ^ 1);
          uint8_t* sb = smem + stage * kStageBytes + kABytes;
          mbar_expect_tx(&op_full[stage], BN * 128);
          const int tp_i = it / p.kblocks;
          const int kb = it - tp_i * p.kblocks;
          const toist_tap tp = p.taps[tp_i];
          if constexpr (MODE == TOIST_GEMM_FWD) {
            tma_load_4d(sb, &tma_b, &op_full[stage], tp.col + kb * kBK, n0, 0, 0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_4d(sb + j * 8192, &tma_b, &op_full[stage], tp.col + n0 + j * 64, kb * kBK, 0, 0);
          }
          if (++stage == S) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    constexpr uint32_t idesc = umma_idesc_bf16(BN, false, kBMN);
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0;; ++i) {
      const int t = next_tile(i);
      if (t < 0) break;
      const int buf = i & 1;
      mbar_wait(&acc_empty[buf], (uint32_t)(((i >> 1) & 1) ^ 1));  // the epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
      for (int it = 0; it < n_iters; ++it) {
        mbar_wait(&op_full[stage], phase);
        tc_fence_after();
        if (i == 0 && it == 0 && lane == 0) trace_stamp(p, 3);
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint32_t sb = sa + kABytes;
#pragma unroll
          for (int ks = 0; ks < kBK / 16; ++ks) {
            const uint64_t da = umma_smem_desc(sa + ks * 32, 16, 1024);
            const uint64_t db = kBMN ? umma_smem_desc(sb + ks * 2048, 8192, 1024) : umma_smem_desc(sb + ks * 32, 16, 1024);
            umma_f16(d_tmem, da, db, idesc, (it > 0 || ks > 0) ? 1u : 0u);
          }
          umma_commit(&op_empty[stage]);
          if (it == n_iters - 1) umma_commit(&acc_full[buf]);
        }
        __syncwarp();
        if (++stage == S) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    if (lane == 0) trace_stamp(p, 4);
    pdl_trigger();  // this CTA starts no more main loops: the next kernel's CTAs may be scheduled as SMs free up
  } else if (warp == 2) {
    // ======================= slab loader: residual / mask slabs of every tile, in the epilogue's order =======================
    if (elect_one()) {
      uint32_t cnt = 0;  // slab buffers handed out so far (same sequence as the epilogue's)
      auto acquire = [&]() -> uint32_t {
        const uint32_t b = cnt % NS;
        mbar_wait(&slab_empty[b], (uint32_t)(((cnt / NS) & 1) ^ 1));
        ++cnt;
        return b;
      };
      for (int i = 0;; ++i) {
        const int t = next_tile(i);
        if (t < 0) break;
        int x0, y0, i0, n0;
        decode(t, x0, y0, i0, n0);
        const int nsl = min(BN / 64, (p.n_cols - n0 + 63) >> 6);
        for (int j = 0; j < nsl; ++j) {
          if (has_res) {
            const uint32_t b = acquire();
            mbar_expect_tx(&slab_full[b], 16384u);
            tma_load_4d(slabs + b * 16384, &tma_res, &slab_full[b], n0 + j * 64, x0, y0, i0);
          }
          if (has_mask) {
            const uint32_t b = acquire();
            mbar_expect_tx(&slab_full[b], 16384u);
            tma_load_4d(slabs + b * 16384, &tma_mask, &slab_full[b], n0 + j * 64, x0, y0, i0);
          }
          if (!has_res && !has_mask) {
            const uint32_t b = acquire();
            mbar_arrive(&slab_full[b]);  // a free buffer for the output slab, nothing to load
          }
        }
      }
    }
  } else {
    // ======================= epilogue (8 warps) =======================
    const int q = warp & 3;
    const int hf = (warp - 3) >> 2;   // column half of each 64-column slab
    const int r = q * 32 + lane;      // row of the tile = TMEM lane
    const bool leader = (warp == 3 && lane == 0);
    const uint32_t rx = (uint32_t)(r & 7);
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool relu = p.act == TOIST_ACT_RELU;
    const bool has_scale = p.col_scale != nullptr || p.alpha != 1.f;
    uint32_t cnt = 0;
    int n_done = 0;
    int pending0 = -1, pending1 = -1;  // slab buffers of the last two TMA stores (leader only), oldest first
    for (int i = 0;; ++i) {
      const int t = next_tile(i);
      if (t < 0) break;
      int x0, y0, i0, n0;
      decode(t, x0, y0, i0, n0);
      const int buf = i & 1;
      const int nsl = min(BN / 64, (p.n_cols - n0 + 63) >> 6);
      // per-column constants of this tile -> shared memory while its main loop is still running (alpha folded into the
      // scale; columns past n_cols are clipped by the TMA store: clamp, do not branch).  The previous user of this
      // half of s_col is tile i - 2, whose epilogue every warp has left (the named barriers of tile i - 1 lie between).
      float* s_scale = s_col + buf * 2 * BN;
      float* s_shift = s_scale + BN;
      {
        const int last = p.n_cols - 1;
        for (int c = (int)threadIdx.x - 96; c < BN; c += 256) {
          const int col = min(n0 + c, last);
          s_scale[c] = p.alpha * (p.col_scale != nullptr ? __ldg(p.col_scale + col) : 1.f);
          s_shift[c] = p.col_shift != nullptr ? __ldg(p.col_shift + col) : 0.f;
        }
      }
      mbar_wait(&acc_full[buf], (uint32_t)((i >> 1) & 1));
      tc_fence_after();
      named_barrier_sync(2, 256);  // s_col of this tile is complete
      if (i == 0 && leader) trace_stamp(p, 5);
      n_done = i + 1;
      const uint32_t t_addr = lane_base + (uint32_t)(buf * BN + hf * 32);
      uint32_t raw_a[16], raw_b[16];
      tmem_ld_32x16(t_addr, raw_a);  // first 16 columns of slab 0
      for (int j = 0; j < nsl; ++j) {
        uint32_t b_res = 0, b_mask = 0, b_out;
        if (has_res) {
          b_res = cnt % NS;
          mbar_wait(&slab_full[b_res], (uint32_t)((cnt / NS) & 1));
          ++cnt;
        }
        if (has_mask) {
          b_mask = cnt % NS;
          mbar_wait(&slab_full[b_mask], (uint32_t)((cnt / NS) & 1));
          ++cnt;
        }
        if (!has_res && !has_mask) {
          b_out = cnt % NS;
          mbar_wait(&slab_full[b_out], (uint32_t)((cnt / NS) & 1));
          ++cnt;
        } else {
          b_out = has_res ? b_res : b_mask;  // the output is written in place over the residual (or mask) slab
        }
        uint8_t* st_res = slabs + b_res * 16384;
        uint8_t* st_mask = slabs + b_mask * 16384;
        uint8_t* st_out = slabs + b_out * 16384;
        auto process = [&](uint32_t (&raw)[16], int half) {  // 16 columns: raw accumulators -> bf16 in the output slab
          const int lc = j * 64 + hf * 32 + half * 16;  // column within the tile
          const float4* sc4 = reinterpret_cast<const float4*>(s_scale + lc);
          const float4* sh4 = reinterpret_cast<const float4*>(s_shift + lc);
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
          const uint32_t row_base = (uint32_t)r * 128u;
          const uint32_t cb = (uint32_t)(hf * 4 + half * 2);  // first 16-byte chunk of this 16-column group
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
            for (int k = 0; k < 16; ++k) v[k] = fmaxf(v[k], 0.f);
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
        // TMEM loads run one 16-column group ahead of the arithmetic
        tmem_ld_wait();
        tmem_ld_32x16(t_addr + (uint32_t)(j * 64 + 16), raw_b);
        process(raw_a, 0);
        tmem_ld_wait();
        if (j + 1 < nsl) {
          tmem_ld_32x16(t_addr + (uint32_t)((j + 1) * 64), raw_a);
        } else {
          // last TMEM read of this tile by this warp: hand the accumulator back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
        process(raw_b, 1);
        fence_proxy_async();          // generic-proxy writes -> visible to the TMA store
        named_barrier_sync(1, 256);   // the eight epilogue warps: the slab is complete, res / mask fully read
        if (leader) {
          if (pending0 >= 0) {  // the store issued two slabs ago has long finished reading its buffer: recycle it
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            mbar_arrive(&slab_empty[pending0]);
          }
          tma_store_4d(&tma_out, st_out, n0 + j * 64, x0, y0, i0);
          tma_store_commit();
          pending0 = pending1;
          pending1 = (int)b_out;
          if (has_res && has_mask) mbar_arrive(&slab_empty[b_mask]);  // only threads read the mask slab
        }
      }
    }
    if (leader) {
      tma_store_wait_read();
      trace_stamp(p, 6);
      if (p.trace != nullptr) {
        const long long cta = (long long)blockIdx.x;
        p.trace[cta * 16 + 7] = n_done;
      }
    }
  }

  // ---------------- teardown
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) trace_stamp(p, 9);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
  if (threadIdx.x == 0) {
    // every CTA has fetched its last tile before it gets here: the last one to finish re-arms the counters
    __threadfence();
    const unsigned int done = atomicAdd(counter + 1, 1u);
    if (done == gridDim.x - 1) {
      counter[0] = 0u;
      counter[1] = 0u;
      __threadfence();
    }
  }
}

}  // namespace toist
